"""Worker of tests/test_superkmer_emul_cpu.py::test_superkmer_routing_gloo (torch.distributed.run, backend gloo): the protocol
of the sharded insert with super-k-mers on the wire (comm.cu pmap_insert_superkmers) with one process per rank and no GPU --
every rank cuts ITS slice of the reads into records (the g++ build of csrc/superkmer.cuh), the records travel to their
minimizer owners, every rank inserts what it received (the oracle's FreqFilter.add over the records as reads).  Checks: the
shards are disjoint, each holds exactly the keys the library's ownership rule (gb_owner_of_minimizer) gives it, and their
union with counts equals the single-map result; then every shard is filtered and the sharded Graph.buildGraph (the g++ build of
csrc/sgraph.cuh over the gloo fabric of tests/gloo_sgraph_worker.py) runs straight from the minimizer-owned shards and must give
the oracle's graph of all reads: the whole multi-GPU path, reads to graph, with processes instead of GPUs."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genome_b200 import capi  # noqa: E402
from genome_b200.dnamap import PairedEndData  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_sgraph_emul_cpu import load_emul, ptr  # noqa: E402
from tests.test_superkmer_emul_cpu import split, to_ragged_bin  # noqa: E402
from tests.gloo_sgraph_worker import GlooFabric, run_build  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = load_emul()
    fab = GlooFabric(rank, world, "sk" + os.environ.get("MASTER_PORT", "0"))
    for k, read_len in ((31, 100), (21, 80)):
        b, n, _ = H.small_reads(30000, read_len, 10, 0.01, seed=40 + k)   # same bytes on every rank
        mine = PairedEndData(b, n // 2).shard(rank, world)
        rec_bytes = 1 + (read_len + 3) // 4
        w, per_owner, recs = split(lib, mine.bin, mine.n_reads, rec_bytes, k, world)
        at, send = 0, []
        for o in range(world):
            send.append(recs[at:at + per_owner[o]].copy())
            at += per_owner[o]
        gathered = [None] * world
        dist.all_gather_object(gathered, send)
        got = np.concatenate([gathered[src][rank] for src in range(world)])
        shard = pyoracle.OracleMap(k)
        shard.insert_reads(to_ragged_bin(got), got.shape[0])
        sk, sv = shard.export_sorted()
        own = np.zeros(sk.size, np.int32)
        capi.check(capi.lib().gb_owner_of_minimizer(ptr(np.ascontiguousarray(sk)), sk.size, k, world, ptr(own)))
        assert np.all(own == rank), "a key landed on a rank that does not own it"
        whole, ow = H.oracle_counts(b, n, k)
        ok, ov = whole.export_sorted()
        all_own = np.zeros(ok.size, np.int32)
        capi.check(capi.lib().gb_owner_of_minimizer(ptr(np.ascontiguousarray(ok)), ok.size, k, world, ptr(all_own)))
        sel = all_own == rank
        assert np.array_equal(sk, ok[sel]) and np.array_equal(sv, ov[sel])
        windows = [None] * world
        dist.all_gather_object(windows, w)
        assert sum(windows) == ow
        # the whole sharded path: filter every shard, then the sharded Graph.buildGraph straight from the minimizer-owned shards
        # (its re-routing finds every key at home) -- against the oracle's graph of all reads
        shard.delete_below(2)
        whole.delete_below(2)
        kept, _ = shard.export()
        og = pyoracle.OracleGraph(whole)
        got, counts, _ = run_build(lib, fab, k, False, kept, rank, world)
        assert counts == og.counts(), (rank, k, counts, og.counts())
        assert got == H.canon_oracle_graph(og), (rank, k)
    fab.close()
    dist.barrier()
    if rank == 0:
        print("SUPERKMER ROUTING OK world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
