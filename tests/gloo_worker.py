"""Worker of tests/test_host_cpu.py::test_multi_rank_routing_gloo (launched by torch.distributed.run, backend gloo)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genome_b200.dnamap import PairedEndData, owner_of  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    k = 21
    b, n, _ = H.small_reads(8000, 80, 12, 0.01, seed=17)  # same bytes on every rank
    data = PairedEndData(b, n // 2)
    mine = data.shard(rank, world)
    keys = pyoracle.extract_canonical(mine.bin, mine.n_reads, k)
    own = owner_of(keys, world)
    send = [keys[own == p] for p in range(world)]
    gathered = [None] * world
    dist.all_gather_object(gathered, send)
    shard = pyoracle.OracleMap(k)
    received = 0
    for src in range(world):
        for x in gathered[src][rank]:
            shard.update1(int(x))
            received += 1
    sk, sv = shard.export_sorted()
    assert np.all(owner_of(sk, world) == rank)
    # PartitionedDNAMap.size = sum over partitions; mapReduce = concatenation
    all_shards = [None] * world
    dist.all_gather_object(all_shards, (sk, sv, received))
    if rank == 0:
        whole = pyoracle.OracleMap(k)
        w = whole.insert_reads(b, n)
        wk, wv = whole.export_sorted()
        ck = np.concatenate([s[0] for s in all_shards])
        cv = np.concatenate([s[1] for s in all_shards])
        o = np.argsort(ck, kind="stable")
        assert sum(s[2] for s in all_shards) == w
        assert np.array_equal(ck[o], wk) and np.array_equal(cv[o], wv)
        assert len(np.unique(ck)) == ck.size  # no key lives on two shards
        print("ROUTING OK", w, wk.size)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
