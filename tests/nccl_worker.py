"""Worker of tests/test_parity_multigpu.py (torch.distributed.run, one rank per GPU): PartitionedDNAMap over the
library's NCCL all-to-all, checked against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genome_b200.dnamap import Communicator, PairedEndData, PartitionedDNAMap, FreqFilter, owner_of, torch_broadcast  # noqa: E402
from genome_b200.graph import Graph  # noqa: E402
from genome_b200 import synth  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from genome_b200 import capi
    capi.tune_from_env()   # GENOME_B200_TUNE of the test that started this world (read by the harness, not by the library)
    comm = Communicator(rank, world, local, torch_broadcast)
    for (k, glen, rl, cov, err, rounds, ragged, cap) in [(31, 60000, 100, 20, 0.01, 3, False, 1 << 20), (21, 20000, 80, 12, 0.02, 2, True, 0),
                                                        (9, 3000, 40, 12, 0.02, 1, False, 0)]:
        b, n, _ = H.small_reads(glen, rl, cov, err, seed=900 + k, ragged=ragged)  # same bytes on every rank
        data = PairedEndData(b, n // 2)
        mine = data.shard(rank, world)
        m = PartitionedDNAMap(k, comm, cap)
        w = m.insert_reads(mine)
        om, ow = H.oracle_counts(b, n, k)
        tw = torch.tensor([w], device="cuda")
        dist.all_reduce(tw)
        assert int(tw.item()) == ow, (int(tw.item()), ow)
        assert m.size == om.size()
        # this shard holds exactly the oracle's keys that it owns
        ok, ov = om.export_sorted()
        sel = m.owner(ok) == rank   # the map's own ownership rule
        gk, gv = m.export_sorted()
        assert np.array_equal(gk, ok[sel]) and np.array_equal(gv, ov[sel])
        assert m.local_size == int(sel.sum())
        # routed lookups: every rank asks about its own mix of present and absent keys
        rng = np.random.default_rng(rank)
        q = np.concatenate([ok[rank::7][:3000], rng.integers(0, 1 << (2 * k), size=1000 + 10 * rank, dtype=np.uint64)])
        counts, found = m.lookup(q)
        for i in range(0, q.size, 13):
            v = om.apply(int(q[i]))
            assert found[i] == (v is not None) and counts[i] == (v or 0)
        # deleteAll + buildGraph over all shards
        m.delete_below(rounds)
        om.delete_below(rounds)
        assert m.size == om.size()
        g = Graph.buildGraph(k, m)
        og = pyoracle.OracleGraph(om)
        H.assert_graph_equal(g, og)
        g.retain_largest(); og.retain_largest()
        g.simplifyGraph(); og.simplify()
        H.assert_graph_equal(g, og)
        g.close()
        m.close()
    # every read the same sequence: a handful of keys, so whole buckets go to ONE owner and, in the single-pass form, overflow their
    # slabs -- the overflow lists take the staged route.  Counts must still be exact.
    k = 21
    one = synth.sample_reads(synth.random_genome(100, 5), 100, 1, 0.0, 6)
    reads = np.repeat(one, 6000, axis=0)
    b = synth.pack_fixed(reads)
    n = reads.shape[0]
    mine = PairedEndData(b, n // 2).shard(rank, world)
    m = PartitionedDNAMap(k, comm, 1 << 16)
    w = m.insert_reads(mine)
    om, ow = H.oracle_counts(b, n, k)
    tw = torch.tensor([w], device="cuda")
    dist.all_reduce(tw)
    assert int(tw.item()) == ow, (int(tw.item()), ow)
    assert m.size == om.size()
    ok, ov = om.export_sorted()
    sel = m.owner(ok) == rank
    gk, gv = m.export_sorted()
    assert np.array_equal(gk, ok[sel]) and np.array_equal(gv, ov[sel])
    m.close()
    # fewer pairs than ranks: some ranks hold no reads at all but still take part in every collective
    k = 15
    b, n, _ = H.small_reads(400, 60, 1, 0.0, seed=77)
    n = 2 * max(1, min(n // 2, world - 1))
    b = b[:n * (1 + 15)]
    mine = PairedEndData(b, n // 2).shard(rank, world)
    m = PartitionedDNAMap(k, comm)
    w = m.insert_reads(mine)
    om, ow = H.oracle_counts(b, n, k)
    tw = torch.tensor([w], device="cuda")
    dist.all_reduce(tw)
    assert int(tw.item()) == ow
    assert m.size == om.size()
    g = Graph.buildGraph(k, m)
    H.assert_graph_equal(g, pyoracle.OracleGraph(om))
    g.close()
    m.close()
    # the device-resident entry point, fixed stride
    k = 31
    b, n, _ = H.small_reads(50000, 100, 10, 0.01, seed=321)
    mine = PairedEndData(b, n // 2).shard(rank, world)
    d = torch.zeros(mine.bin.size + 16, dtype=torch.uint8, device="cuda")
    d[:mine.bin.size].copy_(torch.from_numpy(mine.bin))
    m = PartitionedDNAMap(k, comm, 1 << 20)
    m.insert_reads_device(d.data_ptr(), mine.bin.size, mine.n_reads)
    om, ow = H.oracle_counts(b, n, k)
    assert m.size == om.size()
    ok, ov = om.export_sorted()
    sel = m.owner(ok) == rank   # the map's own ownership rule
    gk, gv = m.export_sorted()
    assert np.array_equal(gk, ok[sel]) and np.array_equal(gv, ov[sel])
    m.close()
    pair_support_over_ranks(comm, rank, world)
    dist.barrier()
    comm.close()
    if rank == 0:
        print("PMAP OK world", world)
    dist.destroy_process_group()


def pair_support_over_ranks(comm, rank, world):
    """GraphSimplifier's pair loop split over the ranks (MapGraph.pairSupport(comm=...)): the graph is built from the shards
    (identical on every rank), every rank walks its slice of the pairs, counts are summed; against the oracle's pathsMap."""
    from genome_b200 import synth
    from genome_b200.simplifier import GraphSimplifier
    k, L = 15, 50
    genome = synth.random_genome(6000, 91)
    n_reads = (int(40 * 6000 / L) // 2) * 2
    reads = synth.sample_reads(genome, L, n_reads, 0.02, 92, insert=(60, 100))
    b = synth.pack_fixed(reads)
    data = PairedEndData(b, n_reads // 2)
    m = PartitionedDNAMap(k, comm)
    m.insert_reads(data.shard(rank, world))
    m.delete_below(2)
    om, _ = H.oracle_counts(b, n_reads, k)
    om.delete_below(2)
    g = Graph.buildGraph(k, m)
    og = pyoracle.OracleGraph(om)
    H.assert_graph_equal(g, og)
    support, bad, walked = g.pairSupport(data, None, (90, 155), comm)
    e1, e2, cnt, obad, owalked = og.pair_support(b, n_reads // 2, 90, 155)
    assert (bad, walked) == (obad, owalked), (bad, walked, obad, owalked)
    assert int(support.sum()) == int(np.sum(cnt)) and int((support > 0).sum()) == len(cnt)
    single = g.pairSupport(data, None, (90, 155))   # every rank alone on all pairs: same counts
    assert np.array_equal(single[0], support) and single[1:] == (bad, walked)
    GraphSimplifier((90, 155), cutoff=3).startup(g, data, comm=comm)
    removed, added = og.split(e1, e2, cnt, 3)
    og.simplify()
    H.assert_graph_equal(g, og)
    g.close()
    m.close()


if __name__ == "__main__":
    main()
