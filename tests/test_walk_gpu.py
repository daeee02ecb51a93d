"""Paired-end path support on the device (gb_graph_pair_support / gb_graph_split_nodes, csrc/walk.cu)
against the oracle's restatement of GraphSimplifier.scala:33-127,188-317.

The device functions are also checked against the oracle on the CPU (tests/test_walk_emul_cpu.py runs the same walk.cuh
code serially); these tests cover the kernels proper (first B200 run: gpurun_out/walk_check.log, all pass)."""
import numpy as np
import pytest

from genome_b200 import synth
from genome_b200.dnamap import FreqFilter, PairedEndData
from genome_b200.graph import Graph
from genome_b200.simplifier import GraphSimplifier
from oracle import pyoracle
from tests import helpers as H
from tests.test_walk_cpu import reads_of, two_chromosomes
from tests.test_walk_emul_cpu import node_signatures

pytestmark = pytest.mark.gpu


def scenario(case):
    k, L = 15, 50
    if case == "two_chromosomes":
        g1, g2 = two_chromosomes(k, 11)
        reads = reads_of([g1, g2], L, 1500, (60, 100), 21)
        return k, synth.pack_fixed(reads), reads.shape[0], 3
    genome = synth.random_genome(6000, 91)
    n_reads = (int(40 * 6000 / L) // 2) * 2
    reads = synth.sample_reads(genome, L, n_reads, 0.02, 92, insert=(60, 100))
    if case == "noisy_ragged":
        lens = np.random.default_rng(93).integers(L // 3, L + 1, size=n_reads)
        return k, synth.pack_ragged([reads[i, :lens[i]] for i in range(n_reads)]), n_reads, 2
    return k, synth.pack_fixed(reads), n_reads, 2


def build_both(k, b, n_reads, rounds):
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n_reads // 2), k, rounds)
    om, _ = H.oracle_counts(b, n_reads, k)
    om.delete_below(rounds)
    return Graph.buildGraph(k, gm), pyoracle.OracleGraph(om)


def edge_key_maps(g, og):
    """edge index (device) and edge id (oracle) -> (start k-mer, spelled bases): unique on an unsplit graph"""
    node_kmer, es, ee, off, bases = g.export()
    mine = {e: (int(node_kmer[es[e]]), bases[int(off[e]):int(off[e + 1])].tobytes()) for e in range(es.size)}
    onk, oid, oes, oee, ooff, obases = og.export()
    by_id = {int(i): int(x) for i, x in zip(oid, onk)}
    theirs = {int(eid): (by_id[int(oes[j])], obases[int(ooff[j]):int(ooff[j + 1])].tobytes()) for j, eid in enumerate(og.edge_ids())}
    return mine, theirs


def canon_support(g, support):
    node_kmer, es, ee, off, bases = g.export()
    out4 = {(int(es[e]), int(bases[int(off[e])])): e for e in range(es.size)}
    keys = {e: (int(node_kmer[es[e]]), bases[int(off[e]):int(off[e + 1])].tobytes()) for e in range(es.size)}
    res = []
    for e1, b in zip(*np.nonzero(support)):
        e2 = out4[(int(ee[e1]), int(b))]
        res.append((keys[int(e1)], keys[e2], int(support[e1, b])))
    return sorted(res)


@pytest.mark.parametrize("case", ["two_chromosomes", "noisy", "noisy_ragged"])
def test_pair_support_matches_oracle(gpu, case):
    k, b, n_reads, rounds = scenario(case)
    g, og = build_both(k, b, n_reads, rounds)
    support, bad, walked = g.pairSupport(PairedEndData(b, n_reads // 2), range_=(90, 155))
    e1, e2, cnt, obad, owalked = og.pair_support(b, n_reads // 2, 90, 155)
    _, theirs = edge_key_maps(g, og)
    want = sorted((theirs[int(a)], theirs[int(c)], int(n)) for a, c, n in zip(e1, e2, cnt))
    assert canon_support(g, support) == want
    assert (bad, walked) == (obad, owalked)
    assert walked > 0 and len(want) > 0


@pytest.mark.parametrize("case,cutoff", [("two_chromosomes", 5), ("two_chromosomes", 10 ** 6), ("noisy", 1), ("noisy", 3)])
def test_split_and_simplify_match_oracle(gpu, case, cutoff):
    k, b, n_reads, rounds = scenario(case)
    g, og = build_both(k, b, n_reads, rounds)
    support, _, _ = g.pairSupport(PairedEndData(b, n_reads // 2), range_=(90, 155))
    e1, e2, cnt, _, _ = og.pair_support(b, n_reads // 2, 90, 155)
    removed, added = g.splitNodes(support, cutoff)
    assert (removed, added) == og.split(e1, e2, cnt, cutoff)
    g.check()
    node_kmer, es, ee, off, bases = g.export()
    got = node_signatures(k, node_kmer, es, ee, off, bases, np.ones(es.size, bool))
    onk, oid, oes, oee, ooff, obases = og.export()
    pos = {int(i): j for j, i in enumerate(oid)}
    want = node_signatures(k, onk, [pos[int(x)] for x in oes], [pos[int(x)] for x in oee], ooff, obases, np.ones(len(oes), bool))
    assert got == want
    g.simplifyGraph()
    og.simplify()
    g.check()
    H.assert_graph_equal(g, og)


def test_graph_simplifier_script(gpu, tmp_path):
    """GraphSimplifier.startup end to end on the shared-k-mer case: both sequences come back as single contigs."""
    k, b, n_reads, rounds = scenario("two_chromosomes")
    g, og = build_both(k, b, n_reads, rounds)
    gs = GraphSimplifier(range_=(90, 155), cutoff=5)
    gs.startup(g, PairedEndData(b, n_reads // 2), outfile=str(tmp_path / "matrices"), contigs=str(tmp_path / "contigs"))
    assert gs.log["edges_before"] == 8 and gs.log["edges_after"] == 4 and gs.log["nodes_added"] == 4
    assert g.counts()[:2] == (8, 4)
    lines = open(tmp_path / "contigs").read().split("\n")
    assert lines[1] == ">abacaba0" and len(lines[0]) > 1300
    assert len(open(tmp_path / "matrices").read().strip().split("\n")) == 2   # X and rc(X): the two 2 x 2 nodes


def test_range_limits(gpu):
    from genome_b200 import capi
    k, b, n_reads, rounds = scenario("two_chromosomes")
    g, _ = build_both(k, b, n_reads, rounds)
    with pytest.raises(capi.GenomeError) as e:
        g.pairSupport(PairedEndData(b, n_reads // 2), range_=(180, 600))
    assert e.value.name == "GB_E_ARG"
    s, bad, walked = g.pairSupport(PairedEndData(b, 0), range_=(90, 155))
    assert s.sum() == 0 and (bad, walked) == (0, 0)
