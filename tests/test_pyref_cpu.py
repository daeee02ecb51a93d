"""The C oracle (oracle/oracle.c) against a second, independent reading of the Scala (oracle/pyref.py: plain Python, written from
the reference's sources alone) on seeded inputs: k-mer arithmetic, the count table before and after deleteAll, Graph.buildGraph,
components, simplifyGraph, getGraphMap.  Neither restatement can be pinned by the reference itself; that two
of them agree bit for bit is the check available without a JVM."""
import random

import numpy as np
import pytest

from oracle import pyoracle, pyref
from tests import helpers as H


@pytest.mark.parametrize("k", [1, 2, 3, 8, 15, 16, 17, 21, 30, 31])
def test_kmer_arithmetic(k):
    rng = np.random.default_rng(k)
    xs = [int(x) for x in rng.integers(0, 1 << (2 * k), size=400, dtype=np.uint64)] + [0, (1 << (2 * k)) - 1]
    for x in xs:
        s = pyref.Seq1(x, k)
        r = s.rev_complement()
        assert r.long == pyoracle.revcomp(x, k)
        assert r.rev_complement() == s
        assert r.bases() == [b ^ 3 for b in reversed(s.bases())]          # A0 G1 C2 T3: the complement is code ^ 3
        assert [pyref.COMPLEMENT[b] for b in s.bases()] == [b ^ 3 for b in s.bases()]
        for variant in (291, 210):
            assert s.hash_code(variant) == pyoracle.hash_(x, variant)
            y = s if s.hash_code(variant) < r.hash_code(variant) else r  # FreqFilter.scala:32
            assert y.long == pyoracle.canonical(x, k, variant)
        if k > 1:
            for b in range(4):                                            # Graph.scala:273, 279
                assert s.drop(1).append(b).long == (x >> 2) | (b << (2 * (k - 1)))
                assert s.take(k - 1).prepend(b).long == ((x << 2) & ((1 << (2 * k)) - 1)) | b


def test_hash_edge_values():
    for v in [0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0x100000000, 0x100000001, (1 << 62) - 1, 0x123456789ABCDEF, 0x3FFFFFFF80000000]:
        for variant in (291, 210):
            assert pyref.long_hash(v, variant) == pyoracle.hash_(v, variant), (hex(v), variant)


CASES = [
    # k, genome, read_len, coverage, err, rounds, ragged
    (31, 1500, 60, 12, 0.01, 2, False),
    (21, 1200, 50, 10, 0.02, 2, True),
    (15, 800, 40, 12, 0.02, 2, False),
    (9, 500, 30, 8, 0.03, 1, True),
    (8, 300, 30, 8, 0.0, 1, False),     # even k: palindromes
    (6, 200, 20, 6, 0.02, 1, False),
    (5, 120, 20, 4, 0.0, 1, True),
    (4, 80, 16, 5, 0.0, 1, False),
    (3, 40, 12, 3, 0.0, 1, False),
    (30, 900, 64, 10, 0.01, 2, False),
]


def _table(freq):
    items = sorted((key[0], v) for key, v in freq.items())
    return [a for a, _ in items], [b for _, b in items]


def _oracle_canonical_components(og):
    node_kmer = og.export()[0]
    nc, label = og.components()
    return sorted(sorted(int(x) for x in node_kmer[label == c]) for c in range(nc))


@pytest.mark.parametrize("k,glen,rl,cov,err,rounds,ragged", CASES)
def test_table_graph_components_simplify(k, glen, rl, cov, err, rounds, ragged):
    b, n, _ = H.small_reads(glen, rl, cov, err, seed=4000 + k, ragged=ragged)
    # ---- the count table, before and after deleteAll(v < rounds)
    om, _ = H.oracle_counts(b, n, k)
    raw = pyref.extract_filtered_kmers(b, n // 2, k, rounds, filter_=False)
    ok, ov = om.export_sorted()
    pk, pv = _table(raw)
    assert pk == [int(x) for x in ok] and pv == [int(x) for x in ov]
    om.delete_below(rounds)
    kept = pyref.extract_filtered_kmers(b, n // 2, k, rounds)
    ok, ov = om.export_sorted()
    pk, pv = _table(kept)
    assert pk == [int(x) for x in ok] and pv == [int(x) for x in ov]
    # ---- Graph.buildGraph
    og = pyoracle.OracleGraph(om)
    pg = pyref.build_graph(k, kept)
    assert (len(pg.nodes), len(pg.edges), sum(len(e.seq) for e in pg.edges.values())) == og.counts()
    want = H.canon_oracle_graph(og)
    got = pg.canonical()
    assert got[0] == want[0], "node sets differ"
    assert got[1] == want[1], "edge multisets differ"
    # ---- components: the same partition of the node set
    comps = sorted(sorted(pg.nodes[i].seq.long for i in c) for c in pg.components())
    assert comps == _oracle_canonical_components(og)
    # ---- getGraphMap: every (k-mer, position) pair, positions named by the canonical form of their node / edge
    kmer, ident, dist = og.graph_map()
    node_kmer, node_id, es, ee, off, bases = og.export()
    by_id = {int(i): int(x) for i, x in zip(node_id, node_kmer)}
    oe = {int(eid): (by_id[int(es[i])], by_id[int(ee[i])], bases[int(off[i]):int(off[i + 1])].tobytes()) for i, eid in enumerate(og.edge_ids())}
    opos = sorted((int(x), ("node", by_id[int(i)]) if d == 0 else ("edge", oe[int(i)], int(d))) for x, i, d in zip(kmer, ident, dist))
    ppos = []
    for x, pos in pyref.graph_positions(pg):
        if pos[0] == "node":
            ppos.append((x, ("node", pg.nodes[pos[1]].seq.long)))
        else:
            e = pg.edges[pos[1]]
            ppos.append((x, ("edge", (pg.nodes[e.start_id].seq.long, pg.nodes[e.end_id].seq.long, bytes(e.seq)), pos[2])))
    assert sorted(ppos) == opos
    # ---- simplifyGraph: the canonical result does not depend on the order of the node sweep, and equals the oracle's
    og.simplify()
    want = H.canon_oracle_graph(og)
    for order_seed in (None, 1, 2):
        g2 = pyref.build_graph(k, kept)
        order = None
        if order_seed is not None:
            order = list(g2.nodes)
            random.Random(order_seed).shuffle(order)
        g2.simplify_graph(order)
        assert g2.canonical() == tuple(want) or list(g2.canonical()) == list(want), order_seed
    # ---- removeEdge on a third of the edges, then simplifyGraph: chains through the nodes left with one in- and one out-edge
    # are merged, nodes left bare are dropped (a fresh unitig graph has nothing to merge, so the sweep above changed little)
    og1 = pyoracle.OracleGraph(om)
    canon = H.canon_oracle_graph(og1)[1]
    pick = set(canon[::3])
    node_kmer, node_id, es, ee, off, bases = og1.export()
    by_id = {int(i): int(x) for i, x in zip(node_id, node_kmer)}
    oids = [int(eid) for i, eid in enumerate(og1.edge_ids())
            if (by_id[int(es[i])], by_id[int(ee[i])], bases[int(off[i]):int(off[i + 1])].tobytes()) in pick]
    oarr = np.array(oids, np.int64)
    assert pyoracle.lib().go_graph_remove_edges(og1.h, oarr.ctypes.data, oarr.size) == len(oids)
    og1.simplify()
    want = H.canon_oracle_graph(og1)
    merged = False
    for order_seed in (None, 3, 4):
        g4 = pyref.build_graph(k, kept)
        for e in [e for e in g4.edges.values() if (g4.nodes[e.start_id].seq.long, g4.nodes[e.end_id].seq.long, bytes(e.seq)) in pick]:
            g4.remove_edge(e)
        before = len(g4.edges)
        order = None
        if order_seed is not None:
            order = list(g4.nodes)
            random.Random(order_seed).shuffle(order)
        g4.simplify_graph(order)
        merged |= len(g4.edges) < before
        assert list(g4.canonical()) == list(want), ("after removeEdge", order_seed)
    assert merged or k > 9   # the tangled small-k graphs do merge; the big-k ones are isolated unitigs whose nodes just drop
    # ---- graph.retain(components.maxBy(_.size)) then simplifyGraph (GraphBuilder.scala:52-54), when the largest is unique
    og2 = pyoracle.OracleGraph(om)
    g3 = pyref.build_graph(k, kept)
    comps = g3.components()
    sizes = sorted((len(c) for c in comps), reverse=True)
    if sizes and (len(sizes) == 1 or sizes[0] > sizes[1]):
        g3.retain(max(comps, key=len))
        og2.retain_largest()
        assert list(g3.canonical()) == list(H.canon_oracle_graph(og2))
        g3.simplify_graph()
        og2.simplify()
        assert list(g3.canonical()) == list(H.canon_oracle_graph(og2))


def test_hash_tie_reads_through_both_restatements():
    """Even-k k-mers whose orientations hash alike (tests/golden/hash_ties.json): both orientations end up stored; the two
    restatements agree on the table and on the graph."""
    for k, kmers in H.hash_ties():
        b, n = H.tie_reads(k, kmers[:1], seed=k, flank=20)
        om, _ = H.oracle_counts(b, n, k)
        om.delete_below(3)
        kept = pyref.extract_filtered_kmers(b, n // 2, k, 3)
        ok, ov = om.export_sorted()
        pk, pv = _table(kept)
        assert pk == [int(x) for x in ok] and pv == [int(x) for x in ov]
        x = kmers[0]
        assert (x, k) in kept and (pyoracle.revcomp(x, k), k) in kept
        og = pyoracle.OracleGraph(om)
        pg = pyref.build_graph(k, kept)
        assert list(pg.canonical()) == list(H.canon_oracle_graph(og))


def test_random_dense_kmer_sets_small_k():
    """Random subsets of the whole k-mer space for k = 2..5 (self-loops, hairpins, palindromes, perfect cycles), inserted in
    canonical orientation: buildGraph and simplifyGraph of both restatements."""
    for k in (2, 3, 4, 5):
        rng = np.random.default_rng(900 + k)
        space = 1 << (2 * k)
        for trial in range(8):
            frac = [0.1, 0.3, 0.6, 0.9][trial % 4]
            xs = np.flatnonzero(rng.random(space) < frac)
            om = pyoracle.OracleMap(k)
            kept = {}
            for x in xs.tolist():
                c = pyoracle.canonical(x, k)
                om.update1(c)
                kept[(c, k)] = kept.get((c, k), 0) + 1
            og = pyoracle.OracleGraph(om)
            pg = pyref.build_graph(k, kept)
            assert list(pg.canonical()) == list(H.canon_oracle_graph(og)), (k, trial)
            og.simplify()
            pg.simplify_graph()
            assert list(pg.canonical()) == list(H.canon_oracle_graph(og)), (k, trial, "simplified")


def test_remove_bubbles_on_fresh_graphs():
    """removeBubbles (Graph.scala:121-149) right after buildGraph, where a node's out-edges sit in Base.fromInt order in the
    reference too (buildEdges adds them in that order) -- later, the reference's order is the insertion order of a small
    immutable Map and depends on the simplifyGraph sweep (the oracle's stated choice is first-base order).
    A repeat-rich genome with substitutions gives parallel edges; then simplifyGraph recompacts."""
    popped = 0
    for k, seed in ((9, 1), (11, 2), (15, 3), (21, 4)):
        rng = np.random.default_rng(seed)
        unit = rng.integers(0, 4, 60).astype(np.uint8)
        variant = unit.copy()
        variant[30] = (variant[30] + 1) % 4                  # one substitution: a bubble between the two copies' flanks
        left, mid, right = (rng.integers(0, 4, 50).astype(np.uint8) for _ in range(3))
        # the two alleles in the same context (a heterozygous site): reads from both haplotypes
        hap1 = np.concatenate([left, unit, right])
        hap2 = np.concatenate([left, variant, right])
        rl = k + 15
        reads = [h[i:i + rl] for h in (hap1, hap2) for i in range(h.size - rl + 1) for _ in range(2)]
        if len(reads) % 2:
            reads.append(reads[-1])
        from genome_b200 import synth
        b = synth.pack_fixed(np.stack(reads))
        n = len(reads)
        om, _ = H.oracle_counts(b, n, k)
        om.delete_below(2)
        kept = pyref.extract_filtered_kmers(b, n // 2, k, 2)
        og = pyoracle.OracleGraph(om)
        pg = pyref.build_graph(k, kept)
        assert list(pg.canonical()) == list(H.canon_oracle_graph(og))
        before = len(pg.edges)
        og.remove_bubbles()
        pg.remove_bubbles()
        popped += before - len(pg.edges)
        assert list(pg.canonical()) == list(H.canon_oracle_graph(og)), (k, "bubbles")
        og.simplify()
        pg.simplify_graph()
        assert list(pg.canonical()) == list(H.canon_oracle_graph(og)), (k, "bubbles + simplify")
    assert popped >= 4       # both strands of at least two of the bubbles
