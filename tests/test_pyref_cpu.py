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


# ---- GraphSimplifier (S/scripts/GraphSimplifier.scala): WalkingActor, the pair loop, the node sweep
def _edge_forms(pg):
    return {e.id: (pg.nodes[e.start_id].seq.long, pg.nodes[e.end_id].seq.long, bytes(e.seq)) for e in pg.edges.values()}


def _oracle_edge_forms(og):
    node_kmer, node_id, es, ee, off, bases = og.export()
    by_id = {int(i): int(x) for i, x in zip(node_id, node_kmer)}
    return {int(eid): (by_id[int(es[i])], by_id[int(ee[i])], bases[int(off[i]):int(off[i + 1])].tobytes()) for i, eid in enumerate(og.edge_ids())}


@pytest.mark.parametrize("k,n,seed", [(5, 300, 1), (6, 600, 2), (4, 120, 4)])
def test_walking_actor(k, n, seed):
    """WalkingActor.receive of both restatements for random position pairs on a small-k tangle (ids differ: positions and
    edge pairs are translated through the canonical form of their node / edge, which is unique in these graphs)."""
    from tests.test_walk_cpu import kmers_of, rand_seq
    om = kmers_of([rand_seq(n, seed)], k)
    og = pyoracle.OracleGraph(om)
    keys, vals = om.export()
    pg = pyref.build_graph(k, {(int(x), k): int(v) for x, v in zip(keys, vals)})
    oforms, pforms = _oracle_edge_forms(og), _edge_forms(pg)
    assert len(set(oforms.values())) == len(oforms) and sorted(oforms.values()) == sorted(pforms.values())
    p_edge = {form: eid for eid, form in pforms.items()}
    p_node = {nd.seq.long: nid for nid, nd in pg.nodes.items()}
    o_node = {int(i): int(x) for x, i in zip(og.export()[0], og.export()[1])}
    kmer, ident, dist = og.graph_map()

    def to_py(i):
        return ("node", p_node[o_node[int(ident[i])]]) if dist[i] == 0 else ("edge", p_edge[oforms[int(ident[i])]], int(dist[i]))

    actor = pyref.WalkingActor(pg, 6, 11)
    rng = np.random.default_rng(seed)
    n_good = 0
    for _ in range(150):
        i, j = rng.integers(0, kmer.size, 2)
        good, pairs = og.walk((int(ident[i]), int(dist[i])), (int(ident[j]), int(dist[j])), 6, 11)
        pgood, ppairs = actor.receive(to_py(i), to_py(j))
        assert pgood == good
        assert sorted((pforms[a], pforms[b]) for a, b in ppairs) == sorted((oforms[a], oforms[b]) for a, b in pairs)
        n_good += good
    assert n_good > 0


@pytest.mark.parametrize("cutoff", [5, 10 ** 6])
def test_pair_loop_and_node_sweep(cutoff):
    """Two sequences sharing one k-mer (tests/test_walk_cpu.py): pathsMap, badPairs, the walked cases, the graph after the node
    sweep and after simplifyGraph -- both restatements."""
    from genome_b200 import synth
    from tests.test_walk_cpu import reads_of, two_chromosomes
    k, L = 15, 50
    g1, g2 = two_chromosomes(k, 11)
    reads = reads_of([g1, g2], L, 400, (60, 100), 21)
    b = synth.pack_fixed(reads)
    n = reads.shape[0]
    om = pyoracle.OracleMap(k)
    om.insert_reads(b, n)
    om.delete_below(2)
    og = pyoracle.OracleGraph(om)
    kept = pyref.extract_filtered_kmers(b, n // 2, k, 2)
    pg = pyref.build_graph(k, kept)
    oforms, pforms = _oracle_edge_forms(og), _edge_forms(pg)
    assert len(set(oforms.values())) == len(oforms) and sorted(oforms.values()) == sorted(pforms.values())
    lo, hi = 90, 155
    e1, e2, cnt, bad, walked = og.pair_support(b, n // 2, lo, hi)
    paths, pbad, pwalked = pyref.pair_support(pg, b, n // 2, k, lo, hi)
    assert (pbad, pwalked) == (bad, walked) and walked > 0
    want = sorted((oforms[int(a)], oforms[int(c)], int(v)) for a, c, v in zip(e1, e2, cnt))
    got = sorted((pforms[a], pforms[c], v) for (a, c), v in paths.items())
    assert got == want and len(got) >= 4
    removed, added = og.split(e1, e2, cnt, cutoff)
    premoved, padded = pyref.split_nodes(pg, paths, cutoff)
    assert (premoved, padded) == (removed, added)
    assert list(pg.canonical()) == list(H.canon_oracle_graph(og))
    og.simplify()
    pg.simplify_graph()
    assert list(pg.canonical()) == list(H.canon_oracle_graph(og))
