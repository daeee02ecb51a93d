"""Paired-end path support (SURVEY 8(f) row 4, S/scripts/GraphSimplifier.scala:33-127,188-317): the oracle's restatement
against an independent brute-force enumeration of walks, and the whole split pipeline on hand-built genomes.  CPU only."""
import sys

import numpy as np
import pytest

from genome_b200 import synth
from oracle import pyoracle

from . import helpers as H


def kmers_of(seqs, k):
    m = pyoracle.OracleMap(k)
    for s in seqs:
        for i in range(len(s) - k + 1):
            x = pyoracle.canonical(synth.kmer_to_int(s[i:i + k]), k)
            if not m.contains(x):
                m.update1(x)
    return m


def rand_seq(n, seed):
    return synth.decode(synth.random_genome(n, seed))


class PyGraph:
    """The oracle graph as plain Python dicts (ids = the oracle's)."""

    def __init__(self, og):
        node_kmer, node_id, es, ee, off, bases = og.export()
        self.k = og.k
        self.kmer = {int(i): int(x) for i, x in zip(node_id, node_kmer)}
        # edge ids: the oracle numbers edges 1.. in creation order and exports live edges in id order; on a freshly built
        # graph every edge is alive
        self.edges = {i + 1: (int(es[i]), int(ee[i]), int(off[i + 1] - off[i])) for i in range(es.size)}
        self.out = {n: [] for n in self.kmer}
        for e, (s, t, ln) in self.edges.items():
            self.out[s].append(e)


def brute_walk(pg, pos1, pos2, lo, hi):
    """Every walk from pos1 to pos2 with dist1 + dist2 in lo..hi, by plain recursion (no memo, no reachability prune):
    the pathEdges set and `good` as WalkingActor.receive defines them (GraphSimplifier.scala:77-126)."""
    if pos2[1] == 0:
        node2, dist2, end_edge = pos2[0], 0, None
    else:
        node2, dist2, end_edge = pg.edges[pos2[0]][0], pos2[1], pos2[0]
    if pos1[1] == 0:
        node0, dist0, start_edge = pos1[0], 0, None
    else:
        node0, dist0, start_edge = pg.edges[pos1[0]][1], pg.edges[pos1[0]][2] - pos1[1], pos1[0]
    pairs = set()

    def dfs(node1, dist1, prev):
        if dist1 + dist2 > hi:
            return False
        cur = False
        if node1 == node2 and lo <= dist1 + dist2 <= hi:
            if prev is not None and end_edge is not None:
                pairs.add((prev, end_edge))
            cur = True
        for e in pg.out[node1]:
            res = dfs(pg.edges[e][1], dist1 + pg.edges[e][2], e)
            if res and prev is not None:
                pairs.add((prev, e))
            cur |= res
        return cur

    old = sys.getrecursionlimit()
    sys.setrecursionlimit(10000)
    try:
        good = dfs(node0, dist0, start_edge)
    finally:
        sys.setrecursionlimit(old)
    return good, sorted(pairs)


@pytest.mark.parametrize("k,n,seed", [(5, 300, 1), (6, 600, 2), (7, 1500, 3), (4, 120, 4)])
def test_walk_matches_brute_force(k, n, seed):
    """Small k on a random sequence: a tangle of short edges and cycles.  go_walk (Dijkstra prune + memo, as the reference)
    must agree with the unpruned, unmemoised enumeration for random position pairs."""
    s = rand_seq(n, seed)
    og = pyoracle.OracleGraph(kmers_of([s], k))
    assert og.check() == 0
    pg = PyGraph(og)
    kmer, ident, dist = og.graph_map()
    rng = np.random.default_rng(seed)
    lo, hi = 6, 11
    n_good = 0
    for _ in range(120):
        i, j = rng.integers(0, kmer.size, 2)
        p1, p2 = (int(ident[i]), int(dist[i])), (int(ident[j]), int(dist[j]))
        good, pairs = og.walk(p1, p2, lo, hi)
        bgood, bpairs = brute_walk(pg, p1, p2, lo, hi)
        assert (good, pairs) == (bgood, bpairs), (p1, p2)
        n_good += good
    assert n_good > 0


def two_chromosomes(k, seed):
    """P X Q and R X S: one shared k-mer X, so X (and rc X) is a node with two in- and two out-edges per strand."""
    x = rand_seq(k, seed)
    g1 = rand_seq(700, seed + 1) + x + rand_seq(700, seed + 2)
    g2 = rand_seq(700, seed + 3) + x + rand_seq(700, seed + 4)
    return g1, g2


def reads_of(genomes, read_len, n_pairs_each, insert, seed):
    parts = []
    for i, g in enumerate(genomes):
        parts.append(synth.sample_reads(synth.encode(g), read_len, 2 * n_pairs_each, 0.0, seed + i, insert=insert))
    return np.concatenate(parts)


def spelled(og):
    """sorted strings start k-mer + edge bases of every edge"""
    nodes, edges = H.canon_oracle_graph(og)
    return sorted(synth.int_to_kmer(u, og.k) + synth.decode(np.frombuffer(seq, np.uint8)) for (u, v, seq) in edges)


def rc(t):
    return t[::-1].translate(str.maketrans("ACGT", "TGCA"))


def test_shared_kmer_is_resolved_by_pair_support():
    """GraphSimplifier's whole sweep on two sequences sharing one k-mer: the 2 x 2 node splits into its two supported
    through-paths and simplifyGraph restores the two sequences (and their reverse complements) as single edges."""
    k, L = 15, 50
    g1, g2 = two_chromosomes(k, 11)
    reads = reads_of([g1, g2], L, 1500, (60, 100), 21)
    b = synth.pack_fixed(reads)
    om = pyoracle.OracleMap(k)
    om.insert_reads(b, reads.shape[0])
    om.delete_below(3)
    og = pyoracle.OracleGraph(om)
    assert og.check() == 0
    assert og.counts()[:2] == (10, 8)  # per strand: 4 ends + X; 4 edges
    # the distance between the two k-mers is D = insert offset + L - k = 95..135.  The reference tests D + k for a pair inside
    # one edge (annotate, 196) but D itself along a walk (94,97): the range has to hold both
    lo, hi = 90, 155
    e1, e2, cnt, bad, walked = og.pair_support(b, reads.shape[0] // 2, lo, hi)
    assert walked > 0 and bad < walked // 2  # bad: mates clamped at the sequence end
    # exactly the four true through-paths are supported (two per strand), with comparable counts
    assert e1.size == 4 and cnt.min() > 20
    removed, added = og.split(e1, e2, cnt, cutoff=5)
    assert (removed, added) == (0, 4)
    assert og.check() == 0
    assert og.counts()[:2] == (14, 8)  # the two original X nodes stay behind, empty
    og.simplify()
    assert og.check() == 0
    assert og.counts()[:2] == (8, 4)
    # (coverage thins out at the sequence ends, so each contig is the sequence minus a few end bases)
    refs = [g1, g2, rc(g1), rc(g2)]
    contigs = spelled(og)
    assert all(len(c) > 1350 for c in contigs)
    assert sorted(next(i for i, r in enumerate(refs) if c in r) for c in contigs) == [0, 1, 2, 3]


def test_unsupported_edges_are_removed():
    """With a cutoff nothing reaches, every in-edge of a 2 x 2 node is alone in its component and every out-edge stays
    uncoloured: all four edges around X go (per strand), the nodes stay until simplifyGraph."""
    k, L = 15, 50
    g1, g2 = two_chromosomes(k, 11)
    reads = reads_of([g1, g2], L, 1500, (60, 100), 21)
    b = synth.pack_fixed(reads)
    om = pyoracle.OracleMap(k)
    om.insert_reads(b, reads.shape[0])
    om.delete_below(3)
    og = pyoracle.OracleGraph(om)
    e1, e2, cnt, bad, walked = og.pair_support(b, reads.shape[0] // 2, 90, 155)
    removed, added = og.split(e1, e2, cnt, cutoff=10 ** 6)
    assert (removed, added) == (8, 0)
    og.simplify()
    assert og.counts() == (0, 0, 0)


def test_pairs_inside_one_edge_are_not_walked():
    """annotate (GraphSimplifier.scala:192-206): on a linear genome every pair has both k-mers on the single edge of its
    strand at a distance inside the range, so nothing is walked and pathsMap stays empty."""
    k, L = 15, 50
    g = rand_seq(2000, 51)
    reads = synth.sample_reads(synth.encode(g), L, 2000, 0.0, 52, insert=(60, 100))
    b = synth.pack_fixed(reads)
    om = pyoracle.OracleMap(k)
    om.insert_reads(b, reads.shape[0])
    om.delete_below(2)
    og = pyoracle.OracleGraph(om)
    assert og.counts()[:2] == (4, 2)
    e1, e2, cnt, bad, walked = og.pair_support(b, reads.shape[0] // 2, 90, 155)
    # walked: only pairs whose mate was clamped at the sequence end (their far k-mer is the end NODE: good when in range,
    # but a node position names no end edge, so nothing is recorded)
    assert e1.size == 0 and 0 < bad < walked < 150
    # pairs whose mate was clamped to the genome end fall outside the range: they are walked, find nothing, and count as bad
    e1, e2, cnt, bad2, walked2 = og.pair_support(b, reads.shape[0] // 2, 200, 250)
    assert e1.size == 0 and bad2 == walked2 > 0
