"""Graph.getGraphMap on the device against the oracle's restatement (Graph.scala:90-119)."""
import os

import numpy as np
import pytest

from genome_b200.dnamap import FreqFilter, PairedEndData
from genome_b200.graph import Graph
from oracle import pyoracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,glen,rl,cov,err,rounds", [(31, 20000, 100, 30, 0.01, 3), (15, 5000, 60, 30, 0.01, 2), (8, 1500, 40, 10, 0.0, 1), (4, 120, 20, 6, 0.0, 1)])
def test_graph_positions_match_oracle(gpu, k, glen, rl, cov, err, rounds):
    b, n, _ = H.small_reads(glen, rl, cov, err, seed=4000 + k)
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, rounds)
    om, _ = H.oracle_counts(b, n, k)
    om.delete_below(rounds)
    g = Graph.buildGraph(k, gm)
    og = pyoracle.OracleGraph(om)

    def canon(kmer, ident, dist, node_kmer, es, ee, off, bases, edge_index):
        # ids are not comparable across implementations: replace them by (start k-mer, end k-mer, seq) of the edge
        out = []
        for x, i, d in zip(kmer.tolist(), ident.tolist(), dist.tolist()):
            if d == 0:
                out.append((x, 0, None))
            else:
                e = edge_index(i)
                out.append((x, d, (int(node_kmer[es[e]]), int(node_kmer[ee[e]]), bases[int(off[e]):int(off[e + 1])].tobytes())))
        return sorted(out, key=lambda t: (t[0], t[1]))

    node_kmer, es, ee, off, bases = g.export()
    gk, gi, gd = g.getGraphMap()
    nn, ne, nb = g.counts()
    assert gk.size == nn + nb - ne
    mine = canon(gk, gi, gd, node_kmer, es, ee, off, bases, lambda i: i)

    onk, oid, oes, oee, ooff, obases = og.export()
    pos = {int(i): j for j, i in enumerate(oid)}
    oes_i = np.array([pos[int(x)] for x in oes], np.int64)
    oee_i = np.array([pos[int(x)] for x in oee], np.int64)
    ok, oi, od = og.graph_map()
    theirs = canon(ok, oi, od, onk, oes_i, oee_i, ooff, obases, lambda i: i - 1)  # fresh oracle graph: edge id = index + 1
    assert mine == theirs


# the logic of gb_graph_map_* is also covered on the CPU by the g++ emulation (tests/test_walk_emul_cpu.py::test_graph_map_get_all_matches_oracle)
@pytest.mark.parametrize("k,glen,rl,cov,err,rounds", [(31, 20000, 100, 30, 0.01, 3), (15, 5000, 60, 30, 0.01, 2), (4, 120, 20, 6, 0.0, 1)])
def test_graph_map_handle(gpu, k, glen, rl, cov, err, rounds):
    """GraphPositionMap.size / getAll / contains against the exported entry list: every entry is found under its k-mer, absent
    k-mers are not, and the CheckGraph property (S/scripts/CheckGraph.scala:47-54) holds: every k-mer of the kept set, in
    both orientations, that lies on the graph is found."""
    b, n, _ = H.small_reads(glen, rl, cov, err, seed=4100 + k)
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, rounds)
    g = Graph.buildGraph(k, gm)
    pk, pi, pd = g.getGraphMap()
    m = g.graphMap()
    assert m.size == pk.size
    want = {}
    for x, i, d in zip(pk.tolist(), pi.tolist(), pd.tolist()):
        want.setdefault(x, []).append((i, d))
    rng = np.random.default_rng(k)
    absent = rng.integers(0, 1 << (2 * k), 1000, dtype=np.uint64)
    keys = np.concatenate([pk, absent])
    counts, ids, dists = m.getAll(keys, 2)
    for j, x in enumerate(keys.tolist()):
        w = want.get(x, [])
        assert counts[j] == len(w)
        assert sorted((int(ids[j, c]), int(dists[j, c])) for c in range(min(len(w), 2))) == sorted(w)[:2] or len(w) > 2
    assert np.array_equal(m.contains(keys), counts > 0)
    # the map is a snapshot: it survives changes to (and the destruction of) the graph it came from
    g.retain_largest()
    g.simplifyGraph()
    g.close()
    assert np.array_equal(m.contains(keys), counts > 0)
    with pytest.raises(Exception):
        m.contains(np.array([1 << (2 * k)], np.uint64))  # longer than k: GB_E_K_RANGE
    m.close()


def test_check_graph_finds_every_genome_kmer(gpu):
    """CheckGraph.startup (S/scripts/CheckGraph.scala:18-56) on an error-free read set that covers the genome: every k-window
    of the genome FASTA is on the graph; windows of an unrelated sequence are reported."""
    from genome_b200 import checkgraph, synth
    k = 21
    genome = synth.random_genome(30000, 77)
    reads = synth.sample_reads(genome, 100, 12000, 0.0, 78)
    b = synth.pack_fixed(reads)
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, reads.shape[0] // 2), k, 1)
    g = Graph.buildGraph(k, gm)
    covered = synth.decode(genome[2000:28000])   # the ends may be under-sampled
    other = synth.decode(synth.random_genome(300, 79))
    stats, missing = checkgraph.check_graph(g, [">chr", covered[:13000], covered[13000:], "ACGTN", other])
    assert stats["max"] > 200 and stats["count"] >= 1
    lens = sorted(set(m[0] for m in missing))
    assert lens == [300] and len(missing) == 300 - k + 1   # only the unrelated line (and none of "ACGTN": no full window)


def test_kmers_calculator(gpu):
    """KmersCalculator (S/scripts/KmersCalculator.scala:16-28): distinct k-windows of a FASTA record, counted on the device."""
    from genome_b200 import checkgraph, synth
    text = synth.decode(synth.random_genome(50000, 5))
    text = text + text[1000:9000]   # a repeat: 8000 - 18 windows seen twice
    lines = [">x"] + [text[i:i + 70] for i in range(0, len(text), 70)]
    n, distinct = checkgraph.kmers_calculator(lines, 19)
    assert n == len(text)
    assert distinct == len({text[i:i + 19] for i in range(len(text) - 18)})
