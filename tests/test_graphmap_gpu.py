"""Graph.getGraphMap on the device against the oracle's restatement (Graph.scala:90-119)."""
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
