"""The reference's driver scripts mirrored over the C ABI, end to end on the device: GraphBuilder.startup
(S/scripts/GraphBuilder.scala:18-59) against the oracle.  Compositions of entry points that have their own parity tests."""
import os

import numpy as np
import pytest

from genome_b200.builder import GraphBuilder, component_histograms
from genome_b200.dnamap import PairedEndData
from oracle import pyoracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k,glen,rl,cov,err", [(31, 20000, 100, 30, 0.01), (15, 5000, 60, 30, 0.01), (9, 4000, 40, 15, 0.03)])
def test_graph_builder_script(gpu, k, glen, rl, cov, err):
    b, n, _ = H.small_reads(glen, rl, cov, err, seed=2000 + k)
    gb = GraphBuilder(k)                       # rounds = 3 like the reference
    g = gb.startup(PairedEndData(b, n // 2))
    om, _ = H.oracle_counts(b, n, k)
    om.delete_below(3)
    og = pyoracle.OracleGraph(om)
    assert gb.log["good_reads_count"] == om.size()
    assert gb.log["total_edges_length"] == og.counts()[2]
    node_kmer, node_id, es, ee, off, bases = og.export()
    nc, label = og.components()
    idx = {int(i): j for j, i in enumerate(node_id)}
    hist, hist2, comp_nodes = component_histograms(nc, label, np.array([idx[int(x)] for x in es], np.int64), off)
    assert gb.log["components_histogram"] == hist and gb.log["components_histogram_2"] == hist2
    assert gb.log["max_component_size"] == (int(comp_nodes.max()) if nc else 0)
    og.retain_largest()
    # the largest component and its strand twin tie: equal up to a global reverse complement
    got = H.canon_gpu_graph(g)
    want = H.canon_oracle_graph(og)
    assert got == want or got == H.rc_graph(want, k)


def test_retain_arbitrary_node_set(gpu):
    """MapGraph.retain(nodesSet) (Graph.scala:161-165): retaining the nodes of the largest component (labels from
    gb_graph_components) must equal gb_graph_retain_largest; retaining everything changes nothing; retaining nothing empties."""
    from genome_b200.dnamap import FreqFilter
    from genome_b200.graph import Graph
    k = 15
    b, n, _ = H.small_reads(5000, 60, 30, 0.01, seed=2015)
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, 2)
    g1, g2, g3 = Graph.buildGraph(k, gm), Graph.buildGraph(k, gm), Graph.buildGraph(k, gm)
    before = H.canon_gpu_graph(g1)
    g1.retain(np.ones(g1.counts()[0], bool))
    assert H.canon_gpu_graph(g1) == before
    nc, label = g2.components()
    sizes = np.bincount(label, minlength=nc)
    node_kmer = g2.export()[0]
    best = [c for c in range(nc) if sizes[c] == sizes.max()]
    pick = min(best, key=lambda c: int(node_kmer[label == c].min()))   # the library's tie rule: smallest node k-mer
    g2.retain(label == pick)
    g3.retain_largest()
    assert H.canon_gpu_graph(g2) == H.canon_gpu_graph(g3)
    g3.retain(np.zeros(g3.counts()[0], bool))
    assert g3.counts() == (0, 0, 0)


def test_graph_edit_mutators(gpu):
    """addNode / addEdge / replaceStart / replaceEnd / removeNode (Graph.scala:172-209) through gb_graph_edit: a round trip that
    must leave the graph as it was, with the intermediate states checked through the export."""
    from genome_b200.dnamap import FreqFilter
    from genome_b200.graph import Graph
    k = 15
    b, n, _ = H.small_reads(5000, 60, 30, 0.01, seed=2015)
    g = Graph.buildGraph(k, FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, 2))
    before = H.canon_gpu_graph(g)
    nn, ne, nb = g.counts()
    assert ne > 0
    x, y = (1 << (2 * k)) - 1, (1 << (2 * k)) - 2   # poly-T and a neighbour: not nodes of a random 5 kbp genome
    assert x not in before[0] and y not in before[0]
    new_nodes, new_edges = g.edit(add_nodes=[x, y], add_edges=[(nn, nn + 1, [0, 1, 2]), (nn + 1, 0, [3])])
    assert new_nodes == [nn, nn + 1] and new_edges == [ne, ne + 1]
    assert g.counts() == (nn + 2, ne + 2, nb + 4)
    node_kmer, es, ee, off, bases = g.export()
    assert (int(node_kmer[nn]), int(node_kmer[nn + 1])) == (x, y)
    assert (int(es[ne]), int(ee[ne]), bases[int(off[ne]):int(off[ne + 1])].tolist()) == (nn, nn + 1, [0, 1, 2])
    assert (int(es[ne + 1]), int(ee[ne + 1]), bases[int(off[ne + 1]):int(off[ne + 2])].tolist()) == (nn + 1, 0, [3])
    assert len(H.canon_gpu_graph(g)[1]) == ne + 2
    # replaceEnd / replaceStart and back
    old_s, old_e = int(es[0]), int(ee[0])
    g.edit(replace=[(0, None, nn)])
    assert int(g.export()[2][0]) == nn and int(g.export()[1][0]) == old_s
    g.edit(replace=[(0, nn + 1, old_e)])
    assert (int(g.export()[1][0]), int(g.export()[2][0])) == (nn + 1, old_e)
    g.edit(replace=[(0, old_s, None)])
    # removeNode refuses a node that still has edges, then succeeds once they are gone
    with pytest.raises(Exception):
        g.edit(remove_nodes=[nn])
    assert g.counts() == (nn + 2, ne + 2, nb + 4)
    g.removeEdges([ne, ne + 1])
    g.edit(remove_nodes=[nn, nn + 1])
    assert g.counts() == (nn, ne, nb)
    assert H.canon_gpu_graph(g) == before
    g.check()


@pytest.mark.parametrize("k,glen,rl,cov,err", [(31, 20000, 100, 30, 0.01), (9, 4000, 40, 15, 0.03)])
def test_graph_file_round_trip(gpu, tmp_path, k, glen, rl, cov, err):
    """graph.write(file) then Graph(file) (Graph.scala:232-261,384-390; GraphBuilder.scala:56 -> GraphSimplifier.scala:34): the Kryo
    file carries the graph from the builder script to the simplifier script.  The graph read back equals the one written and
    behaves the same under simplifyGraph."""
    from genome_b200 import formats
    from genome_b200.dnamap import FreqFilter
    from genome_b200.graph import Graph
    b, n, _ = H.small_reads(glen, rl, cov, err, seed=2000 + k)
    g = Graph.buildGraph(k, FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, 2))
    path = str(tmp_path / "graph")
    g.write(path)
    nodes, edges = formats.read_kryo_graph(open(path, "rb").read())
    assert (len(nodes), len(edges), sum(e[3].size for e in edges)) == g.counts()
    g2 = Graph.apply(path)
    assert g2.counts() == g.counts()
    assert H.canon_gpu_graph(g2) == H.canon_gpu_graph(g)
    g2.check()
    g.simplifyGraph()
    g2.simplifyGraph()
    assert H.canon_gpu_graph(g2) == H.canon_gpu_graph(g)
