"""The reference's driver scripts mirrored over the C ABI, end to end on the device: GraphBuilder.startup
(S/scripts/GraphBuilder.scala:18-59) against the oracle.  Compositions of entry points that have their own parity tests; written
after this round's GPU budget was spent, so opt-in until run on a B200."""
import os

import numpy as np
import pytest

from genome_b200.builder import GraphBuilder, component_histograms
from genome_b200.dnamap import PairedEndData
from oracle import pyoracle
from tests import helpers as H

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.environ.get("GENOME_B200_UNVALIDATED"), reason="not yet run on a B200 (set GENOME_B200_UNVALIDATED=1)")]


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
