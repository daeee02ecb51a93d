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
