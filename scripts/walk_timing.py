"""Paired-end path support at BASELINE configs[1] shape (C2: 4.6 Mbp, 1 % errors, 100 bp, 30x, k = 31) with mates placed so that
the k-mer distance falls in the reference's range 180..250: wall-clock of gb_graph_pair_support (host scan + H2D + kernels) and
of the node sweep.  Writes gpurun_out/walk_timing.json.  Usage: python scripts/walk_timing.py [scale]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genome_b200 import capi, synth  # noqa: E402
from genome_b200.dnamap import FreqFilter, PairedEndData  # noqa: E402
from genome_b200.graph import Graph  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
k, L, cov, err = 31, 100, 30, 0.01
G = int(4_600_000 * scale)
genome = synth.random_genome(G, 0x5EED0002)
n_reads = (int(cov * G / L) // 2) * 2
parts = []
for i, lo in enumerate(range(0, n_reads, 1 << 19)):
    m = min(1 << 19, n_reads - lo)
    parts.append(synth.pack_fixed(synth.sample_reads(genome, L, m, err, 900 + i, insert=(180 - (L - k), 250 - (L - k)))))
b = np.concatenate(parts)
data = PairedEndData(b, n_reads // 2)
res = dict(workload="C2 shape x %.2f, mates at k-mer distance 180..250" % scale, pairs=n_reads // 2, k=k)

t0 = time.perf_counter()
kmers = FreqFilter.extractFilteredKmers(data, k, 3)
g = Graph.buildGraph(k, kmers)
g.retain_largest()                      # GraphBuilder.scala:52-54: what GraphSimplifier loads
res["build_s"] = time.perf_counter() - t0
res["graph"] = g.counts()
launches0 = capi.lib().gb_launch_count()
for rep in ("cold", "warm"):
    t0 = time.perf_counter()
    support, bad, walked = g.pairSupport(data, range_=(180, 250))
    res["pair_support_%s_s" % rep] = time.perf_counter() - t0
res["device_ns"] = {k_: v for k_, v in g.stats().items() if k_.startswith("pair_support")}
res["pair_support_launches"] = (capi.lib().gb_launch_count() - launches0) // 2
res.update(bad_pairs=bad, walked_cases=walked, supported_edge_pairs=int((support > 0).sum()), support_max=int(support.max()) if support.size else 0)
res["pairs_per_s_warm"] = (n_reads // 2) / res["pair_support_warm_s"]
for cutoff in (200, 20):
    if cutoff == 20 or support.max() >= 200:
        t0 = time.perf_counter()
        removed, added = g.splitNodes(support, cutoff)
        g.simplifyGraph()
        res["split_simplify_s"] = time.perf_counter() - t0
        res.update(cutoff=cutoff, edges_removed=removed, nodes_added=added, after=g.counts())
        break
g.check()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "walk_timing.json"), "w") as f:
    json.dump(res, f)
print(json.dumps(res))
