"""CPU side of scripts/walk_timing.py: the same C2-shaped workload through the oracle (test infrastructure), one thread for the pair
loop.  Output committed as profiles/r1_pair_support_c2shape_cpu_oracle.json."""
import json, sys, time
import numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genome_b200 import synth
from oracle import pyoracle
k, L, cov, err = 31, 100, 30, 0.01
G = 4_600_000
genome = synth.random_genome(G, 0x5EED0002)
n_reads = (int(cov * G / L) // 2) * 2
parts = []
for i, lo in enumerate(range(0, n_reads, 1 << 19)):
    m = min(1 << 19, n_reads - lo)
    parts.append(synth.pack_fixed(synth.sample_reads(genome, L, m, err, 900 + i, insert=(180 - (L - k), 250 - (L - k)))))
b = np.concatenate(parts)
om = pyoracle.OracleMap(k, 8)
t0 = time.time(); om.insert_reads(b, n_reads, threads=8); om.delete_below(3); t_ins = time.time() - t0
t0 = time.time(); og = pyoracle.OracleGraph(om); og.retain_largest(); t_g = time.time() - t0
t0 = time.time(); e1, e2, cnt, bad, walked = og.pair_support(b, n_reads // 2, 180, 250); t_ps = time.time() - t0
t0 = time.time(); removed, added = og.split(e1, e2, cnt, 20); og.simplify(); t_sp = time.time() - t0
print(json.dumps(dict(where="build container CPU (not the GPU box), oracle port, 1 thread for the pair loop", insert_filter_s=t_ins, build_retain_s=t_g,
                      pair_support_s=t_ps, pairs=n_reads // 2, pairs_per_s=(n_reads // 2) / t_ps, bad_pairs=bad, walked_cases=walked,
                      supported_edge_pairs=int(e1.size), support_max=int(cnt.max()), split_simplify_s=t_sp, edges_removed=removed, nodes_added=added, after=og.counts())))
