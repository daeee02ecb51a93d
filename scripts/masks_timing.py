"""Graph.buildGraph on one GPU, phase times from the library's CUDA events (gb_map_phase_ns): membership probes (masks_kernel,
one lane per probe), list ranking (jump_kernel), whole build; then the sharded build over P virtual ranks (PartsOp + ProbeOp +
CombineOp).  One JSON object on stdout.
The variant comparisons kept under profiles/ (r2r_masks_timing.json, r2s_masks_timing.json, r2t_graph_variants_timing.json) were
made with this script at the commits that still carried the older forms behind gb_tune keys (masks_flat: one thread per k-mer;
rank_sublists: sublist walks): scripts/r2r_last_call.sh, r2s_call.sh, r2t_call.sh are the commands of those runs.
Usage: python scripts/masks_timing.py [C2] [scale] [reps]      (reps = 1: a single build, for ncu)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genome_b200 import synth  # noqa: E402
from genome_b200.dnamap import ArrayDNAMap  # noqa: E402
from genome_b200.graph import Graph  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
K = 31
b, n, _ = synth.make_config(cfg, scale=scale)
m = ArrayDNAMap(K, int(b.size * 1.2))
m.insert_reads(b, n)
m.delete_below(3)
out = {"workload": cfg, "scale": scale, "kept_kmers": m.size}
tail = slice(min(2, reps - 1), None)   # the first passes settle the arenas
probes, builds, ranking = [], [], []
for rep in range(reps):
    g = Graph.buildGraph(K, m)
    ph = m.phase_ns()
    probes.append(ph["graph_masks_ns"] * 1e-6)
    ranking.append(ph["graph_rank_ns"] * 1e-6)
    builds.append(g.stats()["build_ns"] * 1e-6)
    launches, counts = g.stats()["jump_launches"], g.counts()
    g.close()
out["single_gpu"] = {"probes_ms_min": min(probes[tail]), "probes_ms_all": probes, "ranking_ms_min": min(ranking[tail]),
                     "jump_launches": launches, "build_ms_min": min(builds[tail]), "counts": counts}
P = int(os.environ.get("SG_P", "8"))
builds = []
for rep in range(reps):
    g = Graph.buildGraphVirtualShards(K, m, P)
    builds.append(g.stats()["build_ns"] * 1e-6)
    same = counts == g.counts()
    g.close()
out["virtual_shards"] = {"P": P, "build_ms_min": min(builds[tail]), "build_ms_all": builds, "same_counts": same}
print(json.dumps(out))
