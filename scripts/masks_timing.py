"""Membership probes of Graph.buildGraph on one GPU: masks_kernel (one thread per stored k-mer) against masks_flat_kernel (one lane
per probe), same filtered table, CUDA-event times from gb_map_phase_ns; then the sharded build over P virtual ranks (its MasksOp
runs as items_kernel<MasksOp>).  One JSON object on stdout.
Usage: python scripts/masks_timing.py [C2] [scale] [reps]      (reps = 1: a single build per variant, for ncu)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genome_b200 import capi, synth  # noqa: E402
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
out = {"workload": cfg, "scale": scale, "kept_kmers": m.size, "variants": []}
ref = None
VARIANTS = ((1, 1),) if os.environ.get("MT_ONLY") == "sublists" else ((0, 0), (1, 0), (1, 1))
for flat, sub in VARIANTS:
    with capi.tuned(masks_flat=flat, rank_sublists=sub):
        probes, builds, ranking = [], [], []
        for rep in range(reps):
            g = Graph.buildGraph(K, m)
            ph = m.phase_ns()
            probes.append(ph["graph_masks_ns"] * 1e-6)
            ranking.append(ph["graph_rank_ns"] * 1e-6)
            builds.append(g.stats()["build_ns"] * 1e-6)
            g_launches = g.stats()["jump_launches"]
            counts = g.counts()
            if ref is None:
                ref = (counts, sorted(int(x) for x in g.export()[0]))
            same = ref == (counts, sorted(int(x) for x in g.export()[0]))
            g.close()
        tail = slice(min(2, reps - 1), None)   # the first passes settle the arenas
        out["variants"].append({"masks_flat": flat, "rank_sublists": sub, "kernel": "masks_flat_kernel" if flat else "masks_kernel",
                                "jump_launches": g_launches,
                                "probes_ms_min": min(probes[tail]), "probes_ms_all": probes, "build_ms_min": min(builds[tail]),
                                "ranking_ms_min": min(ranking[tail]), "counts": counts, "same_graph_as_first": same})
P = int(os.environ.get("SG_P", "8"))
out["virtual_shards"] = []
for flat in (0, 1):
    if (reps == 1 and flat == 0) or os.environ.get("MT_ONLY") == "sublists":
        continue                     # the ncu passes profile one form only
    with capi.tuned(masks_flat=flat):
        builds = []
        for rep in range(reps):
            g = Graph.buildGraphVirtualShards(K, m, P)
            builds.append(g.stats()["build_ns"] * 1e-6)
            same = ref[0] == g.counts()
            g.close()
    out["virtual_shards"].append({"P": P, "masks_flat": flat, "probes": "PartsOp + ProbeOp + CombineOp" if flat else "MasksOp",
                                  "build_ms_min": min(builds[min(2, reps - 1):]), "build_ms_all": builds, "same_counts": same})
print(json.dumps(out))
