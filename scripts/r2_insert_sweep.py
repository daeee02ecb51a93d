"""Insert experiments of round 2 on one GPU (tuning run, C2 = BASELINE configs[1]):
  1. gb_bench_smem_upsert: the shared-memory upsert rate (design question of DESIGN.md 3.1)
  2. gb_bench_random_atomics: R_gups of SURVEY 8(d) with the re-specified kernel
  3. phase times of the L2-blocked insert under the gb_tune variants (prefetch, slice size, counted vs single pass)
One JSON object per line on stdout."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from genome_b200 import capi, synth
from genome_b200.dnamap import ArrayDNAMap

L = capi.lib()
for slots_log2, kpb, ctas in [(13, 13000, 2), (12, 6500, 4), (11, 3250, 8), (13, 13000, 1)]:
    n_buckets = int(96.6e6 // kpb)
    ns, dk = C.c_int64(), C.c_int64()
    capi.check(L.gb_bench_smem_upsert(0, slots_log2, kpb, n_buckets, ctas, 5, C.byref(ns), C.byref(dk)))
    n = kpb * n_buckets
    print(json.dumps({"bench": "smem_upsert", "slots": 1 << slots_log2, "keys_per_bucket": kpb, "buckets": n_buckets, "ctas_per_sm": ctas,
                      "keys": n, "distinct": dk.value, "load": dk.value / n_buckets / (1 << slots_log2), "ms": ns.value * 1e-6,
                      "gkeys_per_s": n / ns.value}), flush=True)

for gb_ in (0.25, 1.0, 1.93, 4.0):
    ns = C.c_int64()
    capi.check(L.gb_bench_random_atomics(0, int(gb_ * 1e9), 96_600_000, 5, C.byref(ns)))
    print(json.dumps({"bench": "random_atomics_u64", "table_gb": gb_, "updates": 96_600_000, "ms": ns.value * 1e-6, "gupdates_per_s": 96.6e6 / ns.value}), flush=True)

for region_mb in (64, 2048):
    for mode, what in ((1, "load"), (2, "red"), (3, "load+red"), (4, "load+cas+red")):
        ns = C.c_int64()
        capi.check(L.gb_bench_l2_requests(0, region_mb << 20, 96_600_000, mode, 5, C.byref(ns)))
        print(json.dumps({"bench": "l2_requests", "region_mb": region_mb, "mode": what, "ms": ns.value * 1e-6, "gupdates_per_s": 96.6e6 / ns.value,
                          "cycles_per_update_per_sm_at_1.965GHz": ns.value * 1e-9 * 1.965e9 * 148 / 96.6e6}), flush=True)

b, n, _ = synth.make_config("C2")
d = torch.zeros(b.size + 16, dtype=torch.uint8, device="cuda")
d[:b.size].copy_(torch.from_numpy(b))
cap = 40_200_000
VARIANTS = [
    ("default (single pass)", {}),
    ("default again", {}),
    ("exp=1: survivors through update_keys_kernel", dict(exp=1)),
    ("slices=64", dict(slice_bits=6)),
    ("slices=16", dict(slice_bits=4)),
    ("counted passes", dict(single_pass=0)),
    ("counted, 2 sub-batches", dict(single_pass=0, batches=2)),
    ("direct (no bucket pass)", dict(insert_path=1)),
]
ref = None
for name, kw in VARIANTS:
    with capi.tuned(**kw):
        m = ArrayDNAMap(31, cap)
        rows = []
        for r in range(6):
            m.clear(cap)
            w = m.insert_reads_device(d.data_ptr(), b.size, n)
            s = m.stats()
            rows.append((s["last_insert_ns"] / 1e6, s["bucket_ns"] / 1e6, s["upsert_ns"] / 1e6))
        size = m.size
        if ref is None:
            ref = size
        assert size == ref, (name, size, ref)
        m.delete_below(3)
        ph = m.phase_ns()
        m.close()
    a = np.array(rows[2:])
    print(json.dumps({"bench": "insert C2", "variant": name, "tune": kw, "insert_ms": float(a[:, 0].mean()), "bucket_ms": float(a[:, 1].mean()),
                      "upsert_ms": float(a[:, 2].mean()), "gkmers_per_s": w / a[:, 0].mean() / 1e6, "distinct": size,
                      "filter_sweep_ms": ph["filter_sweep_ns"] / 1e6, "filter_reinsert_ms": ph["filter_reinsert_ns"] / 1e6}), flush=True)
