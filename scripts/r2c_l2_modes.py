"""Why does 'load, then a dependent red' cost more than a load plus a red?  (profiles/r2b_insert_sweep.jsonl: 0.34 + 0.44 ms
apart, 1.40 ms together on an L2-resident region.)  gb_bench_l2_requests variants; one JSON object per line."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genome_b200 import capi

L = capi.lib()
N = 96_600_000
NAMES = {1: "load", 2: "red", 3: "load->red", 4: "load->cas->red", 5: "load + independent red", 6: "load->red on another slot",
         7: "atom.add with return", 8: "load16->red", 9: "load->store"}


def run(region_mb, base, kpt_log2=0, persist=0):
    ns = C.c_int64()
    capi.check(L.gb_bench_l2_requests(0, region_mb << 20, N, base + 100 * kpt_log2 + 1000 * persist, 5, C.byref(ns)))
    print(json.dumps({"bench": "l2_requests", "region_mb": region_mb, "mode": NAMES[base], "in_flight_per_thread": 1 << (kpt_log2 or 2),
                      "persistent_ctas_per_sm": persist, "ms": ns.value * 1e-6, "gupdates_per_s": N / ns.value,
                      "cycles_per_update_per_sm": ns.value * 1e-9 * 1.965e9 * 148 / N}), flush=True)


for base in range(1, 10):
    run(64, base)
for base in (1, 2, 3, 4):
    for k in (1, 3, 4):
        run(64, base, k)
for base in (1, 2, 3, 4, 5):
    for persist in (4, 8):
        for k in (1, 2, 3):
            run(64, base, k, persist)
for base in (3, 4):
    run(2048, base, 2, 8)
