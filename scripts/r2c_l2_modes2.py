"""Second set of gb_bench_l2_requests variants: is the cost of mixing loads and reds on the SM side or in the L2?"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genome_b200 import capi

L = capi.lib()
N = 96_600_000
NAMES = {1: "load", 2: "red", 3: "load->red", 5: "load + independent red", 10: "load half A -> red half B", 11: "even SMs load, odd SMs red (half the work each)",
         12: "even warps load, odd warps red (half the work each)", 13: "SoA: load keys[] -> red counts[]", 14: "line-blocked: load key -> red count in another sector of the line",
         4: "load->cas->red", 15: "SoA: load -> cas keys[] -> red counts[]", 16: "line-blocked: load -> cas -> red"}
for region_mb in (64, 2048):
    for base in (3, 13, 14, 4, 15, 16):
        ns = C.c_int64()
        capi.check(L.gb_bench_l2_requests(0, region_mb << 20, N, base, 5, C.byref(ns)))
        print(json.dumps({"bench": "l2_requests", "region_mb": region_mb, "mode": NAMES[base], "ms": ns.value * 1e-6, "gupdates_per_s": N / ns.value,
                          "cycles_per_update_per_sm": ns.value * 1e-9 * 1.965e9 * 148 / N}), flush=True)
