"""Upsert ablations on C2 (timing only; the tables of the ablated runs are wrong by construction).  exp bits 6-7: 1 = staged keys +
hashing + table key loads, 2 = staged keys + hashing only, 3 = loads + one red per key (no compare-and-swap)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genome_b200 import capi, synth
from genome_b200.dnamap import ArrayDNAMap
b, n, _ = synth.make_config("C2")
d = torch.zeros(b.size + 16, dtype=torch.uint8, device="cuda"); d[:b.size].copy_(torch.from_numpy(b))
cap = 40_200_000
for name, kw in [("full upsert, table cleared first", dict(lazy_clear=0)), ("staged keys + hashing only", dict(lazy_clear=0, exp=2 << 6)),
                 ("+ table key loads", dict(lazy_clear=0, exp=1 << 6)), ("+ a red per key, no CAS", dict(lazy_clear=0, exp=3 << 6)),
                 ("full upsert, clear fused (slice-wise init)", dict(lazy_clear=1))]:
    with capi.tuned(**kw):
        m = ArrayDNAMap(31, cap)
        rows = []
        for r in range(6):
            m.clear(cap)
            m.insert_reads_device(d.data_ptr(), b.size, n)
            s = m.stats()
            rows.append((s["last_insert_ns"] / 1e6, s["bucket_ns"] / 1e6, s["upsert_ns"] / 1e6))
        m.close()
    a = np.array(rows[2:])
    print(json.dumps({"variant": name, "insert_ms": float(a[:, 0].mean()), "bucket_ms": float(a[:, 1].mean()), "upsert_ms": float(a[:, 2].mean())}), flush=True)
