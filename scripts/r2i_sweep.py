"""C2 on one GPU: keys per thread x CTAs per SM of the slab upsert (gb_tune exp >> 8), and the counting table's slots per key.
Every variant must leave the same table (size, sum of counts, checksums) as the default."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genome_b200 import capi, synth
from genome_b200.dnamap import ArrayDNAMap
b, n, _ = synth.make_config("C2")
d = torch.zeros(b.size + 16, dtype=torch.uint8, device="cuda"); d[:b.size].copy_(torch.from_numpy(b))
cap = 40_200_000
ref = None
NAMES = {0: "4 keys/thread, 6 CTAs/SM (default)", 1: "2 keys, 6 CTAs", 2: "3 keys, 6 CTAs", 3: "6 keys, 5 CTAs", 4: "2 keys, 8 CTAs", 5: "3 keys, 8 CTAs", 6: "4 keys, 8 CTAs"}
for kw in [dict(exp=e << 8) for e in range(7)] + [dict(count_cap_x10=25), dict(count_cap_x10=20), dict(count_cap_x10=20, slice_bits=4), dict(count_cap_x10=15)]:
    with capi.tuned(**kw):
        m = ArrayDNAMap(31, cap)
        rows = []
        for r in range(6):
            m.clear(cap)
            t0 = torch.cuda.Event(enable_timing=True)
            w = m.insert_reads_device(d.data_ptr(), b.size, n)
            s = m.stats()
            rows.append((s["last_insert_ns"] / 1e6, s["bucket_ns"] / 1e6, s["upsert_ns"] / 1e6))
        ek, ev = m.export()
        sig = (int(m.size), int(ev.astype(np.int64).sum()), int(np.bitwise_xor.reduce(ek)), int((ek.astype(np.uint64) * ev.astype(np.uint64)).sum() & 0xFFFFFFFFFFFF))
        ref = ref or sig
        table_gb = s["table_bytes"] / 1e9
        m.delete_below(3)
        ph = m.phase_ns()
        m.close()
    a = np.array(rows[2:])
    print(json.dumps({"tune": kw, "what": NAMES.get(kw.get("exp", 0) >> 8) if "exp" in kw else "counting table slots per key x10", "table_gb": table_gb,
                      "insert_ms": float(a[:, 0].mean()), "bucket_ms": float(a[:, 1].mean()), "upsert_ms": float(a[:, 2].mean()),
                      "filter_sweep_ms": ph["filter_sweep_ns"] / 1e6, "same_table_as_default": sig == ref}), flush=True)
