"""Upsert variants of the slab path on C2 (gb_tune exp bits: 2 = CAS first, 4 = persistent + next-item prefetch, 8 / 16 = register
allocation for 6 / 8 CTAs per SM).  Every variant must leave the same table (size, sum of counts, checksum of keys) as the default."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from genome_b200 import capi, synth
from genome_b200.dnamap import ArrayDNAMap

b, n, _ = synth.make_config("C2")
d = torch.zeros(b.size + 16, dtype=torch.uint8, device="cuda")
d[:b.size].copy_(torch.from_numpy(b))
cap = 40_200_000
ref = None
variants = [int(x) for x in sys.argv[1:]] or [0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22]
for e in variants:
    with capi.tuned(exp=e):
        m = ArrayDNAMap(31, cap)
        rows = []
        for r in range(6):
            m.clear(cap)
            w = m.insert_reads_device(d.data_ptr(), b.size, n)
            s = m.stats()
            rows.append((s["last_insert_ns"] / 1e6, s["bucket_ns"] / 1e6, s["upsert_ns"] / 1e6))
        ek, ev = m.export()
        sig = (int(m.size), int(ev.astype(np.int64).sum()), int(np.bitwise_xor.reduce(ek)), int((ek.astype(np.uint64) * ev.astype(np.uint64)).sum() & 0xFFFFFFFFFFFF))
        if ref is None:
            ref = sig
        m.close()
    a = np.array(rows[2:])
    print(json.dumps({"bench": "insert C2", "exp": e, "cas_first": bool(e & 2), "persistent": bool(e & 4), "ctas_per_sm": 8 if e & 16 else 6 if e & 8 else 4,
                      "insert_ms": float(a[:, 0].mean()), "bucket_ms": float(a[:, 1].mean()), "upsert_ms": float(a[:, 2].mean()),
                      "same_table_as_default": sig == ref}), flush=True)
