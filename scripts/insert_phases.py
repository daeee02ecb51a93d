"""Phase times of one insert of a named config on one GPU (debugging aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genome_b200 import synth
from genome_b200.dnamap import ArrayDNAMap
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
b, n, _ = synth.make_config(cfg)
d = torch.zeros(b.size + 16, dtype=torch.uint8, device="cuda"); d[:b.size].copy_(torch.from_numpy(b))
cap = 40_000_000 if cfg == "C2" else 6_000_000
m = ArrayDNAMap(31, cap)
for r in range(reps):
    m.clear(cap)
    w = m.insert_reads_device(d.data_ptr(), b.size, n)
    s = m.stats()
    print("insert %.3f ms  bucket %.3f ms  upsert %.3f ms  (%.2f G k-mers/s) size %d cap %d" % (s["last_insert_ns"]/1e6, s["bucket_ns"]/1e6, s["upsert_ns"]/1e6, w/s["last_insert_ns"], m.size, s["capacity"]))
