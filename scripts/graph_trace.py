import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genome_b200 import synth
from genome_b200.dnamap import ArrayDNAMap
from genome_b200.graph import Graph
cov = int(sys.argv[1]); ks = [int(x) for x in sys.argv[2].split(",")]
b, n, _ = synth.make_config("C2", coverage=cov)
d = torch.zeros(b.size + 16, dtype=torch.uint8, device="cuda"); d[:b.size].copy_(torch.from_numpy(b))
for k in ks:
    m = ArrayDNAMap(k, int(4.6e6 + n * 100 * 0.01 * (k - 9)) + 1_000_000)
    m.insert_reads_device(d.data_ptr(), b.size, n); m.delete_below(3); m.sync()
    print("k", k, "kept", m.size, file=sys.stderr)
    t = time.perf_counter()
    g = Graph.buildGraph(k, m)
    print("host build ms %.2f" % ((time.perf_counter() - t) * 1e3), g.stats(), g.counts(), file=sys.stderr)
    g.close(); m.close()
