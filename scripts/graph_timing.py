"""Times every stage of the graph path on one GPU (debugging aid; not part of the product)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genome_b200 import synth
from genome_b200.dnamap import ArrayDNAMap, PairedEndData
from genome_b200.graph import Graph

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
b, n, genome = synth.make_config(cfg, scale=scale)
m = ArrayDNAMap(31, int(b.size * 1.2))


def T(label, f):
    m.sync()
    t = time.perf_counter()
    r = f()
    m.sync()
    print("%-28s %9.3f ms" % (label, (time.perf_counter() - t) * 1e3), r if r is not None else "")
    return r


for rep in range(2):
    print("--- pass", rep)
    T("clear", lambda: m.clear(int(b.size * 1.2)))
    T("insert_reads(host)", lambda: m.insert_reads(b, n))
    T("size", lambda: m.size)
    T("delete_below", lambda: m.delete_below(3))
    g = T("buildGraph", lambda: Graph.buildGraph(31, m))
    print(g.stats(), g.counts())
    T("components", lambda: g.components()[0])
    T("retain_largest", lambda: g.retain_largest())
    T("simplifyGraph", lambda: g.simplifyGraph())
    T("removeBubbles", lambda: g.removeBubbles())
    T("clipTips", lambda: g.clipTips(62))
    T("simplifyGraph", lambda: g.simplifyGraph())
    T("check", lambda: g.check())
    print(g.counts())
    g.close()
