"""Writes tests/golden/small_k15.json from the ORACLE (the reference cannot be built or imported here: Scala 2.9.1 +
Akka/Kryo SNAPSHOT jars, no JVM in the image).  The fixture therefore guards against drift of the oracle and lets the
GPU tests check one fixed input against committed outputs; it does not pin the oracle to the reference."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from tests import helpers as H  # noqa: E402

k, rounds = 15, 2
b, n, _ = H.small_reads(1200, 50, 14, 0.015, seed=20261018)
m = pyoracle.OracleMap(k)
w = m.insert_reads(b, n)
keys, vals = m.export_sorted()
m.delete_below(rounds)
g = pyoracle.OracleGraph(m)
nodes, edges = H.canon_oracle_graph(g)
g.retain_largest()
g.simplify()
_, edges2 = H.canon_oracle_graph(g)
out = dict(k=k, rounds=rounds, n_reads=n, windows=w, bin_hex=b.tobytes().hex(), keys=[int(x) for x in keys],
           counts=[int(x) for x in vals], nodes=nodes, edges=[[u, v, s.hex()] for (u, v, s) in edges],
           edges_after_retain_simplify=[[u, v, s.hex()] for (u, v, s) in edges2])
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_k15.json"), "w") as f:
    json.dump(out, f)
print("nodes", len(nodes), "edges", len(edges), "after", len(edges2), "keys", len(keys))
