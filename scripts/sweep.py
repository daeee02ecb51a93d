"""BASELINE.json configs[4]: k-mer insert sweep k in {21,25,31} x coverage in {10,30,100} on the C2 genome (1 GPU), plus one
large case that exercises the multi-pass paths (configs[2] scaled to one GPU).  Writes JSON lines to stdout."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from genome_b200 import synth
from genome_b200.dnamap import ArrayDNAMap
from genome_b200.graph import Graph

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def run(name, b, n, k, cap, rounds=3, graph=True, reps=3):
    d = torch.zeros(b.size + 16, dtype=torch.uint8, device="cuda")
    d[:b.size].copy_(torch.from_numpy(b))
    m = ArrayDNAMap(k, cap)
    best = None
    for _ in range(reps):
        m.clear(cap)
        w = m.insert_reads_device(d.data_ptr(), b.size, n)
        s = m.stats()
        if best is None or s["last_insert_ns"] < best["last_insert_ns"]:
            best = s
        distinct = m.size
    keys, vals = (None, None)
    total = None
    if w <= 200_000_000:
        keys, vals = m.export()
        total = int(vals.astype(np.int64).sum())
        assert total == w and len(np.unique(keys)) == keys.size
    m.delete_below(rounds)
    kept = m.size
    out = dict(case=name, k=k, reads=n, kmer_instances=w, distinct=distinct, kept=kept, table_gb=best["table_bytes"] / 1e9,
               insert_ms=best["last_insert_ns"] / 1e6, bucket_ms=best["bucket_ns"] / 1e6, upsert_ms=best["upsert_ns"] / 1e6,
               gkmers_per_s=w / best["last_insert_ns"], sum_counts_ok=total == w if total is not None else None,
               hbm_frac_algorithmic=(16.0 * w + b.size + 8.0 * distinct) / (best["last_insert_ns"] * 1e-9) / 1e9 / PEAK)
    if graph:
        t = time.perf_counter()
        g = Graph.buildGraph(k, m)
        torch.cuda.synchronize()
        out["graph_build_ms"] = (time.perf_counter() - t) * 1e3
        nn, ne, nb = g.counts()
        g.check()
        st = g.stats()
        out.update(nodes=nn, edges=ne, edge_bases=nb, jump_launches=st["jump_launches"], cycle_vertices=st["cycle_vertices"],
                   build_kernels_ms=st["build_ns"] / 1e6)
        # sum of edge lengths = oriented kept k-mers - nodes + edges - unreached (isolated / perfect cycles)
        out["edge_length_identity_slack"] = (2 * kept - nn + ne) - nb
        t = time.perf_counter()
        nc, _ = g.components()
        g.retain_largest()
        g.simplifyGraph()
        torch.cuda.synchronize()
        out["components_retain_simplify_ms"] = (time.perf_counter() - t) * 1e3
        out["components"] = nc
        out["after"] = g.counts()
        g.close()
    m.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "sweep"
    if which in ("sweep", "all"):
        for cov in (10, 30, 100):
            b, n, _ = synth.make_config("C2", coverage=cov)
            for k in (21, 25, 31):
                run("C2 cov=%d" % cov, b, n, k, int(4.6e6 + n * 100 * 0.01 * (k - 9)) + 1_000_000)
    if which == "huge":
        # configs[3] (1 Gbp, 150 bp, 40x, 0.5% errors) scaled to 1/20 on ONE GPU: a multi-GB table, several key-staging passes
        b, n, _ = synth.make_config("C4", scale=0.05)
        run("C4 x0.05 (50 Mbp, 0.5% errors, 150 bp, 40x)", b, n, 31, 330_000_000, reps=2)
    if which in ("large", "all"):
        # configs[2] (100 Mbp, 5% repeats, 150 bp, 50x) scaled to 1/5 so that reads + table fit the time budget of one call
        b, n, _ = synth.make_config("C3", scale=0.2)
        run("C3 x0.2 (20 Mbp, 5% repeats, 150 bp, 50x)", b, n, 31, 24_000_000, reps=2)
