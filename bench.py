#!/usr/bin/env python
"""bench.py -- k-mers inserted/s (FreqFilter.extractFilteredKmers) and graph build + simplify time on N B200s.

A step = one pass of the hot path's insert stage over the whole synthetic read set: a fresh table (clear), bulk
FreqFilter.add over every read (canonical k-mer extraction + hash-table upsert), deleteAll(v < 3).  `value` is
k-mer instances inserted per second with the `.bin` stream already resident in HBM; `e2e` is the same through the
host-buffer C-ABI calls (H2D of the stream and D2H of the size inside the timed region; pinned staging memory as
INTEGRATION.md binds it, and a pageable buffer beside it).  Graph.buildGraph, components, retain and simplifyGraph are
timed on the table the last step left behind and reported under "graph" with their own roofline entries.

Parity comes first: before anything is timed the table, the kept set and the graph of the workload are compared with
the oracle (N = 1: the whole workload; N > 1: every shard against the oracle's keys it owns, on a bounded prefix of
every rank's reads) and the line carries "parity_checked"; a mismatch aborts the run.

N = 1: BASELINE.json configs[1] (4.6 Mbp genome, 1% substitutions, 100 bp reads at 30x, k = 31).
N > 1 (torchrun): weak scaling -- every rank holds 1.38 M reads of an N x 4.6 Mbp genome, the table is one
shard per GPU (PartitionedDNAMap), k-mers are routed by the library; additionally BASELINE.json configs[2] (the 100 Mbp
genome over the N GPUs) is run for a few steps and reported under "named_configs".

`--impl reference`: the CPU restatement of the reference algorithm (oracle/, "port": the Scala/Akka reference cannot
be built here) on all host cores, the same whole read set per step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # first calls must not pay lazy kernel loading (10..400 ms stalls)

K = 31
COVERAGE = None   # --coverage: BASELINE configs[4] sweeps only; None = the named config's coverage
ROUNDS = 3  # GraphBuilder.scala:30
WORKLOADS = {
    # name: (synth config, scale, note)
    "C2": dict(cfg="C2", desc="4.6 Mbp random genome, 1% substitutions, 100 bp reads at 30x, k=31 (BASELINE configs[1])"),
    "C1": dict(cfg="C1", desc="4.6 Mbp random genome, error-free 100 bp reads at 30x, k=31 (BASELINE configs[0])"),
    # BASELINE configs[2]: the genome is scale x world x 100 Mbp, so `--workload C3 --scale 0.125 --gpus 8` (0.25 at 4, 0.5 at 2)
    # is the named 100 Mbp genome sharded over the GPUs
    "C3": dict(cfg="C3", desc="100 Mbp-class genome x scale with 5% interspersed repeats, error-free 150 bp reads at 50x, k=31 (BASELINE configs[2])"),
    # BASELINE configs[3] shape (150 bp reads at 40x, 0.5% errors); use --scale 0.05: 50 Mbp of genome and a 16 GB shard per GPU
    "C4": dict(cfg="C4", desc="1 Gbp-class random genome x scale, 0.5% substitutions, 150 bp reads at 40x, k=31 (BASELINE configs[3] shape)"),
}


def make_workload(name, rank, world, scale):
    """This rank's reads: the genome is `world` times the config's genome (same seed on every rank), the reads
    are this rank's own 30x / world share of it -- fixed work per GPU."""
    from genome_b200 import synth
    c = dict(synth.CONFIGS[WORKLOADS[name]["cfg"]])
    G = int(c["genome"] * scale) * world
    genome = synth.random_genome(G, c["seed"])
    if c["repeats"] > 0:
        genome = synth.add_repeats(genome, c["repeats"], c["seed"] + 1)   # same seed on every rank: the same genome
    n_reads = (int((COVERAGE or c["coverage"]) * G / c["read_len"] / world) // 2) * 2
    parts, done, i = [], 0, 0
    while done < n_reads:
        m = min(1 << 20, n_reads - done)
        parts.append(synth.pack_fixed(synth.sample_reads(genome, c["read_len"], m, c["err"], c["seed"] + 100 + 1000 * rank + i)))
        done += m
        i += 1
    b = np.concatenate(parts)
    windows = n_reads * (c["read_len"] - K + 1)
    # expected distinct k-mers of the whole job: genome + ~ (k windows per error, capped by read geometry)
    distinct_total = G + int(n_reads * world * c["read_len"] * c["err"] * 22)
    return b, n_reads, windows, G, distinct_total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (t, r) in self.rows if t_begin is None or (t_begin <= t <= t_end + 0.06)]
        window = "timed region"
        if not inside:  # region shorter than one sampling period: use the samples of warm-up + timed region
            inside, window = [r for (_, r) in self.rows], "warm-up + timed region (timed region shorter than a sampling period)"
        for r in inside:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def graph_traffic(workload, scale, k, root=ROOT):
    """DRAM bytes per build of the graph kernels from the committed ncu capture (profiles/graph_kernels_traffic.json), used only
    when it was taken on this workload: {kernel: (bytes, launches, source)}.  Never fails the bench: anything odd -> {}."""
    try:
        with open(os.path.join(root, "profiles", "graph_kernels_traffic.json")) as f:
            t = json.load(f)
        if t.get("workload") != workload or t.get("scale") != scale or t.get("k") != k:
            return {}
        return {name: (float(v["dram_bytes_read"]) + float(v["dram_bytes_write"]), int(v.get("launches", 1)), v.get("source"))
                for name, v in t.get("kernels", {}).items()}
    except Exception:
        return {}


def canon_graph(node_kmer, es, ee, off, bases, node_id=None):
    """Sorted node k-mers and sorted edge multiset keyed by node k-mers (ids are not reproducible)."""
    if node_id is None:
        nk = [int(x) for x in node_kmer]
        name = lambda i: nk[int(i)]
    else:
        by_id = {int(i): int(x) for i, x in zip(node_id, node_kmer)}
        nk = list(by_id.values())
        name = lambda i: by_id[int(i)]
    edges = sorted((name(es[i]), name(ee[i]), bases[int(off[i]):int(off[i + 1])].tobytes()) for i in range(es.size))
    return sorted(nk), edges


def oracle_pass(b, n_reads, threads):
    """One FreqFilter pass of the CPU port over the whole read set: ArrayDNAMap layout/probing/rescale, `threads`
    single-threaded partitions fed by extractor threads through in-memory buckets (no Kryo/TCP: strictly faster than
    the real reference).  Returns the map BEFORE deleteAll and the insert time."""
    from oracle import pyoracle
    m = pyoracle.OracleMap(K, partitions=threads)
    t0 = time.perf_counter()
    m.insert_reads(b, n_reads, threads=threads)
    return m, time.perf_counter() - t0


def cpu_reference_steps(b, n_reads, windows, threads, steps, warmup):
    """--impl reference: `steps` whole passes (insert + deleteAll) of the CPU port."""
    from oracle import pyoracle
    pyoracle.build()
    times = []
    for i in range(warmup + steps):
        m, t_ins = oracle_pass(b, n_reads, threads)
        t0 = time.perf_counter()
        m.delete_below(ROUNDS)
        dt = t_ins + time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        del m
    t = float(np.mean(times))
    return windows / t, t


def parity_check(rank, world, cores, b, n_reads, comm, make_map, Graph, dist, torch):
    """The CUDA path against the oracle BEFORE anything is timed.  N = 1: the whole workload -- sorted (k-mer, count)
    table, kept set after deleteAll(v < 3), node set and edge multiset of Graph.buildGraph.  N > 1: a bounded prefix of
    every rank's reads through the sharded map; every shard must hold exactly the oracle's keys it owns (the map's own
    owner function), then deleteAll(v < 2) and the graph built over the shards.  Returns (report, cpu numbers or None)."""
    from oracle import pyoracle
    pyoracle.build()
    rec = b.size // n_reads
    cpu = None
    if world == 1:
        threads, sb, sn, rounds = cores, b, n_reads, ROUNDS
        what = "whole workload: table, kept set, graph (nodes + edge multiset)"
    else:
        per = min(n_reads, 32768)
        mine = torch.from_numpy(np.ascontiguousarray(b[:per * rec])).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        sb, sn, rounds = torch.cat(parts).cpu().numpy(), per * world, 2
        threads = max(1, cores // world)
        what = "first %d reads of every rank through the sharded map: shard == oracle keys it owns, kept set, graph" % per
    oracle_pass(sb, sn, threads)  # warm the allocator / page cache: the timed pass below is the CPU baseline's
    om, t_ins = oracle_pass(sb, sn, threads)
    ow = pyoracle.count_windows(sb, sn, K)
    m = make_map(max(1024, int(ow * 0.4)))
    if world == 1:
        gw = m.insert_reads(sb, sn)
    else:
        gw = m.insert_reads(np.ascontiguousarray(b[:(sn // world) * rec]), sn // world)
        t = torch.tensor([gw], device="cuda")
        dist.all_reduce(t)
        gw = int(t.item())
    ok_ = gw == ow and m.size == om.size()
    okk, okv = om.export_sorted()
    gk, gv = m.export_sorted()
    if world > 1:
        sel = m.owner(okk) == rank
        okk, okv = okk[sel], okv[sel]
    table_ok = ok_ and np.array_equal(gk, okk) and np.array_equal(gv, okv)
    del gk, gv, okk, okv
    t0 = time.perf_counter()
    om.delete_below(rounds)
    t_del = time.perf_counter() - t0
    m.delete_below(rounds)
    okk, okv = om.export_sorted()
    gk, gv = m.export_sorted()
    if world > 1:
        sel = m.owner(okk) == rank
        okk, okv = okk[sel], okv[sel]
    kept_ok = m.size == om.size() and np.array_equal(gk, okk) and np.array_equal(gv, okv)
    kept = om.size()
    del gk, gv, okk, okv
    g = Graph.buildGraph(K, m)
    t0 = time.perf_counter()
    og = pyoracle.OracleGraph(om)
    t1 = time.perf_counter()
    gn, ge = canon_graph(*g.export())
    node_kmer, node_id, es, ee, off, bases = og.export()
    on, oe = canon_graph(node_kmer, es, ee, off, bases, node_id)
    graph_ok = g.counts() == og.counts() and gn == on and ge == oe
    del gn, ge, on, oe
    t2 = time.perf_counter()
    og.components()
    og.retain_largest()
    og.simplify()
    t3 = time.perf_counter()
    g.retain_largest()
    g.simplifyGraph()
    simp_ok = g.counts() == og.counts()
    g.close()
    m.close()
    if world == 1:
        cpu = {"value": ow / (t_ins + t_del), "unit": "k-mers/s", "cores": threads, "kind": "port",
               "sample": "the whole workload: %d reads, %d k-mer instances, %d partitions/threads, second of two passes "
                         "(insert %.2f s + deleteAll %.2f s)" % (sn, ow, threads, t_ins, t_del),
               "graph": {"kept_kmers": kept, "build_ms": (t1 - t0) * 1e3, "components_retain_simplify_ms": (t3 - t2) * 1e3, "cores": 1}}
    flags = {"table": bool(table_ok), "kept_set": bool(kept_ok), "graph": bool(graph_ok), "after_retain_simplify": bool(simp_ok)}
    all_ok = all(flags.values())
    if world > 1:
        t = torch.tensor([1 if all_ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        all_ok = bool(t.item())
    report = {"ok": all_ok, "against": "oracle/ (CPU restatement of the reference; parity unpinned, DESIGN.md section 6)", "what": what,
              "kmer_instances": int(ow), "kept_kmers": int(kept), "rank0": flags}
    return report, cpu


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale (debugging only; 1.0 is the named config)")
    ap.add_argument("--k", type=int, default=31, help="BASELINE configs[4] sweeps only; 31 is every named config's k")
    ap.add_argument("--coverage", type=float, default=None, help="BASELINE configs[4] sweeps only; default = the named config's")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the parity check and the CPU baseline that comes out of it (tuning runs only)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-named", action="store_true", help="N > 1: skip the extra BASELINE configs[2] run")
    args = ap.parse_args()
    global K, COVERAGE
    K, COVERAGE = args.k, args.coverage
    if not 1 <= K <= 31:
        raise SystemExit("--k must be in 1..31")

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    desc = WORKLOADS[args.workload]["desc"]
    if K != 31 or COVERAGE is not None:
        desc += " -- SWEEP VARIANT (BASELINE configs[4]): k=%d, coverage %s" % (K, "as named" if COVERAGE is None else "%gx" % COVERAGE)

    if args.impl == "reference":
        if rank != 0:
            return 0
        b, n_reads, windows, G, _ = make_workload(args.workload, 0, 1, args.scale)
        v, t = cpu_reference_steps(b, n_reads, windows, cores, args.steps, max(args.warmup, 1))
        line = {
            "impl": "reference", "metric": "k-mers inserted/s (FreqFilter.extractFilteredKmers: insert + deleteAll)",
            "value": v, "unit": "k-mers/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, desc), "k": K, "rounds": ROUNDS, "reads_per_gpu": n_reads,
                       "kmer_instances_per_gpu": windows, "genome_bp": G, "scale": args.scale},
            "cpu_baseline": {"value": v, "unit": "k-mers/s", "cores": cores, "kind": "port",
                             "sample": "the whole workload per step: %d reads, %d k-mer instances, %d partitions/threads" % (n_reads, windows, cores)},
            "e2e": {"value": v, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from genome_b200 import capi
    from genome_b200.dnamap import ArrayDNAMap, PartitionedDNAMap, Communicator, torch_broadcast
    from genome_b200.graph import Graph

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tuning = capi.tune_from_env()  # GENOME_B200_TUNE (tuning runs only; the defaults are the measured choices)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    L = capi.lib()
    comm = Communicator(rank, world, local_rank, torch_broadcast) if world > 1 else None

    def make_map(cap):
        return PartitionedDNAMap(K, comm, cap) if world > 1 else ArrayDNAMap(K, cap, device=local_rank)

    def run_workload(name, scale, steps, warmup, with_e2e, with_graph, identities=False):
        """Timed loop over one synthetic workload; returns a dict of raw measurements (all ranks call it together)."""
        b, n_reads, windows, G, distinct_total = make_workload(name, rank, world, scale)
        cap = int(distinct_total / world * 1.15) + 1024  # distinct keys expected on this shard
        out = {"b": b, "n_reads": n_reads, "windows": windows, "G": G, "cap": cap}
        hp = C.c_void_p()
        capi.check(L.gb_host_alloc(b.size, C.byref(hp)))
        pinned = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(b.size,))
        pinned[:] = b
        d_bin = torch.empty(b.size + 16, dtype=torch.uint8, device="cuda")
        d_bin[:b.size].copy_(torch.from_numpy(b))
        torch.cuda.synchronize()
        m = make_map(cap)
        insert_stats = {}

        def step_device():
            m.clear(cap)
            w = m.insert_reads_device(d_bin.data_ptr(), b.size, n_reads)
            insert_stats.update(m.stats())  # table size and insert-kernel time before the filter shrinks the table
            m.delete_below(ROUNDS)
            return w

        def step_host(buf):
            m.clear(cap)
            w = m.insert_reads(buf, n_reads)
            m.delete_below(ROUNDS)
            return w, m.size  # the size read is the step's device->host result

        if identities and cap <= 250_000_000:  # beyond that the export alone is GBs per rank
            # size-independent properties at full size: the counts of all shards add up to the k-windows issued, no key twice
            m.clear(cap)
            assert m.insert_reads_device(d_bin.data_ptr(), b.size, n_reads) == windows
            ek, ev = m.export()
            tot = sum_over_ranks(float(ev.astype(np.int64).sum()))
            out["identities"] = {"sum_of_counts_equals_kmer_instances": bool(tot == sum_over_ranks(float(windows))),
                                 "keys_distinct_on_this_shard": bool(np.unique(ek).size == ek.size),
                                 "distinct_keys_total": int(sum_over_ranks(float(ek.size)))}
            del ek, ev
        sampler = ClockSampler(local_rank)
        sampler.start()
        for _ in range(warmup):
            assert step_device() == windows
        barrier()
        launches0 = L.gb_launch_count()
        insert_ns = []
        barrier()
        t0 = time.perf_counter()
        m.timer_start()
        for _ in range(steps):
            step_device()
            insert_ns.append(insert_stats["last_insert_ns"])
        dev_ns = m.timer_stop()
        barrier()
        t1 = time.perf_counter()
        out["launches"] = L.gb_launch_count() - launches0
        out["clocks"] = sampler.stop(t0, t1)
        out["step_s"] = max_over_ranks(dev_ns * 1e-9 / steps)
        out["wall_step_s"] = max_over_ranks((t1 - t0) / steps)
        out["total_windows"] = sum_over_ranks(float(windows))
        out["kept_total"] = m.size
        out["table_bytes"] = insert_stats["table_bytes"]
        out["insert_ns"] = insert_ns
        out["phase"] = m.phase_ns()  # of the last step: bucket pass, upsert, filter sweep, survivors' re-insert
        if with_e2e:
            for kind, buf in (("pinned", pinned), ("pageable", b)):
                for _ in range(2):
                    step_host(buf)
                barrier()
                t0 = time.perf_counter()
                for _ in range(steps):
                    step_host(buf)
                barrier()
                out["e2e_%s_s" % kind] = max_over_ranks((time.perf_counter() - t0) / steps)
        if with_graph:
            # graph stage on the filtered table: two passes settle the scratch arenas (growth, then consolidation), then timed
            # passes; device time from the library's own CUDA events, wall time beside it (collective for N > 1)
            build, simp, kern = [], [], []
            reps = 5
            for rep in range(2 + reps):
                step_device()
                barrier()
                t0 = time.perf_counter()
                g = Graph.buildGraph(K, m)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                nn, ne, nb = g.counts()
                nc, _ = g.components()
                g.retain_largest()
                g.simplifyGraph()
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                if rep >= 2:
                    build.append(max_over_ranks((t1 - t0) * 1e3))
                    simp.append(max_over_ranks((t2 - t1) * 1e3))
                    kern.append(g.stats()["build_ns"] * 1e-6)
                gs = g.stats()
                after = g.counts()
                g.close()
            ph = m.phase_ns()
            # sum of edge lengths = (2 kept - oriented terminals - oriented k-mers that are isolated or on perfect
            # cycles) + edges; isolated k-mers are not counted by the build, so the identity is reported as its slack (>= 0, even)
            slack = (2 * out["kept_total"] - nn + ne) - nb - gs["cycle_vertices"]
            out["graph"] = {"build_ms": float(np.median(build)), "build_ms_all": build, "build_kernels_ms": float(np.median(kern)),
                            "edge_length_identity_slack_isolated_kmers": int(slack),
                            "components_retain_simplify_ms": float(np.median(simp)), "timed_passes": reps,
                            "nodes": nn, "edges": ne, "edge_bases": nb, "components": nc, "jump_launches": gs["jump_launches"],
                            "after_simplify": after, "kept_kmers": out["kept_total"],
                            "sharding": ("single GPU" if world == 1 else
                                         "sharded: minimizer owners, rank-local list ranking, segment list (no replica)"
                                         if capi.get_tune("pgraph_sharded") else "replicated after all-gather of the shards")}
            out["graph_phase"] = ph
        out["m"], out["d_bin"], out["hp"] = m, d_bin, hp
        return out

    def release(r):
        r["m"].close()
        capi.check(L.gb_host_free(r["hp"]))
        r["d_bin"] = None

    # ---------------- parity first (and, at N = 1, the CPU baseline that falls out of the oracle run)
    parity, cpu = None, None
    if not args.no_cpu_baseline:
        b0, n0, _, _, _ = make_workload(args.workload, rank, world, args.scale)
        parity, cpu = parity_check(rank, world, cores, b0, n0, comm, make_map, Graph, dist, torch)
        del b0
        if not parity["ok"]:
            if rank == 0:
                print(json.dumps({"parity_checked": False, "parity": parity}))
            raise SystemExit("bench.py: the CUDA path differs from the oracle -- nothing is timed")

    R = run_workload(args.workload, args.scale, args.steps, args.warmup, True, not args.no_graph, identities=args.workload != "C2" or args.scale != 1.0)
    m, b, n_reads, windows, cap = R["m"], R["b"], R["n_reads"], R["windows"], R["cap"]
    value = R["total_windows"] / R["step_s"]
    table_bytes = R["table_bytes"]
    insert_ns = R["insert_ns"]
    e2e = {"value": R["total_windows"] / R["e2e_pinned_s"], "unit": "k-mers/s", "h2d_bytes_per_step": int(b.size), "d2h_bytes_per_step": 8 + 4 * 32,
           "ms_per_step": R["e2e_pinned_s"] * 1e3,
           "host_buffer": "pinned (gb_host_alloc, the buffer INTEGRATION.md's binding passes)",
           "pageable": {"value": R["total_windows"] / R["e2e_pageable_s"], "ms_per_step": R["e2e_pageable_s"] * 1e3,
                        "host_buffer": "pageable (a plain numpy array; what a JNA Array[Byte] argument becomes)"}}

    # ---------------- roofline of the dominant kernel (insert), live CUDA-event durations from inside the library
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    roofline = None
    roofline_graph = None
    d_bin = R["d_bin"]
    if world == 1:
        m.clear(cap)
        m.insert_reads_device(d_bin.data_ptr(), b.size, n_reads)
        distinct = m.size  # size before the filter = distinct keys
        st = m.stats()
        ins_s = float(np.mean(insert_ns)) * 1e-9
        # 16 B table read/write per instance + input stream + 8 B first-touch key write per distinct key
        algo_bytes = 16.0 * windows + float(b.size) + 8.0 * distinct
        ns = C.c_int64()
        capi.check(L.gb_bench_random_atomics(local_rank, table_bytes, windows, 5, C.byref(ns)))
        gups = windows / (ns.value * 1e-9)
        partitioned = st["upsert_ns"] > 0
        single_pass = bool(capi.get_tune("single_pass")) and windows >= capi.get_tune("single_pass_min")
        if partitioned:
            dom, dom_s = ("insert_slabs_kernel" if single_pass else "insert_keys_kernel"), st["upsert_ns"] * 1e-9
            path = ("L2-blocked, single pass: bucket_slabs_kernel + insert_slabs_kernel" if single_pass
                    else "L2-blocked: part_count_kernel + part_scatter_kernel + insert_keys_kernel")
        else:
            dom, dom_s, path = "insert_reads_kernel", ins_s, "direct: insert_reads_kernel"
        achieved = algo_bytes / dom_s / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_kmer": algo_bytes / windows,
                    "algorithmic_bytes": "SURVEY 8(d): 16 B table + input stream + 8 B x distinct/instances, per k-mer instance, x instances per launch",
                    "kernel_ms": dom_s * 1e3, "insert_path": path,
                    # the same launch against what it has to move given the design (the staged 8-byte key is read back too)
                    "kernel_bytes_per_kmer_incl_staged_key": (algo_bytes + 8.0 * windows - float(b.size)) / windows if partitioned else None,
                    "insert_ms": ins_s * 1e3, "insert_kmers_per_s": windows / ins_s,
                    "insert_frac": algo_bytes / ins_s / 1e9 / peak,
                    "step_frac": algo_bytes / R["step_s"] / 1e9 / peak,
                    "phases_ms": {"bucket pass": st["bucket_ns"] * 1e-6, "upsert": st["upsert_ns"] * 1e-6,
                                  "deleteAll sweep": R["phase"]["filter_sweep_ns"] * 1e-6, "deleteAll survivors into their table": R["phase"]["filter_reinsert_ns"] * 1e-6,
                                  "rest of the step (clear, counters, launch gaps)": (R["step_s"] - ins_s) * 1e3 - (R["phase"]["filter_sweep_ns"] + R["phase"]["filter_reinsert_ns"]) * 1e-6}
                    if partitioned else None,
                    "random_access_ceiling": {"kernel": "one random 64-bit atomicAdd per element over an array of the table's size (SURVEY 8d R_gups)",
                                              "updates_per_s": gups, "insert_vs_ceiling": (windows / ins_s) / gups},
                    # one 32 B sector in and one dirty sector out per insert = 64 B/instance is the THEORETICAL
                    # random-access roofline (peak / 64 B inserts/s)
                    "insert_vs_64B_sector_roofline": (windows / ins_s) * 64.0 / (peak * 1e9),
                    "table_bytes": table_bytes, "distinct_keys": distinct}
        tpath = os.path.join(ROOT, "profiles", "insert_keys_traffic.json")
        if partitioned and os.path.exists(tpath):
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from a committed ncu --set full
            # capture; used only when it was taken on this workload with this insert path
            t = json.load(open(tpath))
            if t.get("workload") == args.workload and t.get("scale") == args.scale and t.get("k") == K and t.get("insert_path") == path:
                roofline["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
                roofline["traffic_source"] = t.get("source")
                roofline["traffic_over_algorithmic"] = roofline["traffic"] / algo_bytes
        if R.get("graph"):
            ph, gph, kept = R["phase"], R["graph_phase"], R["kept_total"]
            gtraffic = graph_traffic(args.workload, args.scale, K)
            def entry(kernel, nbytes, ns_, note):
                sec = ns_ * 1e-9
                e = {"kernel": kernel, "bound": "hbm", "algorithmic_bytes": nbytes, "ms": ns_ * 1e-6,
                     "achieved": nbytes / sec / 1e9 if sec > 0 else None, "peak": peak, "unit": "GB/s",
                     "frac": nbytes / sec / 1e9 / peak if sec > 0 else None, "traffic": None, "note": note}
                t = gtraffic.get(kernel.split(" ")[0])
                if t and nbytes > 0:   # ncu: dram__bytes_read.sum + dram__bytes_write.sum of the kernel's launches in one build
                    e["traffic"], e["traffic_launches"], e["traffic_source"] = t
                    e["traffic_over_algorithmic"] = t[0] / nbytes
                return e
            roofline_graph = [
                entry("compact_survivors_kernel (deleteAll sweep)", 12.0 * ph["slots_swept"] + 12.0 * kept, ph["filter_sweep_ns"],
                      "12 B read per slot (key and count arrays; the vertex ids are not touched) + 12 B written per survivor"),
                entry("masks_kernel (Graph.buildGraph membership probes)", 73.0 * kept, gph["graph_masks_ns"], "SURVEY 8(d): 8 probes x 8 B + 8 B own key + 1 B mask per kept k-mer"),
                entry("jump_kernel x %d (list ranking)" % gph["graph_jump_launches"], 16.0 * 2 * kept * gph["graph_jump_launches"], gph["graph_rank_ns"],
                      "SURVEY 8(d): 16 B per oriented vertex and jump round"),
            ]
    else:
        # per-rank insert of the sharded map: whole-insert algorithmic bytes (16 B table + stream per instance) against the
        # slowest rank's event time; the first-touch term needs the distinct count before the filter, which the timed loop does not keep
        ins_s = max_over_ranks(float(np.mean(insert_ns)) * 1e-9)
        algo_bytes = 16.0 * windows + float(b.size)
        single_pass = bool(capi.get_tune("single_pass")) and windows // 2 >= capi.get_tune("single_pass_min")
        if rank == 0:
            roofline = {"bound": "hbm",
                        "kernel": ("sharded insert, single pass: bucket_slabs_kernel<PEER> (NVLink stores into slabs of the owners' inboxes) + insert_slabs_kernel<INBOX>"
                                   if single_pass else
                                   "sharded insert: part_count + part_scatter<PEER> (NVLink stores into the owners' inboxes) + insert_keys_kernel"),
                        "achieved": algo_bytes / ins_s / 1e9, "peak": peak, "unit": "GB/s", "frac": algo_bytes / ins_s / 1e9 / peak,
                        "traffic": None, "peak_source": peak_src, "insert_ms": ins_s * 1e3,
                        "insert_kmers_per_s_per_gpu": windows / ins_s,
                        "insert_vs_64B_sector_roofline": (windows / ins_s) * 64.0 / (peak * 1e9)}
            roofline["nvlink_bytes_out_per_gpu"] = 8.0 * windows * (world - 1) / world
            roofline["nvlink_gbs_out_per_gpu"] = 8.0 * windows * (world - 1) / world / ins_s / 1e9
            roofline["nvlink_frac_of_900_per_direction"] = roofline["nvlink_gbs_out_per_gpu"] / 900.0
    graph = R.get("graph")
    launches, clocks = R["launches"], R["clocks"]
    step_s, wall_step_s, G = R["step_s"], R["wall_step_s"], R["G"]
    release(R)

    # ---------------- N > 1: BASELINE configs[2] as named (the 100 Mbp genome over the N GPUs), a few steps
    named = None
    if world > 1 and not args.no_named and args.workload == "C2" and args.scale == 1.0:
        Rn = run_workload("C3", 1.0 / world, 5, 3, False, not args.no_graph, identities=True)
        named = {"C3": {"workload": "BASELINE configs[2]: 100 Mbp genome with 5% interspersed repeats, 150 bp reads at 50x, k=31, one shard per GPU",
                        "genome_bp": Rn["G"], "kmer_instances_total": Rn["total_windows"], "ms_per_step": Rn["step_s"] * 1e3,
                        "value": Rn["total_windows"] / Rn["step_s"], "unit": "k-mers/s", "steps": 5, "kept_kmers": Rn["kept_total"],
                        "table_bytes_per_gpu": Rn["table_bytes"], "identities": Rn.get("identities"), "graph": Rn.get("graph")}}
        release(Rn)

    if rank == 0:
        line = {
            "metric": "k-mers inserted/s (FreqFilter.extractFilteredKmers: insert + deleteAll)",
            "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_s * 1e3, "wall_ms_per_step": wall_step_s * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, desc), "k": K, "rounds": ROUNDS, "reads_per_gpu": n_reads,
                       "kmer_instances_per_gpu": windows, "genome_bp": G, "scale": args.scale,
                       "sharding": ("one table" if world == 1 else
                                    "hash-prefix shard per GPU, keys stored into the owners' NVLink inboxes by the bucket pass"),
                       "tuning": tuning or "defaults",
                       "l2": "table (%.2f GB) is re-initialised and randomly written every step: far larger than the 126 MB L2" % (table_bytes / 1e9)},
            "parity_checked": bool(parity and parity["ok"]), "parity": parity, "identities": R.get("identities"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "graph": graph,
        }
        if roofline:
            line["roofline"] = roofline
        if roofline_graph:
            line["roofline_graph"] = roofline_graph
        if cpu:
            line["cpu_baseline"] = cpu
        if named:
            line["named_configs"] = named
        print(json.dumps(line))
    if comm:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
