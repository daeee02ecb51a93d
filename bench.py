#!/usr/bin/env python
"""bench.py -- k-mers inserted/s (FreqFilter.extractFilteredKmers) and graph build + simplify time on N B200s.

A step = one pass of the hot path's insert stage over the whole synthetic read set: a fresh table (clear), bulk
FreqFilter.add over every read (canonical k-mer extraction + hash-table upsert), deleteAll(v < 3).  `value` is
k-mer instances inserted per second with the `.bin` stream already resident in HBM; `e2e` is the same through the
host-buffer C-ABI calls (H2D of the stream and D2H of the size inside the timed region).  Graph.buildGraph,
components, retain and simplifyGraph are timed once per run on the table the last step left behind and reported
under "graph" (they are part of BASELINE.json's metric, not of the k-mers/s figure).

N = 1: BASELINE.json configs[1] (4.6 Mbp genome, 1% substitutions, 100 bp reads at 30x, k = 31).
N > 1 (torchrun): weak scaling -- every rank holds 1.38 M reads of an N x 4.6 Mbp genome, the table is one
hash shard per GPU (PartitionedDNAMap), k-mers are routed by the library's NCCL all-to-all.

`--impl reference`: the CPU restatement of the reference algorithm (oracle/, "port": the Scala/Akka reference cannot
be built here) on all host cores, same workload.
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


def cpu_reference(b, n_reads, windows, threads, steps, warmup, sample_reads):
    """The oracle's faithful-cost mode: ArrayDNAMap layout/probing/rescale, P single-threaded partitions fed by
    extractor threads through in-memory buckets (no Kryo/TCP: strictly faster than the real reference)."""
    from oracle import pyoracle
    pyoracle.build()
    n = min(n_reads, sample_reads)
    rec = b.size // n_reads
    sb = b[:n * rec]
    w = pyoracle.count_windows(sb, n, K)
    times = []
    for i in range(warmup + steps):
        m = pyoracle.OracleMap(K, partitions=threads)
        t0 = time.perf_counter()
        m.insert_reads(sb, n, threads=threads)
        m.delete_below(ROUNDS)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        del m
    t = float(np.mean(times))
    # the graph stage of the same sample, serial like Graph.buildGraph's driver loop (Graph.scala:367-374 is chunk.par, but
    # every probe is a blocking remote ask there): build, components + retain + simplify
    m = pyoracle.OracleMap(K, partitions=1)
    m.insert_reads(sb, n, threads=0)
    m.delete_below(ROUNDS)
    t0 = time.perf_counter()
    g = pyoracle.OracleGraph(m)
    t1 = time.perf_counter()
    g.components()
    g.retain_largest()
    g.simplify()
    t2 = time.perf_counter()
    cpu_reference.graph = {"kept_kmers": m.size(), "build_ms": (t1 - t0) * 1e3, "components_retain_simplify_ms": (t2 - t1) * 1e3, "cores": 1}
    return w / t, t, n, w


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale (debugging only; 1.0 is the named config)")
    ap.add_argument("--k", type=int, default=31, help="BASELINE configs[4] sweeps only; 31 is every named config's k")
    ap.add_argument("--coverage", type=float, default=None, help="BASELINE configs[4] sweeps only; default = the named config's")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
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
        sample = min(n_reads, 400_000)
        v, t, n, w = cpu_reference(b, n_reads, windows, cores, args.steps, max(args.warmup, 1), sample)
        line = {
            "impl": "reference", "metric": "k-mers inserted/s (FreqFilter.extractFilteredKmers: insert + deleteAll)",
            "value": v, "unit": "k-mers/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, desc), "k": K, "rounds": ROUNDS},
            "cpu_baseline": {"value": v, "unit": "k-mers/s", "cores": cores, "kind": "port",
                             "sample": "first %d of %d reads (%d k-mer instances) per step" % (n, n_reads, w),
                             "graph": getattr(cpu_reference, "graph", None)},
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

    b, n_reads, windows, G, distinct_total = make_workload(args.workload, rank, world, args.scale)
    cap = int(distinct_total / world * 1.15) + 1024  # distinct keys expected on this shard
    L = capi.lib()

    # pinned host copy (e2e leg) and resident device copy (kernel leg) of this rank's `.bin` stream
    hp = C.c_void_p()
    capi.check(L.gb_host_alloc(b.size, C.byref(hp)))
    pinned = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(b.size,))
    pinned[:] = b
    d_bin = torch.empty(b.size + 16, dtype=torch.uint8, device="cuda")
    d_bin[:b.size].copy_(torch.from_numpy(b))
    torch.cuda.synchronize()

    if world > 1:
        comm = Communicator(rank, world, local_rank, torch_broadcast)
        m = PartitionedDNAMap(K, comm, cap)
    else:
        comm = None
        m = ArrayDNAMap(K, cap, device=local_rank)

    insert_stats = {}

    def step_device():
        m.clear(cap)
        w = m.insert_reads_device(d_bin.data_ptr(), b.size, n_reads)
        insert_stats.update(m.stats())  # table size and insert-kernel time before the filter shrinks the table
        m.delete_below(ROUNDS)
        return w

    def step_host():
        m.clear(cap)
        w = m.insert_reads(pinned, n_reads)
        m.delete_below(ROUNDS)
        return w, m.size  # the size read is the step's device->host result

    # ---------------- kernel leg: inputs resident in HBM
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        assert step_device() == windows
    barrier()
    launches0 = L.gb_launch_count()
    insert_ns = []
    table_bytes = 0
    barrier()
    t0 = time.perf_counter()
    m.timer_start()
    for _ in range(args.steps):
        step_device()
        insert_ns.append(insert_stats["last_insert_ns"])
    dev_ns = m.timer_stop()
    barrier()
    t1 = time.perf_counter()
    wall = t1 - t0
    launches = L.gb_launch_count() - launches0
    clocks = sampler.stop(t0, t1)
    step_s = max_over_ranks(dev_ns * 1e-9 / args.steps)
    wall_step_s = max_over_ranks(wall / args.steps)
    total_windows = sum_over_ranks(float(windows))
    value = total_windows / step_s
    kept_total = m.size
    table_bytes = insert_stats["table_bytes"]

    # ---------------- e2e leg: host buffers through the reference-facing calls
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    e2e = {"value": total_windows / e2e_s, "unit": "k-mers/s", "h2d_bytes_per_step": int(b.size), "d2h_bytes_per_step": 8 + 4 * 32,
           "ms_per_step": e2e_s * 1e3}

    # ---------------- graph stage on the filtered table (timed once; collective for N > 1)
    graph = None
    if not args.no_graph:
        for rep in range(3):  # two passes settle the scratch arenas (growth, then consolidation); the third is timed
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
            if rep < 2:
                g.close()
        graph = {"build_ms": max_over_ranks((t1 - t0) * 1e3), "build_kernels_ms": g.stats()["build_ns"] * 1e-6,
                 "components_retain_simplify_ms": max_over_ranks((t2 - t1) * 1e3),
                 "nodes": nn, "edges": ne, "edge_bases": nb, "components": nc, "jump_launches": g.stats()["jump_launches"],
                 "after_simplify": g.counts(), "kept_kmers": kept_total,
                 "sharding": ("single GPU" if world == 1 else
                              "sharded: minimizer owners, rank-local list ranking, segment list (GENOME_B200_PGRAPH=sharded)"
                              if os.environ.get("GENOME_B200_PGRAPH") == "sharded" else "replicated after all-gather of the shards")}
        g.close()

    # ---------------- roofline of the dominant kernel (insert), live CUDA-event durations from inside the library
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    roofline = None
    gups = None
    if world == 1:
        step_device()  # size before the filter = distinct keys
        m.clear(cap)
        m.insert_reads_device(d_bin.data_ptr(), b.size, n_reads)
        distinct = m.size
        st = m.stats()
        ins_s = float(np.mean(insert_ns)) * 1e-9
        # 16 B table read/write per instance + input stream + 8 B first-touch key write per distinct key
        algo_bytes = 16.0 * windows + float(b.size) + 8.0 * distinct
        ns = C.c_int64()
        capi.check(L.gb_bench_random_atomics(local_rank, table_bytes, windows, 5, C.byref(ns)))
        gups = windows / (ns.value * 1e-9)
        partitioned = st["upsert_ns"] > 0
        if partitioned:
            # the insert is three kernels; the dominant one is the slice-ordered upsert (insert_keys_kernel)
            up_s, bk_s = st["upsert_ns"] * 1e-9, st["bucket_ns"] * 1e-9
            dom, dom_s = "insert_keys_kernel", up_s
            # algorithmic bytes of that launch: 8 B key read + 16 B table read/write per instance + first-touch writes
            dom_bytes = 24.0 * windows + 8.0 * distinct
        else:
            dom, dom_s, dom_bytes = "insert_reads_kernel", ins_s, algo_bytes
        achieved = dom_bytes / dom_s / 1e9
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_kmer": dom_bytes / windows, "kernel_ms": dom_s * 1e3,
                    "insert_path": ("partitioned (L2-blocked), single pass: part_scatter<SLABS> + insert_keys (GENOME_B200_COUNTLESS)"
                                    if partitioned and os.environ.get("GENOME_B200_COUNTLESS") else
                                    "partitioned (L2-blocked): part_count + part_scatter + insert_keys" if partitioned else "direct: insert_reads_kernel"),
                    "insert_ms": ins_s * 1e3, "insert_kmers_per_s": windows / ins_s,
                    "insert_algorithmic_bytes_per_kmer": algo_bytes / windows,
                    "insert_frac": algo_bytes / ins_s / 1e9 / peak,
                    "phases_ms": {"bucket (count+offsets+scatter)": st["bucket_ns"] * 1e-6, "upsert": st["upsert_ns"] * 1e-6} if partitioned else None,
                    "random_access_ceiling_kmers_per_s": gups, "insert_vs_random_access_ceiling": (windows / ins_s) / gups,
                    # one 32 B sector in and one dirty sector out per insert = 64 B/instance is the THEORETICAL
                    # random-access roofline (peak / 64 B inserts/s); the L2-blocked path may exceed it, which is its point
                    "insert_vs_64B_sector_roofline": (windows / ins_s) * 64.0 / (peak * 1e9),
                    "table_bytes": table_bytes, "distinct_keys": distinct}
        if partitioned and args.workload == "C2" and args.scale == 1.0:
            # dram__bytes_read.sum + dram__bytes_write.sum of one insert_keys_kernel launch on this workload
            roofline["traffic"] = 2.569e9 + 0.803e9
            roofline["traffic_source"] = "profiles/insert_r1e_ncu_full_summary.csv (ncu --set full, same workload)"
    else:
        # per-rank insert of the sharded map (part_count + part_scatter<PEER> + insert_keys overlapped): whole-insert
        # algorithmic bytes (16 B table + stream per instance) against the slowest rank's event time; the first-touch
        # term needs the distinct count before the filter, which the timed loop does not keep
        ins_s = max_over_ranks(float(np.mean(insert_ns)) * 1e-9)
        algo_bytes = 16.0 * windows + float(b.size)
        superkmer = os.environ.get("GENOME_B200_WIRE") == "superkmer"
        if rank == 0 and superkmer:
            # 16-byte records of ~10 windows each: the exact wire volume is not kept by the timed loop; 1.6 B per window is the
            # CPU-measured figure for k = 31, P = 8 (tests/test_superkmer_emul_cpu.py), stated as an estimate
            roofline = {"bound": "hbm", "kernel": "sharded insert, super-k-mer wire: split (count + emit) + NCCL exchange + local partitioned insert",
                        "achieved": algo_bytes / ins_s / 1e9, "peak": peak, "unit": "GB/s", "frac": algo_bytes / ins_s / 1e9 / peak,
                        "traffic": None, "peak_source": peak_src, "insert_ms": ins_s * 1e3,
                        "insert_kmers_per_s_per_gpu": windows / ins_s,
                        "nvlink_bytes_out_per_gpu_estimate": 1.6 * windows * (world - 1) / world}
        elif rank == 0:
            roofline = {"bound": "hbm", "kernel": "sharded insert: part_count + part_scatter<PEER> (NVLink stores) + insert_keys_kernel",
                        "achieved": algo_bytes / ins_s / 1e9, "peak": peak, "unit": "GB/s", "frac": algo_bytes / ins_s / 1e9 / peak,
                        "traffic": None, "peak_source": peak_src, "insert_ms": ins_s * 1e3,
                        "insert_kmers_per_s_per_gpu": windows / ins_s,
                        "nvlink_bytes_out_per_gpu": 8.0 * windows * (world - 1) / world,
                        "nvlink_gbs_out_per_gpu": 8.0 * windows * (world - 1) / world / ins_s / 1e9}

    # ---------------- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, t, n, w = cpu_reference(b, n_reads, windows, cores, 2, 1, 400_000)
        cpu = {"value": v, "unit": "k-mers/s", "cores": cores, "kind": "port",
               "sample": "first %d of %d reads (%d k-mer instances), %d partitions/threads, mean of 2 passes" % (n, n_reads, w, cores),
               "graph": getattr(cpu_reference, "graph", None)}

    if rank == 0:
        line = {
            "metric": "k-mers inserted/s (FreqFilter.extractFilteredKmers: insert + deleteAll)",
            "value": value, "unit": "k-mers/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_s * 1e3, "wall_ms_per_step": wall_step_s * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, desc), "k": K, "rounds": ROUNDS, "reads_per_gpu": n_reads,
                       "kmer_instances_per_gpu": windows, "genome_bp": G, "scale": args.scale,
                       "sharding": ("one table" if world == 1 else
                                    "minimizer-owner shard per GPU, 16-byte super-k-mer records over NCCL (GENOME_B200_WIRE=superkmer)"
                                    if os.environ.get("GENOME_B200_WIRE") == "superkmer" else "hash-prefix shard per GPU, NCCL all-to-all"),
                       "l2": "table (%.2f GB) is re-initialised and randomly written every step: far larger than the 126 MB L2" % (table_bytes / 1e9)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "graph": graph,
        }
        if roofline:
            line["roofline"] = roofline
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    m.close()
    if comm:
        comm.close()
    capi.check(L.gb_host_free(hp))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
