"""Host logic and the C-ABI boundary, without a GPU: the library loads, exports every symbol include/*.h declares,
validates arguments, and fails loudly (no CPU fallback) when asked to compute."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from genome_b200 import capi, synth
from genome_b200.dnamap import PairedEndData, owner_of, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "genome_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), "libgenome_b200.so does not export " + s
    # the ctypes table and the header agree
    assert sorted(capi.SIGNATURES) == syms


def test_header_compiles_as_c():
    src = '#include "genome_b200.h"\nint main(void) { return GB_OK; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", "/dev/null"],
                       input=src, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = capi.lib().gb_map_create(31, 0, 0, 0, C.byref(h))
    assert rc == -4 and not h.value  # GB_E_CUDA
    assert b"CUDA" in capi.lib().gb_last_error()


def test_argument_validation_needs_no_device():
    L = capi.lib()
    h = C.c_void_p()
    assert L.gb_map_create(32, 0, 0, 0, C.byref(h)) == -2      # GB_E_K_RANGE: k = 32 is broken in the reference
    assert L.gb_map_create(0, 0, 0, 0, C.byref(h)) == -2
    assert L.gb_map_create(31, -5, 0, 0, C.byref(h)) == -1
    assert L.gb_map_create(31, 0, 0, 0, None) == -1
    assert L.gb_map_size(None, None) == -1
    assert L.gb_graph_counts(None, None, None, None) == -1
    assert L.gb_version() >= 100


def test_owner_is_a_partition():
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 1 << 62, size=200000, dtype=np.uint64)
    for parts in (1, 2, 3, 4, 8):
        o = owner_of(keys, parts)
        assert o.min() >= 0 and o.max() < parts
        assert np.array_equal(o, owner_of(keys, parts))  # deterministic
        cnt = np.bincount(o, minlength=parts)
        assert cnt.min() > 0.9 * keys.size / parts  # balanced
    assert np.all(owner_of(keys, 1) == 0)


def test_shard_range_and_read_sharding():
    for count in (0, 1, 7, 100, 101):
        for world in (1, 2, 3, 8):
            spans = [shard_range(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    genome = synth.random_genome(3000, 1)
    reads = synth.sample_reads(genome, 50, 40, 0.0, 2, insert=(20, 60))
    fixed = PairedEndData(synth.pack_fixed(reads), 20)
    rag = PairedEndData(synth.pack_ragged([reads[i, :10 + i] for i in range(40)]), 20)
    for data in (fixed, rag):
        for world in (1, 2, 3):
            parts = [data.shard(r, world) for r in range(world)]
            assert sum(p.count for p in parts) == data.count
            assert np.array_equal(np.concatenate([p.bin for p in parts]), data.bin)
            for p in parts:
                p.record_offsets()  # every slice is a well-formed stream of whole pairs
    with pytest.raises(ValueError):
        PairedEndData(fixed.bin[:-5], 20).record_offsets()


def test_multi_rank_routing_gloo():
    """world_size 2 over gloo: the sharded insert's host-visible logic -- every rank extracts the canonical k-mers of
    its slice of the pair stream, routes them by the library's owner function, and the union of the shards equals the
    single-map result (the extraction is done by the oracle here: no GPU in this test)."""
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", script], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "ROUTING OK" in r.stdout


def build_cpp_driver(tmp, name="graph_builder"):
    exe = os.path.join(tmp, name)
    libdir = os.path.join(ROOT, "genome_b200")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-O1", "-o", exe, os.path.join(ROOT, "hostcpp", name + ".cpp"),
                        "-L" + libdir, "-lgenome_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = build_cpp_driver(str(tmp_path))
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    b = synth.pack_fixed(synth.sample_reads(synth.random_genome(2000, 1), 50, 20, 0.0, 2, insert=(20, 60)))
    p = tmp_path / "r.bin"
    p.write_bytes(b.tobytes())
    r = subprocess.run([exe, str(p), "10", "21"], capture_output=True, text=True)
    assert r.returncode == 1 and "CUDA" in r.stderr


def test_cpp_graph_simplifier_driver_compiles_and_fails_loudly_without_gpu(tmp_path):
    """hostcpp/graph_simplifier.cpp: GraphBuilder + GraphSimplifier (pair support, node split) over the C++ mirror."""
    exe = build_cpp_driver(str(tmp_path), "graph_simplifier")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    b = synth.pack_fixed(synth.sample_reads(synth.random_genome(2000, 1), 50, 20, 0.0, 2, insert=(20, 60)))
    p = tmp_path / "r.bin"
    p.write_bytes(b.tobytes())
    r = subprocess.run([exe, str(p), "10", "21", "5"], capture_output=True, text=True)
    assert r.returncode == 1 and "CUDA" in r.stderr


def test_every_entry_point_has_a_binding_line_and_a_ctypes_signature():
    """include/genome_b200.h is the boundary: every function it declares must appear in the Scala/JNA trait of INTEGRATION.md
    (the binding a maintainer of the reference adds) and in the ctypes table the tests and bench.py call through."""
    import re
    from genome_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "genome_b200.h")).read()
    names = re.findall(r"^[a-z_0-9 \*]*\b(gb_[a-z_0-9]+)\(", hdr, flags=re.M)
    assert len(names) >= 55 and len(set(names)) == len(names)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing_doc = [n for n in names if "def %s(" % n not in doc]
    assert not missing_doc, missing_doc
    missing_sig = [n for n in names if n not in capi.SIGNATURES and n not in ("gb_last_error", "gb_launch_count", "gb_version")]
    assert not missing_sig, missing_sig


def test_ctypes_signatures_have_the_headers_arity():
    """A ctypes table that disagrees with the header on the NUMBER of arguments corrupts the call silently: compare them."""
    import re
    from genome_b200 import capi
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "genome_b200.h")).read(), flags=re.S)
    protos = re.findall(r"^\s*(?:const\s+)?[a-z_0-9]+[ \*]+\b(gb_[a-z_0-9]+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.M | re.S)
    assert len(protos) >= 55
    for name, args in protos:
        n = 0 if args.strip() in ("", "void") else len(args.split(","))
        if name in capi.SIGNATURES:
            assert len(capi.SIGNATURES[name][1]) == n, name


def test_bench_reads_graph_kernel_traffic_only_for_its_workload():
    """bench.py's roofline_graph entries take their DRAM traffic from the committed ncu capture, and only on the workload it
    was taken on; a missing or odd file must not fail the bench."""
    import importlib
    old = sys.argv
    sys.argv = ["bench.py"]
    try:
        bench = importlib.import_module("bench")
    finally:
        sys.argv = old
    t = bench.graph_traffic("C2", 1.0, 31)
    assert set(t) == {"masks_kernel", "jump_kernel"} and t["jump_kernel"][1] == 4
    assert 1.0e9 < t["masks_kernel"][0] < 2.0e9
    assert bench.graph_traffic("C2", 0.5, 31) == {} and bench.graph_traffic("C1", 1.0, 31) == {}
    assert bench.graph_traffic("C2", 1.0, 31, root="/nonexistent") == {}
