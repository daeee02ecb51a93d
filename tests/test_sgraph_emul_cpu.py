"""The sharded Graph.buildGraph (genome_b200/csrc/sgraph.cuh: minimizer re-routing, per-rank index / masks / list ranking,
segment list, global assembly) compiled with g++ and run over P in-process ranks (tests/emul/sgraph_emul.cpp) against the
oracle's restatement of Graph.scala:269-382.  The per-item functors AND the orchestration are the code the CUDA library
runs; only the launches (serial loops here), the memory and the fabric differ.  The device runs are in
tests/test_sgraph_gpu.py (virtual shards on one GPU) and tests/test_parity_multigpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from genome_b200 import synth
from oracle import pyoracle
from tests import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def load_emul():
    """Builds (when stale) and loads tests/emul/sgraph_emul.cpp; also used by device tests that need the CPU side of a format."""
    src = os.path.join(HERE, "emul", "sgraph_emul.cpp")
    out = os.path.join(HERE, "_build", "libsgraph_emul.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src, os.path.join(HERE, "emul", "superkmer.cuh")] + [os.path.join(ROOT, "genome_b200", "csrc", f) for f in ("sgraph.cuh", "sgraph_fabric.cuh", "common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
                               "-pthread", "-Wl,-Bsymbolic", "-I" + cuda_inc, "-o", out, src])   # -Bsymbolic: the library binds its own
        # gb::sg::* symbols; the product library (loaded RTLD_GLOBAL by genome_b200.capi) exports the CUDA versions of them
    return C.CDLL(out)


@pytest.fixture(scope="module")
def emul():
    return load_emul()


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def sharded_build(emul, k, keys, P, dual=False, split="even", seed=0, threads=False):
    """Runs the emulated build with the kept keys dealt to P ranks (`split`: even / skewed / all on the last rank) and
    returns (canonical graph as in tests/helpers.py, stats dict)."""
    keys = np.ascontiguousarray(keys, np.uint64)
    rng = np.random.default_rng(seed)
    keys = keys[rng.permutation(keys.size)]
    if split == "even":
        cuts = [keys.size * r // P for r in range(P + 1)]
    elif split == "skewed":
        cuts = sorted(rng.integers(0, keys.size + 1, P - 1).tolist())
        cuts = [0] + cuts + [keys.size]
    else:
        cuts = [0] * P + [keys.size]
    off = np.array(cuts, np.uint64)
    out = np.zeros(8, np.uint64)
    fn = emul.emul_sharded_build_threads if threads else emul.emul_sharded_build
    rc = fn(k, int(dual), 0, P, ptr(keys), ptr(off), ptr(out), None, None, None, None, None)
    assert rc == 0, rc
    N, E, B = int(out[0]), int(out[1]), int(out[2])
    node_kmer = np.zeros(max(N, 1), np.uint64)
    es = np.zeros(max(E, 1), np.uint32)
    ee = np.zeros(max(E, 1), np.uint32)
    eo = np.zeros(E + 1, np.uint64)
    words = np.zeros((B + 15) // 16 + 1, np.uint32)
    rc = fn(k, int(dual), 0, P, ptr(keys), ptr(off), ptr(out), ptr(node_kmer), ptr(es), ptr(ee), ptr(eo), ptr(words))
    assert rc == 0
    bases = np.zeros(words.size * 16, np.uint8)
    for j in range(16):
        bases[j::16] = (words >> np.uint32(2 * j)) & 3
    nk = [int(x) for x in node_kmer[:N]]
    assert int(eo[E]) == B and np.all(np.diff(eo.astype(np.int64)) > 0)
    edges = sorted((nk[int(es[i])], nk[int(ee[i])], bases[int(eo[i]):int(eo[i + 1])].tobytes()) for i in range(E))
    stats = dict(kept=int(out[3]), segments=int(out[4]), cycle_vertices=int(out[5]), jump_rounds=int(out[6]), seg_rounds=int(out[7]))
    return (sorted(nk), edges), (N, E, B), stats


def oracle_graph(b, n, k, rounds):
    om, _ = H.oracle_counts(b, n, k)
    om.delete_below(rounds)
    return om, pyoracle.OracleGraph(om)


@pytest.mark.parametrize("k", [3, 4, 8, 15, 16, 21, 31])
@pytest.mark.parametrize("P", [1, 2, 5, 8])
def test_neighbour_owner_is_the_owner_of_the_neighbour(emul, k, P):
    """The incremental owner of the 8 neighbours (one m-mer hash each, from the parts of x) equals the owner computed from
    scratch, and a k-mer shares its owner with its reverse complement."""
    rng = np.random.default_rng(100 * k + P)
    full, incr, full_rc = np.zeros(9, np.uint32), np.zeros(8, np.uint32), np.zeros(9, np.uint32)
    xs = rng.integers(0, 1 << (2 * k), 300, dtype=np.uint64).tolist()
    xs += [0, (1 << (2 * k)) - 1, int("01" * k, 2)]   # poly-A, poly-T, poly-G: every m-mer equal
    seen = set()
    for x in xs:
        emul.emul_owners(k, P, C.c_uint64(x), ptr(full), ptr(incr))
        assert np.array_equal(full[1:], incr), (k, P, x)
        assert full.max() < P
        emul.emul_owners(k, P, C.c_uint64(pyoracle.revcomp(x, k)), ptr(full_rc), ptr(incr))
        assert full_rc[0] == full[0]
        seen.add(int(full[0]))
    if k >= 15:
        assert len(seen) == P   # every rank owns something


GRAPH_CASES = [
    # k, genome, read_len, coverage, err, rounds   (tests/test_parity_gpu.py GRAPH_CASES)
    (31, 20000, 100, 30, 0.0, 3),
    (31, 20000, 100, 30, 0.01, 3),
    (21, 30000, 100, 25, 0.02, 2),
    (15, 5000, 60, 30, 0.01, 2),
    (11, 3000, 50, 20, 0.0, 1),
    (9, 4000, 40, 15, 0.03, 1),
    (8, 1500, 40, 10, 0.0, 1),   # even k: palindromes
    (6, 600, 30, 10, 0.02, 1),
    (4, 120, 20, 6, 0.0, 1),
    (5, 300, 20, 4, 0.0, 1),
    (3, 40, 12, 3, 0.0, 1),
]


@pytest.mark.parametrize("k,glen,rl,cov,err,rounds", GRAPH_CASES)
def test_sharded_build_matches_oracle(emul, k, glen, rl, cov, err, rounds):
    b, n, _ = H.small_reads(glen, rl, cov, err, seed=2000 + k)
    om, og = oracle_graph(b, n, k, rounds)
    want = H.canon_oracle_graph(og)
    keys, _ = om.export()
    for P, split in [(1, "even"), (2, "even"), (3, "skewed"), (8, "even"), (8, "last"), (16, "skewed")]:
        got, counts, st = sharded_build(emul, k, keys, P, split=split, seed=P)
        assert counts == og.counts(), (P, split)
        assert got == want, (P, split)
        assert st["kept"] == keys.size
        if P == 1:
            assert st["segments"] == 0


def test_segments_are_a_small_fraction(emul):
    """The point of minimizer ownership: chains break into rank-local runs ~(k - m + 2) / 2 vertices long, so the segment
    list that crosses the fabric is an order of magnitude shorter than the vertex list (hash ownership: 7 of 8 links)."""
    k = 31
    b, n, _ = H.small_reads(60000, 100, 30, 0.0, seed=77)
    om, og = oracle_graph(b, n, k, 2)
    keys, _ = om.export()
    got, counts, st = sharded_build(emul, k, keys, 8)
    assert got == H.canon_oracle_graph(og)
    interior = 2 * keys.size
    assert 0 < st["segments"] < interior / 6, st
    assert st["seg_rounds"] >= 1


def test_noncanonical_keys_both_orientations(emul):
    """Keys stored as given (update without canonicalisation): both orientations of a k-mer may be stored; the smaller one
    is the vertex, the other one is nothing (common.cuh find_oriented / is_secondary)."""
    k = 9
    rng = np.random.default_rng(3)
    genome = synth.random_genome(2000, 99)
    fw = np.array([synth.kmer_to_int(synth.decode(genome[i:i + k])) for i in range(genome.size - k + 1)], np.uint64)
    rc = np.array([pyoracle.revcomp(int(x), k) for x in fw], np.uint64)
    pick = rng.random(fw.size)
    keys = np.unique(np.concatenate([fw[pick < 0.6], rc[pick > 0.4]]))
    om = pyoracle.OracleMap(k)
    for x in keys:
        om.update1(int(x))
    og = pyoracle.OracleGraph(om)
    stored, _ = om.export()
    for P in (1, 4, 8):
        got, counts, _ = sharded_build(emul, k, stored, P, dual=True)
        assert counts == og.counts()
        assert got == H.canon_oracle_graph(og)


@pytest.mark.parametrize("P", [1, 2, 8])
def test_perfect_cycle_is_dropped(emul, P):
    """A circular sequence with no branch has no terminal k-mer (Graph.scala:375): with P > 1 the cycle runs through several
    ranks, so it is the SEGMENT list that fails to resolve."""
    k = 11
    genome = synth.random_genome(500, 8)
    circ = np.concatenate([genome, genome[:k - 1]])
    keys = np.array([pyoracle.canonical(synth.kmer_to_int(synth.decode(circ[i:i + k])), k) for i in range(genome.size)], np.uint64)
    got, counts, st = sharded_build(emul, k, keys, P)
    assert counts == (0, 0, 0)
    assert st["cycle_vertices"] == 2 * genome.size
    if P > 1:
        assert st["segments"] > 0


def test_cycle_next_to_a_real_graph(emul):
    """A perfect cycle and a branching component side by side: the cycle vanishes, the rest is exact."""
    k = 15
    g1 = synth.random_genome(400, 22)
    circ = np.concatenate([g1, g1[:k - 1]])
    cyc = [pyoracle.canonical(synth.kmer_to_int(synth.decode(circ[i:i + k])), k) for i in range(g1.size)]
    b, n, _ = H.small_reads(3000, 50, 20, 0.01, seed=5)
    om, _ = H.oracle_counts(b, n, k)
    om.delete_below(2)
    for x in cyc:
        om.update1(int(x))
    og = pyoracle.OracleGraph(om)
    keys, _ = om.export()
    for P in (1, 3, 8):
        got, counts, st = sharded_build(emul, k, keys, P)
        assert got == H.canon_oracle_graph(og)
        assert st["cycle_vertices"] == 2 * g1.size


def test_empty_and_tiny_inputs(emul):
    for P in (1, 4):
        got, counts, _ = sharded_build(emul, 15, np.zeros(0, np.uint64), P)
        assert counts == (0, 0, 0)
        # one isolated k-mer: (0, 0) vertices are no nodes (Graph.scala:323)
        got, counts, _ = sharded_build(emul, 15, np.array([pyoracle.canonical(12345, 15)], np.uint64), P)
        assert counts == (0, 0, 0)
        # two overlapping k-mers: 4 nodes, 2 edges of length 1
        x = synth.kmer_to_int("ACGTTGCAAGGCTTA")
        y = synth.kmer_to_int("CGTTGCAAGGCTTAC")
        om = pyoracle.OracleMap(15)
        for v in (x, y):
            om.update1(pyoracle.canonical(v, 15))
        og = pyoracle.OracleGraph(om)
        keys, _ = om.export()
        got, counts, _ = sharded_build(emul, 15, keys, P)
        assert counts == og.counts() == (4, 2, 2)
        assert got == H.canon_oracle_graph(og)


@pytest.mark.parametrize("k,glen,rl,cov,err,rounds", [c for c in GRAPH_CASES if c[0] in (31, 21, 9, 8, 4)])
def test_one_thread_per_rank_matches_oracle(emul, k, glen, rl, cov, err, rounds):
    """The one-process-per-GPU control flow: P threads, each driving ONE rank through its own Fabric object, its own copy of
    the global graph arrays, real sums and exchanges between the copies (ThreadFabric in tests/emul/sgraph_emul.cpp).  Every
    rank must end with the same graph, equal to the oracle's; vertex numbering is deterministic here, so the emulated ranks'
    copies are compared byte for byte."""
    b, n, _ = H.small_reads(glen, rl, cov, err, seed=2000 + k)
    om, og = oracle_graph(b, n, k, rounds)
    want = H.canon_oracle_graph(og)
    keys, _ = om.export()
    for P, split in [(2, "even"), (4, "skewed"), (8, "even"), (8, "last")]:
        got, counts, st = sharded_build(emul, k, keys, P, split=split, seed=P, threads=True)
        assert counts == og.counts(), (P, split)
        assert got == want, (P, split)


def test_one_thread_per_rank_cycle_and_dual(emul):
    k = 11
    genome = synth.random_genome(500, 8)
    circ = np.concatenate([genome, genome[:k - 1]])
    keys = np.array([pyoracle.canonical(synth.kmer_to_int(synth.decode(circ[i:i + k])), k) for i in range(genome.size)], np.uint64)
    got, counts, st = sharded_build(emul, k, keys, 4, threads=True)
    assert counts == (0, 0, 0) and st["cycle_vertices"] == 2 * genome.size
    got, counts, st = sharded_build(emul, 15, np.zeros(0, np.uint64), 3, threads=True)
    assert counts == (0, 0, 0)


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_thread_per_rank_under_sanitizers(tmp_path, sanitizer):
    """One thread per rank (tests/emul/sgraph_tsan_main.cpp) under -fsanitize=thread: a read of a peer's window that no Fabric
    barrier orders after the peer's write is a reported data race -- this checks the BARRIER PLACEMENT of sg::build.  The first
    run found one (CloseOp reading an interior successor's entry while its owner's FinalizeOp rewrites it -- benign for
    whole-word accesses, removed anyway by deciding "terminal or not" on the immutable mask byte).  Under
    -fsanitize=address,undefined the same run checks every index the functors compute (the harness allocates exact sizes):
    an overrun here would be an overrun in the kernels."""
    exe = str(tmp_path / "sg_san")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=" + sanitizer, "-fno-sanitize-recover=undefined", "-pthread",
                        "-I" + cuda_inc, "-o", exe, os.path.join(HERE, "emul", "sgraph_tsan_main.cpp")], capture_output=True, text=True)
    if r.returncode != 0 and "san" in (r.stderr or "").lower() and "cannot find" in r.stderr:
        pytest.skip("sanitizer runtime not installed: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    for k, glen, rl, cov, err, rounds, P in [(31, 20000, 100, 30, 0.01, 3, 8), (8, 1500, 40, 10, 0.0, 1, 3), (11, 3000, 50, 20, 0.0, 1, 16)]:
        b, n, _ = H.small_reads(glen, rl, cov, err, seed=2000 + k)
        om, og = oracle_graph(b, n, k, rounds)
        keys, _ = om.export()
        path = str(tmp_path / ("keys_%d.bin" % k))
        np.ascontiguousarray(keys, np.uint64).tofile(path)
        r = subprocess.run([exe, str(k), str(P), path], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66", ASAN_OPTIONS="detect_leaks=0"))
        if "unexpected memory mapping" in r.stderr:
            pytest.skip("the sanitizer cannot run under this kernel's address-space layout")
        assert "WARNING: ThreadSanitizer" not in r.stderr and "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[:3000]
        assert r.returncode == 0, (r.returncode, r.stderr[-500:])
        assert r.stdout.strip() == "rc 0 nodes %d edges %d bases %d" % og.counts()


@pytest.mark.parametrize("k", [2, 3, 4, 5, 6, 7])
def test_random_dense_kmer_sets(emul, k):
    """Random subsets of the whole k-mer space for small k: dense tangles with self-loops, hairpins (x followed by rc(x)),
    palindromes (even k) and perfect cycles that read sets rarely produce; stored canonically (FreqFilter.add) or as given."""
    rng = np.random.default_rng(700 + k)
    space = 1 << (2 * k)
    for trial in range(12):
        frac = [0.05, 0.2, 0.5, 0.9][trial % 4]
        xs = np.flatnonzero(rng.random(space) < frac).astype(np.uint64)
        if xs.size == 0:
            continue
        dual = trial % 3 == 2
        om = pyoracle.OracleMap(k)
        for x in xs.tolist():
            om.update1(x if dual else pyoracle.canonical(x, k))
        og = pyoracle.OracleGraph(om)
        keys, _ = om.export()
        want = H.canon_oracle_graph(og)
        P = int(rng.integers(1, 9))
        for threads in (False, True):
            got, counts, _ = sharded_build(emul, k, keys, P, dual=dual, split=["even", "skewed", "last"][trial % 3], seed=trial, threads=threads)
            assert counts == og.counts(), (k, trial, P, dual, threads)
            assert got == want, (k, trial, P, dual, threads)


@pytest.mark.parametrize("world", [2, 3])
def test_one_process_per_rank_gloo(emul, world):
    """world_size 2 and 3 over gloo: one PROCESS per rank, the Fabric's collectives done by torch.distributed, the peer windows
    in /dev/shm and mapped at different addresses in every process (tests/gloo_sgraph_worker.py) -- the CPU stand-in for NCCL +
    CUDA IPC.  Every rank checks its copy of the graph against the oracle."""
    import sys
    script = os.path.join(HERE, "gloo_sgraph_worker.py")
    port = str(29750 + world)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", port, script], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-5000:]
    assert "SGRAPH GLOO OK world %d" % world in r.stdout


def test_golden_fixture(emul):
    """The committed fixture tests/golden/small_k15.json (keys after the filter -> nodes and edges): the sharded build must
    reproduce the committed graph, for every rank count."""
    import json
    fx = json.load(open(os.path.join(HERE, "golden", "small_k15.json")))
    k, rounds = fx["k"], fx["rounds"]
    keys = np.array([x for x, c in zip(fx["keys"], fx["counts"]) if c >= rounds], np.uint64)
    want = (sorted(fx["nodes"]), sorted((u, v, bytes.fromhex(s)) for u, v, s in fx["edges"]))
    for P in (1, 2, 8):
        for threads in (False, True):
            got, counts, _ = sharded_build(emul, k, keys, P, threads=threads)
            assert got == want


@pytest.mark.parametrize("strands", ["both", "forward"])
def test_sharded_build_with_hash_tie_kmers(emul, strands):
    """Even-k k-mers whose two orientations hash alike (tests/golden/hash_ties.json): reads of both strands store BOTH orientations
    (FreqFilter.scala:32, tie => rcx) without the map being `dual`, so the probes cannot pick a canonical orientation, the
    numerically smaller stored orientation is the vertex and the other one is secondary; reads of one strand store only the
    orientation the reads do NOT spell.  The sharded build must give the oracle's graph either way."""
    for k, kmers in H.hash_ties():
        b, n = H.tie_reads(k, kmers, seed=k, strands=strands)
        om, og = oracle_graph(b, n, k, 3)
        keys, _ = om.export()
        stored = set(int(x) for x in keys)
        for x in kmers:
            assert pyoracle.revcomp(x, k) in stored and (x in stored) == (strands == "both")
        want = H.canon_oracle_graph(og)
        assert og.counts()[1] > 0
        for P, split in [(1, "even"), (2, "even"), (3, "skewed"), (8, "even")]:
            got, counts, st = sharded_build(emul, k, keys, P, split=split, seed=P)
            assert counts == og.counts(), (k, P, split)
            assert got == want, (k, P, split)
