"""The device code of the paired-end path support (genome_b200/csrc/walk.cuh: the functions the kernels of walk.cu run one
item per thread) compiled with g++ and run serially (tests/emul/walk_emul.cpp) against the oracle's restatement of
GraphSimplifier.scala:33-127,188-317.  This checks the walk / split LOGIC where no GPU exists; the kernels themselves (launch
geometry, atomics, tiered retry) are covered by tests/test_walk_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from genome_b200 import synth
from oracle import pyoracle

from .test_walk_cpu import kmers_of, rand_seq, reads_of, two_chromosomes

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NONE32 = 0xFFFFFFFF


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "emul", "walk_emul.cpp")
    sanitize = bool(os.environ.get("GB_EMUL_SANITIZE"))   # set by test_walk_code_under_address_sanitizer's child process
    out = os.path.join(HERE, "_build", "libwalk_emul_asan.so" if sanitize else "libwalk_emul.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src] + [os.path.join(ROOT, "genome_b200", "csrc", f) for f in ("walk.cuh", "common.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function"] +
                              (["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"] if sanitize else []) +
                              ["-Wl,-Bsymbolic", "-I" + cuda_inc, "-o", out, src])   # own symbols first: the product library
        # (loaded RTLD_GLOBAL by genome_b200.capi) exports functions of the same names
    return C.CDLL(out)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class DeviceLayout:
    """An oracle graph in the array layout of the device-resident graph (graph_types.cuh): indices instead of ids."""

    def __init__(self, og):
        node_kmer, node_id, es, ee, off, bases = og.export()
        self.k = og.k
        self.node_idx = {int(i): n for n, i in enumerate(node_id)}
        self.edge_ids = og.edge_ids()
        self.edge_idx = {int(i): n for n, i in enumerate(self.edge_ids)}
        self.node_kmer = np.ascontiguousarray(node_kmer, np.uint64)
        self.edge_start = np.array([self.node_idx[int(x)] for x in es], np.uint32)
        self.edge_end = np.array([self.node_idx[int(x)] for x in ee], np.uint32)
        self.edge_off = np.ascontiguousarray(off, np.uint64)
        nb = bases.size
        padded = np.zeros(((nb + 15) // 16 + 2) * 16, np.uint32)
        padded[:nb] = bases
        q = padded.reshape(-1, 16)
        self.bases_codes = bases
        self.bases = np.zeros(q.shape[0], np.uint32)
        for j in range(16):
            self.bases |= q[:, j] << np.uint32(2 * j)
        self.N, self.E = self.node_kmer.size, self.edge_start.size
        self.out4 = np.full((max(self.N, 1), 4), NONE32, np.uint32)
        for e in range(self.E):
            self.out4[self.edge_start[e], bases[int(off[e])]] = e

    def positions(self, og):
        kmer, ident, dist = og.graph_map()
        idx = np.array([self.node_idx[int(i)] if d == 0 else self.edge_idx[int(i)] for i, d in zip(ident, dist)], np.uint32)
        return np.ascontiguousarray(kmer, np.uint64), idx, dist.astype(np.uint32)

    def graph_args(self):
        return [self.k, C.c_uint64(self.N), C.c_uint64(self.E), ptr(self.node_kmer), ptr(self.edge_start), ptr(self.edge_end),
                ptr(self.edge_off), ptr(self.bases)]

    def support_from_oracle(self, e1, e2, cnt):
        """pathsMap triples (oracle ids) -> the dense support[4 * e1 + b] array of the C ABI"""
        s = np.zeros(4 * self.E, np.uint32)
        for a, b, c in zip(e1, e2, cnt):
            ia, ib = self.edge_idx[int(a)], self.edge_idx[int(b)]
            assert self.edge_end[ia] == self.edge_start[ib]
            s[4 * ia + int(self.bases_codes[int(self.edge_off[ib])])] = c
        return s


def record_offsets(b, n_reads):
    off = np.empty(n_reads + 1, np.uint64)
    pos = 0
    for r in range(n_reads):
        off[r] = pos
        pos += 1 + (int(b[pos]) + 3) // 4
    off[n_reads] = pos
    return off


def emul_support(emul, lay, og, b, n_pairs, lo, hi, lmax=4096):
    pk, pi, pd = lay.positions(og)
    b = np.ascontiguousarray(b, np.uint8)
    off = record_offsets(b, 2 * n_pairs)
    support = np.zeros(4 * max(lay.E, 1), np.uint32)
    bad, walked, over = C.c_uint64(), C.c_uint64(), C.c_uint64()
    rc = emul.emul_pair_support(*lay.graph_args(), C.c_uint64(pk.size), ptr(pk), ptr(pi), ptr(pd), ptr(b), ptr(off),
                                C.c_uint64(n_pairs), lo, hi, lmax, ptr(support), C.byref(bad), C.byref(walked), C.byref(over))
    assert rc == 0
    return support[:4 * lay.E], bad.value, walked.value, over.value


@pytest.mark.parametrize("k,n,seed", [(5, 300, 1), (6, 600, 2), (7, 1500, 3), (4, 120, 4), (9, 4000, 5)])
def test_bitset_walk_matches_oracle_dfs(emul, k, n, seed):
    """walk_one (forward / backward bitset fixed points on a local edge table) against the oracle's memoised DFS with its
    Dijkstra prune, on tangled small-k graphs, for random position pairs and several ranges."""
    s = rand_seq(n, seed)
    og = pyoracle.OracleGraph(kmers_of([s], k))
    lay = DeviceLayout(og)
    pk, pi, pd = lay.positions(og)
    kmer, ident, dist = og.graph_map()
    rng = np.random.default_rng(seed)
    emit = np.zeros(4 * lay.E, np.uint8)
    n_good = n_over = 0
    for trial in range(400):
        i, j = rng.integers(0, kmer.size, 2)
        lo = int(rng.integers(0, 30))
        hi = lo + int(rng.integers(0, 40))
        good, pairs = og.walk((int(ident[i]), int(dist[i])), (int(ident[j]), int(dist[j])), lo, hi)
        want = sorted((lay.edge_idx[a], lay.edge_idx[b]) for a, b in pairs)
        lmax = 4096 if trial % 4 else 6
        r = emul.emul_walk(*lay.graph_args(), int(pi[i]), int(pd[i]), int(pi[j]), int(pd[j]), lo, hi, lmax, ptr(emit))
        if r < 0:
            assert lmax == 6   # a small table may overflow, a large one never does here
            n_over += 1
            continue
        got = sorted((int(x) // 4, int(lay.out4[lay.edge_end[int(x) // 4], int(x) % 4])) for x in np.flatnonzero(emit))
        assert (bool(r), got) == (good, want), (trial, lo, hi)
        n_good += good
    assert n_good > 5 and n_over > 0


def test_long_range_bitset_words(emul):
    """Distances beyond one 64-bit word: a long linear sequence with a few forks, range up to 500."""
    k = 11
    s = rand_seq(3000, 77)
    alt = s[:1000] + ("A" if s[1000] != "A" else "G") + s[1001:]       # a bubble
    tip = s[2000 - k + 1:2000] + ("C" if s[2000] != "C" else "T") + rand_seq(40, 78)
    og = pyoracle.OracleGraph(kmers_of([s, alt, tip], k))
    lay = DeviceLayout(og)
    pk, pi, pd = lay.positions(og)
    kmer, ident, dist = og.graph_map()
    rng = np.random.default_rng(7)
    emit = np.zeros(4 * lay.E, np.uint8)
    n_good = 0
    for trial in range(600):
        i = int(rng.integers(0, kmer.size))
        # a partner a few hundred entries further along the same strand's listing: often reachable
        j = min(kmer.size - 1, i + int(rng.integers(0, 500)))
        lo = int(rng.integers(0, 400))
        hi = min(511, lo + int(rng.integers(0, 200)))
        good, pairs = og.walk((int(ident[i]), int(dist[i])), (int(ident[j]), int(dist[j])), lo, hi)
        want = sorted((lay.edge_idx[a], lay.edge_idx[b]) for a, b in pairs)
        r = emul.emul_walk(*lay.graph_args(), int(pi[i]), int(pd[i]), int(pi[j]), int(pd[j]), lo, hi, 256, ptr(emit))
        assert r >= 0
        got = sorted((int(x) // 4, int(lay.out4[lay.edge_end[int(x) // 4], int(x) % 4])) for x in np.flatnonzero(emit))
        assert (bool(r), got) == (good, want), (trial, lo, hi)
        n_good += good
    assert n_good > 10


def noisy_graph(k, genome_len, read_len, coverage, err, seed, insert, rounds=2, ragged=False):
    genome = synth.random_genome(genome_len, seed)
    n_reads = (int(coverage * genome_len / read_len) // 2) * 2
    reads = synth.sample_reads(genome, read_len, n_reads, err, seed + 1, insert=insert)
    if ragged:
        rng = np.random.default_rng(seed + 2)
        lens = rng.integers(read_len // 3, read_len + 1, size=n_reads)
        b = synth.pack_ragged([reads[i, :lens[i]] for i in range(n_reads)])
    else:
        b = synth.pack_fixed(reads)
    om = pyoracle.OracleMap(k)
    om.insert_reads(b, n_reads)
    om.delete_below(rounds)
    return pyoracle.OracleGraph(om), b, n_reads


@pytest.mark.parametrize("case", ["two_chromosomes", "noisy", "noisy_ragged", "after_retain_simplify"])
def test_pair_support_matches_oracle(emul, case):
    k, L = 15, 50
    if case == "two_chromosomes":
        g1, g2 = two_chromosomes(k, 11)
        reads = reads_of([g1, g2], L, 1500, (60, 100), 21)
        b, n_reads = synth.pack_fixed(reads), reads.shape[0]
        om = pyoracle.OracleMap(k)
        om.insert_reads(b, n_reads)
        om.delete_below(3)
        og = pyoracle.OracleGraph(om)
    else:
        # 2 % errors kept at count >= 2: tips, bubbles and short edges around every surviving error
        og, b, n_reads = noisy_graph(k, 6000, L, 40, 0.02, 91, (60, 100), rounds=2, ragged=case == "noisy_ragged")
        if case == "after_retain_simplify":
            og.retain_largest()
            og.simplify()
    lay = DeviceLayout(og)
    e1, e2, cnt, bad, walked = og.pair_support(b, n_reads // 2, 90, 155)
    want = lay.support_from_oracle(e1, e2, cnt)
    got, gbad, gwalked, over = emul_support(emul, lay, og, b, n_reads // 2, 90, 155)
    assert over == 0
    assert (gbad, gwalked) == (bad, walked)
    assert np.array_equal(got, want)
    assert walked > 0 and want.sum() > 0
    # the tiers: with a small table some cases overflow and commit nothing; the rest must still be exact subsets
    got_small, _, walked_small, over_small = emul_support(emul, lay, og, b, n_reads // 2, 90, 155, lmax=3)
    assert walked_small + over_small == walked
    assert np.all(got_small <= want)


def node_signatures(k, node_kmer, edge_start, edge_end, edge_off, bases, alive):
    """multiset of (node k-mer, sorted in-edge spellings, sorted out-edge spellings): which edges share which node copy"""
    ins = [[] for _ in node_kmer]
    outs = [[] for _ in node_kmer]
    for e in range(len(edge_start)):
        if not alive[e]:
            continue
        key = (int(node_kmer[edge_start[e]]), bases[int(edge_off[e]):int(edge_off[e + 1])].tobytes())
        outs[edge_start[e]].append(key)
        ins[edge_end[e]].append(key)
    return sorted((int(node_kmer[v]), tuple(sorted(ins[v])), tuple(sorted(outs[v]))) for v in range(len(node_kmer)))


@pytest.mark.parametrize("case,cutoff", [("two_chromosomes", 5), ("two_chromosomes", 10 ** 6), ("noisy", 1), ("noisy", 3), ("noisy", 8)])
def test_split_matches_oracle(emul, case, cutoff):
    k, L = 15, 50
    if case == "two_chromosomes":
        g1, g2 = two_chromosomes(k, 11)
        reads = reads_of([g1, g2], L, 1500, (60, 100), 21)
        b, n_reads = synth.pack_fixed(reads), reads.shape[0]
        om = pyoracle.OracleMap(k)
        om.insert_reads(b, n_reads)
        om.delete_below(3)
        og = pyoracle.OracleGraph(om)
    else:
        og, b, n_reads = noisy_graph(k, 6000, L, 40, 0.02, 91, (60, 100), rounds=2)
    lay = DeviceLayout(og)
    e1, e2, cnt, bad, walked = og.pair_support(b, n_reads // 2, 90, 155)
    support = lay.support_from_oracle(e1, e2, cnt)
    es, ee = lay.edge_start.copy(), lay.edge_end.copy()
    node_kmer2 = np.zeros(5 * lay.N, np.uint64)
    kill = np.zeros(lay.E, np.uint32)
    emul.emul_split.restype = C.c_int64
    added = emul.emul_split(lay.k, C.c_uint64(lay.N), C.c_uint64(lay.E), ptr(lay.node_kmer), ptr(es), ptr(ee), ptr(lay.edge_off),
                            ptr(lay.bases), ptr(support), cutoff, ptr(node_kmer2), ptr(kill))
    assert added >= 0
    removed, oadded = og.split(e1, e2, cnt, cutoff)
    assert (int(kill.sum()), added) == (removed, oadded)
    assert og.check() == 0
    got = node_signatures(k, node_kmer2[:lay.N + added], es, ee, lay.edge_off, lay.bases_codes, kill == 0)
    lay2 = DeviceLayout(og)
    want = node_signatures(k, lay2.node_kmer, lay2.edge_start, lay2.edge_end, lay2.edge_off, lay2.bases_codes, np.ones(lay2.E, bool))
    assert got == want
    if case == "noisy":
        assert added > 0


@pytest.mark.parametrize("case", ["fresh", "after_split"])
def test_graph_map_get_all_matches_oracle(emul, case):
    """The DNAMap[GraphPosition] handle (gb_graph_map_create / gb_graph_map_get_all): putNew of every getGraphMap entry and
    getAll / contains per key (posmap_lookup), against a dict built from the oracle's graph_map.  After a node split the
    copies of a node share its k-mer: multimap semantics (putNew never dedupes)."""
    k, L = 15, 50
    og, b, n_reads = noisy_graph(k, 6000, L, 40, 0.02, 91, (60, 100), rounds=2)
    if case == "after_split":
        e1, e2, cnt, _, _ = og.pair_support(b, n_reads // 2, 90, 155)
        _, added = og.split(e1, e2, cnt, 3)
        assert added > 0
    lay = DeviceLayout(og)
    pk, pi, pd = lay.positions(og)
    want = {}
    for x, i, d in zip(pk.tolist(), pi.tolist(), pd.tolist()):
        want.setdefault(x, []).append((i, d))
    rng = np.random.default_rng(5)
    absent = rng.integers(0, 1 << (2 * k), 500, dtype=np.uint64)
    keys = np.concatenate([pk[rng.permutation(pk.size)[:2000]], absent])
    max_per = 3
    counts = np.zeros(keys.size, np.uint32)
    ids = np.full((keys.size, max_per), NONE32, np.uint32)
    dists = np.full((keys.size, max_per), NONE32, np.uint32)
    emul.emul_graph_map_get_all(C.c_uint64(pk.size), ptr(pk), ptr(pi), ptr(pd), C.c_uint64(keys.size), ptr(keys), max_per, ptr(ids),
                                ptr(dists), ptr(counts))
    multi = 0
    for j, x in enumerate(keys.tolist()):
        w = want.get(x, [])
        assert counts[j] == len(w)
        got = sorted((int(ids[j, c]), int(dists[j, c])) for c in range(min(len(w), max_per)))
        if len(w) <= max_per:
            assert got == sorted(w)
        else:
            assert set(got) <= set(w)
        assert np.all(ids[j, len(w):] == NONE32)
        multi += len(w) > 1
    assert (multi > 0) == (case == "after_split")
    # contains only: no output arrays
    c2 = np.zeros(keys.size, np.uint32)
    emul.emul_graph_map_get_all(C.c_uint64(pk.size), ptr(pk), ptr(pi), ptr(pd), C.c_uint64(keys.size), ptr(keys), 0, None, None, ptr(c2))
    assert np.array_equal(c2, counts)


def test_pair_support_is_additive_over_pair_slices(emul):
    """The multi-GPU form of the pair loop (MapGraph.pairSupport(comm=...)): every rank walks its own contiguous slice of the
    pairs on its copy of the graph and the counts are summed.  Pairs are independent (GraphSimplifier.scala:213-248), so the
    sum over PairedEndData.shard slices must equal the whole -- checked here through the emulated device code, with the
    slices cut by the same host code the ranks use (fixed and ragged records)."""
    from genome_b200.dnamap import PairedEndData
    k, L = 15, 50
    for ragged in (False, True):
        og, b, n_reads = noisy_graph(k, 6000, L, 40, 0.02, 91, (60, 100), rounds=2, ragged=ragged)
        lay = DeviceLayout(og)
        whole, bad, walked, over = emul_support(emul, lay, og, b, n_reads // 2, 90, 155)
        assert over == 0 and walked > 0
        data = PairedEndData(b, n_reads // 2)
        for world in (2, 3, 8):
            acc = np.zeros_like(whole)
            tb = tw = tp = 0
            for rank in range(world):
                mine = data.shard(rank, world)
                s, sb, sw, so = emul_support(emul, lay, og, mine.bin, mine.count, 90, 155)
                assert so == 0
                acc += s
                tb, tw, tp = tb + sb, tw + sw, tp + mine.count
            assert tp == data.count
            assert (tb, tw) == (bad, walked)
            assert np.array_equal(acc, whole)
        # takeFirst, then the slices of that prefix
        first = data.take(data.count // 3)
        wf, bf, wkf, _ = emul_support(emul, lay, og, first.bin, first.count, 90, 155)
        acc = np.zeros_like(wf)
        for rank in range(4):
            mine = first.shard(rank, 4)
            acc += emul_support(emul, lay, og, mine.bin, mine.count, 90, 155)[0]
        assert np.array_equal(acc, wf)
        assert first.count == data.count // 3 and data.take(10 ** 9) is data


def test_walk_code_under_address_sanitizer():
    """The walk / split device functions (walk.cuh) once more, with the harness built with -fsanitize=address,undefined and the
    sanitizer runtime preloaded into a child pytest: every index these functions compute on the local edge tables, the bitsets
    and the caller's arrays is checked (numpy's buffers come from the intercepted malloc).  An overrun here is an overrun in
    walk_cases_kernel / split_nodes_kernel."""
    import sys
    if os.environ.get("GB_EMUL_SANITIZE"):
        pytest.skip("this IS the sanitized child")
    libs = [subprocess.run(["gcc", "-print-file-name=" + n], capture_output=True, text=True).stdout.strip() for n in ("libasan.so", "libubsan.so")]
    if not all(os.path.isabs(p) and os.path.exists(p) for p in libs):
        pytest.skip("no sanitizer runtime next to gcc")
    env = dict(os.environ, GB_EMUL_SANITIZE="1", LD_PRELOAD=" ".join(libs), ASAN_OPTIONS="detect_leaks=0:verify_asan_link_order=0")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-p", "no:cacheprovider",
                        "-k", "bitset_walk or long_range or pair_support_matches or split_matches or graph_map_get_all"],
                       capture_output=True, text=True, env=env, timeout=1500, cwd=ROOT)
    assert "ERROR: AddressSanitizer" not in r.stderr + r.stdout and "runtime error" not in r.stderr + r.stdout, (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
