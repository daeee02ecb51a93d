"""Parity of the CUDA path (through the C ABI) with the oracle on the same seeded inputs.  Bit-exact: integer work."""
import numpy as np
import pytest

from genome_b200 import synth
from genome_b200.dnamap import ArrayDNAMap, FreqFilter, PairedEndData
from genome_b200.graph import Graph
from genome_b200 import capi
from oracle import pyoracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["direct", "partitioned", "single-pass"])
def insert_path(request):
    """Every insert path against the oracle.  Small tables take the direct kernel by default; the L2-blocked path (the one
    bench.py times and every BASELINE config takes) is forced here, once with the counted bucket pass (part_count +
    part_scatter: also what the sharded map runs) and once with the single-pass one (part_scatter<SLABS>, the default for
    large batches), both followed by insert_keys_kernel."""
    kw = {"direct": dict(insert_path=1), "partitioned": dict(insert_path=2, single_pass=0),
          "single-pass": dict(insert_path=2, single_pass=1, single_pass_min=1)}[request.param]
    with capi.tuned(**kw):
        yield request.param


def gpu_map_from(bin_bytes, n_reads, k, min_capacity=0):
    m = ArrayDNAMap(k, min_capacity)
    w = m.insert_reads(bin_bytes, n_reads)
    return m, w


@pytest.mark.parametrize("k,read_len,ragged,err", [
    (31, 100, False, 0.0), (31, 100, False, 0.01), (21, 100, True, 0.01), (25, 150, False, 0.005),
    (31, 255, True, 0.0), (15, 36, False, 0.02), (8, 50, True, 0.0), (4, 30, False, 0.0), (1, 10, False, 0.0),
    (30, 64, True, 0.01), (16, 100, False, 0.0),
])
def test_insert_counts_match_oracle(gpu, insert_path, k, read_len, ragged, err):
    b, n, _ = H.small_reads(20000, read_len, 12, err, seed=1000 + k, ragged=ragged)
    om, ow = H.oracle_counts(b, n, k)
    gm, gw = gpu_map_from(b, n, k)
    assert gw == ow == pyoracle.count_windows(b, n, k)
    assert (gm.stats()["upsert_ns"] > 0) == (insert_path != "direct" and gw > 0)
    assert gm.size == om.size()
    gk, gv = gm.export_sorted()
    ok, ov = om.export_sorted()
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    assert int(gv.sum()) == gw
    # deleteAll(v < rounds)
    gm.delete_below(3)
    om.delete_below(3)
    gk, gv = gm.export_sorted()
    ok, ov = om.export_sorted()
    assert gm.size == om.size()
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)


def test_insert_grows_from_small_table(gpu, insert_path):
    b, n, _ = H.small_reads(200000, 100, 8, 0.01, seed=7)
    om, ow = H.oracle_counts(b, n, 31)
    gm, gw = gpu_map_from(b, n, 31, min_capacity=0)
    assert gm.stats()["grows"] > 0
    gk, gv = gm.export_sorted()
    ok, ov = om.export_sorted()
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)


def test_presized_equals_grown(gpu, insert_path):
    b, n, _ = H.small_reads(100000, 100, 10, 0.01, seed=8)
    a, _ = gpu_map_from(b, n, 31, min_capacity=0)
    c, _ = gpu_map_from(b, n, 31, min_capacity=4_000_000)
    assert c.stats()["grows"] == 0
    ak, av = a.export_sorted()
    ck, cv = c.export_sorted()
    assert np.array_equal(ak, ck) and np.array_equal(av, cv)


def test_empty_and_short_inputs(gpu):
    m = ArrayDNAMap(31)
    assert m.insert_reads(np.zeros(0, np.uint8), 0) == 0
    assert m.size == 0
    # reads shorter than k are skipped (FreqFilter.scala:29); a zero-length read is one byte
    b = synth.pack_ragged([np.zeros(0, np.uint8), synth.encode("ACGT"), synth.encode("A" * 30)])
    assert m.insert_reads(b, 3) == 0
    assert m.size == 0
    keys, vals = m.export()
    assert keys.size == 0
    g = Graph.buildGraph(31, m)
    assert g.counts() == (0, 0, 0)
    g.simplifyGraph()
    g.retain_largest()
    assert g.components()[0] == 0


def test_truncated_stream_is_an_error(gpu):
    m = ArrayDNAMap(31)
    b = synth.pack_fixed(synth.sample_reads(synth.random_genome(1000, 1), 100, 4, 0, 2, insert=(50, 100)))
    with pytest.raises(capi.GenomeError) as e:
        m.insert_reads(b[:-3], 4)
    assert e.value.name == "GB_E_ARG"


def test_k_range(gpu):
    for k in (0, 32, 33, -1):
        with pytest.raises(capi.GenomeError) as e:
            ArrayDNAMap(k)
        assert e.value.name == "GB_E_K_RANGE"


def test_lookup_update_and_masks(gpu):
    k = 21
    b, n, _ = H.small_reads(30000, 80, 10, 0.01, seed=11)
    om, _ = H.oracle_counts(b, n, k)
    gm, _ = gpu_map_from(b, n, k)
    gm.delete_below(2)
    om.delete_below(2)
    rng = np.random.default_rng(5)
    keys, _ = om.export()
    q = np.concatenate([keys[:2000], rng.integers(0, 1 << (2 * k), size=2000, dtype=np.uint64)])
    counts, found = gm.lookup(q)
    for i in range(q.size):
        v = om.apply(int(q[i]))
        assert found[i] == (v is not None)
        assert counts[i] == (v or 0)
    # neighbour masks of stored keys and of their reverse complements
    L = pyoracle.lib()
    qq = np.concatenate([keys[:1500], np.array([pyoracle.revcomp(int(x), k) for x in keys[:500]], np.uint64)])
    out, inn = gm.neighbour_masks(qq)
    for i in range(qq.size):
        x = int(qq[i])
        eo = sum(1 << bb for bb in range(4) if om.contains(L.go_append(x, k, bb)) or om.contains(pyoracle.revcomp(L.go_append(x, k, bb), k)))
        ei = sum(1 << bb for bb in range(4) if om.contains(L.go_prepend(x, k, bb)) or om.contains(pyoracle.revcomp(L.go_prepend(x, k, bb), k)))
        assert out[i] == eo and inn[i] == ei
    # update(key, v) and update(key, 1, _ + 1) on explicit keys
    m2 = ArrayDNAMap(k)
    o2 = pyoracle.OracleMap(k)
    ks = rng.integers(0, 1 << (2 * k), size=5000, dtype=np.uint64)
    ks = np.concatenate([ks, ks[:1000], ks[:10]])
    m2.update_counts(ks)
    for x in ks:
        o2.update1(int(x))
    uk = np.unique(ks)[:100]
    m2.update(uk, np.arange(100, dtype=np.int32) + 1000)
    for i, x in enumerate(uk):
        o2.update(int(x), 1000 + i)
    a, bv = m2.export_sorted()
    c, d = o2.export_sorted()
    assert np.array_equal(a, c) and np.array_equal(bv, d)
    with pytest.raises(capi.GenomeError):
        m2.update_counts(np.array([1 << (2 * k)], np.uint64))
    # apply / contains assert key.length == k (ArrayDNAMap.scala:182-206): a query with bits above 2k is an error, and the
    # empty-slot sentinel is not a key
    for bad in (1 << (2 * k), 0xFFFFFFFFFFFFFFFF):
        with pytest.raises(capi.GenomeError) as e:
            m2.lookup(np.array([bad], np.uint64))
        assert e.value.name == "GB_E_K_RANGE"
        with pytest.raises(capi.GenomeError) as e:
            m2.neighbour_masks(np.array([bad], np.uint64))
        assert e.value.name == "GB_E_K_RANGE"


GRAPH_CASES = [
    # k, genome, read_len, coverage, err, rounds
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
def test_build_graph_matches_oracle(gpu, k, glen, rl, cov, err, rounds):
    b, n, _ = H.small_reads(glen, rl, cov, err, seed=2000 + k)
    data = PairedEndData(b, n // 2)
    gm = FreqFilter.extractFilteredKmers(data, k, rounds)
    om, _ = H.oracle_counts(b, n, k)
    om.delete_below(rounds)
    assert gm.size == om.size()
    g = Graph.buildGraph(k, gm)
    og = pyoracle.OracleGraph(om)
    assert og.check() == 0
    g.check()
    assert g.counts() == og.counts()
    H.assert_graph_equal(g, og)
    # components: same partition of the node set
    nc, label = g.components()
    onc, olabel = og.components()
    assert nc == onc
    gn = g.export()[0]
    on = og.export()[0]
    gsets = sorted(sorted(int(x) for x in gn[label == c]) for c in range(nc))
    osets = sorted(sorted(int(x) for x in on[olabel == c]) for c in range(onc))
    assert gsets == osets


@pytest.mark.parametrize("k,glen,rl,cov,err,rounds", GRAPH_CASES)
def test_graph_operators_match_oracle(gpu, k, glen, rl, cov, err, rounds):
    b, n, _ = H.small_reads(glen, rl, cov, err, seed=3000 + k)

    def fresh():
        gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, rounds)
        om, _ = H.oracle_counts(b, n, k)
        om.delete_below(rounds)
        return Graph.buildGraph(k, gm), pyoracle.OracleGraph(om)

    # simplifyGraph on the fresh graph (a no-op for a unitig graph except isolated loops) and after edits
    g, og = fresh()
    g.simplifyGraph(); og.simplify()
    H.assert_graph_equal(g, og)
    # removeBubbles, then simplify
    g, og = fresh()
    g.removeBubbles(); og.remove_bubbles()
    g.check()
    H.assert_graph_equal(g, og)
    g.simplifyGraph(); og.simplify()
    H.assert_graph_equal(g, og)
    assert og.check() == 0
    # tip clipping (extension), then simplify
    g, og = fresh()
    r = g.clipTips(2 * k); orr = og.clip_tips(2 * k)
    assert r == orr
    H.assert_graph_equal(g, og)
    g.simplifyGraph(); og.simplify()
    H.assert_graph_equal(g, og)
    # retain(maxBy size), then simplify
    g, og = fresh()
    g.retain_largest(); og.retain_largest()
    H.assert_graph_equal(g, og)
    g.simplifyGraph(); og.simplify()
    H.assert_graph_equal(g, og)


def test_remove_edges_then_simplify(gpu):
    k = 15
    b, n, _ = H.small_reads(8000, 60, 30, 0.02, seed=77)
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, 2)
    om, _ = H.oracle_counts(b, n, k)
    om.delete_below(2)
    g = Graph.buildGraph(k, gm)
    og = pyoracle.OracleGraph(om)
    # remove the same edges on both sides: pick by canonical key
    gn, ge = H.canon_gpu_graph(g)
    pick = set(ge[::3])
    node_kmer, es, ee, off, bases = g.export()
    gidx = [i for i in range(es.size) if (int(node_kmer[es[i]]), int(node_kmer[ee[i]]), bases[int(off[i]):int(off[i + 1])].tobytes()) in pick]
    onk, oid, oes, oee, ooff, obases = og.export()
    by_id = {int(i): int(x) for i, x in zip(oid, onk)}
    # oracle edge ids = position among ALL edges ever created (all alive here) + 1
    oidx = [i + 1 for i in range(oes.size) if (by_id[int(oes[i])], by_id[int(oee[i])], obases[int(ooff[i]):int(ooff[i + 1])].tobytes()) in pick]
    g.removeEdges(np.array(gidx, np.uint32))
    oarr = np.array(oidx, np.int64)  # keep the array alive across the call
    assert pyoracle.lib().go_graph_remove_edges(og.h, oarr.ctypes.data, oarr.size) == len(gidx)
    H.assert_graph_equal(g, og)
    g.simplifyGraph(); og.simplify()
    H.assert_graph_equal(g, og)
    g.check()


def test_noncanonical_keys_both_orientations(gpu):
    """Keys pushed through update as they are: both orientations of a k-mer can be stored (like a hash tie,
    SURVEY Q3); contains() probes both, nodes are a SET of oriented k-mers."""
    k = 9
    rng = np.random.default_rng(3)
    genome = synth.random_genome(2000, 99)
    fw = np.array([synth.kmer_to_int(synth.decode(genome[i:i + k])) for i in range(genome.size - k + 1)], np.uint64)
    rc = np.array([pyoracle.revcomp(int(x), k) for x in fw], np.uint64)
    pick = rng.random(fw.size)
    keys = np.concatenate([fw[pick < 0.6], rc[pick > 0.4]])  # overlap: both orientations for 20%
    gm = ArrayDNAMap(k)
    om = pyoracle.OracleMap(k)
    gm.update_counts(keys)
    for x in keys:
        om.update1(int(x))
    assert gm.size == om.size()
    g = Graph.buildGraph(k, gm)
    og = pyoracle.OracleGraph(om)
    H.assert_graph_equal(g, og)
    g.simplifyGraph(); og.simplify()
    H.assert_graph_equal(g, og)


@pytest.mark.parametrize("strands", ["both", "forward"])
def test_hash_tie_kmers_in_real_reads(gpu, insert_path, strands):
    """x with hash(x) == hash(rc x), x != rc x (even k >= 18 only, tests/golden/hash_ties.json): a read of x stores rc(x) and a
    read of rc(x) stores x (FreqFilter.scala:32: tie => rcx).  Reads of both strands leave BOTH orientations in the table of a map
    that is not `dual`: the insert's tie rule, contains() probing both orientations and the primary / secondary orientation of
    Graph.buildGraph (single GPU and sharded) are exercised with real reads; reads of one strand store only the orientation
    the reads do not spell."""
    for k, kmers in H.hash_ties():
        b, n = H.tie_reads(k, kmers, seed=k, strands=strands)
        om, ow = H.oracle_counts(b, n, k)
        gm, gw = gpu_map_from(b, n, k)
        assert gw == ow
        gk, gv = gm.export_sorted()
        ok, ov = om.export_sorted()
        assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
        stored = set(int(x) for x in gk)
        for x in kmers:
            assert pyoracle.revcomp(x, k) in stored and (x in stored) == (strands == "both")
        _, found = gm.lookup(np.array(kmers, np.uint64))     # apply(x): exact orientation, no canonicalisation
        assert found.all() if strands == "both" else not found.any()
        gm.delete_below(3)
        om.delete_below(3)
        gk, gv = gm.export_sorted()
        ok, ov = om.export_sorted()
        assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
        og = pyoracle.OracleGraph(om)
        g = Graph.buildGraph(k, gm)
        assert og.counts()[1] > 0
        H.assert_graph_equal(g, og)
        g.check()
        for P in (2, 3):
            H.assert_graph_equal(Graph.buildGraphVirtualShards(k, gm, P), og)
        g.simplifyGraph(); og.simplify()
        H.assert_graph_equal(g, og)


def test_canonical_rule_on_random_kmers(gpu):
    """The canonical choice (FreqFilter.scala:31-32) on all orientations of random k-mers for several k."""
    for kk in (16, 17, 21, 31):
        rng = np.random.default_rng(kk)
        xs = rng.integers(0, 1 << (2 * kk), size=4000, dtype=np.uint64)
        reads = [np.array([(int(x) >> (2 * i)) & 3 for i in range(kk)], np.uint8) for x in xs]
        b = synth.pack_ragged(reads)
        gm = ArrayDNAMap(kk)
        gm.insert_reads(b, len(reads))
        gk, gv = gm.export_sorted()
        exp = {}
        for x in xs:
            c = pyoracle.canonical(int(x), kk)
            exp[c] = exp.get(c, 0) + 1
        assert [int(x) for x in gk] == sorted(exp)
        assert [int(v) for v in gv] == [exp[x] for x in sorted(exp)]


def test_perfect_cycle_is_dropped(gpu):
    """A circular sequence with no branch has no terminal k-mer: buildGraph yields nothing (Graph.scala:375)."""
    k = 11
    genome = synth.random_genome(500, 8)  # a seed without a repeated 10-mer on either strand
    circ = np.concatenate([genome, genome[:k - 1]])
    keys = np.array([pyoracle.canonical(synth.kmer_to_int(synth.decode(circ[i:i + k])), k) for i in range(genome.size)], np.uint64)
    gm = ArrayDNAMap(k)
    om = pyoracle.OracleMap(k)
    gm.update_counts(keys)
    for x in keys:
        om.update1(int(x))
    g = Graph.buildGraph(k, gm)
    og = pyoracle.OracleGraph(om)
    assert om.size() == genome.size and og.counts() == (0, 0, 0)
    assert g.counts() == (0, 0, 0)
    assert g.stats()["cycle_vertices"] == 2 * genome.size


def test_error_free_linear_genome_is_two_edges(gpu):
    """an error-free random linear genome gives exactly 4 nodes / 2 edges, and the two edges spell the
    genome and its reverse complement."""
    k = 31
    glen = 300000
    b, n, genome = H.small_reads(glen, 100, 40, 0.0, seed=42)
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, 1)
    g = Graph.buildGraph(k, gm)
    nn, ne, nb = g.counts()
    kept = gm.size
    nodes, edges = H.canon_gpu_graph(g)
    # coverage 40 leaves no gap with overwhelming probability; if the ends are uncovered the covered span is shorter
    assert (nn, ne) == (4, 2)
    assert nb == 2 * (kept - 1)
    (u1, v1, s1), (u2, v2, s2) = edges
    spell = lambda u, s: synth.int_to_kmer(u, k) + synth.decode(np.frombuffer(s, np.uint8))
    a, c = spell(u1, s1), spell(u2, s2)
    gs = synth.decode(genome)
    rc = synth.decode(H.revcomp_codes(genome))
    assert (a in gs and c in rc) or (a in rc and c in gs)
    assert len(a) == kept + k - 1
    # retain keeps one strand, simplify leaves it alone
    g.retain_largest()
    assert g.counts()[:2] == (2, 1)
    g.simplifyGraph()
    assert g.counts()[:2] == (2, 1)


def test_golden_fixture_on_gpu(gpu):
    """The committed fixture (tests/golden/small_k15.json, written by make_golden.py from the oracle)."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_k15.json")))
    b = np.frombuffer(bytes.fromhex(gold["bin_hex"]), np.uint8)
    m = ArrayDNAMap(gold["k"])
    assert m.insert_reads(b, gold["n_reads"]) == gold["windows"]
    keys, vals = m.export_sorted()
    assert [int(x) for x in keys] == gold["keys"] and [int(x) for x in vals] == gold["counts"]
    m.delete_below(gold["rounds"])
    g = Graph.buildGraph(gold["k"], m)
    nodes, edges = H.canon_gpu_graph(g)
    assert nodes == gold["nodes"]
    assert [[u, v, s.hex()] for (u, v, s) in edges] == gold["edges"]
    g.retain_largest()
    g.simplifyGraph()
    _, edges = H.canon_gpu_graph(g)
    assert [[u, v, s.hex()] for (u, v, s) in edges] == gold["edges_after_retain_simplify"]


def test_device_resident_stream_and_clear(gpu):
    """gb_map_insert_reads_device (fixed stride and with explicit offsets) and gb_map_clear."""
    import torch
    k = 25
    b, n, _ = H.small_reads(30000, 100, 10, 0.01, seed=55)
    om, ow = H.oracle_counts(b, n, k)
    ok, ov = om.export_sorted()
    d = torch.zeros(b.size + 16, dtype=torch.uint8, device="cuda")
    d[:b.size].copy_(torch.from_numpy(b))
    m = ArrayDNAMap(k, 1 << 20)
    for rep in range(2):
        assert m.insert_reads_device(d.data_ptr(), b.size, n) == ow
        gk, gv = m.export_sorted()
        assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
        m.clear(1 << 20)
        assert m.size == 0
    off = torch.from_numpy(PairedEndData(b, n // 2).record_offsets().astype(np.int64)).cuda()
    assert m.insert_reads_device(d.data_ptr(), b.size, n, off.data_ptr()) == ow
    gk, gv = m.export_sorted()
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    # ragged stream with offsets
    rb, rn, _ = H.small_reads(20000, 120, 8, 0.01, seed=56, ragged=True)
    rm, rw = H.oracle_counts(rb, rn, k)
    d2 = torch.zeros(rb.size + 16, dtype=torch.uint8, device="cuda")
    d2[:rb.size].copy_(torch.from_numpy(rb))
    off2 = torch.from_numpy(PairedEndData(rb, rn // 2).record_offsets().astype(np.int64)).cuda()
    m.clear(0)
    assert m.insert_reads_device(d2.data_ptr(), rb.size, rn, off2.data_ptr()) == rw
    gk, gv = m.export_sorted()
    rk, rv = rm.export_sorted()
    assert np.array_equal(gk, rk) and np.array_equal(gv, rv)
    # a ragged stream without offsets is refused
    m.clear(0)
    with pytest.raises(capi.GenomeError):
        m.insert_reads_device(d2.data_ptr(), rb.size, rn)


def test_full_size_properties_c1(gpu):
    """BASELINE configs[0] at full size (4.6 Mbp, 100 bp error-free reads at 30x, k = 31) through size-independent
    properties: sum of counts = windows; the kept set is the genome's k-mers; 4 nodes / 2 edges spelling the genome
    span and its reverse complement; deleteAll is idempotent."""
    k = 31
    b, n, genome = synth.make_config("C1")
    m = ArrayDNAMap(k, 6_000_000)
    w = m.insert_reads(b, n)
    assert w == n * (100 - k + 1)
    keys, vals = m.export()
    assert int(vals.astype(np.int64).sum()) == w
    assert len(np.unique(keys)) == keys.size == m.size
    m.delete_below(3)
    s1 = m.size
    m.delete_below(3)
    assert m.size == s1
    g = Graph.buildGraph(k, m)
    nn, ne, nb = g.counts()
    # 30x coverage with a threshold of 3 leaves a few gaps: every component is a strand pair of one linear contig
    nc, label = g.components()
    assert nn == 2 * nc and ne == nc and nn % 4 == 0
    # sum of edge lengths = oriented kept k-mers - oriented node k-mers + edges; no isolated k-mers
    assert nb == (2 * s1 - nn) + ne
    g.retain_largest()
    g.simplifyGraph()
    assert g.counts()[:2] == (2, 1)
    node_kmer, es, ee, off, bases = g.export()
    contig = synth.int_to_kmer(int(node_kmer[es[0]]), k) + synth.decode(bases)
    gs = synth.decode(genome)
    assert contig in gs or contig in synth.decode(H.revcomp_codes(genome))


@pytest.mark.parametrize("cfg", ["C1", "C2"])
def test_full_size_oracle_equality(gpu, cfg):
    """BASELINE configs[0] and configs[1] at FULL size (4.6 Mbp, 100 bp x 30x, k = 31; C2 with 1 % errors = the bench
    workload, L2-blocked insert path) against the oracle: sorted (k-mer, count) table, kept set after deleteAll(v < 3),
    node set and edge multiset of Graph.buildGraph, and again after retain + simplifyGraph.  About 90 s of oracle time."""
    k = 31
    b, n, _ = synth.make_config(cfg)
    om, ow = H.oracle_counts(b, n, k)
    gm = ArrayDNAMap(k, 34_000_000 if cfg == "C2" else 6_000_000)
    assert gm.insert_reads(b, n) == ow == n * (100 - k + 1)
    assert gm.stats()["upsert_ns"] > 0, "the L2-blocked insert path (the benchmarked one) must be the one tested here"
    gk, gv = gm.export_sorted()
    ok, ov = om.export_sorted()
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    del gk, gv, ok, ov
    gm.delete_below(3)
    om.delete_below(3)
    assert gm.size == om.size()
    gk, gv = gm.export_sorted()
    ok, ov = om.export_sorted()
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    g = Graph.buildGraph(k, gm)
    og = pyoracle.OracleGraph(om)
    assert g.counts() == og.counts()
    g.check()
    H.assert_graph_equal(g, og)
    g.retain_largest(); og.retain_largest()
    g.simplifyGraph(); og.simplify()
    assert g.counts() == og.counts()
    H.assert_graph_equal(g, og)


def test_cpp_host_driver(gpu, tmp_path):
    """hostcpp/graph_builder.cpp (GraphBuilder.startup over the C++ mirror) prints the oracle's numbers."""
    from tests.test_host_cpu import build_cpp_driver
    import subprocess
    exe = build_cpp_driver(str(tmp_path))
    k = 21
    b, n, _ = H.small_reads(20000, 80, 20, 0.01, seed=404)
    p = tmp_path / "r.bin"
    p.write_bytes(b.tobytes())
    r = subprocess.run([exe, str(p), str(n // 2), str(k)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    om, _ = H.oracle_counts(b, n, k)
    om.delete_below(3)
    og = pyoracle.OracleGraph(om)
    assert "Good reads count: %d" % om.size() in r.stdout
    assert "Total edges length: %d" % og.counts()[2] in r.stdout
    og.retain_largest()
    og.simplify()
    assert "Graph nodes: %d" % og.counts()[0] in r.stdout
    assert "Node map: %d" % og.graph_map()[0].size in r.stdout


def test_build_twice_and_after_queries(gpu):
    """Graph.buildGraph reuses the filter's vertex array: it must survive queries, a second build, a repeated deleteAll,
    and be dropped by any mutation."""
    k = 21
    b, n, _ = H.small_reads(30000, 100, 20, 0.01, seed=808)
    gm = FreqFilter.extractFilteredKmers(PairedEndData(b, n // 2), k, 3)
    om, _ = H.oracle_counts(b, n, k)
    om.delete_below(3)
    og = pyoracle.OracleGraph(om)
    g1 = Graph.buildGraph(k, gm)
    H.assert_graph_equal(g1, og)
    keys, _ = gm.export()
    gm.lookup(keys[:1000])
    gm.neighbour_masks(keys[:1000])
    g2 = Graph.buildGraph(k, gm)
    H.assert_graph_equal(g2, og)
    gm.delete_below(3)  # removes nothing, but re-compacts the survivors in another order
    g3 = Graph.buildGraph(k, gm)
    H.assert_graph_equal(g3, og)
    gm.delete_below(5)
    om.delete_below(5)
    H.assert_graph_equal(Graph.buildGraph(k, gm), pyoracle.OracleGraph(om))
    extra = np.array([pyoracle.canonical(int(x) ^ 0x155, k) for x in keys[:50]], np.uint64)
    gm.update_counts(extra)
    for x in extra:
        om.update1(int(x))
    H.assert_graph_equal(Graph.buildGraph(k, gm), pyoracle.OracleGraph(om))
