"""Host-side formats and the oracle's getGraphMap restatement (row 3): CPU only."""
import numpy as np
import pytest

from genome_b200 import formats, synth
from genome_b200 import formats as F
from oracle import pyoracle
from tests import helpers as H


def fastq(pairs, n):
    lines = []
    for i, (a, b) in enumerate(pairs):
        lines += ["@r%d" % i, a + b, "+", "I" * (len(a) + len(b))]
    return [l + "\n" for l in lines]


def test_convert2bin_follows_the_reference():
    n, k = 36, 23
    rng = np.random.default_rng(0)
    rd = lambda m: "".join("AGCT"[c] for c in rng.integers(0, 4, size=m))
    pairs = [(rd(36), rd(36)), (rd(20) + "N" + rd(15), rd(36)), (rd(36), "N" + rd(35)), (rd(36), rd(30))]
    b, count, kmers, short = formats.convert2bin(fastq(pairs, n), n, k)
    assert count == 4
    reads = formats.read_bin(b, 8)
    # takeWhile: cut at the first non-base; lower case is not a base either
    assert [len(r) for r in reads] == [36, 36, 20, 36, 36, 0, 36, 30]
    assert synth.decode(reads[2]) == pairs[1][0][:20]
    assert synth.decode(reads[7]) == pairs[3][1]
    assert kmers == sum(max(0, len(r) - k + 1) for r in reads)
    assert short == 2  # the pairs with a read shorter than k: (20, 36) and (36, 0)
    assert b.size == sum(1 + (len(r) + 3) // 4 for r in reads)
    # the stream feeds the counting path unchanged
    assert pyoracle.count_windows(b, 8, k) == kmers
    # truncated input: a record without its quality line is dropped
    b2, count2, _, _ = formats.convert2bin(fastq(pairs, n)[:-1], n, k)
    assert count2 == 3 and np.array_equal(b2, b[:b2.size])
    assert formats.convert2bin([], n, k)[1] == 0


def test_read_bin_roundtrip_and_truncation():
    rng = np.random.default_rng(1)
    reads = [rng.integers(0, 4, size=int(m), dtype=np.uint8) for m in [0, 1, 4, 5, 100, 255]]
    b = synth.pack_ragged(reads)
    back = formats.read_bin(b, len(reads))
    assert all(np.array_equal(x, y) for x, y in zip(back, reads))
    try:
        formats.read_bin(b[:-1], len(reads))
        assert False
    except ValueError:
        pass


def test_contigs_file(tmp_path):
    p = tmp_path / "contigs"
    formats.write_contigs([synth.encode("ACGT"), synth.encode("GG")], str(p))
    assert p.read_text() == "ACGT\n>abacaba0\nGG\n>abacaba1\n"


def test_graph_map_restatement():
    """Graph.getGraphMap (Graph.scala:90-119): nodes + interior edge k-mers, each oriented k-mer exactly once; CheckGraph's
    invariant (CheckGraph.scala:48-55): every k-mer of the genome that survived the filter is a key, in read orientation."""
    k = 15
    b, n, genome = H.small_reads(6000, 60, 30, 0.01, seed=12)
    m, _ = H.oracle_counts(b, n, k)
    m.delete_below(2)
    g = pyoracle.OracleGraph(m)
    kmer, ident, dist = g.graph_map()
    nn, ne, nb = g.counts()
    assert kmer.size == nb + nn - ne            # `total`, Graph.scala:97
    assert len(set(kmer.tolist())) == kmer.size  # putNew never collides here: every oriented k-mer has one position
    assert int((dist == 0).sum()) == nn
    # positions are consistent with the edges: the k-mer at (edge, dist) is the window of start + seq that ends at seq[dist - 1]
    node_kmer, node_id, es, ee, off, bases = g.export()
    by_id = {int(i): int(x) for i, x in zip(node_id, node_kmer)}
    # oracle edge ids: position among all edges ever created + 1; a fresh graph has none removed
    for e in range(0, es.size, max(1, es.size // 50)):
        s = synth.int_to_kmer(by_id[int(es[e])], k) + synth.decode(bases[int(off[e]):int(off[e + 1])])
        sel = (ident == e + 1) & (dist > 0)
        assert int(sel.sum()) == int(off[e + 1] - off[e]) - 1  # one entry per interior k-mer of the edge
        for km, d in zip(kmer[sel].tolist(), dist[sel].tolist()):
            assert synth.int_to_kmer(km, k) == s[d:d + k]
    keys = set(kmer.tolist())
    gs = synth.decode(genome)
    found = missing = 0
    for i in range(len(gs) - k + 1):
        x = synth.kmer_to_int(gs[i:i + k])
        if m.contains(x) or m.contains(pyoracle.revcomp(x, k)):
            if x in keys:
                found += 1
            else:
                missing += 1  # only isolated (0,0) k-mers and perfect cycles are absent (Graph.scala:375)
    assert found > 0 and missing <= 0.001 * found


def test_paired_end_header_round_trip_and_known_bytes():
    """PairedEndData.write / PairedEndData(f) (S/data/PairedEndData.scala:14-18,38-41): the Java object stream of the header.
    The java.io.File part is checked against the JDK's well-known serial form (class descriptor with serialVersionUID
    0x042DA4450E0DE4FF, flags SC_SERIALIZABLE | SC_WRITE_METHOD, one String field `path`, separator char as block data)."""
    b = F.write_paired_end_header(690000, 200, "/data/ecoli.bin")
    assert b[:4] == bytes.fromhex("aced0005")
    assert F.read_paired_end_header(b) == (690000, 200, "/data/ecoli.bin")
    file_part = bytes.fromhex("7372000c6a6176612e696f2e46696c65042da4450e0de4ff0300014c0004706174687400124c6a6176612f6c616e672f537472696e673b7870"
                              "74000f2f646174612f65636f6c692e62696e7702002f78")
    assert b.endswith(file_part)
    # class descriptor: name, serialVersionUID 1 (@SerialVersionUID(1L), PairedEndData.scala:11), SC_SERIALIZABLE, 3 fields
    name = b"ru.ifmo.genome.data.PairedEndData"
    assert b[4:6] == b"\x73\x72" and b[6:8] == len(name).to_bytes(2, "big") and b[8:8 + len(name)] == name
    assert b[8 + len(name):8 + len(name) + 11] == (1).to_bytes(8, "big") + b"\x02\x00\x03"
    # count is a Long: values beyond 32 bits survive
    assert F.read_paired_end_header(F.write_paired_end_header(1 << 40, -1, "x"))[:2] == (1 << 40, -1)


def test_paired_end_header_reader_is_order_independent():
    """A stream a different JVM could have written: object field first in the descriptor is impossible for ObjectStreamClass,
    but back references (TC_REFERENCE) to an earlier type string and a superclass-free second object are legal; fields are
    found by name."""
    import struct
    utf = lambda s: struct.pack(">H", len(s)) + s.encode()
    out = bytearray(bytes.fromhex("aced0005"))
    out += b"\x73\x72" + utf("ru.ifmo.genome.data.PairedEndData") + struct.pack(">q", 1) + b"\x02" + struct.pack(">H", 4)
    out += b"I" + utf("insert") + b"J" + utf("count")                      # other order than ours
    out += b"L" + utf("aux") + b"\x74" + utf("Ljava/io/File;")             # handle 0x7e0001 = this type string
    out += b"L" + utf("bin") + b"\x71" + struct.pack(">I", 0x7E0001)       # reference to it
    out += b"\x78\x70"
    out += struct.pack(">iq", 321, 99)
    out += b"\x70"                                                          # aux = null
    out += b"\x73\x72" + utf("java.io.File") + struct.pack(">q", 301077366599181567) + b"\x03" + struct.pack(">H", 1)
    out += b"L" + utf("path") + b"\x74" + utf("Ljava/lang/String;") + b"\x78\x70"
    out += b"\x74" + utf("C:\\reads.bin") + b"\x77\x02" + struct.pack(">H", ord("\\")) + b"\x78"
    assert F.read_paired_end_header(bytes(out)) == (99, 321, "C:\\reads.bin")


@pytest.mark.parametrize("bad", ["magic", "truncated", "class"])
def test_paired_end_header_errors(bad):
    b = bytearray(F.write_paired_end_header(5, 200, "a.bin"))
    if bad == "magic":
        b[0] = 0
    elif bad == "truncated":
        b = b[:-6]
    else:
        b[10] ^= 1   # another class name
    with pytest.raises(ValueError):
        F.read_paired_end_header(bytes(b))


def test_paired_end_data_files(tmp_path):
    """Convert2bin's two outputs (`<name>.bin` + the header object `<name>`, Convert2bin.scala:18,83) written and read back
    through the PairedEndData mirror."""
    from genome_b200.dnamap import PairedEndData
    genome = synth.random_genome(3000, 3)
    reads = synth.sample_reads(genome, 40, 200, 0.0, 4)
    d = PairedEndData(synth.pack_fixed(reads), 100, insert=200)
    d.write(str(tmp_path / "reads"), str(tmp_path / "reads.bin"))
    e = PairedEndData.apply(str(tmp_path / "reads"))
    assert (e.count, e.insert) == (100, 200) and np.array_equal(e.bin, d.bin)


def test_checkgraph_host_logic():
    """CheckGraph.scala:37-41,46-47 and N50.scala:14-31: the host-side parts (no device)."""
    from genome_b200 import checkgraph as CG
    st = CG.contig_stats([10, 201, 200, 500, 300, 1000])
    assert st == dict(count=4, size=2001, n50=500, max=1000)   # contigs(4 / 2) of [201, 300, 500, 1000]
    with pytest.raises(IndexError):
        CG.contig_stats([5, 200])
    k = 5
    line = "ACGTNACGTAC"
    keys, starts, short = CG.line_windows(line, k)
    want = [i for i in range(len(line) - k + 1) if "N" not in line[i:i + k]]
    assert starts.tolist() == want and short == 0
    assert [synth.int_to_kmer(int(x), k) for x in keys] == [line[i:i + k] for i in want]
    assert CG.line_windows("ACG", k)[2] == 1 and CG.line_windows("ANG", k)[2] == 0 and CG.line_windows("", k)[2] == 0
    # N50: pairs sorted by length; the crossing pair; the trailing pair without a comma is invisible to the reference's regex
    out, pairs, ge100 = CG.n50("hist: Map(100 -> 2, 50 -> 4, 300 -> 1, 7 -> 9)")
    assert pairs == [(50, 4), (100, 2), (300, 1)] and ge100 == 3
    assert out == [(100, 2)]   # total 700: 200 < 350 <= 400


def test_kmers_calculator_windows():
    """KmersCalculator.scala:23-27, host part: header dropped, lines concatenated, windows in reading order (not canonical)."""
    from genome_b200 import checkgraph as CG
    n, keys, short = CG.sequence_windows([">chr1 test\n", "ACGTAC\n", "GTTA\n"], 4)
    text = "ACGTACGTTA"
    assert n == 10 and not short
    assert [synth.int_to_kmer(int(x), 4) for x in keys] == [text[i:i + 4] for i in range(7)]
    assert len(set(keys.tolist())) == len({text[i:i + 4] for i in range(7)}) == 6   # ACGT twice
    assert CG.sequence_windows([">h", "ACG"], 4) == (3, pytest.approx(np.zeros(1)), True)
    assert CG.sequence_windows([">h"], 4)[0] == 0
    with pytest.raises(ValueError):
        CG.sequence_windows([">h", "ACNT"], 2)


def test_graph_builder_histograms_on_oracle_graph():
    """GraphBuilder.scala:41-47: the two component histograms, computed from exported arrays; checked here on the oracle's graph
    against a direct per-component tally."""
    from genome_b200.builder import component_histograms
    b, n, _ = H.small_reads(4000, 40, 15, 0.03, seed=2009)
    om, _ = H.oracle_counts(b, n, 9)
    om.delete_below(1)
    og = pyoracle.OracleGraph(om)
    node_kmer, node_id, es, ee, off, bases = og.export()
    nc, label = og.components()
    idx = {int(i): j for j, i in enumerate(node_id)}
    es_i = np.array([idx[int(x)] for x in es], np.int64)
    hist, hist2, comp_nodes = component_histograms(nc, label, es_i, off)
    sizes, lens = {}, {}
    for c in range(nc):
        nodes = np.flatnonzero(label == c)
        sizes[len(nodes)] = sizes.get(len(nodes), 0) + 1
        total = sum(int(off[e + 1] - off[e]) for e in range(es_i.size) if label[es_i[e]] == c)
        lens[total] = lens.get(total, 0) + 1
    assert hist == sorted(sizes.items()) and hist2 == sorted(lens.items())
    assert sum(c for _, c in hist) == nc and int(comp_nodes.max()) == max(sizes)
    assert component_histograms(0, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(1, np.uint64))[:2] == ([], [])


def test_convert2bin_cuts_reads_to_their_quality_segment():
    """Convert2bin.scala:61-62 zips the bases with the quality string, so a read is as long as the shorter of the two."""
    import io
    from genome_b200 import formats
    n, k = 8, 3
    fq = "@r\nACGTACGTGGCCAATT\n+\nIIIIIIIIIIII\n"   # 16 bases, 12 qualities: mate 1 keeps 8, mate 2 keeps 4
    b, pairs, kmers, short = formats.convert2bin(io.StringIO(fq), n, k)
    reads = formats.read_bin(b, 2)
    assert pairs == 1 and [len(r) for r in reads] == [8, 4]
    assert kmers == (8 - k + 1) + (4 - k + 1) and short == 0


# ---- the Kryo `graph` file (Graph.scala:232-261, Node.scala:14-37): formats.write_kryo_graph / read_kryo_graph
def _zz(v):
    """Kryo's writeLong(v, false) for small non-negative v: zig-zag then one varint byte"""
    assert 0 <= v < 64
    return bytes([2 * v])


def test_kryo_graph_bytes_of_a_two_node_graph():
    """The stream spelled out by hand from the reference's call sites and Kryo 2.x's wire rules (formats.py header): a graph
    GT -> TA with one edge `A` (k = 2).  Marker 1 before every object, the class name once, name id afterwards."""
    k = 2
    gt = 1 | 3 << 2      # G = 1 at bits 0-1, T = 3 at bits 2-3
    ta = 3 | 0 << 2
    got = F.write_kryo_graph(k, [gt, ta], [0], [1], [np.array([0], np.uint8)])
    name = bytearray(b"ru.ifmo.genome.dna.Long1DNASeq")
    name[-1] |= 0x80
    want = b"\x01"                                           # writeObject(out, this): new object
    want += b"\x00\x00\x00\x02"                              # out.writeInt(nodes.size)
    want += b"\x01" + (1).to_bytes(8, "big")                 # node 1: marker, out.writeLong(id)
    want += b"\x01\x00" + bytes(name)                        # writeClass: NAME + 2, name id 0, the name
    want += b"\x01" + bytes([k]) + bytes([2 * gt])           # marker, len: Byte, long: zig-zag varlong
    want += b"\x00\x00\x00\x00"                              # no in-edges
    want += b"\x00\x00\x00\x01" + b"\x00" + (1).to_bytes(8, "big")   # one out-edge: base A -> edge 1
    want += b"\x01" + (2).to_bytes(8, "big")                 # node 2
    want += b"\x01\x00"                                      # the class again: name id only
    want += b"\x01" + bytes([k]) + bytes([2 * ta])
    want += b"\x00\x00\x00\x01" + (1).to_bytes(8, "big")     # in-edge 1
    want += b"\x00\x00\x00\x00"
    want += b"\x00\x00\x00\x01"                              # out.writeInt(edges.size)
    want += b"\x01" + _zz(2) + _zz(1)                        # edge: marker, endId, id (FieldSerializer: fields by name)
    want += b"\x01\x00" + b"\x01" + bytes([1]) + _zz(0)      # seq: class by id, marker, len 1, long 0
    want += _zz(1)                                           # startId
    assert got == want
    nodes, edges = F.read_kryo_graph(got)
    assert [(n[0], n[1].tolist(), n[2], n[3]) for n in nodes] == [(1, [1, 3], [], [(0, 1)]), (2, [3, 0], [1], [])]
    assert [(e[0], e[1], e[2], e[3].tolist()) for e in edges] == [(1, 1, 2, [0])]


def test_kryo_varints():
    assert F._kvar(0, 32) == b"\x00" and F._kvar(127, 32) == b"\x7f" and F._kvar(128, 32) == b"\x80\x01"
    assert F._kvar(0xFFFFFFFF, 32) == b"\xff\xff\xff\xff\x0f"
    assert F._kzig(-1 & 0xFFFFFFFF, 32) == b"\x01" and F._kzig(1, 32) == b"\x02"
    assert len(F._kvar((1 << 64) - 1, 64)) == 9 and F._kvar((1 << 64) - 1, 64)[-1] == 0xFF
    assert F._kzig(1 << 63, 64) == b"\xff" * 9            # Long.MinValue -> all ones
    rng = np.random.default_rng(5)
    for bits in (32, 64):
        vals = [0, 1, (1 << bits) - 1, 1 << (bits - 1), (1 << (bits - 1)) - 1] + [int(x) >> s for x in rng.integers(0, 1 << 63, 200) for s in (0, 20, 45)]
        for v in vals:
            v &= (1 << bits) - 1
            for enc, dec in ((F._kvar, "var"), (F._kzig, "zig")):
                r = F._KryoReader(enc(v, bits))
                assert getattr(r, dec)(bits) == v and r.pos == len(r.b)


@pytest.mark.parametrize("k", [1, 5, 16, 31])
def test_kryo_graph_round_trip(k):
    """Every sequence class of DNASeq.newBuilder.result (<= 32 bases, <= 64, longer) and lengths around the word borders."""
    rng = np.random.default_rng(k)
    n_nodes = 7
    node_kmer = rng.integers(0, 1 << (2 * k), n_nodes, dtype=np.uint64)
    lens = [1, 31, 32, 33, 63, 64, 65, 66, 67, 68, 200, 4097]
    es = rng.integers(0, n_nodes, len(lens)).astype(np.uint32)
    ee = rng.integers(0, n_nodes, len(lens)).astype(np.uint32)
    seqs = [rng.integers(0, 4, ln).astype(np.uint8) for ln in lens]
    used = {}
    for e in range(len(lens)):       # a Map[Base, Long] per node: distinct first bases among a node's out-edges
        taken = used.setdefault(int(es[e]), set())
        free = [b for b in range(4) if b not in taken]
        if not free:
            es[e] = next(n for n in range(n_nodes) if len(used.setdefault(n, set())) < 4)
            taken = used[int(es[e])]
            free = [b for b in range(4) if b not in taken]
        seqs[e][0] = free[0]
        taken.add(free[0])
    data = F.write_kryo_graph(k, node_kmer, es, ee, seqs)
    nodes, edges = F.read_kryo_graph(data)
    k2, nk2, es2, ee2, seqs2 = F.kryo_graph_arrays(nodes, edges)
    assert k2 == k and np.array_equal(nk2, node_kmer) and np.array_equal(es2, es) and np.array_equal(ee2, ee)
    assert all(np.array_equal(a, b) for a, b in zip(seqs, seqs2))
    # the node records agree with the edge list (what Graph.addEdge maintains)
    for i, (nid, _, ins, outs) in enumerate(nodes):
        assert nid == i + 1
        assert sorted(ins) == sorted(e + 1 for e in range(len(lens)) if ee[e] == i)
        assert sorted(outs) == sorted((int(seqs[e][0]), e + 1) for e in range(len(lens)) if es[e] == i)
    # all three class names appear exactly once in the stream
    for cls in ("Long1DNASeq", "Long2DNASeq", "ArrayDNASeq"):
        assert data.count(cls.encode()[:-1]) == 1


def test_kryo_graph_rejects_what_the_reference_cannot_have_written():
    good = F.write_kryo_graph(3, [5, 9], [0], [1], [np.array([2, 1], np.uint8)])
    F.read_kryo_graph(good)
    with pytest.raises(ValueError, match="truncated"):
        F.read_kryo_graph(good[:-1])
    with pytest.raises(ValueError, match="after the graph"):
        F.read_kryo_graph(good + b"\x00")
    with pytest.raises(ValueError, match="reference marker"):
        F.read_kryo_graph(b"\x00" + good[1:])            # a null MapGraph
    with pytest.raises(ValueError, match="same first base"):
        F.write_kryo_graph(3, [5, 9], [0, 0], [1, 1], [np.array([2], np.uint8), np.array([2, 3], np.uint8)])
    with pytest.raises(ValueError, match="does not fit"):
        F.write_kryo_graph(3, [5, 9], [0], [2], [np.array([2], np.uint8)])
    nodes, edges = F.read_kryo_graph(good)
    with pytest.raises(ValueError, match="does not hold"):
        F.kryo_graph_arrays(nodes, [(1, 1, 7, edges[0][3])])
    assert F.kryo_graph_arrays([], [])[0] == 0


def test_cpp_kryo_codec_reproduces_the_python_bytes(tmp_path):
    """hostcpp/kryo_graph.hpp (MapGraph::write / Graph::apply of the C++ host mirror): decode + encode of the files formats.py
    writes gives the same bytes, for every sequence class; malformed files are refused."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "kryo_codec")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-O1", "-fsanitize=address,undefined", "-o", exe,
                        os.path.join(root, "tests", "emul", "kryo_codec_main.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rng = np.random.default_rng(77)
    for k in (1, 5, 16, 31):
        n_nodes = 6
        node_kmer = rng.integers(0, 1 << (2 * k), n_nodes, dtype=np.uint64)
        lens = [1, 2, 31, 32, 33, 63, 64, 65, 67, 68, 129, 1000]
        es = np.array([i % n_nodes for i in range(len(lens))], np.uint32)      # <= 2 out-edges per node: give them distinct first bases
        ee = rng.integers(0, n_nodes, len(lens)).astype(np.uint32)
        seqs = [rng.integers(0, 4, ln).astype(np.uint8) for ln in lens]
        for e in range(len(lens)):
            seqs[e][0] = e // n_nodes
        data = F.write_kryo_graph(k, node_kmer, es, ee, seqs)
        src, dst = tmp_path / ("g%d.kryo" % k), tmp_path / ("g%d.out" % k)
        src.write_bytes(data)
        r = subprocess.run([exe, str(src), str(dst)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout.strip() == "k=%d nodes=%d edges=%d" % (k, n_nodes, len(lens))
        assert dst.read_bytes() == data
        bad = tmp_path / "bad.kryo"
        bad.write_bytes(data[:-3])
        r = subprocess.run([exe, str(bad), str(dst)], capture_output=True, text=True)
        assert r.returncode == 1 and "truncated" in r.stderr
        bad.write_bytes(data + b"\x01")
        r = subprocess.run([exe, str(bad), str(dst)], capture_output=True, text=True)
        assert r.returncode == 1 and "after the graph" in r.stderr
    # the hand-spelled stream of test_kryo_graph_bytes_of_a_two_node_graph
    two = F.write_kryo_graph(2, [1 | 3 << 2, 3], [0], [1], [np.array([0], np.uint8)])
    src = tmp_path / "two.kryo"
    src.write_bytes(two)
    r = subprocess.run([exe, str(src), str(tmp_path / "two.out")], capture_output=True, text=True)
    assert r.returncode == 0 and (tmp_path / "two.out").read_bytes() == two
