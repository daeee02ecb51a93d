"""Host-side file formats either side of the hot path; plain host I/O, no device work.

  convert2bin     S/scripts/Convert2bin.scala:15-92   FASTQ (two reads of n bases per sequence line) -> `.bin` stream
  read_bin        S/data/PairedEndData.scala:20-36    `.bin` stream -> list of base-code arrays
  write_contigs   S/scripts/GraphSimplifier.scala:338-347  edges -> the `contigs` text file
  write_paired_end_header / read_paired_end_header
                  S/data/PairedEndData.scala:12-18,38-41   the header object file (java.io.ObjectOutputStream)
  write_kryo_graph / read_kryo_graph
                  S/data/graph/Graph.scala:232-261,384-390, Node.scala:14-37, Edge.scala:11   the Kryo `graph` file
Paths relative to /root/reference, S/ = src/main/scala/ru/ifmo/genome/.  The JDK's Object Serialization Stream Protocol
is a published grammar, and the class pins `@SerialVersionUID(1L)`; what no JVM here could confirm is the set of field names
scalac 2.9.1 emits for `class PairedEndData(val count: Long, val insert: Int, val bin: File)` (taken as `count`, `insert`,
`bin`), so the writer's bytes are "parity unpinned"; the reader is a generic parser of the grammar and does not depend on it.
"""
import struct

import numpy as np

from . import synth

_CODE = {c: i for i, c in enumerate(synth.BASES)}  # Base.fromChar: upper-case A G C T only (S/dna/Base.scala:20)


def convert2bin(lines, n=36, k=23):
    """Convert2bin.processRead over an iterable of FASTQ lines.  Every record is 4 lines; the sequence line holds BOTH reads
    (`line.splitAt(n)`), each read is cut at its first character that is not a base (`takeWhile`, lines 40-41) and
    written as 1 length byte + (len + 3) / 4 packed bytes (lines 35-38).  Returns (bin bytes, pairs, kmers, short_reads):
    `kmers` counts the k-windows of reads with len >= k (45-47), `short_reads` the pairs with a read shorter than k (67-69).
    A truncated record ends the input like the reference's null checks (52-58)."""
    it = iter(lines)
    reads, pairs, kmers, short = [], 0, 0, 0

    def filtered(s):
        nonlocal kmers
        codes = []
        for ch in s:
            if ch not in _CODE:
                break
            codes.append(_CODE[ch])
        if len(codes) >= k:
            kmers += len(codes) - k + 1
        return np.array(codes, np.uint8)

    while True:
        header = next(it, None)
        if header is None:
            break
        line = next(it, None)
        if line is None:
            break
        line = line.rstrip("\r\n")
        next(it, None)                       # '+' line
        quality = next(it, None)             # split like the bases; its VALUES are unused (the quality filter is commented out, 42-43)
        if quality is None:
            # in.readLine().splitAt(n) on null throws in the reference; a truncated last record is dropped here
            break
        quality = quality.rstrip("\r\n")
        q1, q2 = quality[:n], quality[n:]
        # `line1 zip quality1` (61-62): a read is cut to the length of its quality segment before takeWhile
        s1, s2 = filtered(line[:n][:len(q1)]), filtered(line[n:][:len(q2)])
        if len(s1) < k or len(s2) < k:
            short += 1
        reads.append(s1)
        reads.append(s2)
        pairs += 1
    return synth.pack_ragged(reads), pairs, kmers, short


def read_bin(bin_bytes, n_reads):
    """PairedEndData.getPairs.read: n_reads records -> list of uint8 code arrays; ValueError on a truncated stream."""
    b = np.ascontiguousarray(bin_bytes, dtype=np.uint8)
    out, pos = [], 0
    for r in range(n_reads):
        if pos >= b.size:
            raise ValueError("truncated .bin stream at read %d" % r)
        ln = int(b[pos])
        bl = (ln + 3) // 4
        if pos + 1 + bl > b.size:
            raise ValueError("truncated .bin stream at read %d" % r)
        packed = b[pos + 1:pos + 1 + bl]
        codes = np.empty(bl * 4, np.uint8)
        for j in range(4):
            codes[j::4] = (packed >> (2 * j)) & 3
        out.append(codes[:ln])
        pos += 1 + bl
    return out


def write_contigs(edge_seqs, path):
    """The `contigs` file of GraphSimplifier.scala:338-347: for edge i its bases on one line, THEN the line `>abacaba<i>`
    (the reference prints the sequence before its header)."""
    with open(path, "w") as f:
        for i, seq in enumerate(edge_seqs):
            f.write(synth.decode(seq) + "\n")
            f.write(">abacaba%d\n" % i)


# ---- java.io.ObjectOutputStream / ObjectInputStream (Object Serialization Stream Protocol, version 5)
_MAGIC = b"\xac\xed\x00\x05"
_TC_NULL, _TC_REFERENCE, _TC_CLASSDESC, _TC_OBJECT, _TC_STRING = 0x70, 0x71, 0x72, 0x73, 0x74
_TC_BLOCKDATA, _TC_ENDBLOCKDATA, _TC_BLOCKDATALONG, _TC_LONGSTRING = 0x77, 0x78, 0x7A, 0x7C
_SC_WRITE_METHOD, _SC_SERIALIZABLE = 0x01, 0x02
_HEADER_CLASS = "ru.ifmo.genome.data.PairedEndData"
_FILE_UID = 301077366599181567   # java.io.File.serialVersionUID
_PRIM = {"B": ">b", "C": ">H", "D": ">d", "F": ">f", "I": ">i", "J": ">q", "S": ">h", "Z": ">?"}


def _utf(s):
    b = s.encode("utf-8")   # modified UTF-8 differs only for NUL and non-BMP characters: not in file paths we write
    return struct.pack(">H", len(b)) + b


def write_paired_end_header(count, insert, bin_path, separator="/"):
    """`new PairedEndData(count, insert, bin).write(f)` (PairedEndData.scala:14-18; called at Convert2bin.scala:83): the bytes
    of `ObjectOutputStream.writeObject(this)`.  Class descriptor fields in ObjectStreamClass order (primitives by name, then
    objects by name); java.io.File carries its path string and, from its writeObject, the separator char as block data."""
    out = bytearray(_MAGIC)
    out += bytes([_TC_OBJECT, _TC_CLASSDESC]) + _utf(_HEADER_CLASS) + struct.pack(">q", 1) + bytes([_SC_SERIALIZABLE])
    out += struct.pack(">H", 3)
    out += b"J" + _utf("count") + b"I" + _utf("insert")
    out += b"L" + _utf("bin") + bytes([_TC_STRING]) + _utf("Ljava/io/File;")
    out += bytes([_TC_ENDBLOCKDATA, _TC_NULL])
    out += struct.pack(">qi", int(count), int(insert))
    out += bytes([_TC_OBJECT, _TC_CLASSDESC]) + _utf("java.io.File") + struct.pack(">q", _FILE_UID)
    out += bytes([_SC_SERIALIZABLE | _SC_WRITE_METHOD]) + struct.pack(">H", 1)
    out += b"L" + _utf("path") + bytes([_TC_STRING]) + _utf("Ljava/lang/String;")
    out += bytes([_TC_ENDBLOCKDATA, _TC_NULL])
    out += bytes([_TC_STRING]) + _utf(str(bin_path))
    out += bytes([_TC_BLOCKDATA, 2]) + struct.pack(">H", ord(separator)) + bytes([_TC_ENDBLOCKDATA])
    return bytes(out)


class _JavaStream:
    """The subset of the grammar a serialised PairedEndData can contain: objects of plain Serializable classes (with or
    without a writeObject annotation), strings, nulls and back references."""

    def __init__(self, data):
        self.b = bytes(data)
        self.pos = 0
        self.handles = []
        if self.b[:4] != _MAGIC:
            raise ValueError("not a Java object stream (bad magic)")
        self.pos = 4

    def take(self, n):
        if self.pos + n > len(self.b):
            raise ValueError("truncated Java object stream")
        v = self.b[self.pos:self.pos + n]
        self.pos += n
        return v

    def u8(self):
        return self.take(1)[0]

    def utf(self, long_form=False):
        n = struct.unpack(">Q", self.take(8))[0] if long_form else struct.unpack(">H", self.take(2))[0]
        return self.take(n).decode("utf-8", "replace")

    def class_desc(self):
        tc = self.u8()
        if tc == _TC_NULL:
            return None
        if tc == _TC_REFERENCE:
            return self.handles[struct.unpack(">I", self.take(4))[0] - 0x7E0000]
        if tc != _TC_CLASSDESC:
            raise ValueError("unsupported class descriptor tag 0x%02x" % tc)
        desc = {"name": self.utf(), "uid": struct.unpack(">q", self.take(8))[0]}
        self.handles.append(desc)
        desc["flags"] = self.u8()
        fields = []
        for _ in range(struct.unpack(">H", self.take(2))[0]):
            code = chr(self.u8())
            name = self.utf()
            if code in "L[":
                self.content()   # the field's type string (TC_STRING or a reference to an earlier one)
            fields.append((code, name))
        desc["fields"] = fields
        while self.content() is not _END:   # classAnnotation
            pass
        desc["super"] = self.class_desc()
        return desc

    def content(self):
        tc = self.u8()
        if tc == _TC_NULL:
            return None
        if tc == _TC_ENDBLOCKDATA:
            return _END
        if tc == _TC_REFERENCE:
            return self.handles[struct.unpack(">I", self.take(4))[0] - 0x7E0000]
        if tc in (_TC_STRING, _TC_LONGSTRING):
            v = self.utf(tc == _TC_LONGSTRING)
            self.handles.append(v)
            return v
        if tc == _TC_BLOCKDATA:
            return self.take(self.u8())
        if tc == _TC_BLOCKDATALONG:
            return self.take(struct.unpack(">I", self.take(4))[0])
        if tc == _TC_OBJECT:
            self.pos -= 1
            return self.object()
        raise ValueError("unsupported stream element 0x%02x" % tc)

    def object(self):
        if self.u8() != _TC_OBJECT:
            raise ValueError("expected an object")
        desc = self.class_desc()
        obj = {"__class__": desc["name"]}
        self.handles.append(obj)
        chain = []
        while desc is not None:
            chain.append(desc)
            desc = desc["super"]
        for d in reversed(chain):   # superclass data first
            if not d["flags"] & _SC_SERIALIZABLE:
                raise ValueError("class %s is not plain Serializable" % d["name"])
            for code, name in d["fields"]:
                if code in _PRIM:
                    fmt = _PRIM[code]
                    obj[name] = struct.unpack(fmt, self.take(struct.calcsize(fmt)))[0]
                else:
                    obj[name] = self.content()
            if d["flags"] & _SC_WRITE_METHOD:
                extra = []
                while True:
                    c = self.content()
                    if c is _END:
                        break
                    extra.append(c)
                obj.setdefault("__annotation__", []).extend(extra)
        return obj


_END = object()


def read_paired_end_header(data):
    """`PairedEndData(f)` (PairedEndData.scala:38-41): ObjectInputStream.readObject of the header file -> (count, insert,
    bin path).  Field lookup is by name, the descriptor order in the file is whatever the writing JVM chose."""
    obj = _JavaStream(data).object()
    if obj["__class__"] != _HEADER_CLASS:
        raise ValueError("serialised object is a %s, not a %s" % (obj["__class__"], _HEADER_CLASS))
    f = obj.get("bin")
    path = f.get("path") if isinstance(f, dict) else None
    if not isinstance(obj.get("count"), int) or not isinstance(obj.get("insert"), int) or not isinstance(path, str):
        raise ValueError("PairedEndData header lacks count / insert / bin")
    return obj["count"], obj["insert"], path


# ---- the Kryo `graph` file (MapGraph.write / Graph(file), S/data/graph/Graph.scala:232-261,384-390)
# The serialiser is a third-party dependency that is absent from /root/reference: com.esotericsoftware.kryo:kryo
# 2.14-SNAPSHOT (project/Build.scala:39).  What follows restates the wire rules of the Kryo 2.x line as published in its
# sources (Kryo.writeObject / writeClassAndObject / writeReferenceOrNull, DefaultClassResolver.writeName, io.Output,
# serializers.FieldSerializer, DefaultArraySerializers.ByteArraySerializer), applied to the reference's own call sites:
#   * `new Kryo()`: references on, registration not required -> every object that is not a primitive wrapper is preceded by
#     a reference marker (0 = null, 1 = first occurrence, n + 2 = back reference to object n); an unregistered class is
#     written as marker 1 (NAME + 2), a per-stream name id (varint) and, the first time only, its name as a Kryo string
#     (ASCII bytes, bit 7 set on the last one).
#   * MapGraph is KryoSerializable: marker, then MapGraph.write(kryo, out) = writeInt(nodes.size) + nodes, writeInt(edges.size)
#     + edges (Output.writeInt / writeLong without a flag: 4 / 8 bytes big-endian).
#   * Node has @DefaultSerializer(NodeSerializer) (Node.scala:14-27): id (8 bytes), writeClassAndObject(seq), in-edge ids,
#     (base byte, out-edge id) pairs.
#   * Edge has no serialiser of its own -> FieldSerializer: fields in name order (endId, id, seq, startId); a long field is a
#     zig-zag varlong (writeLong(v, false)), an int field a zig-zag varint, a byte field one byte, `seq: DNASeq` (not final)
#     is writeClassAndObject.
#   * DNASeq subclasses carry @DefaultSerializer(FieldSerializer) (DNASeq.scala:39,73,171): Long1DNASeq = len, long;
#     Long2DNASeq = len, long1, long2; ArrayDNASeq = data (byte[]: reference marker, varint length + 1, the bytes), length.
#     Which class holds a sequence is the builder's rule (DNASeq.scala:262-270): <= 32 bases Long1, <= 64 Long2, else Array.
# PARITY UNPINNED: no Kryo jar and no JVM here, and the reference holds no graph file; the reader below accepts what this writer
# produces and any of the three sequence classes in any order, and fails loudly on anything else.
_KRYO_SEQ = "ru.ifmo.genome.dna."
_L1, _L2, _AR = _KRYO_SEQ + "Long1DNASeq", _KRYO_SEQ + "Long2DNASeq", _KRYO_SEQ + "ArrayDNASeq"


def _kvar(v, bits):
    """Output.writeInt(v, true) / writeLong(v, true): 7 bits per byte, low groups first, bit 7 = more; the last byte of a
    full-width value carries the 8 remaining bits (5 bytes for an int, 9 for a long)."""
    v &= (1 << bits) - 1
    out = bytearray()
    for _ in range(4 if bits == 32 else 8):
        if v >> 7 == 0:
            out.append(v)
            return bytes(out)
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _kzig(v, bits):
    """writeInt(v, false) / writeLong(v, false): (v << 1) ^ (v >> bits-1) on the signed value, then the varint."""
    v &= (1 << bits) - 1
    signed = v - (1 << bits) if v >> (bits - 1) else v
    return _kvar(((signed << 1) ^ (signed >> (bits - 1))) & ((1 << bits) - 1), bits)


def _kstring(s):
    """Output.writeString for a non-empty ASCII string of 2..63 characters (a class name)."""
    b = bytearray(s.encode("ascii"))
    if not 1 < len(b) < 64:
        raise ValueError("class name outside Kryo's ASCII fast path: %r" % s)
    b[-1] |= 0x80
    return bytes(b)


class _KryoWriter:
    def __init__(self):
        self.out = bytearray()
        self.name_id = {}

    def class_name(self, name):
        self.out.append(1)                      # NAME + 2
        if name in self.name_id:
            self.out += _kvar(self.name_id[name], 32)
            return
        self.name_id[name] = len(self.name_id)
        self.out += _kvar(self.name_id[name], 32) + _kstring(name)

    def seq(self, codes):
        """writeClassAndObject(out, seq: DNASeq)"""
        codes = np.asarray(codes, np.uint8)
        n = int(codes.size)
        if n <= 64:
            words = [0, 0]
            for i, c in enumerate(codes.tolist()):
                words[i >> 5] |= c << (2 * (i & 31))
            if n <= 32:
                self.class_name(_L1)
                self.out += bytes([1, n]) + _kzig(words[0], 64)
            else:
                self.class_name(_L2)
                self.out += bytes([1, n]) + _kzig(words[0], 64) + _kzig(words[1], 64)
            return
        if n >= 1 << 31:
            raise ValueError("sequence longer than an Int")
        pad = np.zeros((n + 3) // 4 * 4, np.uint8)
        pad[:n] = codes
        data = (pad[0::4] | (pad[1::4] << 2) | (pad[2::4] << 4) | (pad[3::4] << 6)).astype(np.uint8)
        self.class_name(_AR)
        self.out += bytes([1, 1]) + _kvar(data.size + 1, 32) + data.tobytes() + _kzig(n, 32)


def write_kryo_graph(k, node_kmer, edge_start, edge_end, edge_seqs):
    """`graph.write(file)` (Graph.scala:232-248) for a graph given as arrays: node i = k-mer node_kmer[i] (base j at bits 2j),
    edge e runs edge_start[e] -> edge_end[e] (node indices) with the base codes edge_seqs[e].  Ids are index + 1 (the
    reference's AtomicLong generators start at 1; ids are not reproducible anyway).  A node's in / out lists follow
    edge order; an out-edge is keyed by the first base of its sequence (Graph.addEdge).  Returns the file's bytes."""
    n_nodes = len(node_kmer)
    ins = [[] for _ in range(n_nodes)]
    outs = [[] for _ in range(n_nodes)]
    for e in range(len(edge_start)):
        s, t = int(edge_start[e]), int(edge_end[e])
        if not (0 <= s < n_nodes and 0 <= t < n_nodes) or len(edge_seqs[e]) == 0:
            raise ValueError("edge %d does not fit the node list" % e)
        outs[s].append((int(edge_seqs[e][0]), e + 1))
        ins[t].append(e + 1)
    w = _KryoWriter()
    w.out.append(1)                              # writeObject(out, this): first occurrence of the MapGraph
    w.out += struct.pack(">i", n_nodes)
    for i in range(n_nodes):
        w.out.append(1)                          # writeObject(out, node)
        w.out += struct.pack(">q", i + 1)
        w.seq([(int(node_kmer[i]) >> (2 * j)) & 3 for j in range(k)])
        w.out += struct.pack(">i", len(ins[i]))
        for e in ins[i]:
            w.out += struct.pack(">q", e)
        if len({b for b, _ in outs[i]}) != len(outs[i]):
            raise ValueError("node %d has two out-edges with the same first base: not a Map[Base, Long]" % i)
        w.out += struct.pack(">i", len(outs[i]))
        for b, e in outs[i]:
            w.out += struct.pack(">bq", b, e)
    w.out += struct.pack(">i", len(edge_start))
    for e in range(len(edge_start)):
        w.out.append(1)                          # writeObject(out, edge): FieldSerializer, fields by name
        w.out += _kzig(int(edge_end[e]) + 1, 64) + _kzig(e + 1, 64)
        w.seq(edge_seqs[e])
        w.out += _kzig(int(edge_start[e]) + 1, 64)
    return bytes(w.out)


class _KryoReader:
    def __init__(self, data):
        self.b = bytes(data)
        self.pos = 0
        self.names = []

    def take(self, n):
        if self.pos + n > len(self.b):
            raise ValueError("truncated Kryo graph file at byte %d" % self.pos)
        v = self.b[self.pos:self.pos + n]
        self.pos += n
        return v

    def fixed(self, fmt):
        return struct.unpack(fmt, self.take(struct.calcsize(fmt)))[0]

    def var(self, bits):
        v, shift = 0, 0
        for i in range(5 if bits == 32 else 9):
            c = self.take(1)[0]
            if i == (4 if bits == 32 else 8):
                return (v | (c << shift)) & ((1 << bits) - 1)
            v |= (c & 0x7F) << shift
            if not c & 0x80:
                return v
            shift += 7

    def zig(self, bits):
        v = self.var(bits)
        return ((v >> 1) ^ -(v & 1)) & ((1 << bits) - 1)

    def marker(self, what):
        m = self.var(32)
        if m != 1:
            raise ValueError("%s at byte %d: reference marker %d (null or a back reference) is not something MapGraph.write emits"
                             % (what, self.pos - 1, m))

    def class_name(self):
        if self.var(32) != 1:
            raise ValueError("registered class id at byte %d: the reference registers no classes" % (self.pos - 1))
        i = self.var(32)
        if i < len(self.names):
            return self.names[i]
        if i != len(self.names):
            raise ValueError("class name id %d out of order at byte %d" % (i, self.pos))
        start = self.pos
        while not self.take(1)[0] & 0x80:
            pass
        raw = bytearray(self.b[start:self.pos])
        if len(raw) == 1:
            raise ValueError("UTF-8 class name at byte %d: not produced for these classes" % start)
        raw[-1] &= 0x7F
        self.names.append(raw.decode("ascii"))
        return self.names[-1]

    def seq(self):
        name = self.class_name()
        self.marker("sequence")
        if name in (_L1, _L2):
            n = self.fixed(">b")
            words = [self.zig(64)] + ([self.zig(64)] if name == _L2 else [])
            if not 0 <= n <= 32 * len(words):
                raise ValueError("%s of length %d" % (name, n))
            return np.array([(words[i >> 5] >> (2 * (i & 31))) & 3 for i in range(n)], np.uint8)
        if name == _AR:
            self.marker("ArrayDNASeq.data")
            ln = self.var(32) - 1
            if ln < 0:
                raise ValueError("null byte[] in an ArrayDNASeq")
            data = np.frombuffer(self.take(ln), np.uint8)
            n = self.zig(32)
            if n > 4 * ln:
                raise ValueError("ArrayDNASeq of length %d in %d bytes" % (n, ln))
            codes = np.empty(4 * ln, np.uint8)
            for j in range(4):
                codes[j::4] = (data >> (2 * j)) & 3
            return codes[:n].copy()
        raise ValueError("unexpected class %r in a graph file" % name)


def read_kryo_graph(data):
    """`Graph(file)` (Graph.scala:250-261,384-390) -> (nodes, edges): nodes = list of (id, seq codes, in-edge ids,
    [(base, out-edge id)]), edges = list of (id, start node id, end node id, seq codes), in file order."""
    r = _KryoReader(data)
    r.marker("MapGraph")
    nodes, edges = [], []
    for _ in range(r.fixed(">i")):
        r.marker("Node")
        nid = r.fixed(">q")
        seq = r.seq()
        ins = [r.fixed(">q") for _ in range(r.fixed(">i"))]
        outs = [(r.fixed(">b"), r.fixed(">q")) for _ in range(r.fixed(">i"))]
        nodes.append((nid, seq, ins, outs))
    for _ in range(r.fixed(">i")):
        r.marker("Edge")
        end = r.zig(64)
        eid = r.zig(64)
        seq = r.seq()
        start = r.zig(64)
        edges.append((eid, start, end, seq))
    if r.pos != len(r.b):
        raise ValueError("%d bytes after the graph" % (len(r.b) - r.pos))
    return nodes, edges


def kryo_graph_arrays(nodes, edges):
    """(nodes, edges) of read_kryo_graph -> (k, node_kmer u64[N], edge_start u32[E], edge_end u32[E], edge_seqs) with ids
    replaced by positions; checks what Graph.read relies on (unique ids, edges naming existing nodes, one k)."""
    idx = {nid: i for i, (nid, _, _, _) in enumerate(nodes)}
    if len(idx) != len(nodes) or len({e[0] for e in edges}) != len(edges):
        raise ValueError("duplicate ids in the graph file")
    ks = {int(seq.size) for _, seq, _, _ in nodes}
    if len(ks) > 1 or (ks and not 1 <= min(ks) <= 31):
        raise ValueError("node sequences of lengths %s" % sorted(ks))
    k = ks.pop() if ks else 0
    node_kmer = np.array([sum(int(c) << (2 * j) for j, c in enumerate(seq.tolist())) for _, seq, _, _ in nodes], np.uint64)
    try:
        es = np.array([idx[e[1]] for e in edges], np.uint32)
        ee = np.array([idx[e[2]] for e in edges], np.uint32)
    except KeyError as ex:
        raise ValueError("edge names node id %s, which the file does not hold" % ex)
    return k, node_kmer, es, ee, [e[3] for e in edges]
