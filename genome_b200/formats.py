"""Host-side file formats either side of the hot path; plain host I/O, no device work.

  convert2bin     S/scripts/Convert2bin.scala:15-92   FASTQ (two reads of n bases per sequence line) -> `.bin` stream
  read_bin        S/data/PairedEndData.scala:20-36    `.bin` stream -> list of base-code arrays
  write_contigs   S/scripts/GraphSimplifier.scala:338-347  edges -> the `contigs` text file
  write_paired_end_header / read_paired_end_header
                  S/data/PairedEndData.scala:12-18,38-41   the header object file (java.io.ObjectOutputStream)
Paths relative to /root/reference, S/ = src/main/scala/ru/ifmo/genome/.  The Kryo graph file is not reproduced (its bytes are
defined by a SNAPSHOT serialiser the reference does not pin).  The header IS: the JDK's Object Serialization Stream Protocol
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
