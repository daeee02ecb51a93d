"""Host-side file formats either side of the hot path; plain host I/O, no device work.

  convert2bin     S/scripts/Convert2bin.scala:15-92   FASTQ (two reads of n bases per sequence line) -> `.bin` stream
  read_bin        S/data/PairedEndData.scala:20-36    `.bin` stream -> list of base-code arrays
  write_contigs   S/scripts/GraphSimplifier.scala:338-347  edges -> the `contigs` text file
Paths relative to /root/reference, S/ = src/main/scala/ru/ifmo/genome/.  The Java-serialised PairedEndData header and the
Kryo graph file are not reproduced (their bytes are defined by JVM serialisers that are not pinned by the reference).
"""
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
        quality = next(it, None)             # read (and split) but unused: the quality filter is commented out (42-43)
        if quality is None:
            # in.readLine().splitAt(n) on null throws in the reference; a truncated last record is dropped here
            break
        s1, s2 = filtered(line[:n]), filtered(line[n:])
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
