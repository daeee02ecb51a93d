"""Host-side mirrors of the two QC scripts that consume the hot path's output (paths relative to /root/reference,
S/ = src/main/scala/ru/ifmo/genome/):

  CheckGraph.startup   S/scripts/CheckGraph.scala:18-56   contig statistics + "is every k-mer of a FASTA on the graph"
  N50                  S/scripts/N50.scala:11-32          the N50 of a logged `length -> count` histogram line
  KmersCalculator      S/scripts/KmersCalculator.scala:16-28   distinct k-windows (as read, not canonical) of a one-record FASTA

CheckGraph's per-window `graphMap.contains(read)()` round trip (line 51) becomes one bulk `GraphPositionMap.contains` per
FASTA line.  Host logic only; the device work is behind MapGraph.graphMap().
"""
import re

import numpy as np

from . import synth

_CODE = {c: i for i, c in enumerate(synth.BASES)}


def contig_stats(edge_lengths):
    """CheckGraph.scala:37-41: edges longer than 200 bases, sorted; `N50` is the reference's `contigs(contigs.size / 2)`, i.e.
    the MEDIAN length (sic).  An empty list raises like the reference's `contigs(0)` / `contigs.last`."""
    contigs = sorted(int(x) for x in edge_lengths if int(x) > 200)
    if not contigs:
        raise IndexError("no contig longer than 200 bases")
    return dict(count=len(contigs), size=sum(contigs), n50=contigs[len(contigs) // 2], max=contigs[-1])


def line_windows(line, k):
    """The k-mers CheckGraph tests for one FASTA line (46-47): `line.sliding(k)`, keeping windows made of A/G/C/T only.
    Scala's `sliding` yields ONE short window when the line is shorter than k; such a window becomes a shorter DNASeq that no
    k-mer key equals, reported here as not-found through its own list.  Returns (keys u64[], window start of each key,
    number of short windows)."""
    n = len(line)
    if n == 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.int64), 0
    if n < k:
        return np.zeros(0, np.uint64), np.zeros(0, np.int64), int(all(c in _CODE for c in line))
    codes = np.array([_CODE.get(c, 4) for c in line], np.uint64)
    bad = np.concatenate([[0], np.cumsum(codes > 3)])
    starts = np.flatnonzero(bad[k:] - bad[:n - k + 1] == 0)
    keys = np.zeros(starts.size, np.uint64)
    for j in range(k):
        keys |= (codes[starts + j] & np.uint64(3)) << np.uint64(2 * j)
    return keys, starts.astype(np.int64), 0


def check_graph(graph, fasta_lines, log=None):
    """CheckGraph.startup on a MapGraph: returns (stats dict, list of (line length, k-mer string) not found on the graph)
    -- the reference logs "Not found <line.length> <read>" for each (52)."""
    k = graph.k
    _, _, _, off, _ = graph.export()
    stats = contig_stats(np.diff(off.astype(np.int64)))
    if log:
        for name in ("count", "size", "n50", "max"):
            log("Contigs %s %d" % ({"n50": "N50"}.get(name, name), stats[name]))
    gm = graph.graphMap()
    missing = []
    try:
        for line in fasta_lines:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                continue
            keys, starts, short = line_windows(line, k)
            if short:
                missing.append((len(line), line))
            if keys.size:
                found = gm.contains(keys)
                for s in starts[~found]:
                    missing.append((len(line), line[int(s):int(s) + k]))
    finally:
        gm.close()
    return stats, missing


_PAIR = re.compile(r"[^\d](\d+) -> (\d+),")


def n50(line):
    """N50.scala:14-31 on one logged histogram line (`Map(len -> count, ...)`).  Returns (the (length, count) pairs at which
    the running sum crosses half of the total — what the script prints first —, the sorted pairs, the number of entries
    with length >= 100).  Like the reference's regex, the last pair of the line is seen only if a comma follows it."""
    pairs = sorted((int(a), int(b)) for a, b in _PAIR.findall(line))
    total = sum(a * b for a, b in pairs)
    out, run = [], 0
    for a, b in pairs:
        if 2 * run < total and 2 * (run + a * b) >= total:
            out.append((a, b))
        run += a * b
    return out, pairs, sum(b for a, b in pairs if a >= 100)


def sequence_windows(fasta_lines, k):
    """KmersCalculator.scala:23-27: `getLines().drop(1)` (only the FIRST line is treated as a header), all characters
    concatenated, every character must be a base (`Base.fromChar` throws otherwise), then `seq.sliding(k)`.  Returns
    (sequence length, k-windows as u64 in reading order).  A sequence shorter than k yields Scala's single short window, which
    is one distinct sequence: reported as one window of value 0 with `short=True` in the third element."""
    it = iter(fasta_lines)
    next(it, None)
    text = "".join(line.rstrip("\r\n") for line in it)
    bad = [c for c in set(text) if c not in _CODE]
    if bad:
        raise ValueError("not a base: %r (Base.fromChar, S/dna/Base.scala:20)" % bad[0])
    n = len(text)
    if n == 0:
        return 0, np.zeros(0, np.uint64), False
    if n < k:
        return n, np.zeros(1, np.uint64), True
    codes = np.array([_CODE[c] for c in text], np.uint64)
    keys = np.zeros(n - k + 1, np.uint64)
    for j in range(k):
        keys |= codes[j:n - k + 1 + j] << np.uint64(2 * j)
    return n, keys, False


def kmers_calculator(fasta_lines, k=19):
    """KmersCalculator (k = 19 hard-coded there, line 21): (sequence length, number of distinct k-windows) with the windows
    counted on the device: a DNAMap of the windows AS READ (`update` without canonicalisation), then `size`."""
    from .dnamap import ArrayDNAMap
    n, keys, short = sequence_windows(fasta_lines, k)
    if short or keys.size == 0:
        return n, int(keys.size)
    with ArrayDNAMap(k, int(keys.size * 2)) as m:
        m.update_counts(keys)
        return n, m.size
