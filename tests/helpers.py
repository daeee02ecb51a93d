"""Canonical forms used to compare the CUDA path with the oracle (ids are not reproducible, so graphs
are compared as sorted node lists and sorted edge multisets keyed by node k-mers)."""
import numpy as np

from genome_b200 import synth
from oracle import pyoracle


def oracle_counts(bin_bytes, n_reads, k, partitions=1, variant=291):
    m = pyoracle.OracleMap(k, partitions, variant)
    w = m.insert_reads(bin_bytes, n_reads)
    return m, w


def canon_oracle_graph(og):
    node_kmer, node_id, es, ee, off, bases = og.export()
    by_id = {int(i): int(km) for i, km in zip(node_id, node_kmer)}
    edges = sorted((by_id[int(es[i])], by_id[int(ee[i])], bases[int(off[i]):int(off[i + 1])].tobytes()) for i in range(es.size))
    return sorted(by_id.values()), edges


def canon_gpu_graph(g):
    node_kmer, es, ee, off, bases = g.export()
    nk = [int(x) for x in node_kmer]
    edges = sorted((nk[int(es[i])], nk[int(ee[i])], bases[int(off[i]):int(off[i + 1])].tobytes()) for i in range(es.size))
    return sorted(nk), edges


def assert_graph_equal(g, og):
    gn, ge = canon_gpu_graph(g)
    on, oe = canon_oracle_graph(og)
    assert len(gn) == len(on), "node count %d != oracle %d" % (len(gn), len(on))
    assert gn == on, "node k-mer sets differ"
    assert len(ge) == len(oe), "edge count %d != oracle %d" % (len(ge), len(oe))
    assert ge == oe, "edge multisets differ"


def revcomp_codes(codes):
    return (np.asarray(codes, np.uint8)[::-1] ^ 3).astype(np.uint8)


def rc_graph(canon, k):
    """The strand twin of a canonical graph: every node k-mer reverse-complemented, every edge reversed.
    Edge (u -> v, seq) of length L spells u + seq; its twin runs rc(v) -> rc(u) and appends rc of the first L bases
    of u + seq (the twin spells rc(u + seq))."""
    nodes, edges = canon
    rn = sorted(pyoracle.revcomp(x, k) for x in nodes)
    re = []
    for (u, v, seq) in edges:
        s = np.frombuffer(seq, np.uint8)
        ucodes = np.array([(u >> (2 * i)) & 3 for i in range(k)], np.uint8)
        full = np.concatenate([ucodes, s])
        twin = revcomp_codes(full[:s.size])
        re.append((pyoracle.revcomp(v, k), pyoracle.revcomp(u, k), twin.tobytes()))
    return rn, sorted(re)


def small_reads(genome_len, read_len, coverage, err, seed, ragged=False):
    genome = synth.random_genome(genome_len, seed)
    n_reads = max(2, (int(coverage * genome_len / read_len) // 2) * 2)
    reads = synth.sample_reads(genome, read_len, n_reads, err, seed + 1, insert=(read_len // 2, read_len))
    if not ragged:
        return synth.pack_fixed(reads), n_reads, genome
    rng = np.random.default_rng(seed + 2)
    lens = rng.integers(0, read_len + 1, size=n_reads)
    lst = [reads[i, :lens[i]] for i in range(n_reads)]
    return synth.pack_ragged(lst), n_reads, genome


def hash_ties(variant=291):
    """[(k, [k-mers x with hash(x) == hash(rc x), x != rc x])] from tests/golden/hash_ties.json (found by brute force,
    tests/golden/find_hash_ties.c); every test that uses them re-checks the property with the oracle's own arithmetic."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hash_ties.json")) as f:
        return [(e["k"], e["kmers"]) for e in json.load(f) if e["variant"] == variant]


def tie_reads(k, kmers, seed, strands="both", copies=3, flank=40):
    """Reads over a genome that embeds the hash-tie k-mers `kmers` between random flanks: every window of k + 12 bases of the
    forward strand (`strands` = "forward"), of the reverse strand ("reverse") or of both, `copies` times each.  A read of x stores
    rc(x) and a read of rc(x) stores x (FreqFilter.scala:32, tie => rcx), so "both" leaves BOTH orientations of a tie k-mer in
    the table and "forward" leaves only rc(x).  Returns (.bin bytes, number of reads)."""
    rng = np.random.default_rng(seed)
    parts = [rng.integers(0, 4, flank).astype(np.uint8)]
    for x in kmers:
        parts.append(np.array([(int(x) >> (2 * i)) & 3 for i in range(k)], np.uint8))
        parts.append(rng.integers(0, 4, flank).astype(np.uint8))
    genome = np.concatenate(parts)
    rl = k + 12
    strands_ = {"forward": [genome], "reverse": [revcomp_codes(genome)], "both": [genome, revcomp_codes(genome)]}[strands]
    reads = [s[i:i + rl] for s in strands_ for i in range(s.size - rl + 1) for _ in range(copies)]
    if len(reads) % 2:
        reads.append(reads[-1])
    return synth.pack_fixed(np.stack(reads)), len(reads)
