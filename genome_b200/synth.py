"""Synthetic read sets in the reference's `.bin` layout.

Layout (S/data/PairedEndData.scala:24-32, S/scripts/Convert2bin.scala:35-38): per read one length byte then
(len+3)/4 bytes, base i in byte i/4 at bits 2(i%4); base code A0 G1 C2 T3 (S/dna/Base.scala:13-16); reads
come in pairs (read, mate).  Deterministic for a given numpy version; the committed fixtures under
tests/golden/ are files, not seeds.
"""
import numpy as np

BASES = "AGCT"  # index = 2-bit code (S/dna/Base.scala:13-18)
_CODE = {c: i for i, c in enumerate(BASES)}


def encode(s):
    return np.array([_CODE[c] for c in s], dtype=np.uint8)


def decode(codes):
    return "".join(BASES[int(c)] for c in codes)


def kmer_to_int(s):
    """Long1DNASeq.long of a string: base i at bits 2i (S/dna/DNASeq.scala:80-85)."""
    v = 0
    for i, c in enumerate(s):
        v |= _CODE[c] << (2 * i)
    return v


def int_to_kmer(v, k):
    return "".join(BASES[(int(v) >> (2 * i)) & 3] for i in range(k))


def random_genome(n, seed):
    return np.random.default_rng(seed).integers(0, 4, size=n, dtype=np.uint8)


def add_repeats(genome, frac, seed, lo=500, hi=5000):
    """Copy random source segments over random targets until `frac` of positions are covered (config C3)."""
    rng = np.random.default_rng(seed)
    g = genome.copy()
    n = g.size
    covered = 0
    while covered < frac * n:
        ln = int(rng.integers(lo, hi + 1))
        if ln >= n:
            break
        s = int(rng.integers(0, n - ln))
        t = int(rng.integers(0, n - ln))
        g[t:t + ln] = genome[s:s + ln]
        covered += ln
    return g


def sample_reads(genome, read_len, n_reads, err, seed, insert=(200, 500)):
    """n_reads (even) reads as (n_reads, read_len) codes.  Read 2i is a uniformly placed window on a random
    strand; read 2i+1 is its mate on the opposite strand `insert` bases downstream (clamped to the genome)."""
    rng = np.random.default_rng(seed)
    n_pairs = n_reads // 2
    G = genome.size
    L = read_len
    p1 = rng.integers(0, G - L + 1, size=n_pairs)
    ins = rng.integers(insert[0], insert[1] + 1, size=n_pairs)
    p2 = np.minimum(p1 + ins, G - L)
    flip = rng.random(n_pairs) < 0.5
    idx = np.arange(L)
    r1 = genome[p1[:, None] + idx]
    r2 = genome[p2[:, None] + idx]
    r2 = (r2[:, ::-1] ^ 3).astype(np.uint8)  # mate: reverse strand
    # strand flip of the whole fragment: swap roles and reverse-complement both
    a = np.where(flip[:, None], r2, r1)
    b = np.where(flip[:, None], r1, r2)
    reads = np.empty((2 * n_pairs, L), np.uint8)
    reads[0::2] = a
    reads[1::2] = b
    if err > 0:
        e = rng.random(reads.shape) < err
        sub = rng.integers(1, 4, size=reads.shape, dtype=np.uint8)
        reads = np.where(e, (reads + sub) & 3, reads).astype(np.uint8)
    return reads


def pack_fixed(reads):
    """(n, L) codes -> `.bin` bytes, every record 1 + ceil(L/4) bytes."""
    n, L = reads.shape
    bl = (L + 3) // 4
    padded = np.zeros((n, bl * 4), np.uint8)
    padded[:, :L] = reads
    q = padded.reshape(n, bl, 4)
    packed = (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)
    out = np.empty((n, 1 + bl), np.uint8)
    out[:, 0] = L
    out[:, 1:] = packed
    return out.reshape(-1)


def pack_ragged(read_list):
    """list of 1-D code arrays (len 0..255 each) -> `.bin` bytes."""
    parts = []
    for r in read_list:
        r = np.asarray(r, np.uint8)
        assert r.size <= 255
        parts.append(np.array([r.size], np.uint8))
        if r.size:
            parts.append(pack_fixed(r[None, :])[1:])
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def windows_fixed(n_reads, read_len, k):
    return n_reads * max(0, read_len - k + 1)


CONFIGS = {
    # name: genome bp, read len, coverage, error rate, repeat fraction  (BASELINE.json configs)
    "C1": dict(genome=4_600_000, read_len=100, coverage=30, err=0.0, repeats=0.0, seed=0x5EED0001),
    "C2": dict(genome=4_600_000, read_len=100, coverage=30, err=0.01, repeats=0.0, seed=0x5EED0002),
    "C3": dict(genome=100_000_000, read_len=150, coverage=50, err=0.0, repeats=0.05, seed=0x5EED0003),
    "C4": dict(genome=1_000_000_000, read_len=150, coverage=40, err=0.005, repeats=0.0, seed=0x5EED0004),
}


def make_config(name, scale=1.0, coverage=None, chunk_reads=1 << 20):
    """Generate a BASELINE.json config (optionally with the genome scaled down): returns (bin_bytes, n_reads, genome)."""
    c = dict(CONFIGS[name])
    G = max(1000, int(c["genome"] * scale))
    cov = coverage if coverage is not None else c["coverage"]
    genome = random_genome(G, c["seed"])
    if c["repeats"] > 0:
        genome = add_repeats(genome, c["repeats"], c["seed"] + 1)
    n_reads = (int(cov * G / c["read_len"]) // 2) * 2
    parts = []
    done = 0
    i = 0
    while done < n_reads:
        m = min(chunk_reads, n_reads - done)
        parts.append(pack_fixed(sample_reads(genome, c["read_len"], m, c["err"], c["seed"] + 100 + i)))
        done += m
        i += 1
    return np.concatenate(parts), n_reads, genome
