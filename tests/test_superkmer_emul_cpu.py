"""The super-k-mer splitter (tests/emul/superkmer.cuh, test infrastructure: the record generator of the device test of
gb_map_insert_records_device) against the oracle's own extraction: the records carry exactly the reads' canonical k-window multiset and
every window of a record has the record's minimizer owner (csrc/sgraph.cuh: the ownership rule of the sharded graph build)."""
import ctypes as C
import os

import numpy as np
import pytest

from genome_b200 import synth
from oracle import pyoracle
from tests import helpers as H
from tests.test_sgraph_emul_cpu import emul, ptr  # noqa: F401  (the fixture builds tests/emul/sgraph_emul.cpp)


def split(emul, b, n_reads, rec_bytes, k, P):
    emul.emul_superkmers.restype = C.c_uint64
    per_owner = np.zeros(P, np.uint64)
    w = emul.emul_superkmers(ptr(b), rec_bytes, C.c_uint64(n_reads), k, P, ptr(per_owner), None)
    total = int(per_owner.sum())
    recs = np.zeros(max(total, 1) * 2, np.uint64)
    w2 = emul.emul_superkmers(ptr(b), rec_bytes, C.c_uint64(n_reads), k, P, ptr(per_owner), ptr(recs))
    assert w2 == w
    return int(w), per_owner.astype(np.int64), recs[:2 * total].view(np.uint8).reshape(total, 16)


def to_ragged_bin(recs):
    """16-byte fixed-stride records -> the contiguous `.bin` stream the oracle reads (padding stripped)"""
    out = []
    for r in recs:
        ln = int(r[0])
        out.append(r[:1 + (ln + 3) // 4])
    return np.concatenate(out) if out else np.zeros(0, np.uint8)


@pytest.mark.parametrize("k,read_len,err", [(31, 100, 0.01), (31, 150, 0.0), (21, 100, 0.02), (25, 36, 0.0), (15, 60, 0.01), (31, 31, 0.0), (8, 40, 0.0)])
@pytest.mark.parametrize("P", [1, 2, 8])
def test_records_carry_exactly_the_reads_windows(emul, k, read_len, err, P):
    genome = synth.random_genome(20000, 300 + k)
    reads = synth.sample_reads(genome, read_len, 3000, err, 301 + k)
    b = synth.pack_fixed(reads)
    rec_bytes = 1 + (read_len + 3) // 4
    w, per_owner, recs = split(emul, b, reads.shape[0], rec_bytes, k, P)
    assert w == reads.shape[0] * (read_len - k + 1) == pyoracle.count_windows(b, reads.shape[0], k)
    lens = recs[:, 0].astype(int)
    assert lens.min() >= k and lens.max() <= 52
    assert int((lens - k + 1).sum()) == w                      # every window in exactly one record
    assert np.all(recs[:, 14:] == 0)                           # padding
    # the same canonical k-mer multiset through the oracle's extraction (FreqFilter.add)
    om_reads, _ = H.oracle_counts(b, reads.shape[0], k)
    rb = to_ragged_bin(recs)
    om_recs, w_recs = H.oracle_counts(rb, recs.shape[0], k)
    assert w_recs == w
    k1, v1 = om_reads.export_sorted()
    k2, v2 = om_recs.export_sorted()
    assert np.array_equal(k1, k2) and np.array_equal(v1, v2)
    # every window of owner o's records has owner o
    full, incr = np.zeros(9, np.uint32), np.zeros(8, np.uint32)
    at = 0
    rng = np.random.default_rng(P)
    for o in range(P):
        block = recs[at:at + per_owner[o]]
        at += per_owner[o]
        for r in block[rng.permutation(block.shape[0])[:60]]:
            ln = int(r[0])
            codes = np.array([(int(r[1 + i // 4]) >> (2 * (i % 4))) & 3 for i in range(ln)], np.uint64)
            for j in range(ln - k + 1):
                x = 0
                for t in range(k):
                    x |= int(codes[j + t]) << (2 * t)
                emul.emul_owners(k, P, C.c_uint64(x), ptr(full), ptr(incr))
                assert int(full[0]) == o
    if P == 8 and k == 31 and read_len >= 100:
        # the point of it: ~11 windows per record on random sequence => ~1.5 bytes per window instead of 8
        assert 16.0 * recs.shape[0] / w < 2.6, 16.0 * recs.shape[0] / w


def test_short_and_boundary_reads(emul):
    k = 21
    reads = [synth.random_genome(n, 7 + n) for n in (0, 5, 20, 21, 22, 72, 73, 74, 100)]   # 72 = 52 + 20: two full records exactly
    width = 100
    arr = np.zeros((len(reads), width), np.uint8)
    b = np.zeros(len(reads) * 26, np.uint8)
    for i, r in enumerate(reads):   # fixed stride 26 with per-record length bytes
        b[26 * i] = r.size
        for j in range(r.size):
            b[26 * i + 1 + j // 4] |= int(r[j]) << (2 * (j % 4))
    w, per_owner, recs = split(emul, b, len(reads), 26, k, 1)
    assert w == sum(max(0, r.size - k + 1) for r in reads)
    lens = sorted(recs[:, 0].astype(int).tolist())
    # one owner: runs break only at 32 windows (52 bases): 21 -> [21]; 22 -> [22]; 72 -> [52, 40]; 73 -> [52, 41]; 74 -> [52, 42]; 100 -> [52, 52, 36]
    assert lens == sorted([21, 22, 52, 40, 52, 41, 52, 42, 52, 52, 36])


def test_library_minimizer_owner_matches_the_emulation(emul):
    """gb_owner_of_minimizer (host arithmetic inside libgenome_b200.so, what gb_pmap_lookup routes by under
    the sharded graph build) against the g++ build of the same header; and the reverse complement shares the owner."""
    from genome_b200 import capi
    rng = np.random.default_rng(11)
    full, incr = np.zeros(9, np.uint32), np.zeros(8, np.uint32)
    for k in (8, 21, 31):
        keys = rng.integers(0, 1 << (2 * k), 400, dtype=np.uint64)
        rc = np.array([pyoracle.revcomp(int(x), k) for x in keys], np.uint64)
        for P in (1, 3, 8):
            own = np.zeros(keys.size, np.int32)
            own_rc = np.zeros(keys.size, np.int32)
            capi.check(capi.lib().gb_owner_of_minimizer(ptr(keys), keys.size, k, P, ptr(own)))
            capi.check(capi.lib().gb_owner_of_minimizer(ptr(rc), rc.size, k, P, ptr(own_rc)))
            assert np.array_equal(own, own_rc) and own.min() >= 0 and own.max() < P
            for x, o in zip(keys[:80].tolist(), own[:80].tolist()):
                emul.emul_owners(k, P, C.c_uint64(x), ptr(full), ptr(incr))
                assert int(full[0]) == o
