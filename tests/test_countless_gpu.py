"""The single-pass bucket pass (per-(bucket, CTA) slabs instead of a count pass, csrc/partition.cu
part_scatter_kernel<..., SLABS>; the default for large batches) against the oracle, forced onto small inputs, and its
corner cases: slab overflow, the overflow list of the chunked host insert, ragged streams, super-k-mer records."""
import os

import numpy as np
import pytest

from genome_b200 import synth
from genome_b200.dnamap import ArrayDNAMap
from oracle import pyoracle
from tests import helpers as H

from genome_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.fixture
def countless():
    # small tables take the direct path otherwise; single_pass_min lets batches of a few 100 k windows in
    with capi.tuned(insert_path=2, single_pass=1, single_pass_min=1):
        yield


def check(b, n, k, cap=0):
    om, ow = H.oracle_counts(b, n, k)
    gm = ArrayDNAMap(k, cap)
    gw = gm.insert_reads(b, n)
    assert gw == ow == pyoracle.count_windows(b, n, k)
    assert gm.size == om.size()
    gk, gv = gm.export_sorted()
    ok, ov = om.export_sorted()
    assert np.array_equal(gk, ok) and np.array_equal(gv, ov)
    # a second batch into the same table (the table is no longer empty; sizes the slabs anew)
    gw2 = gm.insert_reads(b, n)
    assert gw2 == ow
    gk, gv = gm.export_sorted()
    assert np.array_equal(gk, ok) and np.array_equal(gv, 2 * ov)
    gm.delete_below(3)
    om.delete_below(2)   # counts doubled: v < 3 on 2v  <=>  v < 2
    assert gm.size == om.size()
    gm.close()


@pytest.mark.parametrize("k,read_len,ragged,err", [
    (31, 100, False, 0.01), (21, 100, True, 0.01), (25, 150, False, 0.005), (31, 255, True, 0.0), (15, 36, False, 0.02),
    (8, 50, True, 0.0), (4, 30, False, 0.0), (1, 10, False, 0.0), (16, 100, False, 0.0),
])
def test_countless_insert_matches_oracle(gpu, countless, k, read_len, ragged, err):
    """The parity cases of the table (tests/test_parity_gpu.py) through the slab bucket pass -- fixed-stride cases also
    through the chunked host insert."""
    b, n, _ = H.small_reads(20000, read_len, 12, err, seed=1000 + k, ragged=ragged)
    check(b, n, k)


def test_countless_overflow_path(gpu, countless):
    """Every read is the same sequence: all windows fall into a handful of buckets, so almost every key overflows its slab and
    is upserted by the bucket pass itself (random access).  Counts and the k-window total must still be exact."""
    k = 21
    one = synth.random_genome(100, 5)
    reads = np.tile(one, (60000, 1))
    check(synth.pack_fixed(reads), reads.shape[0], k)
    # and a mono-base stream: ONE key in total
    reads = np.zeros((40000, 80), np.uint8)
    check(synth.pack_fixed(reads), reads.shape[0], k)


def test_countless_overflow_list(gpu, countless):
    """The chunked host insert keeps overflowing keys in a list (it may not touch the table before the whole stream is
    verified).  4 % identical reads among random ones: their windows overflow the slabs of a few buckets but fit the list
    (1/16 of the keys); 30 %: the list fills up, the call falls back to the ordinary path.  Exact either way."""
    k = 25
    genome = synth.random_genome(400000, 9)
    for frac in (0.04, 0.3):
        reads = synth.sample_reads(genome, 100, 40000, 0.0, 10)
        n_same = int(frac * reads.shape[0])
        reads[:n_same] = reads[0]
        check(synth.pack_fixed(reads), reads.shape[0], k, cap=1 << 22)


def test_countless_ragged_stream_falls_back(gpu, countless):
    """A stream whose first records have equal lengths but that is ragged further on: the chunked path verifies on the device
    chunk by chunk and must leave the table untouched when it finds out."""
    k = 21
    genome = synth.random_genome(300000, 11)
    reads = synth.sample_reads(genome, 100, 30000, 0.0, 12)
    lst = [reads[i] for i in range(reads.shape[0])]
    lst[25000] = lst[25000][:60]   # one short record, in the last chunk
    lst += [reads[0], reads[1]]    # two spare records: the stream stays long enough for a fixed-stride attempt on 30 000 reads
    b = synth.pack_ragged(lst)
    n = reads.shape[0]
    assert n * 26 <= b.size
    check(b, n, k, cap=1 << 22)


def test_countless_at_size(gpu, countless):
    """1.5 M reads of a 5 Mbp genome with 1 % errors (the C2 shape): enough keys for real slabs (hundreds per (bucket, CTA)),
    compared through size-independent properties and against the counted bucket pass."""
    k = 31
    b, n, _ = synth.make_config("C2", scale=0.25)
    gm = ArrayDNAMap(k, int(b.size * 1.2))
    w = gm.insert_reads(b, n)
    gk, gv = gm.export_sorted()
    assert int(gv.astype(np.int64).sum()) == w and np.unique(gk).size == gk.size
    with capi.tuned(single_pass=0):
        ref = ArrayDNAMap(k, int(b.size * 1.2))
        assert ref.insert_reads(b, n) == w
    rk, rv = ref.export_sorted()
    assert np.array_equal(gk, rk) and np.array_equal(gv, rv)


@pytest.mark.parametrize("k,P", [(31, 8), (21, 2)])
def test_superkmer_records_insert_like_the_reads(gpu, k, P):
    """The receiving end of the super-k-mer wire format (csrc/superkmer.cuh), on one GPU: the reads are cut into 16-byte
    records on the CPU (the g++ build of the splitting code), the records go through gb_map_insert_records_device -- fixed
    stride, per-record lengths -- and the table must equal the oracle's table of the READS.  Also with the single-pass bucket
    pass, whose upsert takes the exact total from the device."""
    import torch
    from tests.test_sgraph_emul_cpu import load_emul
    from tests.test_superkmer_emul_cpu import split
    lib = load_emul()
    genome = synth.random_genome(300000, 21)
    reads = synth.sample_reads(genome, 100, 60000, 0.01, 22)
    b = synth.pack_fixed(reads)
    w, per_owner, recs = split(lib, b, reads.shape[0], 26, k, P)
    om, ow = H.oracle_counts(b, reads.shape[0], k)
    assert w == ow
    ok, ov = om.export_sorted()
    d = torch.zeros(recs.size + 16, dtype=torch.uint8, device="cuda")
    d[:recs.size].copy_(torch.from_numpy(np.ascontiguousarray(recs).reshape(-1)))
    for env in (dict(insert_path=1), dict(insert_path=2, single_pass=0), dict(insert_path=2, single_pass=1, single_pass_min=1)):
        with capi.tuned(**env):
            gm = ArrayDNAMap(k, 1 << 22)
            assert gm.insert_records_device(d.data_ptr(), recs.size, 16, recs.shape[0], 52) == w
            gk, gv = gm.export_sorted()
            assert np.array_equal(gk, ok) and np.array_equal(gv, ov), env
            gm.close()
    # a record longer than max_len is an argument error, the table stays as it was
    gm = ArrayDNAMap(k, 1 << 20)
    with pytest.raises(Exception):
        gm.insert_records_device(d.data_ptr(), recs.size, 16, recs.shape[0], 40)
    assert gm.size == 0
