// superkmer.cuh -- TEST INFRASTRUCTURE: cuts reads into 16-byte super-k-mer records, the generator behind the device test of
// gb_map_insert_records_device (tests/test_countless_gpu.py) and tests/test_superkmer_emul_cpu.py.  A read is cut into RUNS of
// consecutive k-windows that share their minimizer owner (csrc/sgraph.cuh), at most SK_MAX_WINDOWS windows per run; each run is one
// record in the reference's `.bin` record layout (S/data/PairedEndData.scala:24-32: 1 length byte, then 4 bases per byte, first
// base in the low bits) at a fixed stride.  The records reproduce the reads' canonical k-window multiset exactly.
// History: round 2 ran this as the WIRE FORMAT of the sharded insert (1.5 bytes per k-window instead of 8: minimizer owners, count +
// emit pass, NCCL exchange, insert of the records).  Measured on 2 B200s (profiles/r2l_multi_2gpu_variants.log): 8.3 ms per 96.6 M
// k-windows per GPU against 2.8 ms for 8-byte keys stored into NVLink inboxes (2.4 ms single pass) -- the minimizer arithmetic and
// the second extraction cost far more than the wire saves.  Removed from the library; the splitter stays here as a record generator.
#pragma once
#include "../../genome_b200/csrc/sgraph.cuh"

namespace gb {
namespace sg {

constexpr int SK_RECORD_BYTES = 16;                      // 1 length byte + 13 bytes of bases + 2 bytes of padding
constexpr int SK_MAX_BASES = 52;                         // 13 bytes x 4 bases
SG_HD int sk_max_windows(int k) { return SK_MAX_BASES - k + 1 > 0 ? SK_MAX_BASES - k + 1 : 0; } // 22 at k = 31

// base i of a `.bin` record (rec points at its length byte)
SG_HD u32 rec_base(const u8 *rec, int i) { return (rec[1 + (i >> 2)] >> (2 * (i & 3))) & 3u; }

// bases [s, s + nb) of the source record as a 16-byte record: w0 = length | bases 0..27 << 8, w1 = bases 28..51
SG_HD void sk_pack(const u8 *src, int s, int nb, u64 *w0, u64 *w1)
{
    u64 a = (u64)nb, b = 0;
    for (int i = 0; i < nb; i++) {
        const u64 c = rec_base(src, s + i);
        if (i < 28) a |= c << (8 + 2 * i); else b |= c << (2 * (i - 28));
    }
    *w0 = a;
    *w1 = b;
}

// Cuts one read into runs.  emit(owner, first_base, n_bases) is called once per run, in read order.  Returns the number of
// k-windows of the read (0 when it is shorter than k: FreqFilter.scala:29 skips such reads).
// The owner of window j is owner_of_kmer of its k-mer: min over its w = k - m + 1 m-mers of the hash of the canonical m-mer,
// kept here as a ring of the last w hashes (one new m-mer per base, rolled forwards and as reverse complement).
template <class Emit>
SG_HD int sk_split_read(const u8 *rec, int k, int m, int P, Emit emit)
{
    const int len = rec[0];
    if (len < k) return 0;
    const int w = k - m + 1, max_windows = sk_max_windows(k);
    const u64 mm = (1ull << (2 * m)) - 1;
    u32 ring[32]; // w <= 31 (k <= 31, m >= 1)
    u64 fwd = 0, rc = 0;
    int run_start = 0, run_windows = 0;
    u32 run_owner = 0;
    for (int i = 0; i < len; i++) {
        const u64 c = rec_base(rec, i);
        fwd = (fwd >> 2) | (c << (2 * (m - 1)));
        rc = ((rc << 2) | (3 - c)) & mm;
        if (i < m - 1) continue;
        const int p = i - (m - 1); // the m-mer that starts at base p is complete
        ring[p % w] = mmer_hash(fwd, rc);
        if (p < w - 1) continue;
        const int j = p - (w - 1); // window j = bases [j, j + k) is complete: its m-mers are p - w + 1 .. p
        u32 h = H_NONE;
        for (int q = 0; q < w; q++) h = ring[q] < h ? ring[q] : h;
        const u32 owner = owner_from_hash(h, P);
        if (run_windows && (owner != run_owner || run_windows == max_windows)) {
            emit(run_owner, run_start, run_windows + k - 1);
            run_windows = 0;
        }
        if (!run_windows) { run_start = j; run_owner = owner; }
        run_windows++;
    }
    if (run_windows) emit(run_owner, run_start, run_windows + k - 1);
    return len - k + 1;
}

// one read per item: counts the records per owner (first pass) ...
// (record r starts at bin + offsets[r], or at bin + r * rec_bytes when offsets == nullptr)
struct SkCountOp {
    const u8 *bin; const u64 *offsets; u32 rec_bytes; int k, m, P; u64 *per_owner; u64 *windows;
    SG_HD void operator()(u64 r) const
    {
        u64 *po = per_owner;
        const u8 *rec = bin + (offsets ? offsets[r] : r * rec_bytes);
        const int nw = sk_split_read(rec, k, m, P, [po](u32 owner, int, int) { at_add64(po + owner, 1); });
        if (nw) at_add64(windows, (u64)nw);
    }
};
// ... and writes them: record number c of owner o goes to out[o] + 2 * c (cursor[o] starts at 0; out[o] holds per_owner[o] records)
struct SkEmitOp {
    const u8 *bin; const u64 *offsets; u32 rec_bytes; int k, m, P; u64 *cursor; u64 *const *out;
    SG_HD void operator()(u64 r) const
    {
        const u8 *rec = bin + (offsets ? offsets[r] : r * rec_bytes);
        u64 *cur = cursor;
        u64 *const *o = out;
        sk_split_read(rec, k, m, P, [rec, cur, o](u32 owner, int s, int nb) {
            u64 w0, w1;
            sk_pack(rec, s, nb, &w0, &w1);
            u64 *dst = o[owner] + 2 * at_add64(cur + owner, 1);
            dst[0] = w0;
            dst[1] = w1;
        });
    }
};

} // namespace sg
} // namespace gb
