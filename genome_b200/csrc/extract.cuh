// extract.cuh -- `.bin` tile staging and canonical k-window extraction shared by the local insert kernel
// (map.cu) and the owner-routing kernels of the sharded map (comm.cu); table upsert helpers.
// Replaces PairedEndData.getPairs.read (S/data/PairedEndData.scala:24-32), seq.sliding(k) and the canonical
// rule of FreqFilter.add (S/data/FreqFilter.scala:29-33); paths relative to /root/reference.
#pragma once
#include <vector>

#include "common.cuh"

namespace gb {

// One CTA stages a tile of TILE_READS consecutive records in shared memory with 128-bit loads, then its
// threads take work items of SEG consecutive k-windows of one read: the first window is cut out of the packed
// bytes with funnel shifts, the rest roll in one base at a time (forward and reverse-complement registers).
constexpr int TILE_READS = 128;
constexpr int SEG = 8;
constexpr int INSERT_THREADS = 256;
constexpr int MAX_REC_BYTES = 65; // 1 length byte + ceil(255 / 4)
constexpr int TILE_SMEM_WORDS = (TILE_READS * MAX_REC_BYTES + 16 + 16) / 4 + 4;

struct __align__(16) ReadTile {
    unsigned int bytes[TILE_SMEM_WORDS];
    unsigned int start[TILE_READS];   // bit position of base 0 of the read inside `bytes`
    unsigned int prefix[TILE_READS + 1]; // exclusive prefix of work items per read
    unsigned short nwin[TILE_READS];
    unsigned int ipr_inv; // every read of the tile has the same number (>= 2) of work items: floor(2^32 / items) + 1, else 0
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long extract_bits(const unsigned int *s, unsigned int bitpos)
{
    unsigned int w = bitpos >> 5, sh = bitpos & 31;
    unsigned int a = s[w], b = s[w + 1], c = s[w + 2];
    unsigned int lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
    return ((unsigned long long)hi << 32) | lo;
}

// stages tile number `tile_idx` of the batch; returns the number of reads in it (<= 0: nothing to do, uniform over the CTA).
// Ends with a __syncthreads(): tile.* is readable by every thread on return.
template <bool FIXED>
__device__ __forceinline__ int stage_tile(ReadTile &tile, const uint8_t *__restrict__ bin, unsigned long long n_bytes,
                                          const unsigned long long *__restrict__ offsets, unsigned int rec_bytes,
                                          long long read0, long long n_reads, int k, long long tile_idx)
{
    const int tid = threadIdx.x;
    const long long r0 = read0 + tile_idx * TILE_READS;
    const int nr = (int)min((long long)TILE_READS, read0 + n_reads - r0);
    if (nr <= 0) return nr;

    const unsigned long long byte0 = FIXED ? (unsigned long long)r0 * rec_bytes : offsets[r0];
    const unsigned long long byte1 = FIXED ? (unsigned long long)(r0 + nr) * rec_bytes : offsets[r0 + nr];
    const unsigned long long base = byte0 & ~15ull;

    // [base, byte1): coalesced 16-byte loads, streaming (read once)
    for (unsigned long long v = tid; base + 16 * v < byte1; v += INSERT_THREADS) {
        unsigned long long g = base + 16 * v;
        uint4 d;
        if (g + 16 <= n_bytes) {
            d = __ldcs(reinterpret_cast<const uint4 *>(bin + g));
        } else {
            unsigned int w[4] = { 0, 0, 0, 0 };
            for (int i = 0; i < 16 && g + i < n_bytes; i++) w[i >> 2] |= (unsigned int)bin[g + i] << (8 * (i & 3));
            d = make_uint4(w[0], w[1], w[2], w[3]);
        }
        reinterpret_cast<uint4 *>(tile.bytes)[v] = d;
    }
    __syncthreads();

    if (tid < TILE_READS) {
        unsigned int nwin = 0, start = 0;
        if (tid < nr) {
            unsigned long long off = FIXED ? (unsigned long long)(r0 + tid) * rec_bytes : offsets[r0 + tid];
            unsigned int rel = (unsigned int)(off - base);
            unsigned int len = (tile.bytes[rel >> 2] >> (8 * (rel & 3))) & 0xFF;
            nwin = len >= (unsigned int)k ? len - k + 1 : 0; // reads shorter than k are skipped (FreqFilter.scala:29)
            start = (rel + 1) * 8;
        }
        tile.start[tid] = start;
        tile.nwin[tid] = (unsigned short)nwin;
    }
    __syncthreads();
    if (tid < 32) {
        // exclusive scan of items per read, 4 reads per lane
        unsigned int c[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            c[j] = (tile.nwin[tid * 4 + j] + SEG - 1) / SEG;
            sum += c[j];
        }
        unsigned int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (tid >= d) incl += t;
        }
        unsigned int run = incl - sum;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            tile.prefix[tid * 4 + j] = run;
            run += c[j];
        }
        if (tid == 31) tile.prefix[TILE_READS] = run;
        // equal-length reads (the usual stream): item -> read by one multiply instead of a search (item_keys)
        const unsigned int c0 = __shfl_sync(0xFFFFFFFFu, c[0], 0);
        const bool same = __all_sync(0xFFFFFFFFu, c[0] == c0 && c[1] == c0 && c[2] == c0 && c[3] == c0);
        if (tid == 0) tile.ipr_inv = same && c0 >= 2 ? (unsigned int)((1ull << 32) / c0) + 1u : 0u;
    }
    __syncthreads();
    return nr;
}

// canonical k-mers of work item `item`: key[0..cnt) valid, the rest are harmless garbage; returns cnt
template <bool V210>
__device__ __forceinline__ int item_keys(const ReadTile &tile, unsigned int item, int k, unsigned long long key[SEG])
{
    int lo = 0;
    if (tile.ipr_inv) {
        lo = (int)__umulhi(item, tile.ipr_inv); // exact for item < 2^32 / items per read
    } else {
        int hi = TILE_READS;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (tile.prefix[mid] <= item) lo = mid; else hi = mid;
        }
    }
    const unsigned long long kmask = (1ull << (2 * k)) - 1;
    const unsigned int p = (item - tile.prefix[lo]) * SEG;
    const int cnt = min((int)SEG, (int)tile.nwin[lo] - (int)p);
    const unsigned int bit0 = tile.start[lo] + 2 * p;
    unsigned long long fwd = extract_bits(tile.bytes, bit0) & kmask;
    unsigned int nxt = (unsigned int)extract_bits(tile.bytes, bit0 + 2 * k);
    unsigned long long rc = revcomp(fwd, k);
#pragma unroll
    for (int j = 0; j < SEG; j++) {
        key[j] = canonical<V210>(fwd, rc);
        unsigned int b = nxt & 3;
        nxt >>= 2;
        fwd = (fwd >> 2) | ((unsigned long long)b << (2 * (k - 1)));
        rc = ((rc << 2) | (3 - b)) & kmask;
    }
    return cnt;
}

// update(key, add, _ + add) (S/ds/ArrayDNAMap.scala:129-150) starting at slot idx whose key was already loaded
// into `cur`.  The caller guarantees the table never fills (map_budget), so the probe always terminates.
// Returns 1 when the key was new.
__device__ __forceinline__ int upsert_add(const Table &table, unsigned long long idx, unsigned long long cur, unsigned long long key, int add)
{
    for (;;) {
        if (cur == key) {
            red_add_s32(table.count + idx, add);
            return 0;
        }
        if (cur == EMPTY_KEY) {
            unsigned long long old = atomicCAS(table.key + idx, EMPTY_KEY, key);
            if (old == EMPTY_KEY) { // claimed: the count word of a free slot already holds 1 (init_table_kernel)
                if (add != 1) red_add_s32(table.count + idx, add - 1);
                return 1;
            }
            if (old == key) {
                red_add_s32(table.count + idx, add);
                return 0;
            }
        }
        idx = next_slot(idx, table.cap);
        cur = load_key(table, idx);
    }
}
#endif

// host-side entry points shared between map.cu and comm.cu
int map_budget(Map *m, int64_t incoming, int64_t *budget);
int map_read_counters(Map *m, unsigned long long out[4]);
int map_zero_counters(Map *m);
int map_rebuild(Map *m, unsigned long long new_cap, bool filter, int min_count);
int map_export_device(Map *m, unsigned long long *d_keys, int *d_vals);
// FreqFilter.add over records at a fixed stride with their own length bytes (<= max_len), already verified; map.cu
int map_insert_records(Map *m, const uint8_t *d_bin, size_t n_bytes, unsigned int rec_bytes, int64_t n_records, unsigned int max_len,
                       int64_t *n_windows);
int map_verify_fixed(Map *m, const uint8_t *d_bin, unsigned int rec, unsigned int len0, int64_t n_reads, unsigned long long *bad);
int map_launch_update_counts(Map *m, const unsigned long long *d_keys, int64_t n, cudaStream_t st);
int map_launch_update_set(Map *m, const unsigned long long *d_keys, const int *d_vals, int64_t n, cudaStream_t st, bool set_vid);
int scan_records(const uint8_t *bin, size_t n_bytes, int64_t n_reads, int k, std::vector<unsigned long long> &off,
                 std::vector<int64_t> &winp);

} // namespace gb
