// partition.cu -- L2-blocked k-mer insertion.
//
// A k-mer table far larger than the 126 MB L2 turns FreqFilter.add (S/data/FreqFilter.scala:28-36, paths relative
// to /root/reference) into one random DRAM access per k-mer instance, and every miss moves a whole 128-byte line
// (ncu, profiles/insert_r1a_ncu_full_summary.csv: 138 B of DRAM reads per k-mer).  Instead:
//   1. part_count / part_scatter stream the read batch twice and write every canonical k-mer (8 B) into the bucket
//      of the table SLICE it hashes to -- sequential traffic;
//   2. insert_key_chunks walks the buckets in slice order, so the CTAs that are resident at any moment all update
//      the same <= 64 MiB slice of the table, which lives in L2; DRAM sees each slice once in and once out.
// The same buckets, with an owner-shard prefix, are what the sharded map sends over NVLink (comm.cu).
#include <math.h>

#include <algorithm>

#include "partition.cuh"

#include "extract.cuh"
#include "scan.cuh"

namespace gb {

__device__ __forceinline__ unsigned int bucket_of(unsigned long long h, unsigned int owners, int lp_bits)
{
    unsigned int slice = lp_bits ? (unsigned int)(h >> (64 - lp_bits)) : 0u;
    return (owner_of(h, owners) << lp_bits) | slice;
}

constexpr int SPREAD = 256; // new-key tallies are spread over this many counters: no same-address atomic storm

// Persistent CTAs: CTA c takes tiles c, c + grid, ... in BOTH passes, so its per-bucket counts of pass 1 are exactly
// the room it needs in pass 2: no global atomics, no shared-memory atomics with a return value, deterministic layout.
constexpr int WARPS = INSERT_THREADS / 32;
constexpr int STAGE_MAX_BUCKETS = 128;           // the staged pass keeps per-bucket bookkeeping in shared memory
constexpr int ROUND_KEYS = INSERT_THREADS * SEG; // keys a CTA stages per round

// keys of round `blk` of a KeySource for this thread: positions blk*ROUND_KEYS + j*256 + tid, j < SEG (coalesced);
// returns how many of them exist (they are the first `cnt`: positions ascend with j)
__device__ __forceinline__ int load_round_keys(const KeySource &ks, unsigned long long blk, unsigned long long key[SEG])
{
    const unsigned long long v0 = blk * ROUND_KEYS;
    int c = 0;
    if (ks.n_chunks > 1) { // last chunk with vstart <= v0
        int lo = 0, hi = ks.n_chunks;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (ks.vstart[mid] <= v0) lo = mid; else hi = mid;
        }
        c = lo;
    }
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < SEG; j++) {
        const unsigned long long v = v0 + (unsigned long long)j * INSERT_THREADS + threadIdx.x;
        if (v < ks.n_total) {
            while (v >= ks.vstart[c + 1]) c++;
            key[j] = __ldcs(ks.keys + ks.off[c] + (v - ks.vstart[c]));
            cnt = j + 1;
        }
    }
    return cnt;
}

template <bool FIXED, bool V210, bool SRC_KEYS>
__global__ void __launch_bounds__(INSERT_THREADS)
part_count_kernel(ReadBatch rb, KeySource ks, int k, unsigned int owners, int lp_bits, unsigned int nb, unsigned int *cta_hist)
{
    __shared__ ReadTile tile;
    __shared__ unsigned int s_hist[MAX_BUCKETS];
    const int tid = threadIdx.x;
    for (unsigned int b = tid; b < MAX_BUCKETS; b += INSERT_THREADS) s_hist[b] = 0;
    __syncthreads();
    if (SRC_KEYS) {
        const unsigned long long n_blk = (ks.n_total + ROUND_KEYS - 1) / ROUND_KEYS;
        for (unsigned long long blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
            unsigned long long key[SEG];
            const int cnt = load_round_keys(ks, blk, key);
#pragma unroll
            for (int j = 0; j < SEG; j++)
                if (j < cnt) atomicAdd(&s_hist[bucket_of(mix64(key[j]), owners, lp_bits)], 1u);
        }
    } else {
        const long long n_tiles = (rb.n_reads + TILE_READS - 1) / TILE_READS;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            stage_tile<FIXED>(tile, rb.bin, rb.n_bytes, rb.offsets, rb.rec_bytes, rb.read0, rb.n_reads, k, t);
            const unsigned int total_items = tile.prefix[TILE_READS];
            for (unsigned int item = tid; item < total_items; item += INSERT_THREADS) {
                unsigned long long key[SEG];
                const int cnt = item_keys<V210>(tile, item, k, key);
#pragma unroll
                for (int j = 0; j < SEG; j++)
                    if (j < cnt) atomicAdd(&s_hist[bucket_of(mix64(key[j]), owners, lp_bits)], 1u); // no return value: a RED
            }
            __syncthreads(); // the tile is overwritten by the next stage_tile
        }
    }
    __syncthreads();
    for (unsigned int b = tid; b < nb; b += INSERT_THREADS) cta_hist[(size_t)blockIdx.x * nb + b] = s_hist[b];
}

// one CTA per bucket: turn the column of per-CTA counts into exclusive offsets inside the bucket
__global__ void __launch_bounds__(256)
part_offsets_kernel(unsigned int *cta_hist, int rows, unsigned int nb, unsigned long long *bucket_total)
{
    const unsigned int b = blockIdx.x;
    const int per = (rows + 255) / 256;
    const int c0 = min(rows, (int)threadIdx.x * per), c1 = min(rows, c0 + per);
    unsigned int sum = 0;
    for (int c = c0; c < c1; c++) sum += cta_hist[(size_t)c * nb + b];
    unsigned long long run = block_alloc(sum, nullptr); // exclusive prefix over the threads of this CTA
    if (threadIdx.x == 255) bucket_total[b] = run + sum;
    for (int c = c0; c < c1; c++) {
        unsigned int v = cta_hist[(size_t)c * nb + b];
        cta_hist[(size_t)c * nb + b] = (unsigned int)run; // a bucket of one batch holds < 2^32 keys
        run += v;
    }
}

__global__ void __launch_bounds__(MAX_BUCKETS)
part_bases_kernel(const unsigned long long *bucket_total, unsigned int nb, unsigned long long *bucket_base)
{
    __shared__ unsigned long long s_tot[MAX_BUCKETS];
    s_tot[threadIdx.x] = threadIdx.x < nb ? bucket_total[threadIdx.x] : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (unsigned int i = 0; i < nb; i++) {
            bucket_base[i] = acc;
            acc += s_tot[i];
        }
        bucket_base[nb] = acc;
    }
}

// Pass 2, a CTA-level multisplit.  A round = 256 items x SEG keys.  Ranks inside a warp come from match.any (lanes
// with the same bucket) plus a warp-private per-bucket counter; a prefix over the CTA's warps and a scan over the
// buckets give every key its place in a staging area sorted by bucket, and the round is flushed with consecutive
// threads storing consecutive keys: runs of ROUND_KEYS / nb keys (512 B at 32 buckets) per bucket instead of 32
// scattered 8-byte stores per warp instruction.  That is what fills NVLink write packets when PEER.
// PEER: the position of a key is (owner, index inside the owner's segment) packed in 32 bits and the store goes to
// the owner's inbox through its peer mapping -- the all-to-all happens inside this kernel, store by store.
// (The single-GPU insert does not use this kernel: bucket_slabs_kernel below needs no count pass.)

// Destination of the single-pass bucket pass (bucket_slabs_kernel; single-GPU path only).  CTA c owns, for every bucket b, the slab
// [(b * grid + c) * slab, + slab) of `out`, sized for its expected share plus 8 sigma; what it wrote goes to count[b * grid + c],
// and the upsert walks the slabs in that (bucket-major = slice) order.  A key that does not fit its slab is upserted by the bucket
// pass itself, with the random access of the direct path: always correct, and only pathological inputs (one bucket far above its
// share within one CTA's tiles) ever take it.
struct SlabOut {
    unsigned int slab = 0;              // keys per (bucket, CTA) slab
    unsigned int *count = nullptr;      // [nb][grid]
    Table table;                        // overflow path
    unsigned long long *spread = nullptr;   // new-key tallies (fold_new_keys_kernel)
    unsigned long long *overflowed = nullptr; // keys that took the overflow path (they count as k-windows too)
    // LIST mode (ovf != nullptr; the chunked host insert): the bucket pass must not touch the table (the stream is not verified
    // yet), so a key that does not fit its slab is appended to ovf[0 .. ovf_cap) instead (cursor = *overflowed); if that region
    // fills up too, *failed is set and the caller falls back to the counted passes.  The slab cursors persist across launches
    // in `count`, so consecutive launches over consecutive read ranges fill the same slabs.
    unsigned long long *ovf = nullptr;
    unsigned long long ovf_cap = 0;
    unsigned int *failed = nullptr;
};


template <bool FIXED, bool V210, bool SRC_KEYS, bool PEER>
__global__ void __launch_bounds__(INSERT_THREADS, 4)
part_scatter_kernel(ReadBatch rb, KeySource ks, int k, unsigned int owners, int lp_bits, unsigned int nb, const unsigned int *cta_off,
                    const unsigned long long *bucket_base, unsigned long long *out, PeerOut peers)
{
    __shared__ ReadTile tile;
    extern __shared__ unsigned int s_dyn[];
    // layout: bcur [nb] | rcnt [WARPS][nb] | bstart [nb + 2] | sdst [ROUND_KEYS] u32 | skey [ROUND_KEYS] u64
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned int *bcur = s_dyn;
    unsigned int *rcnt = s_dyn + nb + (size_t)warp * nb;
    unsigned int *rcnt_all = s_dyn + nb;
    unsigned int *bstart = s_dyn + nb + (size_t)WARPS * nb;
    unsigned int *sdst = bstart + nb + 2; // + 2: keeps skey 8-byte aligned for even nb (nb is a power of two or owners << lp)
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(sdst + ROUND_KEYS); // offset 10 nb + 2 + ROUND_KEYS words: even
    // position of the CTA's next key of bucket b, relative to out[0] (a batch holds < 2^32 keys)
    for (unsigned int b = tid; b < nb; b += INSERT_THREADS) {
        unsigned int pos = (unsigned int)bucket_base[b] + cta_off[(size_t)blockIdx.x * nb + b];
        if (PEER) { // relative to the owner's segment, owner in the top bits
            const unsigned int o = b >> lp_bits;
            pos = (pos - (unsigned int)bucket_base[o << lp_bits]) | (o << P2P_REL_BITS);
        }
        bcur[b] = pos;
    }
    auto store = [&](unsigned int pos, unsigned long long key) {
        if (PEER) peers.base[pos >> P2P_REL_BITS][pos & ((1u << P2P_REL_BITS) - 1)] = key;
        else out[pos] = key;
    };
    const unsigned int lt = (1u << lane) - 1;
    __syncthreads();
    // one round: the CTA's (up to) ROUND_KEYS keys, key[0..cnt) per thread, go out sorted by bucket
    auto do_round = [&](const unsigned long long (&key)[SEG], int cnt) {
        for (unsigned int b = lane; b < nb; b += 32) rcnt[b] = 0;
        __syncwarp();
        unsigned int bk[SEG], rk[SEG];
        // all SEG match.any first: they are independent, and their latency (a quarter of this kernel's stall samples when each
        // was consumed at once, profiles/r2b) overlaps
#pragma unroll
        for (int j = 0; j < SEG; j++) {
            bk[j] = j < cnt ? bucket_of(mix64(key[j]), owners, lp_bits) : 0xFFFFFFFFu;
            rk[j] = __match_any_sync(0xFFFFFFFFu, bk[j]);
        }
#pragma unroll
        for (int j = 0; j < SEG; j++) {
            const bool valid = j < cnt;
            const unsigned int b = bk[j], peers_mask = rk[j];
            const unsigned int rank = __popc(peers_mask & lt);
            const unsigned int base = valid ? rcnt[b] : 0;
            __syncwarp();
            if (valid && rank == 0) rcnt[b] = base + __popc(peers_mask);
            __syncwarp();
            rk[j] = base + rank; // rank among this warp's keys of bucket b in this round
        }
        __syncthreads();
        // per bucket: exclusive prefix over the warps (in place) and the round's total
        if (tid < (int)nb) {
            unsigned int acc = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) {
                unsigned int v = rcnt_all[(size_t)w * nb + tid];
                rcnt_all[(size_t)w * nb + tid] = acc;
                acc += v;
            }
            bstart[tid] = acc; // total, turned into the start below
        }
        __syncthreads();
        if (warp == 0) { // exclusive scan of up to 128 totals: 4 consecutive buckets per lane
            unsigned int v[4], sum = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned int b = lane * 4 + q;
                v[q] = b < nb ? bstart[b] : 0;
                sum += v[q];
            }
            unsigned int incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned int x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += x;
            }
            unsigned int run = incl - sum;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const unsigned int b = lane * 4 + q;
                if (b < nb) bstart[b] = run;
                run += v[q];
            }
            if (lane == 31) bstart[nb] = run;
        }
        __syncthreads();
        const unsigned int round_total = bstart[nb];
#pragma unroll
        for (int j = 0; j < SEG; j++)
            if (j < cnt) {
                const unsigned int in_bucket = rcnt[bk[j]] + rk[j];
                const unsigned int idx = bstart[bk[j]] + in_bucket;
                skey[idx] = key[j];
                unsigned int pos = bcur[bk[j]] + in_bucket;
                sdst[idx] = pos;
            }
        __syncthreads();
        for (unsigned int idx = tid; idx < round_total; idx += INSERT_THREADS) store(sdst[idx], skey[idx]);
        if (tid < (int)nb) bcur[tid] += bstart[tid + 1] - bstart[tid];
        __syncthreads();
    };
    if (SRC_KEYS) {
        const unsigned long long n_blk = (ks.n_total + ROUND_KEYS - 1) / ROUND_KEYS;
        for (unsigned long long blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
            unsigned long long key[SEG];
            const int cnt = load_round_keys(ks, blk, key);
            do_round(key, cnt);
        }
    } else {
        const long long n_tiles = (rb.n_reads + TILE_READS - 1) / TILE_READS;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            stage_tile<FIXED>(tile, rb.bin, rb.n_bytes, rb.offsets, rb.rec_bytes, rb.read0, rb.n_reads, k, t);
            const unsigned int total_items = tile.prefix[TILE_READS];
            for (unsigned int item0 = 0; item0 < total_items; item0 += INSERT_THREADS) { // uniform over the CTA
                const unsigned int item = item0 + tid;
                unsigned long long key[SEG];
                int cnt = 0;
                if (item < total_items) cnt = item_keys<V210>(tile, item, k, key);
                do_round(key, cnt);
            }
        }
    }
}

// ---------------------------------------------------------------- the single-pass bucket pass (single GPU)
// One pass over the reads, no count pass: CTA c owns, for every bucket b, the slab [(b * grid + c) * slab, + slab) of `out`
// (SlabOut above).  Same CTA-level multisplit as part_scatter_kernel, rebuilt around what ncu showed there (profiles/
// insert_r2c_*: 157 instructions per k-mer, the ADU pipe -- match.any and barriers -- 69 % busy):
//   * the lanes of a warp that share a bucket are found with one ballot per bucket BIT (vote + one 3-input logic op) instead of
//     match.any;
//   * three CTA barriers per round instead of six: warp 0 alone turns the per-warp counts into positions (prefix over the warps,
//     scan over the buckets, slab cursors) between two of them;
//   * the staging position of a key and its destination differ by a per-bucket constant that warp 0 leaves in shared memory, so
//     the flush is a flat loop -- consecutive threads store consecutive staged keys -- with no per-bucket bookkeeping (a run-wise
//     flush, one warp per bucket run, cost 39 instructions per key at 64 keys per run: profiles/r2j).
// LPB = log2(buckets) at compile time (the ballots and the shared-memory indexing unroll), -1 = run time (any nb <= 128).
// PEER: sharded form (PeerSlabs); LPB then counts the bits of the whole bucket id (owner | slice), lp_bits_rt the slice bits.
template <bool FIXED, bool V210, int LPB, bool PEER>
__global__ void __launch_bounds__(INSERT_THREADS, 4) // measured r2j: 5 / 6 CTAs per SM (51 / 42 registers, spills): 0.618 / 0.641 against 0.594 ms
bucket_slabs_kernel(ReadBatch rb, int k, int lp_bits_rt, unsigned int nb_rt, unsigned long long *out, SlabOut so, PeerSlabs ps)
{
    __shared__ ReadTile tile;
    extern __shared__ unsigned int s_dyn[];
    // layout: rcnt [WARPS][nb] | delta [nb] | lim [nb] | bcur [nb] | total [1] | sdst [ROUND_KEYS] | (pad) | skey [ROUND_KEYS] u64
    const int lp_bits = LPB >= 0 && !PEER ? LPB : lp_bits_rt; // slice bits
    const unsigned int nb = LPB >= 0 && !PEER ? (1u << (LPB >= 0 ? LPB : 0)) : nb_rt;
    int id_bits = LPB >= 0 ? LPB : lp_bits; // bits of a bucket id: one ballot each
    if (LPB < 0 && PEER) { id_bits = 0; while ((1u << id_bits) < nb) id_bits++; }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned int *rcnt_all = s_dyn;
    unsigned int *rcnt = rcnt_all + (size_t)warp * nb;
    unsigned int *delta = s_dyn + (size_t)WARPS * nb; // destination - staging position, per bucket and round
    unsigned int *lim = delta + nb;                   // end of this CTA's slab of the bucket
    unsigned int *bcur = lim + nb;
    unsigned int *total = bcur + nb;
    unsigned int *sdst = total + 1;
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(s_dyn + (((size_t)(WARPS + 3) * nb + 1 + ROUND_KEYS + 1) & ~(size_t)1));
    const unsigned int slab = so.slab, grid = gridDim.x, cta = blockIdx.x;
    for (unsigned int b = tid; b < nb; b += INSERT_THREADS) {
        if (PEER) { // position = owner | index inside my region of the owner's inbox (slice-major slabs); a launch starts empty slabs
            const unsigned int o = b >> lp_bits, l = b & ((1u << lp_bits) - 1);
            bcur[b] = (o << P2P_REL_BITS) | ((l * grid + cta) * slab); // region < 2^P2P_REL_BITS keys: checked by the host
        } else {
            bcur[b] = (b * grid + cta) * slab + so.count[(size_t)b * grid + cta]; // < 2^32: checked by the host
        }
        lim[b] = (PEER ? bcur[b] : (b * grid + cta) * slab) + slab;
    }
    const unsigned int lt = (1u << lane) - 1;
    __syncthreads();
    const long long n_tiles = (rb.n_reads + TILE_READS - 1) / TILE_READS;
    for (long long t = cta; t < n_tiles; t += grid) {
        stage_tile<FIXED>(tile, rb.bin, rb.n_bytes, rb.offsets, rb.rec_bytes, rb.read0, rb.n_reads, k, t);
        const unsigned int total_items = tile.prefix[TILE_READS];
        for (unsigned int item0 = 0; item0 < total_items; item0 += INSERT_THREADS) { // uniform over the CTA
            const unsigned int item = item0 + tid;
            unsigned long long key[SEG];
            int cnt = 0;
            if (item < total_items) cnt = item_keys<V210>(tile, item, k, key);
            for (unsigned int b = lane; b < nb; b += 32) rcnt[b] = 0;
            unsigned int bk[SEG], pm[SEG], rk[SEG];
            if (__any_sync(0xFFFFFFFFu, cnt > 0)) { // a warp without items (the tail round of a tile) only keeps the barriers
#pragma unroll
            for (int j = 0; j < SEG; j++) {
                const bool valid = j < cnt;
                unsigned int b;
                if (PEER) {
                    const unsigned long long h = mix64(key[j]);
                    b = (owner_of(h, ps.owners) << lp_bits) | (lp_bits ? (unsigned int)(h >> (64 - lp_bits)) : 0u);
                } else {
                    b = lp_bits ? (unsigned int)(mix64(key[j]) >> (64 - lp_bits)) : 0u;
                }
                unsigned int peers = __ballot_sync(0xFFFFFFFFu, valid);
                if (LPB >= 0) {
#pragma unroll
                    for (int bit = 0; bit < (LPB >= 0 ? LPB : 0); bit++) {
                        const bool one = (b >> bit) & 1u;
                        const unsigned int bal = __ballot_sync(0xFFFFFFFFu, one);
                        peers &= one ? bal : ~bal;
                    }
                } else {
                    for (int bit = 0; bit < id_bits; bit++) { // uniform trip count
                        const bool one = (b >> bit) & 1u;
                        const unsigned int bal = __ballot_sync(0xFFFFFFFFu, one);
                        peers &= one ? bal : ~bal;
                    }
                }
                bk[j] = b;
                pm[j] = valid ? peers : 0u;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < SEG; j++) {
                const bool valid = pm[j] != 0;
                const unsigned int rank = __popc(pm[j] & lt);
                const unsigned int base = valid ? rcnt[bk[j]] : 0;
                __syncwarp();
                if (valid && rank == 0) rcnt[bk[j]] = base + __popc(pm[j]);
                __syncwarp();
                rk[j] = base + rank; // rank among this warp's keys of the bucket in this round
            }
            }
            __syncthreads(); // (A) every warp's counts are final; the previous round's flush is over
            if (warp == 0) {
                unsigned int carry = 0;
                for (unsigned int b0 = 0; b0 < nb; b0 += 32) {
                    const unsigned int b = b0 + lane;
                    unsigned int tot = 0, pre[WARPS];
                    if (b < nb) {
#pragma unroll
                        for (int w = 0; w < WARPS; w++) {
                            pre[w] = tot; // exclusive prefix over the warps
                            tot += rcnt_all[(size_t)w * nb + b];
                        }
                    }
                    unsigned int incl = tot;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned int x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                        if (lane >= d) incl += x;
                    }
                    if (b < nb) {
                        const unsigned int st = carry + incl - tot;
#pragma unroll
                        for (int w = 0; w < WARPS; w++) rcnt_all[(size_t)w * nb + b] = st + pre[w]; // where warp w's keys of b start in the staging area
                        const unsigned int cur = bcur[b];
                        delta[b] = cur - st;
                        bcur[b] = min(cur + tot, lim[b]); // a full slab stays full: no 32-bit wrap
                    }
                    carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
                }
                if (lane == 0) *total = carry;
            }
            __syncthreads(); // (B)
#pragma unroll
            for (int j = 0; j < SEG; j++)
                if (j < cnt) {
                    const unsigned int at = rcnt[bk[j]] + rk[j], dst = at + delta[bk[j]];
                    skey[at] = key[j];
                    sdst[at] = dst < lim[bk[j]] ? dst : 0xFFFFFFFFu; // beyond the slab
                }
            __syncthreads(); // (C) the round is staged, sorted by bucket
            const unsigned int round_total = *total;
#pragma unroll
            for (int i = 0; i < SEG; i++) {
                const unsigned int at = (unsigned int)i * INSERT_THREADS + tid;
                if (at >= round_total) break;
                const unsigned int dst = sdst[at];
                const unsigned long long key1 = skey[at];
                if (dst != 0xFFFFFFFFu) {
                    if (PEER) ps.keys[dst >> P2P_REL_BITS][dst & ((1u << P2P_REL_BITS) - 1)] = key1;
                    else out[dst] = key1;
                    continue;
                }
                if (so.ovf) { // beyond the slab (pathological inputs only), LIST mode
                    const unsigned long long pos = atomicAdd(so.overflowed, 1ull);
                    if (pos < so.ovf_cap) so.ovf[pos] = key1; else *so.failed = 1u;
                } else { // upserted right here, random access
                    const unsigned long long i1 = slot_of(mix64(key1), so.table.cap);
                    if (upsert_add(so.table, i1, load_key(so.table, i1), key1, 1)) atomicAdd(&so.spread[cta & (SPREAD - 1)], 1ull);
                    atomicAdd(so.overflowed, 1ull);
                }
            }
            // no barrier here: the next round touches only warp-private counters before its barrier (A)
        }
        __syncthreads(); // the tile is overwritten by the next stage_tile; skey / bstart by the next round
    }
    __syncthreads();
    for (unsigned int b = tid; b < nb; b += INSERT_THREADS) {
        if (PEER) {
            const unsigned int o = b >> lp_bits, l = b & ((1u << lp_bits) - 1), used = bcur[b] - (lim[b] - slab);
            ps.cnt[o][l * grid + cta] = used;
            if (used) atomicAdd(&ps.owner_total[o], (unsigned long long)used);
        } else {
            so.count[(size_t)b * grid + cta] = min(bcur[b] - (b * grid + cta) * slab, slab);
        }
    }
}

template <bool FIXED, bool V210>
static void launch_bucket_slabs(int grid, size_t smem, cudaStream_t st, const ReadBatch &rb, int k, int lp_bits, unsigned int nb, unsigned long long *out,
                                const SlabOut &so)
{
    PeerSlabs none;
    memset(&none, 0, sizeof none);
#define GB_BS(L) bucket_slabs_kernel<FIXED, V210, L, false><<<grid, INSERT_THREADS, smem, st>>>(rb, k, lp_bits, nb, out, so, none)
    if (FIXED && !V210) { // the common stream shape gets the unrolled forms
        switch (nb == (1u << lp_bits) ? lp_bits : -1) {
        case 3: GB_BS(3); return;
        case 4: GB_BS(4); return;
        case 5: GB_BS(5); return;
        case 6: GB_BS(6); return;
        case 7: GB_BS(7); return;
        default: break;
        }
    }
    GB_BS(-1);
#undef GB_BS
}

static size_t slabs_smem(unsigned int nb) { return ((((size_t)(WARPS + 3) * nb + 1 + ROUND_KEYS + 1) & ~(size_t)1) + 2) * 4 + (size_t)ROUND_KEYS * 8; }

// ---------------------------------------------------------------- bulk upsert from key ranges
constexpr int IK_THREADS = 256;
constexpr int IK_PER_THREAD = 3; // 3 keys per thread in 32 registers = 8 CTAs per SM: the measurements are at insert_slabs_kernel below

// No CTA barrier anywhere: every thread finds its own chunk, every warp adds its new-key count to one of SPREAD
// global counters (fold_new_keys_kernel sums them into counters[0] afterwards).
// Measured and dropped (round 2, profiles/r2a_bench_*.json, r2b_insert_sweep.jsonl): claiming a new key together with its count
// by one 128-bit compare-and-swap (1.865 ms against 1.834 ms), CAS-first probing, 2 / 8 / 16 keys per thread, prefetching the
// next table slice into L2 while the current one is filled (2.22 against 2.15 ms).
// DEV_TOTAL: n_total is only an upper bound (it sized the grid); the exact number of keys is vstart[n_chunks].
// Slot indices are 32-bit (IdxT) whenever the table has fewer than 2^32 slots.
template <bool DEV_TOTAL, typename IdxT>
__global__ void __launch_bounds__(IK_THREADS, 8)
insert_keys_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ vstart,
                   const unsigned long long *__restrict__ off, int n_chunks, unsigned long long n_total, Table table, unsigned long long *spread)
{
    const unsigned long long cap = table.cap;
    constexpr int IK_PER_CTA = IK_THREADS * IK_PER_THREAD;
    const unsigned long long v0 = (unsigned long long)blockIdx.x * IK_PER_CTA;
    if (DEV_TOTAL) {
        n_total = min(n_total, vstart[n_chunks]);
        if (v0 >= n_total) return;
    }
    int c = 0;
    if (n_chunks > 1) { // last chunk with vstart <= v0 (uniform over the CTA: served from L1)
        int lo = 0, hi = n_chunks;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (vstart[mid] <= v0) lo = mid; else hi = mid;
        }
        c = lo;
    }
    unsigned long long key[IK_PER_THREAD], cur[IK_PER_THREAD];
    IdxT idx[IK_PER_THREAD];
    bool ok[IK_PER_THREAD];
#pragma unroll
    for (int j = 0; j < IK_PER_THREAD; j++) {
        unsigned long long v = v0 + (unsigned long long)j * IK_THREADS + threadIdx.x;
        ok[j] = v < n_total;
        if (ok[j]) {
            while (v >= vstart[c + 1]) c++; // chunks ascend with v; empty chunks are skipped
            key[j] = __ldcs(keys + off[c] + (v - vstart[c]));
            idx[j] = (IdxT)slot_of(mix64(key[j]), cap);
        }
    }
    int nk = 0;
    unsigned long long old[IK_PER_THREAD];
#pragma unroll
    for (int j = 0; j < IK_PER_THREAD; j++)
        if (ok[j]) cur[j] = load_key(table, idx[j]);
    // the CAS round trips of a thread's keys overlap: all of them are issued before the first result is used
#pragma unroll
    for (int j = 0; j < IK_PER_THREAD; j++) {
        old[j] = cur[j];
        if (ok[j] && cur[j] == EMPTY_KEY) old[j] = atomicCAS(table.key + idx[j], EMPTY_KEY, key[j]);
    }
#pragma unroll
    for (int j = 0; j < IK_PER_THREAD; j++) {
        if (!ok[j]) continue;
        const bool claimed = cur[j] == EMPTY_KEY && old[j] == EMPTY_KEY;
        if (claimed) {
            nk++; // the count word of a free slot already holds this key's 1 (init_table_kernel)
        } else if (old[j] == key[j]) {
            red_add_s32(table.count + idx[j], 1);
        } else {
            // the slot belongs to another key: linear probing from the next slot (rare at load <= 0.5)
            unsigned long long nx = next_slot(idx[j], cap);
            nk += upsert_add(table, nx, load_key(table, nx), key[j], 1);
        }
    }
    nk = __reduce_add_sync(0xFFFFFFFFu, nk);
    if ((threadIdx.x & 31) == 0 && nk)
        atomicAdd(&spread[(blockIdx.x * (IK_THREADS / 32) + (threadIdx.x >> 5)) & (SPREAD - 1)], (unsigned long long)nk);
}

__global__ void __launch_bounds__(SPREAD)
fold_new_keys_kernel(unsigned long long *spread, unsigned long long *counters)
{
    unsigned long long v = spread[threadIdx.x];
    spread[threadIdx.x] = 0;
    // block-wide sum: warp reduce then one atomic per warp (8 atomics)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&counters[0], v);
}

// The same upsert over the SLABS of the single-pass bucket pass: slab number s = (bucket, CTA of the bucket pass) holds count[s]
// keys at keys[s * slab], and CTA (s, part) takes keys [part * 768, ...) of it.  The launch is sized for full slabs; a CTA
// whose part lies beyond count[s] leaves at once.  No chunk table, no search: measured on C2 the search over the 19 k-entry
// chunk table cost the generic kernel 0.2 ms (2.15 against 1.95 ms).  Slabs are bucket-major, so slice order is kept.
// The kernel is latency bound (ncu r2c: 29 long-scoreboard stall cycles per issue): three dependent round trips per key -- staged
// key, table key, compare-and-swap -- so it runs at full occupancy: 3 keys per thread in 32 registers, 8 CTAs per SM (measured
// r2i, keys per thread x CTAs per SM on C2: 4 x 6 1.326 ms, 3 x 8 1.303, 2 x 8 1.356, 4 x 8 (spills) 1.458, 6 x 5 1.531).  Measured and dropped (profiles/r2d_upsert_variants.jsonl, r2g_upsert_ablation.jsonl): the
// compare-and-swap as the probe (no preceding load): 1.73 ms; a persistent grid fetching the next item's keys early: 9 ms -- the
// CTAs drift apart and the slice being filled no longer stays in L2; the clear fused into the insert (slice-wise initialisation
// right before a slice is filled, bounded probing, overflow list): 2.39 ms against 0.23 + 1.52 -- lines that are already in L2 do
// not make the upsert faster, its bound is the L2's handling of the requests, not DRAM.  What the parts cost on C2: staged keys +
// hashing 0.20 ms, + table key loads 0.47, + one red per key 0.87, + compare-and-swap of the 31 % new keys 1.52.
// Slot indices are 32-bit (IdxT) whenever the table has fewer than 2^32 slots.
constexpr int IS_PER_THREAD = 3;
// INBOX (sharded map): the slabs are the ones P source ranks stored into this rank's inbox (PeerSlabs); they are walked slice-major --
// slab number s = ((slice * P + source) * grid + CTA) -- so that the resident CTAs still share one table slice.
template <typename IdxT, bool INBOX>
__global__ void __launch_bounds__(IK_THREADS, 8)
insert_slabs_kernel(const unsigned long long *__restrict__ keys, const unsigned int *__restrict__ count, unsigned int slab,
                    unsigned int ctas_per_slab, Table table, unsigned long long *spread, InboxSlabs in)
{
    constexpr int IK_PER_CTA = IK_THREADS * IS_PER_THREAD;
    const unsigned long long cap = table.cap;
    const unsigned int s = blockIdx.x / ctas_per_slab, part = blockIdx.x - s * ctas_per_slab;
    const unsigned long long *slab_keys;
    unsigned int filled;
    if (INBOX) {
        const unsigned int cta = s % in.grid, ls = s / in.grid, src_rank = ls % in.sources, l = ls / in.sources;
        const unsigned long long *region = keys + (size_t)src_rank * in.region_cap;
        slab_keys = region + (size_t)(l * in.grid + cta) * slab;
        filled = reinterpret_cast<const unsigned int *>(region + in.cnt_off)[l * in.grid + cta];
    } else {
        slab_keys = keys + (size_t)s * slab;
        filled = count[s];
    }
    const unsigned int n = min(filled, slab), v0 = part * IK_PER_CTA;
    if (v0 >= n) return;
    const unsigned long long *src = slab_keys + v0;
    unsigned long long key[IS_PER_THREAD], cur[IS_PER_THREAD], old[IS_PER_THREAD];
    IdxT idx[IS_PER_THREAD];
    bool ok[IS_PER_THREAD];
#pragma unroll
    for (int j = 0; j < IS_PER_THREAD; j++) {
        const unsigned int i = (unsigned int)j * IK_THREADS + threadIdx.x;
        ok[j] = v0 + i < n;
        if (ok[j]) {
            key[j] = __ldcs(src + i);
            idx[j] = (IdxT)slot_of(mix64(key[j]), cap);
        }
    }
#pragma unroll
    for (int j = 0; j < IS_PER_THREAD; j++)
        if (ok[j]) cur[j] = load_key(table, idx[j]);
    // the CAS round trips of a thread's keys overlap: all of them are issued before the first result is used
#pragma unroll
    for (int j = 0; j < IS_PER_THREAD; j++) {
        old[j] = cur[j];
        if (ok[j] && cur[j] == EMPTY_KEY) old[j] = atomicCAS(table.key + idx[j], EMPTY_KEY, key[j]);
    }
    int nk = 0;
#pragma unroll
    for (int j = 0; j < IS_PER_THREAD; j++) {
        if (!ok[j]) continue;
        const bool claimed = cur[j] == EMPTY_KEY && old[j] == EMPTY_KEY;
        if (claimed) {
            nk++; // the count word of a free slot already holds this key's 1 (init_table_kernel)
        } else if (old[j] == key[j]) {
            red_add_s32(table.count + idx[j], 1);
        } else {
            // the slot belongs to another key: linear probing from the next slot (rare at load <= 0.5)
            const unsigned long long nx = next_slot(idx[j], cap);
            nk += upsert_add(table, nx, load_key(table, nx), key[j], 1);
        }
    }
    nk = __reduce_add_sync(0xFFFFFFFFu, nk);
    if ((threadIdx.x & 31) == 0 && nk)
        atomicAdd(&spread[(blockIdx.x * (IK_THREADS / 32) + (threadIdx.x >> 5)) & (SPREAD - 1)], (unsigned long long)nk);
}

// ---------------------------------------------------------------- host side

int PartWork::ensure(cudaStream_t st)
{
    if (cta_hist) return GB_OK;
    grid = SM_COUNT * 4; // persistent over the tiles; 4-5 CTAs of 256 threads fit the shared memory of an SM
    owner_stream = st;
    GB_CUDA(cudaMalloc((void **)&cta_hist, (size_t)grid * MAX_BUCKETS * sizeof(unsigned int)));
    GB_CUDA(cudaMalloc((void **)&bucket_base, (MAX_BUCKETS + 1) * sizeof(unsigned long long)));
    GB_CUDA(cudaMalloc((void **)&bucket_total, MAX_BUCKETS * sizeof(unsigned long long)));
    return GB_OK;
}

void PartWork::release()
{
    if (cta_hist) cudaFree(cta_hist);
    if (bucket_base) cudaFree(bucket_base);
    if (bucket_total) cudaFree(bucket_total);
    cta_hist = nullptr;
    bucket_base = bucket_total = nullptr;
}

static int finish_count(PartWork &w, unsigned int nb, cudaStream_t st)
{
    part_offsets_kernel<<<nb, 256, 0, st>>>(w.cta_hist, w.grid, nb, w.bucket_total);
    GB_LAUNCHED();
    part_bases_kernel<<<1, MAX_BUCKETS, 0, st>>>(w.bucket_total, nb, w.bucket_base);
    GB_LAUNCHED();
    return GB_OK;
}

int part_count(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, cudaStream_t st)
{
    GB_TRY(w.ensure(st));
    const unsigned int nb = (unsigned int)pl.nb();
    if (nb > MAX_BUCKETS) { set_error("internal: %u buckets", nb); return GB_E_ARG; }
    const bool fixed = rb.offsets == nullptr;
    KeySource none;
#define GB_PC(F, V) part_count_kernel<F, V, false><<<w.grid, INSERT_THREADS, 0, st>>>(rb, none, k, (unsigned int)pl.owners, pl.lp_bits, nb, w.cta_hist)
    if (fixed) { if (v210) GB_PC(true, true); else GB_PC(true, false); }
    else { if (v210) GB_PC(false, true); else GB_PC(false, false); }
#undef GB_PC
    GB_LAUNCHED();
    return finish_count(w, nb, st);
}

int part_count_keys(const KeySource &ks, const PartLayout &pl, PartWork &w, cudaStream_t st)
{
    GB_TRY(w.ensure(st));
    const unsigned int nb = (unsigned int)pl.nb();
    if (nb > MAX_BUCKETS) { set_error("internal: %u buckets", nb); return GB_E_ARG; }
    ReadBatch none;
    part_count_kernel<true, false, true><<<w.grid, INSERT_THREADS, 0, st>>>(none, ks, 0, (unsigned int)pl.owners, pl.lp_bits, nb, w.cta_hist);
    GB_LAUNCHED();
    return finish_count(w, nb, st);
}

static size_t scatter_smem(unsigned int nb) { return ((size_t)nb * (2 + WARPS) + 4 + ROUND_KEYS) * 4 + (size_t)ROUND_KEYS * 8; }

static int launch_scatter(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, unsigned long long *out,
                          const PeerOut *peers, cudaStream_t st)
{
    const unsigned int nb = (unsigned int)pl.nb();
    const bool fixed = rb.offsets == nullptr;
    if (nb > STAGE_MAX_BUCKETS) { set_error("internal: %u buckets exceed the staged bucket pass", nb); return GB_E_ARG; }
    const size_t smem = scatter_smem(nb);
    PeerOut po;
    memset(&po, 0, sizeof po);
    if (peers) po = *peers;
    KeySource none;
#define GB_PS(F, V, P) part_scatter_kernel<F, V, false, P><<<w.grid, INSERT_THREADS, smem, st>>>(rb, none, k, (unsigned int)pl.owners, pl.lp_bits, nb, w.cta_hist, w.bucket_base, out, po)
#define GB_PS2(F, V) do { if (peers) GB_PS(F, V, true); else GB_PS(F, V, false); } while (0)
    if (fixed) { if (v210) GB_PS2(true, true); else GB_PS2(true, false); }
    else { if (v210) GB_PS2(false, true); else GB_PS2(false, false); }
#undef GB_PS2
#undef GB_PS
    GB_LAUNCHED();
    return GB_OK;
}

int part_scatter(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, unsigned long long *out, cudaStream_t st)
{
    return launch_scatter(rb, k, v210, pl, w, out, nullptr, st);
}

int part_scatter_peers(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, const PeerOut &peers, cudaStream_t st)
{
    return launch_scatter(rb, k, v210, pl, w, nullptr, &peers, st);
}

int part_scatter_keys(const KeySource &ks, const PartLayout &pl, PartWork &w, unsigned long long *out, cudaStream_t st)
{
    const unsigned int nb = (unsigned int)pl.nb();
    if (nb > STAGE_MAX_BUCKETS) { set_error("internal: %u buckets exceed the staged bucket pass", nb); return GB_E_ARG; }
    ReadBatch none;
    PeerOut po;
    memset(&po, 0, sizeof po);
    part_scatter_kernel<true, false, true, false><<<w.grid, INSERT_THREADS, scatter_smem(nb), st>>>(none, ks, 0, (unsigned int)pl.owners, pl.lp_bits, nb,
                                                                                             w.cta_hist, w.bucket_base, out, po);
    GB_LAUNCHED();
    return GB_OK;
}

// ---- the single-pass variant (gb_tune single_pass, the default for large batches)
// keys per (bucket, CTA) slab.  A CTA takes every grid-th tile, so what it can see is bounded by its number of tiles, not by
// total / grid (a batch with fewer tiles than CTAs leaves most CTAs idle and the busy ones with a whole tile each):
// cta_keys = tiles per CTA x reads per tile x windows per read; the slab is that share of one bucket plus 8 standard
// deviations of a binomial draw, rounded up to 64 keys (512 B).  0 = the slabs would not fit 32-bit positions.
unsigned int slab_keys_for(unsigned long long cta_keys, unsigned int nb, int grid)
{
    const double e = (double)cta_keys / (double)nb;
    const unsigned long long slab = (((unsigned long long)(e + 8.0 * sqrt(e + 1.0)) + 64) + 63) / 64 * 64;
    if (slab * nb * (unsigned long long)grid >= 0xFFF00000ull) return 0; // positions are 32-bit, with room for one round above a limit
    return (unsigned int)slab;
}
// upper bound of the keys one CTA of a `grid`-CTA bucket pass sees in a batch of n_reads reads with at most `windows` k-windows
unsigned long long slab_cta_keys(long long n_reads, unsigned long long windows, int grid)
{
    if (n_reads <= 0) return 0;
    const unsigned long long tiles = ((unsigned long long)n_reads + TILE_READS - 1) / TILE_READS;
    const unsigned long long per_cta = (tiles + (unsigned long long)grid - 1) / (unsigned long long)grid;
    const unsigned long long per_read = (windows + (unsigned long long)n_reads - 1) / (unsigned long long)n_reads;
    return per_cta * TILE_READS * per_read;
}

// after the slab bucket pass: counters[3] (k-windows) += keys in slabs + keys that left through the overflow path.
// ovf_cap > 0 (LIST mode): desc = { 0, keys in the overflow list, 0 } = the chunk table of that list (vstart[0], vstart[1], off[0]).
__global__ void __launch_bounds__(1024)
slab_finish_kernel(const unsigned int *count, unsigned int n_slabs, unsigned int slab, const unsigned long long *overflowed,
                   unsigned long long *counters, unsigned long long ovf_cap, unsigned long long *desc)
{
    unsigned long long sum = 0;
    for (unsigned int c = threadIdx.x; c < n_slabs; c += 1024) sum += min(count[c], slab);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_down_sync(0xFFFFFFFFu, sum, d);
    __shared__ unsigned long long s_part[32];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (int i = 0; i < 32; i++) acc += s_part[i];
        const unsigned long long extra = ovf_cap ? min(*overflowed, ovf_cap) : *overflowed;
        atomicAdd(&counters[3], acc + extra);
        if (desc) { desc[0] = 0; desc[1] = ovf_cap ? extra : 0; desc[2] = 0; }
    }
}

// pass 2 without pass 1: `out` holds nb * grid slabs of `slab` keys, w.cta_hist their fill counts ([bucket][CTA]).
// `spread` / `overflowed` / the table are for the keys that do not fit their slab.
int part_scatter_slabs(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, unsigned long long *out, unsigned int slab,
                       Map *m, cudaStream_t st)
{
    GB_TRY(w.ensure(st));
    const unsigned int nb = (unsigned int)pl.nb();
    if (nb > STAGE_MAX_BUCKETS || pl.owners != 1) { set_error("internal: slab bucket pass with %u buckets, %d owners", nb, pl.owners); return GB_E_ARG; }
    if (!m->d_spread) {
        GB_CUDA(cudaMalloc((void **)&m->d_spread, SPREAD * 8));
        GB_CUDA(cudaMemsetAsync(m->d_spread, 0, SPREAD * 8, st));
    }
    const bool fixed = rb.offsets == nullptr;
    const size_t smem = slabs_smem(nb);
    const unsigned int n_chunks = nb * (unsigned int)w.grid;
    SlabOut so;
    so.slab = slab;
    so.count = w.cta_hist; // [nb][grid] here (the counted passes use it as [grid][nb])
    so.table = m->view();
    so.spread = m->d_spread;
    so.overflowed = w.bucket_total; // one word is enough; the counted passes are not running
    GB_CUDA(cudaMemsetAsync(so.overflowed, 0, 8, st));
    GB_CUDA(cudaMemsetAsync(so.count, 0, (size_t)n_chunks * 4, st)); // the kernel resumes from these cursors
#define GB_PSS(F, V) launch_bucket_slabs<F, V>(w.grid, smem, st, rb, k, pl.lp_bits, nb, out, so)
    if (fixed) { if (v210) GB_PSS(true, true); else GB_PSS(true, false); }
    else { if (v210) GB_PSS(false, true); else GB_PSS(false, false); }
#undef GB_PSS
    GB_LAUNCHED();
    slab_finish_kernel<<<1, 1024, 0, st>>>(so.count, n_chunks, slab, so.overflowed, m->d_counters, 0ull, nullptr);
    GB_LAUNCHED();
    return GB_OK;
}

// LIST mode, one launch per read range (the chunked host insert of map.cu).  begin: zero the cursors; range: one bucket-pass
// launch over rb; end: k-window total and the 3-word chunk table of the overflow list in d_desc.
// out = nb * grid slabs | overflow region of ovf_cap keys.  w.bucket_total[0] = overflow cursor, [1] = failed flag.
int slab_list_begin(const PartLayout &pl, PartWork &w, cudaStream_t st)
{
    GB_TRY(w.ensure(st));
    GB_CUDA(cudaMemsetAsync(w.bucket_total, 0, 16, st));
    GB_CUDA(cudaMemsetAsync(w.cta_hist, 0, (size_t)pl.nb() * w.grid * 4, st));
    return GB_OK;
}
int slab_list_range(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, unsigned long long *out, unsigned int slab,
                    unsigned long long ovf_cap, cudaStream_t st)
{
    const unsigned int nb = (unsigned int)pl.nb();
    if (nb > STAGE_MAX_BUCKETS || pl.owners != 1 || rb.offsets) { set_error("internal: slab list pass misuse"); return GB_E_ARG; }
    SlabOut so;
    so.slab = slab;
    so.count = w.cta_hist;
    so.overflowed = w.bucket_total;
    so.ovf = out + (size_t)slab * nb * (size_t)w.grid;
    so.ovf_cap = ovf_cap;
    so.failed = reinterpret_cast<unsigned int *>(w.bucket_total + 1);
    const size_t smem = slabs_smem(nb);
    if (v210) launch_bucket_slabs<true, true>(w.grid, smem, st, rb, k, pl.lp_bits, nb, out, so);
    else launch_bucket_slabs<true, false>(w.grid, smem, st, rb, k, pl.lp_bits, nb, out, so);
    GB_LAUNCHED();
    return GB_OK;
}
int slab_list_end(const PartLayout &pl, PartWork &w, unsigned int slab, unsigned long long ovf_cap, unsigned long long *d_desc, Map *m, cudaStream_t st)
{
    const unsigned int n_chunks = (unsigned int)pl.nb() * (unsigned int)w.grid;
    slab_finish_kernel<<<1, 1024, 0, st>>>(w.cta_hist, n_chunks, slab, w.bucket_total, m->d_counters, ovf_cap, d_desc);
    GB_LAUNCHED();
    return GB_OK;
}

__global__ void make_desc_kernel(const unsigned long long *total, unsigned long long *desc, unsigned long long *counters)
{
    desc[0] = 0;      // vstart[0]
    desc[1] = *total; // vstart[1]
    desc[2] = 0;      // off[0]
    atomicAdd(&counters[3], *total);
}

int make_single_chunk(const unsigned long long *d_total, unsigned long long *d_desc, unsigned long long *d_counters, cudaStream_t st)
{
    make_desc_kernel<<<1, 1, 0, st>>>(d_total, d_desc, d_counters);
    GB_LAUNCHED();
    return GB_OK;
}

int insert_slabs(Map *m, const unsigned long long *d_keys, const unsigned int *d_count, unsigned int slab, unsigned int n_slabs, cudaStream_t st)
{
    if (!n_slabs || !slab) return GB_OK;
    m->kept_valid = false;
    if (!m->d_spread) {
        GB_CUDA(cudaMalloc((void **)&m->d_spread, SPREAD * 8));
        GB_CUDA(cudaMemsetAsync(m->d_spread, 0, SPREAD * 8, st));
    }
    const unsigned int per = (slab + IK_THREADS * IS_PER_THREAD - 1) / (IK_THREADS * IS_PER_THREAD);
    const unsigned long long work = (unsigned long long)n_slabs * per;
    if (work >= 0x7FFFFFFFull) { set_error("internal: %llu slab CTAs", work); return GB_E_ARG; }
    const Table t = m->view();
    InboxSlabs none;
    if (t.cap < (1ull << 32)) insert_slabs_kernel<unsigned int, false><<<(unsigned int)work, IK_THREADS, 0, st>>>(d_keys, d_count, slab, per, t, m->d_spread, none);
    else insert_slabs_kernel<unsigned long long, false><<<(unsigned int)work, IK_THREADS, 0, st>>>(d_keys, d_count, slab, per, t, m->d_spread, none);
    GB_LAUNCHED();
    fold_new_keys_kernel<<<1, SPREAD, 0, st>>>(m->d_spread, m->d_counters);
    GB_LAUNCHED();
    return GB_OK;
}

// upsert of the slabs P source ranks stored into this rank's inbox (one buffer set), slice-major
int insert_inbox_slabs(Map *m, const unsigned long long *d_inbox, const InboxSlabs &in, unsigned int slab, cudaStream_t st)
{
    if (!slab || !in.sources || !in.slices) return GB_OK;
    m->kept_valid = false;
    if (!m->d_spread) {
        GB_CUDA(cudaMalloc((void **)&m->d_spread, SPREAD * 8));
        GB_CUDA(cudaMemsetAsync(m->d_spread, 0, SPREAD * 8, st));
    }
    const unsigned int per = (slab + IK_THREADS * IS_PER_THREAD - 1) / (IK_THREADS * IS_PER_THREAD);
    const unsigned long long work = (unsigned long long)in.slices * in.sources * in.grid * per;
    if (work >= 0x7FFFFFFFull) { set_error("internal: %llu slab CTAs", work); return GB_E_ARG; }
    const Table t = m->view();
    if (t.cap < (1ull << 32)) insert_slabs_kernel<unsigned int, true><<<(unsigned int)work, IK_THREADS, 0, st>>>(d_inbox, nullptr, slab, per, t, m->d_spread, in);
    else insert_slabs_kernel<unsigned long long, true><<<(unsigned int)work, IK_THREADS, 0, st>>>(d_inbox, nullptr, slab, per, t, m->d_spread, in);
    GB_LAUNCHED();
    fold_new_keys_kernel<<<1, SPREAD, 0, st>>>(m->d_spread, m->d_counters);
    GB_LAUNCHED();
    return GB_OK;
}

// the single-pass bucket pass of a sharded map: reads [read0, +n_reads) of a fixed-stride stream into the owners' inbox slabs (ps);
// keys beyond a slab are appended to the LOCAL list ovf[0 .. ovf_cap) (cursor *d_cursor, zeroed here; *d_failed set if it overflows)
int bucket_slabs_peers(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, unsigned int slab, const PeerSlabs &ps,
                       unsigned long long *ovf, unsigned long long ovf_cap, unsigned long long *d_cursor, unsigned int *d_failed, cudaStream_t st)
{
    GB_TRY(w.ensure(st));
    const unsigned int nb = (unsigned int)pl.nb();
    if (nb > STAGE_MAX_BUCKETS || rb.offsets) { set_error("internal: sharded slab bucket pass with %u buckets / a ragged stream", nb); return GB_E_ARG; }
    if ((unsigned long long)(1u << pl.lp_bits) * (unsigned long long)w.grid * slab >= (1ull << P2P_REL_BITS)) { set_error("internal: inbox region beyond 2^%d keys", P2P_REL_BITS); return GB_E_ARG; }
    SlabOut so;
    so.slab = slab;
    so.overflowed = d_cursor;
    so.ovf = ovf;
    so.ovf_cap = ovf_cap;
    so.failed = d_failed;
    GB_CUDA(cudaMemsetAsync(d_cursor, 0, 8, st));
    GB_CUDA(cudaMemsetAsync(d_failed, 0, 4, st));
    const size_t smem = slabs_smem(nb);
    int bits = 0;
    while ((1u << bits) < nb) bits++;
#define GB_BP(V, L) bucket_slabs_kernel<true, V, L, true><<<w.grid, INSERT_THREADS, smem, st>>>(rb, k, pl.lp_bits, nb, nullptr, so, ps)
    if (v210) GB_BP(true, -1);
    else switch (bits) {
        case 3: GB_BP(false, 3); break;
        case 4: GB_BP(false, 4); break;
        case 5: GB_BP(false, 5); break;
        case 6: GB_BP(false, 6); break;
        case 7: GB_BP(false, 7); break;
        default: GB_BP(false, -1); break;
    }
#undef GB_BP
    GB_LAUNCHED();
    return GB_OK;
}

int insert_key_chunks(Map *m, const unsigned long long *d_keys, const unsigned long long *d_vstart, const unsigned long long *d_off,
                      int n_chunks, unsigned long long n_total, cudaStream_t st, bool total_is_upper_bound)
{
    if (!n_total) return GB_OK;
    m->kept_valid = false;
    if (!m->d_spread) {
        GB_CUDA(cudaMalloc((void **)&m->d_spread, SPREAD * 8));
        GB_CUDA(cudaMemsetAsync(m->d_spread, 0, SPREAD * 8, st));
    }
    const unsigned int grid = (unsigned int)((n_total + IK_THREADS * IK_PER_THREAD - 1) / (IK_THREADS * IK_PER_THREAD));
    const Table t = m->view();
#define GB_IK(D, I) insert_keys_kernel<D, I><<<grid, IK_THREADS, 0, st>>>(d_keys, d_vstart, d_off, n_chunks, n_total, t, m->d_spread)
    if (t.cap < (1ull << 32)) { if (total_is_upper_bound) GB_IK(true, unsigned int); else GB_IK(false, unsigned int); }
    else { if (total_is_upper_bound) GB_IK(true, unsigned long long); else GB_IK(false, unsigned long long); }
#undef GB_IK
    GB_LAUNCHED();
    fold_new_keys_kernel<<<1, SPREAD, 0, st>>>(m->d_spread, m->d_counters);
    GB_LAUNCHED();
    return GB_OK;
}

} // namespace gb
