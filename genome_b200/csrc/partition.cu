// partition.cu -- L2-blocked k-mer insertion.
//
// A k-mer table far larger than the 126 MB L2 turns FreqFilter.add (S/data/FreqFilter.scala:28-36, paths relative
// to /root/reference) into one random DRAM access per k-mer instance, and every miss moves a whole 128-byte line
// (ncu, profiles/r1a: 138 B of DRAM reads per k-mer).  Instead:
//   1. part_count / part_scatter stream the read batch twice and write every canonical k-mer (8 B) into the bucket
//      of the table SLICE it hashes to -- sequential traffic;
//   2. insert_key_chunks walks the buckets in slice order, so the CTAs that are resident at any moment all update
//      the same 32 MiB slice of the table, which lives in L2; DRAM sees each slice once in and once out.
// The same buckets, with an owner-shard prefix, are what the sharded map sends over NVLink (comm.cu).
#include "partition.cuh"

#include "extract.cuh"
#include "scan.cuh"

namespace gb {

__device__ __forceinline__ unsigned int bucket_of(unsigned long long h, unsigned int owners, int lp_bits)
{
    unsigned int slice = lp_bits ? (unsigned int)(h >> (64 - lp_bits)) : 0u;
    return (owner_of(h, owners) << lp_bits) | slice;
}

// persistent CTAs: CTA c takes tiles c, c + grid, ... in BOTH passes, so its per-bucket counts of pass 1 are exactly
// the room it needs in pass 2 -- no global atomics, deterministic layout.
template <bool FIXED, bool V210>
__global__ void __launch_bounds__(INSERT_THREADS)
part_count_kernel(ReadBatch rb, int k, unsigned int owners, int lp_bits, unsigned int nb, unsigned int *cta_hist)
{
    __shared__ ReadTile tile;
    __shared__ unsigned int s_hist[MAX_BUCKETS];
    const int tid = threadIdx.x;
    for (int b = tid; b < MAX_BUCKETS; b += INSERT_THREADS) s_hist[b] = 0;
    const long long n_tiles = (rb.n_reads + TILE_READS - 1) / TILE_READS;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        stage_tile<FIXED>(tile, rb.bin, rb.n_bytes, rb.offsets, rb.rec_bytes, rb.read0, rb.n_reads, k, t);
        const unsigned int total_items = tile.prefix[TILE_READS];
        for (unsigned int item = tid; item < total_items; item += INSERT_THREADS) {
            unsigned long long key[SEG];
            const int cnt = item_keys<V210>(tile, item, k, key);
#pragma unroll
            for (int j = 0; j < SEG; j++)
                if (j < cnt) atomicAdd(&s_hist[bucket_of(mix64(key[j]), owners, lp_bits)], 1u);
        }
        __syncthreads(); // the tile is overwritten by the next stage_tile
    }
    __syncthreads();
    for (unsigned int b = tid; b < nb; b += INSERT_THREADS) cta_hist[(size_t)blockIdx.x * nb + b] = s_hist[b];
}

// one CTA per bucket: turn the column of per-CTA counts into exclusive offsets inside the bucket
__global__ void __launch_bounds__(256)
part_offsets_kernel(unsigned int *cta_hist, int grid, unsigned int nb, unsigned long long *bucket_total)
{
    const unsigned int b = blockIdx.x;
    const int per = (grid + 255) / 256;
    const int c0 = threadIdx.x * per, c1 = min(grid, c0 + per);
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

template <bool FIXED, bool V210>
__global__ void __launch_bounds__(INSERT_THREADS)
part_scatter_kernel(ReadBatch rb, int k, unsigned int owners, int lp_bits, unsigned int nb, const unsigned int *cta_hist,
                    const unsigned long long *bucket_base, unsigned long long *out)
{
    __shared__ ReadTile tile;
    __shared__ unsigned long long s_base[MAX_BUCKETS];
    __shared__ unsigned int s_cur[MAX_BUCKETS];
    const int tid = threadIdx.x;
    for (unsigned int b = tid; b < MAX_BUCKETS; b += INSERT_THREADS) {
        s_cur[b] = 0;
        s_base[b] = b < nb ? bucket_base[b] + cta_hist[(size_t)blockIdx.x * nb + b] : 0;
    }
    const long long n_tiles = (rb.n_reads + TILE_READS - 1) / TILE_READS;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        stage_tile<FIXED>(tile, rb.bin, rb.n_bytes, rb.offsets, rb.rec_bytes, rb.read0, rb.n_reads, k, t);
        const unsigned int total_items = tile.prefix[TILE_READS];
        for (unsigned int item = tid; item < total_items; item += INSERT_THREADS) {
            unsigned long long key[SEG];
            const int cnt = item_keys<V210>(tile, item, k, key);
#pragma unroll
            for (int j = 0; j < SEG; j++)
                if (j < cnt) {
                    unsigned int b = bucket_of(mix64(key[j]), owners, lp_bits);
                    out[s_base[b] + atomicAdd(&s_cur[b], 1u)] = key[j];
                }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- bulk upsert from key ranges
constexpr int IK_THREADS = 256;
constexpr int IK_PER_THREAD = 8;
constexpr int IK_PER_CTA = IK_THREADS * IK_PER_THREAD;

__global__ void __launch_bounds__(IK_THREADS)
insert_keys_kernel(const unsigned long long *__restrict__ keys, const unsigned long long *__restrict__ vstart,
                   const unsigned long long *__restrict__ off, int n_chunks, unsigned long long n_total, Slot *table, int bits,
                   unsigned long long *counters)
{
    __shared__ int s_chunk0;
    __shared__ unsigned int s_new;
    const unsigned long long v0 = (unsigned long long)blockIdx.x * IK_PER_CTA;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_chunks; // last chunk with vstart <= v0
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (vstart[mid] <= v0) lo = mid; else hi = mid;
        }
        s_chunk0 = lo;
        s_new = 0;
    }
    __syncthreads();
    const unsigned long long tmask = (1ull << bits) - 1;
    unsigned long long key[IK_PER_THREAD], idx[IK_PER_THREAD], cur[IK_PER_THREAD];
    bool ok[IK_PER_THREAD];
    int c = s_chunk0;
#pragma unroll
    for (int j = 0; j < IK_PER_THREAD; j++) {
        unsigned long long v = v0 + (unsigned long long)j * IK_THREADS + threadIdx.x;
        ok[j] = v < n_total;
        if (ok[j]) {
            while (v >= vstart[c + 1]) c++; // chunks ascend with v; empty chunks are skipped
            key[j] = __ldcs(keys + off[c] + (v - vstart[c]));
            idx[j] = slot_of(mix64(key[j]), bits);
        }
    }
#pragma unroll
    for (int j = 0; j < IK_PER_THREAD; j++)
        if (ok[j]) cur[j] = load_key(table + idx[j]);
    int nk = 0;
#pragma unroll
    for (int j = 0; j < IK_PER_THREAD; j++)
        if (ok[j]) nk += upsert_add(table, tmask, idx[j], cur[j], key[j], 1);
    nk = __reduce_add_sync(0xFFFFFFFFu, nk);
    if ((threadIdx.x & 31) == 0 && nk) atomicAdd(&s_new, (unsigned int)nk);
    __syncthreads();
    if (threadIdx.x == 0 && s_new) atomicAdd(&counters[0], (unsigned long long)s_new);
}

// ---------------------------------------------------------------- host side

int PartWork::ensure(cudaStream_t st)
{
    if (cta_hist) return GB_OK;
    grid = SM_COUNT * 8; // 8 resident CTAs of 256 threads per SM (37 registers); persistent over the tiles
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

int part_count(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, cudaStream_t st)
{
    GB_TRY(w.ensure(st));
    const unsigned int nb = (unsigned int)pl.nb();
    if (nb > MAX_BUCKETS) { set_error("internal: %u buckets", nb); return GB_E_ARG; }
    const bool fixed = rb.offsets == nullptr;
#define GB_PC(F, V) part_count_kernel<F, V><<<w.grid, INSERT_THREADS, 0, st>>>(rb, k, (unsigned int)pl.owners, pl.lp_bits, nb, w.cta_hist)
    if (fixed) { if (v210) GB_PC(true, true); else GB_PC(true, false); }
    else { if (v210) GB_PC(false, true); else GB_PC(false, false); }
#undef GB_PC
    GB_LAUNCHED();
    part_offsets_kernel<<<nb, 256, 0, st>>>(w.cta_hist, w.grid, nb, w.bucket_total);
    GB_LAUNCHED();
    part_bases_kernel<<<1, MAX_BUCKETS, 0, st>>>(w.bucket_total, nb, w.bucket_base);
    GB_LAUNCHED();
    return GB_OK;
}

int part_scatter(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, unsigned long long *out, cudaStream_t st)
{
    const unsigned int nb = (unsigned int)pl.nb();
    const bool fixed = rb.offsets == nullptr;
#define GB_PS(F, V) part_scatter_kernel<F, V><<<w.grid, INSERT_THREADS, 0, st>>>(rb, k, (unsigned int)pl.owners, pl.lp_bits, nb, w.cta_hist, w.bucket_base, out)
    if (fixed) { if (v210) GB_PS(true, true); else GB_PS(true, false); }
    else { if (v210) GB_PS(false, true); else GB_PS(false, false); }
#undef GB_PS
    GB_LAUNCHED();
    return GB_OK;
}

int insert_key_chunks(Map *m, const unsigned long long *d_keys, const unsigned long long *d_vstart, const unsigned long long *d_off,
                      int n_chunks, unsigned long long n_total, cudaStream_t st)
{
    if (!n_total) return GB_OK;
    unsigned long long grid = (n_total + IK_PER_CTA - 1) / IK_PER_CTA;
    insert_keys_kernel<<<(unsigned int)grid, IK_THREADS, 0, st>>>(d_keys, d_vstart, d_off, n_chunks, n_total, m->table, m->bits, m->d_counters);
    GB_LAUNCHED();
    return GB_OK;
}

} // namespace gb
