// scan.cuh -- device-wide exclusive prefix sums and stream-ordered temporaries (no library calls).
#pragma once
#include "common.cuh"

namespace gb {

// temporary device array: from the current arena (common.cuh), else a plain cudaMalloc
template <typename T>
struct Tmp {
    T *p = nullptr;
    cudaStream_t s = nullptr;
    bool owned = false;
    Tmp() = default;
    Tmp(const Tmp &) = delete;
    Tmp &operator=(const Tmp &) = delete;
    ~Tmp() { release(); }
    int alloc(size_t n, cudaStream_t stream)
    {
        release();
        s = stream;
        const size_t bytes = (n ? n : 1) * sizeof(T);
        if (tl_arena) { owned = false; return tl_arena->alloc((void **)&p, bytes); }
        owned = true;
        GB_CUDA(cudaMalloc((void **)&p, bytes));
        return GB_OK;
    }
    int zero(size_t n)
    {
        GB_CUDA(cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), s));
        return GB_OK;
    }
    int fill_ff(size_t n)
    {
        GB_CUDA(cudaMemsetAsync(p, 0xFF, (n ? n : 1) * sizeof(T), s));
        return GB_OK;
    }
    void release()
    {
        if (p && owned) { cudaStreamSynchronize(s); cudaFree(p); }
        p = nullptr;
    }
};

#ifdef __CUDACC__
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// in-place exclusive scan of one 1024-element tile per CTA; tile totals go to tile_sums (optional)
static __global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_kernel(unsigned long long *data, unsigned long long n, unsigned long long *tile_sums)
{
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_TILE + (unsigned long long)threadIdx.x * SCAN_ITEMS;
    unsigned long long v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        v[j] = base + j < n ? data[base + j] : 0;
        sum += v[j];
    }
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < SCAN_THREADS / 32 ? s_warp[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < SCAN_THREADS / 32) s_warp[lane] = wi - w;
        if (lane == SCAN_THREADS / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    unsigned long long run = s_warp[warp] + incl - sum;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        if (base + j < n) data[base + j] = run;
        run += v[j];
    }
}

static __global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(unsigned long long *data, unsigned long long n, const unsigned long long *tile_offsets)
{
    const unsigned long long base = (unsigned long long)blockIdx.x * SCAN_TILE + (unsigned long long)threadIdx.x * SCAN_ITEMS;
    const unsigned long long o = tile_offsets[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++)
        if (base + j < n) data[base + j] += o;
}

// data[0..n) -> exclusive prefix sums in place; *d_total (device, optional) = sum of all inputs
static int exclusive_scan_u64(unsigned long long *data, unsigned long long n, unsigned long long *d_total, cudaStream_t s)
{
    if (n == 0) {
        if (d_total) GB_CUDA(cudaMemsetAsync(d_total, 0, 8, s));
        return GB_OK;
    }
    unsigned long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        scan_tile_kernel<<<1, SCAN_THREADS, 0, s>>>(data, n, d_total);
        GB_LAUNCHED();
        return GB_OK;
    }
    Tmp<unsigned long long> sums;
    GB_TRY(sums.alloc(tiles, s));
    scan_tile_kernel<<<(unsigned int)tiles, SCAN_THREADS, 0, s>>>(data, n, sums.p);
    GB_LAUNCHED();
    GB_TRY(exclusive_scan_u64(sums.p, tiles, d_total, s));
    scan_add_kernel<<<(unsigned int)tiles, SCAN_THREADS, 0, s>>>(data, n, sums.p);
    GB_LAUNCHED();
    return GB_OK;
}

// per-CTA aggregated allocation of `count` consecutive indices from a global counter (all threads call);
// with counter == nullptr it is a plain CTA-wide exclusive prefix sum
__device__ __forceinline__ unsigned long long block_alloc(unsigned int count, unsigned long long *counter)
{
    __shared__ unsigned int s_wsum[32];
    __shared__ unsigned long long s_base;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    unsigned int incl = count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    __syncthreads(); // protects s_wsum / s_base against a previous call
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = lane < nwarp ? s_wsum[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned int t = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= d) wi += t;
        }
        s_wsum[lane] = wi - w;
        if (lane == 31) s_base = (counter && wi) ? atomicAdd(counter, (unsigned long long)wi) : 0;
    }
    __syncthreads();
    return s_base + s_wsum[warp] + incl - count;
}
#endif

} // namespace gb
