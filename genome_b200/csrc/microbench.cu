// microbench.cu -- measurement kernels behind gb_bench_* (diagnostics of include/genome_b200.h; nothing here is on the data path).
// gb_bench_smem_upsert answers one design question with a number: how fast is update(key, 1, _ + 1) when the
// table slice lives in SHARED memory -- one CTA per fine bucket, the slice initialised on chip, every key of the bucket upserted
// with shared-memory atomics, the slice streamed out once -- compared with the L2-atomics upsert of partition.cu.
#include <algorithm>

#include "common.cuh"

namespace gb {

// synthetic bucket contents with C2's multiplicity profile: 31 % of the instances are keys seen once, the rest hit a pool of
// keys_per_bucket / 20 recurring keys (about 14 instances each)
__global__ void __launch_bounds__(256)
bench_fill_bucket_keys_kernel(unsigned long long *keys, unsigned long long keys_per_bucket, unsigned long long n_buckets, unsigned long long seed)
{
    const unsigned long long n = keys_per_bucket * n_buckets;
    const unsigned long long pool = keys_per_bucket / 20 ? keys_per_bucket / 20 : 1;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long b = i / keys_per_bucket, r = mix64(i * 0x9E3779B97F4A7C15ull + seed);
        const bool once = (r & 1023) < 317;
        const unsigned long long id = once ? (1ull << 40) + i : (b << 20) + (r >> 10) % pool;
        keys[i] = mix64(id + seed) >> 2; // 62-bit key, never EMPTY_KEY
    }
}

template <int KPT>
__global__ void __launch_bounds__(512)
bench_smem_upsert_kernel(const unsigned long long *__restrict__ keys, unsigned long long keys_per_bucket, unsigned long long n_buckets,
                         int slots_log2, Slot *out, unsigned long long *new_keys)
{
    extern __shared__ unsigned long long s_mem[];
    const unsigned int slots = 1u << slots_log2, mask = slots - 1;
    unsigned long long *skey = s_mem;
    unsigned int *scnt = reinterpret_cast<unsigned int *>(s_mem + slots);
    unsigned int nk = 0;
    for (unsigned long long b = blockIdx.x; b < n_buckets; b += gridDim.x) {
        for (unsigned int i = threadIdx.x; i < slots; i += blockDim.x) { skey[i] = EMPTY_KEY; scnt[i] = 0; }
        __syncthreads();
        const unsigned long long *kb = keys + b * keys_per_bucket;
        for (unsigned long long i0 = 0; i0 < keys_per_bucket; i0 += (unsigned long long)blockDim.x * KPT) {
            unsigned long long key[KPT];
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const unsigned long long i = i0 + (unsigned long long)j * blockDim.x + threadIdx.x;
                key[j] = i < keys_per_bucket ? __ldcs(kb + i) : EMPTY_KEY;
            }
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                if (key[j] == EMPTY_KEY) continue;
                unsigned int s = (unsigned int)(mix64(key[j]) >> 40) & mask;
                for (;;) {
                    unsigned long long cur = skey[s];
                    if (cur == EMPTY_KEY) {
                        cur = atomicCAS(&skey[s], EMPTY_KEY, key[j]);
                        if (cur == EMPTY_KEY) { nk++; cur = key[j]; }
                    }
                    if (cur == key[j]) { atomicAdd(&scnt[s], 1u); break; }
                    s = (s + 1) & mask;
                }
            }
        }
        __syncthreads();
        uint4 *o = reinterpret_cast<uint4 *>(out + b * slots);
        for (unsigned int i = threadIdx.x; i < slots; i += blockDim.x) {
            const unsigned long long kk = skey[i];
            o[i] = make_uint4((unsigned int)kk, (unsigned int)(kk >> 32), scnt[i], NONE32);
        }
        __syncthreads();
    }
    nk = __reduce_add_sync(0xFFFFFFFFu, nk);
    if ((threadIdx.x & 31) == 0 && nk) atomicAdd(new_keys, (unsigned long long)nk);
}

// request-rate probe of the L2-atomics upsert's access pattern, hashing and probing stripped away: every update touches one random
// 16-byte slot of a region that fits L2.  mode 1: 8-byte load only; 2: 4-byte red.add only; 3: load then red (what the upsert does
// for a key that is already there); 4: load, then 64-bit CAS, then red (a new key).  4 updates per thread in flight like the upsert.
__global__ void __launch_bounds__(256)
bench_l2_requests_kernel(Slot *table, unsigned long long slots, long long n, int mode, unsigned long long seed, unsigned long long *sink)
{
    const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x);
    const long long stride = (long long)gridDim.x * 256;
    unsigned long long idx[4], cur[4], acc = 0;
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const long long i = i0 + j * stride;
        ok[j] = i < n;
        idx[j] = __umul64hi(mix64((unsigned long long)i * 0x9E3779B97F4A7C15ull + seed), slots);
    }
    if (mode != 2) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (ok[j]) {
                asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(cur[j]) : "l"(&table[idx[j]].key));
                acc += cur[j];
            }
    }
    if (mode == 4) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (ok[j]) acc += atomicCAS(&table[idx[j]].key, cur[j], cur[j] + 1);
    }
    if (mode >= 2) {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (ok[j]) red_add_s32(&table[idx[j]].count, mode == 2 ? 1 : (int)(cur[j] & 1) + 1);
    }
    if (acc == 0x123456789ull) *sink = acc; // keeps the loads alive
}

} // namespace gb

using namespace gb;

extern "C" int gb_bench_l2_requests(int device, size_t region_bytes, int64_t n_updates, int mode, int iters, int64_t *ns_per_iter)
{
    if (!ns_per_iter || region_bytes < 1024 || n_updates <= 0 || iters <= 0 || mode < 1 || mode > 4) { set_error("bad arguments"); return GB_E_ARG; }
    GB_CUDA(cudaSetDevice(device));
    const unsigned long long slots = region_bytes / sizeof(Slot);
    Slot *t = nullptr;
    unsigned long long *sink = nullptr;
    GB_CUDA(cudaMalloc((void **)&t, slots * sizeof(Slot)));
    GB_CUDA(cudaMalloc((void **)&sink, 8));
    GB_CUDA(cudaMemset(t, 0, slots * sizeof(Slot)));
    cudaEvent_t e0, e1;
    GB_CUDA(cudaEventCreate(&e0));
    GB_CUDA(cudaEventCreate(&e1));
    const unsigned int grid = (unsigned int)((n_updates + 1023) / 1024);
    bench_l2_requests_kernel<<<grid, 256>>>(t, slots, n_updates, mode, 1, sink);
    GB_LAUNCHED();
    GB_CUDA(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) {
        bench_l2_requests_kernel<<<grid, 256>>>(t, slots, n_updates, mode, 2 + i, sink);
        GB_LAUNCHED();
    }
    GB_CUDA(cudaEventRecord(e1));
    GB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    GB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ns_per_iter = (int64_t)(ms * 1e6 / iters);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(t);
    cudaFree(sink);
    return GB_OK;
}

extern "C" int gb_bench_smem_upsert(int device, int slots_log2, int64_t keys_per_bucket, int64_t n_buckets, int ctas_per_sm, int iters,
                                    int64_t *ns_per_iter, int64_t *distinct_keys)
{
    if (!ns_per_iter || slots_log2 < 8 || slots_log2 > 14 || keys_per_bucket < 1 || n_buckets < 1 || iters < 1 || ctas_per_sm < 1) {
        set_error("bad arguments");
        return GB_E_ARG;
    }
    GB_CUDA(cudaSetDevice(device));
    const size_t slots = (size_t)1 << slots_log2, smem = slots * 12;
    const size_t n = (size_t)keys_per_bucket * (size_t)n_buckets;
    unsigned long long *keys = nullptr, *d_new = nullptr;
    Slot *out = nullptr;
    GB_CUDA(cudaMalloc((void **)&keys, n * 8));
    GB_CUDA(cudaMalloc((void **)&out, slots * (size_t)n_buckets * sizeof(Slot)));
    GB_CUDA(cudaMalloc((void **)&d_new, 8));
    GB_CUDA(cudaMemset(d_new, 0, 8));
    bench_fill_bucket_keys_kernel<<<SM_COUNT * 8, 256>>>(keys, (unsigned long long)keys_per_bucket, (unsigned long long)n_buckets, 12345);
    GB_LAUNCHED();
    GB_CUDA(cudaFuncSetAttribute(bench_smem_upsert_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    GB_CUDA(cudaEventCreate(&e0));
    GB_CUDA(cudaEventCreate(&e1));
    const unsigned int grid = (unsigned int)std::min<int64_t>(n_buckets, (int64_t)SM_COUNT * ctas_per_sm);
    bench_smem_upsert_kernel<4><<<grid, 512, smem>>>(keys, (unsigned long long)keys_per_bucket, (unsigned long long)n_buckets, slots_log2, out, d_new);
    GB_LAUNCHED();
    unsigned long long h_new = 0;
    GB_CUDA(cudaMemcpy(&h_new, d_new, 8, cudaMemcpyDeviceToHost));
    GB_CUDA(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) {
        bench_smem_upsert_kernel<4><<<grid, 512, smem>>>(keys, (unsigned long long)keys_per_bucket, (unsigned long long)n_buckets, slots_log2, out, d_new);
        GB_LAUNCHED();
    }
    GB_CUDA(cudaEventRecord(e1));
    GB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    GB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ns_per_iter = (int64_t)(ms * 1e6 / iters);
    if (distinct_keys) *distinct_keys = (int64_t)h_new;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(keys);
    cudaFree(out);
    cudaFree(d_new);
    return GB_OK;
}
