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

} // namespace gb

using namespace gb;

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
