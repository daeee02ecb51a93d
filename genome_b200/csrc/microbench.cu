// microbench.cu -- measurement kernels behind gb_bench_* (diagnostics of include/genome_b200.h; nothing here is on the data path).
// gb_bench_smem_upsert answers one design question with a number: how fast is update(key, 1, _ + 1) when the
// table slice lives in SHARED memory -- one CTA per fine bucket, the slice initialised on chip, every key of the bucket upserted
// with shared-memory atomics, the slice streamed out once -- compared with the L2-atomics upsert of partition.cu.
#include <algorithm>

#include "common.cuh"

namespace gb {

// the 16-byte record-per-slot layout of rounds 1 and 2a, kept here as the thing the measurements compare against
struct __align__(16) Rec {
    unsigned long long key;
    int count;
    unsigned int vid;
};
__device__ __forceinline__ unsigned long long load_rec_key16(const Rec *p)
{
    unsigned int lo, hi, c, v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(hi), "=r"(c), "=r"(v) : "l"(p));
    return (((unsigned long long)hi << 32) | lo) + c + v;
}

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
                         int slots_log2, Rec *out, unsigned long long *new_keys)
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
// 16-byte slot of a region.  `mode` = base + 100 * log2(updates in flight per thread; 0 = the upsert's 4) + 1000 * P
// (P > 0: a persistent grid of P CTAs per SM, software-pipelined: the loads of round r + 1 are issued before the atomics of round r).
// base 1: 8-byte load only; 2: 4-byte red.add only; 3: load, then a red that depends on the loaded value (what the upsert does for a
// key that is already there); 4: load, then 64-bit CAS, then red (a new key); 5: load and an INDEPENDENT red on the same slot;
// 6: load, then a dependent red on ANOTHER random slot; 7: atom.add with a return value, no load; 8: 16-byte load then dependent red;
// 9: load, then a dependent plain 4-byte store; 10: load from the first half of the region, dependent red into the second half (no
// sector is both read and written); 11: like 5, but the SMs with an even %smid only load and the odd ones only red; 12: the same
// split by warp parity inside every CTA; 13: structure-of-arrays addressing -- 8-byte load from a key array, dependent red into a
// separate 4-byte count array; 14: line-blocked layout -- a 128-byte line holds 8 keys | 8 counts | 8 vertex ids, load the key, dependent
// red on the count in the same line but another sector; 15 / 16: layouts of 13 / 14 with a 64-bit CAS on the key between the two.
__device__ __forceinline__ unsigned int smid()
{
    unsigned int r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}

__device__ __forceinline__ unsigned long long *blocked_key(Rec *table, unsigned long long i)
{
    return reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(table) + (i >> 3) * 128 + (i & 7) * 8);
}

template <int KPT>
__device__ __forceinline__ void l2_round(Rec *table, unsigned long long slots, const long long (&i)[KPT], const bool (&ok)[KPT], int mode,
                                         unsigned long long seed, unsigned long long (&idx)[KPT], unsigned long long (&cur)[KPT])
{
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        idx[j] = __umul64hi(mix64((unsigned long long)i[j] * 0x9E3779B97F4A7C15ull + seed), slots);
        cur[j] = 1;
    }
    if (mode == 2 || mode == 7) return;
    if (mode == 10) {
#pragma unroll
        for (int j = 0; j < KPT; j++) idx[j] >>= 1; // first half
    }
    if ((mode == 11 && (smid() & 1)) || (mode == 12 && ((threadIdx.x >> 5) & 1))) return; // this SM / warp only writes
#pragma unroll
    for (int j = 0; j < KPT; j++)
        if (ok[j]) {
            if (mode == 8) cur[j] = load_rec_key16(table + idx[j]);
            else if (mode == 13 || mode == 15) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(cur[j]) : "l"(reinterpret_cast<unsigned long long *>(table) + idx[j]));
            else if (mode == 14 || mode == 16) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(cur[j]) : "l"(blocked_key(table, idx[j])));
            else asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(cur[j]) : "l"(&table[idx[j]].key));
        }
}

template <int KPT>
__device__ __forceinline__ unsigned long long l2_finish(Rec *table, unsigned long long slots, const bool (&ok)[KPT], int mode,
                                                        const unsigned long long (&idx)[KPT], const unsigned long long (&cur)[KPT])
{
    unsigned long long acc = 0;
    if (mode == 4 || mode == 15 || mode == 16) {
#pragma unroll
        for (int j = 0; j < KPT; j++)
            if (ok[j]) {
                unsigned long long *kp = mode == 4 ? &table[idx[j]].key : mode == 15 ? reinterpret_cast<unsigned long long *>(table) + idx[j] : blocked_key(table, idx[j]);
                acc += atomicCAS(kp, cur[j], cur[j] + 1);
            }
    }
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        if (!ok[j]) continue;
        const int dep = (int)(cur[j] & 1) + 1;
        switch (mode) {
        case 1: acc += cur[j]; break;
        case 2: red_add_s32(&table[idx[j]].count, 1); break;
        case 5: red_add_s32(&table[idx[j]].count, 1); break; // cur[j] is consumed only by the caller's sink
        case 6: red_add_s32(&table[slots - 1 - idx[j]].count, dep); break;
        case 7: acc += (unsigned int)atomicAdd(&table[idx[j]].count, 1); break;
        case 9: asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(&table[idx[j]].count), "r"(dep) : "memory"); break;
        case 10: red_add_s32(&table[(slots >> 1) + idx[j]].count, dep); break;
        case 11: if (smid() & 1) red_add_s32(&table[idx[j]].count, 1); break;
        case 12: if ((threadIdx.x >> 5) & 1) red_add_s32(&table[idx[j]].count, 1); break;
        case 14: case 16: red_add_s32(reinterpret_cast<int *>(reinterpret_cast<char *>(table) + (idx[j] >> 3) * 128 + 64 + (idx[j] & 7) * 4), dep); break;
        case 15:
        case 13: red_add_s32(reinterpret_cast<int *>(reinterpret_cast<unsigned long long *>(table) + slots) + idx[j], dep); break;
        default: red_add_s32(&table[idx[j]].count, dep); break; // 3, 4, 8
        }
        if (mode == 5) acc += cur[j];
    }
    return acc;
}

template <int KPT>
__global__ void __launch_bounds__(256)
bench_l2_requests_kernel(Rec *table, unsigned long long slots, long long n, int mode, unsigned long long seed, unsigned long long *sink)
{
    const long long stride = (long long)gridDim.x * 256;
    long long i[KPT];
    unsigned long long idx[KPT], cur[KPT];
    bool ok[KPT];
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        i[j] = ((long long)blockIdx.x * 256 + threadIdx.x) + j * stride;
        ok[j] = i[j] < n;
    }
    l2_round<KPT>(table, slots, i, ok, mode, seed, idx, cur);
    const unsigned long long acc = l2_finish<KPT>(table, slots, ok, mode, idx, cur);
    if (acc == 0x123456789ull) *sink = acc; // keeps the loads alive
}

// persistent, software-pipelined form: round r + 1's loads are in flight while round r's atomics are issued
template <int KPT>
__global__ void __launch_bounds__(256)
bench_l2_requests_pipe_kernel(Rec *table, unsigned long long slots, long long n, int mode, unsigned long long seed, unsigned long long *sink)
{
    const long long per_round = (long long)gridDim.x * 256 * KPT;
    long long i[KPT], i2[KPT];
    unsigned long long idx[KPT], cur[KPT], idx2[KPT], cur2[KPT], acc = 0;
    bool ok[KPT], ok2[KPT];
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        i[j] = ((long long)blockIdx.x * 256 + threadIdx.x) + (long long)j * gridDim.x * 256;
        ok[j] = i[j] < n;
    }
    l2_round<KPT>(table, slots, i, ok, mode, seed, idx, cur);
    for (long long base = 0; base < n; base += per_round) {
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            i2[j] = i[j] + per_round;
            ok2[j] = i2[j] < n;
        }
        l2_round<KPT>(table, slots, i2, ok2, mode, seed, idx2, cur2);
        acc += l2_finish<KPT>(table, slots, ok, mode, idx, cur);
#pragma unroll
        for (int j = 0; j < KPT; j++) { i[j] = i2[j]; ok[j] = ok2[j]; idx[j] = idx2[j]; cur[j] = cur2[j]; }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

} // namespace gb

using namespace gb;

template <int KPT>
static void launch_l2_requests(Rec *t, unsigned long long slots, long long n, int base, int persist, unsigned long long seed, unsigned long long *sink)
{
    if (persist) bench_l2_requests_pipe_kernel<KPT><<<SM_COUNT * persist, 256>>>(t, slots, n, base, seed, sink);
    else bench_l2_requests_kernel<KPT><<<(unsigned int)((n + 256 * KPT - 1) / (256 * KPT)), 256>>>(t, slots, n, base, seed, sink);
}

extern "C" int gb_bench_l2_requests(int device, size_t region_bytes, int64_t n_updates, int mode, int iters, int64_t *ns_per_iter)
{
    const int base = mode % 100, kpt_log2 = (mode / 100) % 10, persist = mode / 1000;
    if (!ns_per_iter || region_bytes < 1024 || n_updates <= 0 || iters <= 0 || base < 1 || base > 16 || kpt_log2 > 4 || persist > 16) {
        set_error("bad arguments");
        return GB_E_ARG;
    }
    GB_CUDA(cudaSetDevice(device));
    const unsigned long long slots = region_bytes / sizeof(Rec);
    Rec *t = nullptr;
    unsigned long long *sink = nullptr;
    GB_CUDA(cudaMalloc((void **)&t, slots * sizeof(Rec)));
    GB_CUDA(cudaMalloc((void **)&sink, 8));
    GB_CUDA(cudaMemset(t, 0, slots * sizeof(Rec)));
    cudaEvent_t e0, e1;
    GB_CUDA(cudaEventCreate(&e0));
    GB_CUDA(cudaEventCreate(&e1));
    auto run = [&](unsigned long long seed) {
        switch (kpt_log2) {
        case 1: launch_l2_requests<2>(t, slots, n_updates, base, persist, seed, sink); break;
        case 3: launch_l2_requests<8>(t, slots, n_updates, base, persist, seed, sink); break;
        case 4: launch_l2_requests<16>(t, slots, n_updates, base, persist, seed, sink); break;
        default: launch_l2_requests<4>(t, slots, n_updates, base, persist, seed, sink); break; // 0 and 2
        }
    };
    run(1);
    GB_LAUNCHED();
    GB_CUDA(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) {
        run(2 + i);
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
    Rec *out = nullptr;
    GB_CUDA(cudaMalloc((void **)&keys, n * 8));
    GB_CUDA(cudaMalloc((void **)&out, slots * (size_t)n_buckets * sizeof(Rec)));
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
