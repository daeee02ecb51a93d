// map.cu -- DNAMap[Int] on one B200: open-addressing HBM hash table, bulk FreqFilter.add, filter, export.
// Replaces S/ds/ArrayDNAMap.scala:62-243 and the insert loop of S/data/FreqFilter.scala:25-58
// (paths relative to /root/reference, S/ = src/main/scala/ru/ifmo/genome/).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "extract.cuh"
#include "partition.cuh"
#include "scan.cuh"

namespace gb {

// ================================================================ errors
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return e == cudaErrorMemoryAllocation ? GB_E_OOM : GB_E_CUDA;
}

thread_local Arena *tl_arena = nullptr;
Tuning g_tune;

// ---- arena block cache (see common.cuh): at most CACHE_SLOTS blocks, best fit, per device
namespace {
struct CachedBlock { void *p; size_t bytes; int device; };
constexpr int CACHE_SLOTS = 16;
CachedBlock g_cache[CACHE_SLOTS];
int g_cache_n = 0;
std::mutex g_cache_mu;
}
void *arena_cache_get(size_t bytes, size_t *got)
{
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_cache_mu);
    int best = -1;
    for (int i = 0; i < g_cache_n; i++)
        if (g_cache[i].device == dev && g_cache[i].bytes >= bytes && (best < 0 || g_cache[i].bytes < g_cache[best].bytes)) best = i;
    if (best < 0 || g_cache[best].bytes > 4 * bytes + (64u << 20)) return nullptr; // do not burn a huge block on a small request
    void *p = g_cache[best].p;
    *got = g_cache[best].bytes;
    g_cache[best] = g_cache[--g_cache_n];
    return p;
}
void arena_cache_put(void *p, size_t bytes)
{
    if (!p) return;
    int dev = 0;
    cudaGetDevice(&dev);
    void *victim = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        if (g_cache_n < CACHE_SLOTS) {
            g_cache[g_cache_n++] = { p, bytes, dev };
        } else { // evict the smallest block
            int small = 0;
            for (int i = 1; i < g_cache_n; i++)
                if (g_cache[i].bytes < g_cache[small].bytes) small = i;
            if (g_cache[small].bytes < bytes) { victim = g_cache[small].p; g_cache[small] = { p, bytes, dev }; }
            else victim = p;
        }
    }
    if (victim) cudaFree(victim);
}
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// one-time per-device settings
int pool_setup(int)
{
    // Random 16-byte slot accesses use one 32-byte sector of a 128-byte L2 line: ask L2 not to fetch the neighbours
    // (a hint; on B200 it changes nothing measurable: ncu r1a shows 128 B per miss either way).
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
    cudaGetLastError();
    return GB_OK;
}

// ================================================================ kernels

// EMPTY table: keys = all ones, counts = 1 (!); with_vid: vertex ids = NONE too.  The count word of a free slot already holds the
// 1 of the key that will claim it, so update(key, 1, _ + 1) of a NEW key is the claiming compare-and-swap alone, without a red
// behind it -- 31 % of C2's k-mer instances are new keys (upsert_add, extract.cuh; every other writer stores the count outright).  n is a multiple of 1024 (cap_for), so every array
// is a whole number of 16-byte vectors.  The vertex ids of a counting table are never read before they are assigned
// (assign_vertices_kernel / place_distinct_kernel write them for every vertex), so the clear of a FreqFilter pass moves 12 bytes
// per slot, not 16.
__global__ void init_table_kernel(Table t, bool with_vid)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x, i0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), first = make_uint4(1u, 1u, 1u, 1u);
    for (unsigned long long i = i0; i < t.cap / 2; i += stride) reinterpret_cast<uint4 *>(t.key)[i] = ones;
    for (unsigned long long i = i0; i < t.cap / 4; i += stride) reinterpret_cast<uint4 *>(t.count)[i] = first;
    if (with_vid)
        for (unsigned long long i = i0; i < t.cap / 4; i += stride) reinterpret_cast<uint4 *>(t.vid)[i] = ones;
}

// ---------------------------------------------------------------- bulk FreqFilter.add
// Tile staging and window extraction live in extract.cuh; here all SEG table probes of a work item are issued
// before the first one is consumed (memory-level parallelism for the random 32-byte sector reads).
template <bool FIXED, bool V210>
__global__ void __launch_bounds__(INSERT_THREADS)
insert_reads_kernel(const uint8_t *__restrict__ bin, unsigned long long n_bytes,
                    const unsigned long long *__restrict__ offsets, unsigned int rec_bytes,
                    long long read0, long long n_reads, int k, Table table, unsigned long long *counters)
{
    __shared__ ReadTile tile;
    __shared__ unsigned int s_newkeys;
    const int tid = threadIdx.x;
    const int nr = stage_tile<FIXED>(tile, bin, n_bytes, offsets, rec_bytes, read0, n_reads, k, blockIdx.x);
    if (nr <= 0) return;
    if (tid == 0) s_newkeys = 0;

    const unsigned int total_items = tile.prefix[TILE_READS];
    int newkeys = 0;

    for (unsigned int item = tid; item < total_items; item += INSERT_THREADS) {
        unsigned long long key[SEG], idx[SEG], cur[SEG];
        const int cnt = item_keys<V210>(tile, item, k, key);
#pragma unroll
        for (int j = 0; j < SEG; j++) idx[j] = slot_of(mix64(key[j]), table.cap);
#pragma unroll
        for (int j = 0; j < SEG; j++)
            if (j < cnt) cur[j] = load_key(table, idx[j]);
#pragma unroll
        for (int j = 0; j < SEG; j++)
            if (j < cnt) newkeys += upsert_add(table, idx[j], cur[j], key[j], 1);
    }

    // one global atomic per CTA for the size counter
    newkeys = __reduce_add_sync(0xFFFFFFFFu, newkeys);
    __syncthreads();
    if ((tid & 31) == 0 && newkeys) atomicAdd(&s_newkeys, (unsigned int)newkeys);
    __syncthreads();
    if (tid == 0) {
        if (s_newkeys) atomicAdd(&counters[0], (unsigned long long)s_newkeys);
        unsigned long long w = 0;
        for (int i = 0; i < nr; i++) w += tile.nwin[i];
        atomicAdd(&counters[3], w);
    }
}

// empirical random-access ceiling (R_gups): ONE random 64-bit atomicAdd (no return value: a RED) per element
// over an array of the table's size; no hashing of keys, no probing, no preceding load; 8 updates per thread in flight.
__global__ void __launch_bounds__(256)
random_atomics_kernel(unsigned long long *words, unsigned long long n_words, long long n, unsigned long long seed)
{
    long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (i0 + j < n) {
            const unsigned long long idx = __umul64hi(mix64((unsigned long long)(i0 + j) * 0x9E3779B97F4A7C15ull + seed), n_words);
            atomicAdd(words + idx, 1ull);
        }
}

// every record of a fixed-stride stream must carry the same length byte; counters[2] != 0 otherwise
__global__ void verify_fixed_kernel(const uint8_t *__restrict__ bin, unsigned int rec_bytes, unsigned int len0,
                                    long long n_reads, unsigned long long *counters)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    int bad = 0;
    for (; i < n_reads; i += stride) bad |= bin[(unsigned long long)i * rec_bytes] != len0;
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicOr(&counters[2], 1ull);
}

// records at a fixed stride with their own length bytes (gb_map_insert_records_device): every length must fit the stride and
// stay at or below max_len; counters[2] != 0 otherwise
__global__ void verify_records_kernel(const uint8_t *__restrict__ bin, unsigned int rec_bytes, unsigned int max_len, long long n_reads,
                                      unsigned long long *counters)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    int bad = 0;
    for (; i < n_reads; i += stride) {
        const unsigned int len = bin[(unsigned long long)i * rec_bytes];
        bad |= len > max_len || 1 + (len + 3) / 4 > rec_bytes;
    }
    if (__any_sync(0xFFFFFFFFu, bad) && (threadIdx.x & 31) == 0) atomicOr(&counters[2], 1ull);
}

// update(key, 1, _ + 1) / update(key, v) for explicit keys
// SET: update(key, v); set_vid: additionally slot.vid = the key's index in `keys` (keys must be distinct then)
template <bool SET>
__global__ void update_keys_kernel(const unsigned long long *__restrict__ keys, const int *__restrict__ vals,
                                   long long n, Table table, unsigned long long *counters, bool set_vid = false,
                                   uint8_t *fp = nullptr)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int nk = 0;
    if (i < n) {
        unsigned long long key = keys[i];
        const unsigned long long h = mix64(key);
        const unsigned long long cap = table.cap;
        unsigned long long idx = slot_of(h, cap);
        if (!SET) {
            nk = upsert_add(table, idx, load_key(table, idx), key, 1);
        } else {
            for (;;) {
                unsigned long long cur = load_key(table, idx);
                if (cur == EMPTY_KEY) {
                    cur = atomicCAS(table.key + idx, EMPTY_KEY, key);
                    if (cur == EMPTY_KEY) { nk = 1; cur = key; }
                }
                if (cur == key) {
                    atomicExch(table.count + idx, vals[i]);
                    if (set_vid) table.vid[idx] = (unsigned int)i;
                    if (fp) fp[idx] = (uint8_t)fp_tag(h);
                    break;
                }
                idx = next_slot(idx, cap);
            }
        }
    }
    nk = __reduce_add_sync(0xFFFFFFFFu, nk);
    if ((threadIdx.x & 31) == 0 && nk) atomicAdd(&counters[0], (unsigned long long)nk);
}

// The survivors of deleteAll (or a replica's gathered keys) into an EMPTY table: distinct keys, index = vertex id.  The table is
// at load <= 1/3 and starts empty, so the compare-and-swap IS the probe (no preceding load); count and vertex id are one 8-byte
// store; two keys per thread in flight.  Replaces update_keys_kernel<true> here: 5 L2 requests per key (load, CAS, exchange,
// two stores) became 3 (measured on C2's 4.6 M survivors: 0.51 ms before).
__global__ void __launch_bounds__(256)
place_distinct_kernel(const unsigned long long *__restrict__ keys, const int *__restrict__ vals, long long n, Table table,
                      uint8_t *fp, unsigned long long *counters)
{
    const unsigned long long cap = table.cap;
    const long long i0 = ((long long)blockIdx.x * 256 + threadIdx.x);
    const long long stride = (long long)gridDim.x * 256;
    unsigned long long key[2], h[2], idx[2], old[2];
    bool ok[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const long long i = i0 + j * stride;
        ok[j] = i < n;
        if (ok[j]) {
            key[j] = keys[i];
            h[j] = mix64(key[j]);
            idx[j] = slot_of(h[j], cap);
        }
    }
#pragma unroll
    for (int j = 0; j < 2; j++)
        if (ok[j]) old[j] = atomicCAS(table.key + idx[j], EMPTY_KEY, key[j]);
    int nk = 0;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        if (!ok[j]) continue;
        const long long i = i0 + j * stride;
        while (old[j] != EMPTY_KEY) { // someone else's slot: linear probing (keys are distinct: old is never this key)
            idx[j] = next_slot(idx[j], cap);
            old[j] = load_key(table, idx[j]);
            if (old[j] == EMPTY_KEY) old[j] = atomicCAS(table.key + idx[j], EMPTY_KEY, key[j]);
        }
        table.count[idx[j]] = vals[i]; // the slot is ours alone
        table.vid[idx[j]] = (unsigned int)i;
        fp[idx[j]] = (uint8_t)fp_tag(h[j]);
        nk++;
    }
    nk = __reduce_add_sync(0xFFFFFFFFu, nk);
    if ((threadIdx.x & 31) == 0 && nk) atomicAdd(&counters[0], (unsigned long long)nk);
}

__global__ void lookup_kernel(const unsigned long long *__restrict__ keys, long long n, Table table, int *counts, uint8_t *found)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long at = probe_find(table, keys[i]);
    const bool f = at >= 0;
    if (counts) counts[i] = f ? load_count(table, (unsigned long long)at) : 0;
    if (found) found[i] = f;
}

// deleteAll(v < min_count) (ArrayDNAMap.scala:164-173,212-215): ONE streaming pass over the table compacts the
// survivors into (key, count) arrays (warp-aggregated cursor), then they are re-inserted into a table sized for
// them (the reference tombstones and rescales when the load drops below 0.3, ArrayDNAMap.scala:217-230).
__global__ void __launch_bounds__(256)
compact_survivors_kernel(Table table, int min_count, unsigned long long *out_keys, int *out_vals, unsigned long long *counters)
{
    const unsigned long long n = table.cap;
    // a CTA iteration covers 1024 consecutive slots (4 per thread) and takes ONE ticket from the global cursor
    const unsigned long long tiles = (n + 1023) / 1024;
    for (unsigned long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        Slot s[4];
        unsigned int c = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned long long i = t * 1024 + (unsigned long long)j * 256 + threadIdx.x;
            s[j].key = EMPTY_KEY;
            s[j].count = 0;
            if (i < n) { s[j].key = __ldcs(table.key + i); s[j].count = __ldcs(table.count + i); }
            c += s[j].key != EMPTY_KEY && s[j].count >= min_count;
        }
        unsigned long long pos = block_alloc(c, &counters[1]);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (s[j].key != EMPTY_KEY && s[j].count >= min_count) {
                out_keys[pos] = s[j].key;
                out_vals[pos] = s[j].count;
                pos++;
            }
    }
}

__global__ void rehash_kernel(Table old_table, int min_count, bool filter, Table new_table)
{
    const unsigned long long n = old_table.cap, new_cap = new_table.cap;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        Slot s;
        s.key = __ldcs(old_table.key + i);
        if (s.key == EMPTY_KEY) continue;
        s.count = __ldcs(old_table.count + i);
        if (filter && s.count < min_count) continue;
        unsigned long long idx = slot_of(mix64(s.key), new_cap);
        for (;;) {
            unsigned long long cur = load_key(new_table, idx);
            if (cur == EMPTY_KEY && atomicCAS(new_table.key + idx, EMPTY_KEY, s.key) == EMPTY_KEY) {
                new_table.count[idx] = s.count;
                break;
            }
            idx = next_slot(idx, new_cap);
        }
    }
}

// mapReduce/foreach export: compaction with one global atomic per CTA
__global__ void export_kernel(Table table, unsigned long long *keys, int *vals, unsigned long long cap, unsigned long long *counters)
{
    const unsigned long long n = table.cap;
    __shared__ unsigned int s_n;
    __shared__ unsigned long long s_base;
    unsigned long long tiles = (n + blockDim.x - 1) / blockDim.x;
    for (unsigned long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        unsigned long long i = t * blockDim.x + threadIdx.x;
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        Slot s;
        s.key = EMPTY_KEY;
        s.count = 0;
        if (i < n) { s.key = __ldcs(table.key + i); s.count = __ldcs(table.count + i); }
        bool live = s.key != EMPTY_KEY;
        unsigned int ballot = __ballot_sync(0xFFFFFFFFu, live);
        unsigned int lane = threadIdx.x & 31, wbase = 0;
        if (lane == 0 && ballot) wbase = atomicAdd(&s_n, __popc(ballot));
        wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
        __syncthreads();
        if (threadIdx.x == 0) s_base = s_n ? atomicAdd(&counters[1], (unsigned long long)s_n) : 0;
        __syncthreads();
        if (live) {
            unsigned long long o = s_base + wbase + __popc(ballot & ((1u << lane) - 1));
            if (o < cap) {
                if (keys) keys[o] = s.key;
                if (vals) vals[o] = s.count;
            }
        }
        __syncthreads();
    }
}

} // namespace gb

using namespace gb;

// ================================================================ host side

namespace gb {

// Table memory.  A FreqFilter pass cycles between a large table (counting) and a small one (after deleteAll), and
// the next pass starts again with a large one: the map keeps both allocations (`table` and `spare`) and swaps them,
// so that the steady state allocates nothing.  Everything happens in stream order on m->stream.
static int init_table(void *t, unsigned long long n, cudaStream_t s, bool with_vid)
{
    init_table_kernel<<<grid_for(n / 2, 256, 32), 256, 0, s>>>(table_view(t, n), with_vid);
    GB_LAUNCHED();
    return GB_OK;
}

void map_retire_table(Map *m, void *t, unsigned long long alloc_cap)
{
    if (!t) return;
    if (!m->spare) { m->spare = t; m->spare_cap = alloc_cap; return; }
    if (alloc_cap > m->spare_cap) { std::swap(t, m->spare); std::swap(alloc_cap, m->spare_cap); }
    cudaStreamSynchronize(m->stream); // the retired table may still be read by a rehash on the stream
    cudaFree(t);
}

// make an EMPTY table of new_cap slots current; the previous one is handed back (still valid on the stream:
// the caller rehashes out of it and then retires it)
int map_swap_table(Map *m, unsigned long long new_cap, void **old_table, unsigned long long *old_alloc_cap)
{
    void *nt = nullptr;
    unsigned long long na = new_cap;
    if (m->spare && m->spare_cap >= new_cap) {
        nt = m->spare;
        na = m->spare_cap;
        m->spare = nullptr;
    } else {
        GB_CUDA(cudaMalloc(&nt, SLOT_BYTES * new_cap));
    }
    GB_TRY(init_table(nt, new_cap, m->stream, false));
    m->kept_valid = false;
    *old_table = m->table;
    *old_alloc_cap = m->alloc_cap;
    m->table = nt;
    m->alloc_cap = na;
    m->cap = new_cap;
    return GB_OK;
}

int map_stage(Map *m, size_t n_u64)
{
    if (n_u64 <= m->stage_cap) return GB_OK;
    if (m->stage) {
        GB_CUDA(cudaStreamSynchronize(m->stream));
        GB_CUDA(cudaFree(m->stage));
    }
    m->stage = nullptr;
    m->stage_cap = 0;
    size_t want = n_u64 + n_u64 / 16 + 1024;
    GB_CUDA(cudaMalloc((void **)&m->stage, want * 8));
    m->stage_cap = want;
    return GB_OK;
}

// rehash into a table of new_cap slots (optionally dropping counts below min_count)
int map_rebuild(Map *m, unsigned long long new_cap, bool filter, int min_count)
{
    void *old = nullptr;
    unsigned long long old_alloc = 0;
    const unsigned long long n = m->cap;
    GB_TRY(map_swap_table(m, new_cap, &old, &old_alloc));
    rehash_kernel<<<grid_for(n, 256, 32), 256, 0, m->stream>>>(table_view(old, n), min_count, filter, m->view());
    GB_LAUNCHED();
    map_retire_table(m, old, old_alloc);
    m->grows++;
    return GB_OK;
}

// make room so that `incoming` further updates can never fill the table (load stays <= 0.9 even if every
// update is a new key), growing by rehash like ArrayDNAMap.rescale (217-230).  Returns how many updates may
// be issued before the next reserve through *budget.
int map_budget(Map *m, int64_t incoming, int64_t *budget)
{
    int64_t cap = (int64_t)m->cap;
    int64_t room = (int64_t)(cap * 0.9) - m->size;
    int64_t min_batch = std::min<int64_t>(incoming, (int64_t)TILE_READS * 256);
    if (m->size * 10 > cap * 7 || room < min_batch) {
        unsigned long long nc = cap_for(std::max<int64_t>(m->size, 1) * 2);
        while ((int64_t)(nc * 0.9) - m->size < min_batch) nc *= 2;
        if (nc > m->cap) GB_TRY(map_rebuild(m, nc, false, 0));
        cap = (int64_t)m->cap;
        room = (int64_t)(cap * 0.9) - m->size;
    }
    *budget = room;
    return GB_OK;
}

int map_launch_update_counts(Map *m, const unsigned long long *d_keys, int64_t n, cudaStream_t st)
{
    if (n <= 0) return GB_OK;
    update_keys_kernel<false><<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(d_keys, nullptr, n, m->view(), m->d_counters);
    GB_LAUNCHED();
    return GB_OK;
}

// set_vid: the table is EMPTY and the keys distinct; their index becomes the slot's vid and the fingerprint array is built
int map_launch_update_set(Map *m, const unsigned long long *d_keys, const int *d_vals, int64_t n, cudaStream_t st, bool set_vid)
{
    uint8_t *fp = nullptr;
    if (set_vid) {
        if (m->fp_cap < m->cap) {
            if (m->fp) { GB_CUDA(cudaStreamSynchronize(st)); GB_CUDA(cudaFree(m->fp)); }
            m->fp = nullptr;
            m->fp_cap = 0;
            GB_CUDA(cudaMalloc((void **)&m->fp, m->cap + m->cap / 8 + 1024));
            m->fp_cap = m->cap + m->cap / 8 + 1024;
        }
        GB_CUDA(cudaMemsetAsync(m->fp, 0, m->cap, st));
        fp = m->fp;
    }
    if (n <= 0) return GB_OK;
    if (set_vid) place_distinct_kernel<<<(unsigned int)((n + 511) / 512), 256, 0, st>>>(d_keys, d_vals, n, m->view(), fp, m->d_counters);
    else update_keys_kernel<true><<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(d_keys, d_vals, n, m->view(), m->d_counters, false, nullptr);
    GB_LAUNCHED();
    return GB_OK;
}

int map_read_counters(Map *m, unsigned long long out[4])
{
    GB_CUDA(cudaMemcpyAsync(out, m->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, m->stream));
    GB_CUDA(cudaStreamSynchronize(m->stream));
    return GB_OK;
}

int map_zero_counters(Map *m)
{
    GB_CUDA(cudaMemsetAsync(m->d_counters, 0, 4 * sizeof(unsigned long long), m->stream));
    return GB_OK;
}

template <bool FIXED>
static int launch_insert(Map *m, const uint8_t *d_bin, size_t n_bytes, const unsigned long long *d_off,
                         unsigned int rec, int64_t read0, int64_t n_reads)
{
    if (n_reads <= 0) return GB_OK;
    unsigned int grid = (unsigned int)((n_reads + TILE_READS - 1) / TILE_READS);
    if (m->v210)
        insert_reads_kernel<FIXED, true><<<grid, INSERT_THREADS, 0, m->stream>>>(
            d_bin, n_bytes, d_off, rec, read0, n_reads, m->k, m->view(), m->d_counters);
    else
        insert_reads_kernel<FIXED, false><<<grid, INSERT_THREADS, 0, m->stream>>>(
            d_bin, n_bytes, d_off, rec, read0, n_reads, m->k, m->view(), m->d_counters);
    GB_LAUNCHED();
    return GB_OK;
}

// counters[2] != 0 <=> some record of the fixed-stride stream has another length byte
int map_verify_fixed(Map *m, const uint8_t *d_bin, unsigned int rec, unsigned int len0, int64_t n_reads, unsigned long long *bad)
{
    GB_TRY(map_zero_counters(m));
    verify_fixed_kernel<<<grid_for((unsigned long long)n_reads, 256), 256, 0, m->stream>>>(d_bin, rec, len0, n_reads, m->d_counters);
    GB_LAUNCHED();
    unsigned long long c[4];
    GB_TRY(map_read_counters(m, c));
    *bad = c[2];
    return GB_OK;
}

// the single-pass bucket pass leaves small batches to the counted passes: below g_tune.single_pass_min k-windows the slabs
// (one per bucket and CTA, each at least 128 keys) would dwarf the batch
static int64_t countless_min() { return std::max<int64_t>(1, g_tune.single_pass_min); }

// 0 = choose by table size, 1 = direct (fused extract + upsert, random access), 2 = partitioned (L2-blocked)
static int insert_mode() { return (int)g_tune.insert_path; }

// L2-blocked insert of reads [read0, read0 + n_reads): bucket the canonical k-mers by table slice (count, scatter),
// upsert slice by slice.  The range is cut into sub-batches and pipelined over two staging halves: the bucket pass of
// sub-batch s + 1 (instruction bound, high-priority stream) overlaps the upsert of sub-batch s (L2 bound, map stream).
// win_upper(r0, r1) = upper bound of the k-windows of reads [r0, r1).
template <typename WinUpper>
static int insert_partitioned(Map *m, const uint8_t *d_bin, size_t n_bytes, const unsigned long long *d_off, unsigned int rec,
                              int64_t read0, int64_t n_reads, WinUpper win_upper, bool bound_is_exact)
{
    cudaStream_t up = m->stream, bk = m->copy_stream;
    if (!m->part) m->part = new PartWork();
    if (!m->part2) m->part2 = new PartWork();
    for (int i = 0; i < 4; i++)
        if (!m->pe[i]) GB_CUDA(cudaEventCreate(&m->pe[i]));
    for (int i = 0; i < 2; i++) {
        if (!m->pready[i]) GB_CUDA(cudaEventCreateWithFlags(&m->pready[i], cudaEventDisableTiming));
        if (!m->pfree[i]) GB_CUDA(cudaEventCreateWithFlags(&m->pfree[i], cudaEventDisableTiming));
    }
    PartWork *work[2] = { m->part, m->part2 };
    GB_TRY(work[0]->ensure(bk));
    GB_TRY(work[1]->ensure(bk));
    PartLayout pl;
    pl.owners = 1;
    pl.lp_bits = slice_bits_for(m->cap, 1);
    if (g_tune.slice_bits >= 0) pl.lp_bits = (int)std::max<long long>(0, std::min<long long>(pl.lp_bits + 3, g_tune.slice_bits));
    while ((1 << pl.lp_bits) > 128) pl.lp_bits--;

    // sub-batches of whole persistent-grid waves of tiles, about four of them
    const int64_t tiles = (n_reads + TILE_READS - 1) / TILE_READS, grid = work[0]->grid;
    // measured on C2 (one B200): 1 / 2 / 4 / 8 sub-batches = 2.95 / 3.13 / 3.20 / 3.93 ms -- on ONE GPU the two passes
    // fight for the same SM slots and L2, so the default is no overlap; gb_tune batches turns the pipeline on
    const int want = (int)std::max<long long>(1, g_tune.batches);
    int64_t per_tiles = std::max<int64_t>(1, (tiles / want + grid / 2) / grid) * grid;
    if (tiles < 2 * grid || want == 1) per_tiles = tiles;
    const int64_t per_reads = per_tiles * TILE_READS;
    const int64_t n_sub = (n_reads + per_reads - 1) / per_reads;
    int64_t half = 0; // keys per staging half
    for (int64_t s = 0; s < n_sub; s++) half = std::max(half, win_upper(read0 + s * per_reads, read0 + std::min(n_reads, (s + 1) * per_reads)));
    // single pass (one sub-batch only): no count pass, per-(bucket, CTA) slabs instead of exact bucket ranges.  Measured on C2
    // (profiles/r2a_bench_*.json): bucket pass 0.995 -> 0.765 ms, upsert over the slab chunks 1.83 -> 1.95 ms
    const unsigned int nbk = (unsigned int)pl.nb();
    const unsigned int slab = g_tune.single_pass && n_sub == 1 && pl.owners == 1 && nbk <= 128 && half >= countless_min()
                                  ? slab_keys_for(slab_cta_keys(n_reads, (unsigned long long)half, (int)grid), nbk, (int)grid) : 0;
    const size_t slab_keys = (size_t)slab * nbk * (size_t)grid, n_slab_chunks = (size_t)nbk * (size_t)grid;
    GB_TRY(map_stage(m, slab ? slab_keys + 8 : (size_t)((n_sub > 1 ? 2 : 1) * (half + 8))));
    if (n_sub == 1) bk = up; // nothing to overlap: one stream, no cross-stream events
    // the bucket stream starts after everything already queued on the map's stream (clear, earlier inserts)
    GB_CUDA(cudaEventRecord(m->pe[2], up));
    if (bk != up) GB_CUDA(cudaStreamWaitEvent(bk, m->pe[2], 0));
    GB_CUDA(cudaEventRecord(m->pe[0], bk));
    m->n_pup = 0;
    for (int64_t s = 0; s < n_sub; s++) {
        const int h = (int)(s & 1);
        const int64_t r0 = read0 + s * per_reads, nr = std::min(per_reads, read0 + n_reads - r0);
        const int64_t wu = win_upper(r0, r0 + nr);
        unsigned long long *keys = m->stage + (size_t)h * (half + 8), *d_desc = keys + half;
        ReadBatch rb;
        rb.bin = d_bin; rb.n_bytes = n_bytes; rb.offsets = d_off; rb.rec_bytes = rec; rb.read0 = r0; rb.n_reads = nr;
        if (s >= 2) GB_CUDA(cudaStreamWaitEvent(bk, m->pfree[h], 0)); // the upsert of sub-batch s - 2 has consumed this half
        if (slab) {
            GB_TRY(part_scatter_slabs(rb, m->k, m->v210, pl, *work[h], keys, slab, m, bk));
        } else {
            GB_TRY(part_count(rb, m->k, m->v210, pl, *work[h], bk));
            GB_TRY(part_scatter(rb, m->k, m->v210, pl, *work[h], keys, bk));
            GB_TRY(make_single_chunk(work[h]->bucket_base + pl.nb(), d_desc, m->d_counters, bk));
        }
        if (bk != up) {
            GB_CUDA(cudaEventRecord(m->pready[h], bk));
            GB_CUDA(cudaStreamWaitEvent(up, m->pready[h], 0));
        }
        if (m->n_pup + 2 <= 16) {
            for (int i = 0; i < 2; i++)
                if (!m->pup[m->n_pup + i]) GB_CUDA(cudaEventCreate(&m->pup[m->n_pup + i]));
            GB_CUDA(cudaEventRecord(m->pup[m->n_pup], up));
        }
        // the buckets are contiguous and already in slice order: one chunk; the launch is sized from the upper bound
        unsigned long long total = (unsigned long long)wu;
        if (slab) {
            // the fill counts are on the device (keys in slabs; overflowed keys were upserted by the bucket pass)
            (void)total;
            GB_TRY(insert_slabs(m, keys, work[h]->cta_hist, slab, (unsigned int)n_slab_chunks, up));
        } else {
        if (!bound_is_exact) { // record lengths are only on the device: fetch the count (stalls the pipeline; rare path)
            GB_CUDA(cudaMemcpyAsync(&total, work[h]->bucket_base + pl.nb(), 8, cudaMemcpyDeviceToHost, bk));
            GB_CUDA(cudaStreamSynchronize(bk));
        }
        GB_TRY(insert_key_chunks(m, keys, d_desc, d_desc + 2, 1, total, up));
        }
        if (m->n_pup + 2 <= 16) {
            GB_CUDA(cudaEventRecord(m->pup[m->n_pup + 1], up));
            m->n_pup += 2;
        }
        GB_CUDA(cudaEventRecord(m->pfree[h], up));
    }
    if (bk != up) GB_CUDA(cudaEventRecord(m->pe[1], bk));
    GB_CUDA(cudaStreamSynchronize(bk));
    GB_CUDA(cudaStreamSynchronize(up));
    float ms = 0;
    // span of the bucket stream (one stream: up to the start of the upsert)
    GB_CUDA(cudaEventElapsedTime(&ms, m->pe[0], bk != up ? m->pe[1] : m->pup[0]));
    m->phase_ns[0] += (int64_t)(ms * 1e6);
    for (int i = 0; i + 1 < m->n_pup; i += 2) {
        GB_CUDA(cudaEventElapsedTime(&ms, m->pup[i], m->pup[i + 1]));
        m->phase_ns[2] += (int64_t)(ms * 1e6);
    }
    return GB_OK;
}

// Core of gb_map_insert_reads*: the stream is on the device.  h_windows_prefix (optional, n_reads+1) gives
// exact per-read window prefix sums for batching in the ragged case.
// lengths_vary (fixed stride only): len0 is the LARGEST record length, window counts derived from it are upper bounds
static int insert_device(Map *m, const uint8_t *d_bin, size_t n_bytes, const unsigned long long *d_off,
                         bool fixed, unsigned int rec, unsigned int len0, int64_t n_reads,
                         const int64_t *h_win_prefix, int64_t *n_windows, bool lengths_vary = false)
{
    GB_TRY(map_zero_counters(m));
    m->kept_valid = false;
    int64_t done = 0;
    int64_t total_ns = 0;
    unsigned long long before[4] = { 0, 0, 0, 0 };
    m->phase_ns[0] = m->phase_ns[1] = m->phase_ns[2] = 0;
    while (done < n_reads) {
        // how many reads fit the budget (whole tiles)
        int64_t left = n_reads - done;
        int64_t win_per_read_max = fixed ? std::max<int64_t>(0, (int64_t)len0 - m->k + 1) : (255 - m->k + 1);
        int64_t want = fixed ? left * win_per_read_max
                             : (h_win_prefix ? h_win_prefix[n_reads] - h_win_prefix[done] : left * win_per_read_max);
        int64_t budget = 0;
        GB_TRY(map_budget(m, want, &budget));
        int64_t take = left;
        if (want > budget) {
            if (fixed) {
                take = win_per_read_max ? budget / win_per_read_max : left;
            } else if (h_win_prefix) {
                const int64_t *b = h_win_prefix + done, *e = h_win_prefix + n_reads + 1;
                take = (std::upper_bound(b, e, h_win_prefix[done] + budget) - b) - 1;
            } else {
                take = budget / win_per_read_max;
            }
            take = std::max<int64_t>(TILE_READS, take / TILE_READS * TILE_READS);
            take = std::min(take, left);
        }
        // the counters were zeroed above and only this call advances them: what they held before this batch is known on the host,
        // and the batch's work is queued behind whatever the stream is still doing (the clear) without a synchronisation
        unsigned long long after[4];
        const int64_t take_windows = fixed ? take * win_per_read_max
                                           : (h_win_prefix ? h_win_prefix[done + take] - h_win_prefix[done] : take * win_per_read_max);
        const int mode = insert_mode();
        const bool partitioned = mode == 2 || (mode == 0 && (SLOT_BYTES * m->cap) > (96u << 20) && take_windows >= (1 << 20));
        GB_CUDA(cudaEventRecord(m->ev0, m->stream));
        if (partitioned) {
            // bounded key staging: at most 2^28 k-mers (2 GiB) per call, pipelined inside
            const int64_t per_read = std::max<int64_t>(1, fixed ? win_per_read_max : (255 - m->k + 1));
            const int64_t max_reads = std::max<int64_t>(TILE_READS, (((int64_t)1 << 28) / per_read) / TILE_READS * TILE_READS);
            auto win_upper = [&](int64_t r0, int64_t r1) -> int64_t {
                if (fixed) return (r1 - r0) * win_per_read_max;
                return h_win_prefix ? h_win_prefix[r1] - h_win_prefix[r0] : (r1 - r0) * per_read;
            };
            for (int64_t o = 0; o < take; o += max_reads) {
                const int64_t sub = std::min(max_reads, take - o);
                GB_TRY(insert_partitioned(m, d_bin, n_bytes, fixed ? nullptr : d_off, rec, done + o, sub, win_upper,
                                          (fixed && !lengths_vary) || h_win_prefix != nullptr));
            }
        } else if (fixed) {
            GB_TRY(launch_insert<true>(m, d_bin, n_bytes, nullptr, rec, done, take));
        } else {
            GB_TRY(launch_insert<false>(m, d_bin, n_bytes, d_off, 0, done, take));
        }
        GB_CUDA(cudaEventRecord(m->ev1, m->stream));
        GB_TRY(map_read_counters(m, after));
        float ms = 0;
        GB_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
        total_ns += (int64_t)(ms * 1e6);
        m->size += (int64_t)(after[0] - before[0]);
        before[0] = after[0];
        done += take;
    }
    unsigned long long c[4];
    GB_TRY(map_read_counters(m, c));
    m->windows += (int64_t)c[3];
    m->last_insert_ns = total_ns;
    if (n_windows) *n_windows = (int64_t)c[3];
    return GB_OK;
}

// gb_map_insert_reads on a fixed-stride HOST stream (single-pass bucket pass): the host-to-device copy is cut into chunks
// (copy stream) and the single-pass bucket pass of chunk c (map stream) runs while chunk c + 1 is still on the wire.  Nothing
// touches the table before every chunk has been verified: keys go to the slabs or to the overflow list (LIST mode of
// partition.cu), and the upsert starts after the last chunk.  *handled = false: conditions not met, or the stream turned
// out ragged / the overflow list filled up -- the table is untouched and the caller takes the ordinary path.
// Measured on C2 (profiles/r2a_bench_*.json): 4.45 -> 3.88 ms per step from a pinned host buffer.
static int insert_host_pipelined(Map *m, const uint8_t *bin, int64_t n_reads, unsigned int rec, unsigned int len0, int64_t *n_windows,
                                 bool *handled)
{
    *handled = false;
    if (!g_tune.single_pass || (int)len0 < m->k) return GB_OK;
    const int64_t per_read = (int64_t)len0 - m->k + 1, want = n_reads * per_read;
    if (want < countless_min() || want > ((int64_t)1 << 28)) return GB_OK; // small: not worth it; large: the ordinary path batches
    int64_t budget = 0;
    GB_TRY(map_budget(m, want, &budget));
    if (want > budget) return GB_OK;
    const int mode = insert_mode();
    if (!(mode == 2 || (mode == 0 && (SLOT_BYTES * m->cap) > (96u << 20)))) return GB_OK;
    if (!m->part) m->part = new PartWork();
    PartWork &w = *m->part;
    GB_TRY(w.ensure(m->stream));
    PartLayout pl;
    pl.owners = 1;
    pl.lp_bits = slice_bits_for(m->cap, 1);
    const unsigned int nb = (unsigned int)pl.nb();
    if (nb > 128) return GB_OK;
    const unsigned long long ovf_cap = (unsigned long long)want / 16 + 65536;
    const int n_copy = (int)std::max<long long>(1, std::min<long long>(16, g_tune.h2d_chunks));
    const int64_t per_chunk = std::max<int64_t>(TILE_READS, ((n_reads + n_copy - 1) / n_copy + TILE_READS - 1) / TILE_READS * TILE_READS);
    // every chunk is its own launch and deals its tiles to the CTAs from CTA 0 on: a CTA sees the sum of its shares
    unsigned long long cta_keys = 0;
    for (int64_t r0 = 0; r0 < n_reads; r0 += per_chunk) {
        const int64_t nr = std::min(per_chunk, n_reads - r0);
        cta_keys += slab_cta_keys(nr, (unsigned long long)(nr * per_read), w.grid);
    }
    const unsigned int slab = slab_keys_for(cta_keys, nb, w.grid); // 0: positions would not fit 32 bits
    if (!slab) return GB_OK;
    const size_t n_chunks = (size_t)nb * (size_t)w.grid, slab_keys = (size_t)slab * n_chunks;
    GB_TRY(map_stage(m, slab_keys + (size_t)ovf_cap + 16));
    unsigned long long *keys = m->stage, *d_desc = keys + slab_keys + ovf_cap;
    const size_t used = (size_t)n_reads * rec;
    DeviceBuf d_bin;
    GB_TRY(d_bin.alloc(used + 16, m->stream));

    cudaEvent_t entry = nullptr, copied[16];
    for (auto &e : copied) e = nullptr;
    struct Guard {
        cudaEvent_t *entry, *copied;
        ~Guard()
        {
            if (*entry) cudaEventDestroy(*entry);
            for (int i = 0; i < 16; i++)
                if (copied[i]) cudaEventDestroy(copied[i]);
        }
    } guard{ &entry, copied };
    GB_CUDA(cudaEventCreateWithFlags(&entry, cudaEventDisableTiming));
    if (!m->pe[0])
        for (int i = 0; i < 4; i++) GB_CUDA(cudaEventCreate(&m->pe[i]));

    GB_TRY(map_zero_counters(m));
    m->phase_ns[0] = m->phase_ns[1] = m->phase_ns[2] = 0;
    GB_CUDA(cudaEventRecord(m->ev0, m->stream));
    GB_TRY(slab_list_begin(pl, w, m->stream));
    // the copy stream may not write arena memory that work already queued on the map's stream could still be reading
    GB_CUDA(cudaEventRecord(entry, m->stream));
    GB_CUDA(cudaStreamWaitEvent(m->copy_stream, entry, 0));
    int ci = 0;
    for (int64_t r0 = 0; r0 < n_reads; r0 += per_chunk, ci++) {
        const int64_t nr = std::min(per_chunk, n_reads - r0);
        GB_CUDA(cudaMemcpyAsync((uint8_t *)d_bin.p + (size_t)r0 * rec, bin + (size_t)r0 * rec, (size_t)nr * rec, cudaMemcpyHostToDevice, m->copy_stream));
        GB_CUDA(cudaEventCreateWithFlags(&copied[ci], cudaEventDisableTiming));
        GB_CUDA(cudaEventRecord(copied[ci], m->copy_stream));
        GB_CUDA(cudaStreamWaitEvent(m->stream, copied[ci], 0));
        verify_fixed_kernel<<<grid_for((unsigned long long)nr, 256), 256, 0, m->stream>>>((const uint8_t *)d_bin.p + (size_t)r0 * rec, rec, len0, nr,
                                                                                       m->d_counters);
        GB_LAUNCHED();
        ReadBatch rb;
        rb.bin = (const uint8_t *)d_bin.p; rb.n_bytes = used; rb.offsets = nullptr; rb.rec_bytes = rec; rb.read0 = r0; rb.n_reads = nr;
        GB_TRY(slab_list_range(rb, m->k, m->v210, pl, w, keys, slab, ovf_cap, m->stream));
    }
    GB_TRY(slab_list_end(pl, w, slab, ovf_cap, d_desc, m, m->stream));
    GB_CUDA(cudaEventRecord(m->pe[0], m->stream));
    // ragged? overflow list full?  (one synchronisation, like the ordinary path's verification)
    unsigned long long c[4], flags[2] = { 0, 0 };
    GB_CUDA(cudaMemcpyAsync(flags, w.bucket_total, 16, cudaMemcpyDeviceToHost, m->stream));
    GB_TRY(map_read_counters(m, c));
    if (c[2] || (unsigned int)flags[1]) {
        GB_TRY(map_zero_counters(m)); // counters[3] was advanced by slab_list_end
        return GB_OK;
    }
    GB_TRY(insert_slabs(m, keys, w.cta_hist, slab, (unsigned int)n_chunks, m->stream));
    // the overflow list (its keys hit any slice: random access for those few); d_desc = its 3-word chunk table
    GB_TRY(insert_key_chunks(m, keys + slab_keys, d_desc, d_desc + 2, 1, ovf_cap, m->stream, true));
    GB_CUDA(cudaEventRecord(m->ev1, m->stream));
    GB_TRY(map_read_counters(m, c));
    float ms = 0;
    GB_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->pe[0]));
    m->phase_ns[0] = (int64_t)(ms * 1e6); // copy + bucket pass, overlapped
    GB_CUDA(cudaEventElapsedTime(&ms, m->pe[0], m->ev1));
    m->phase_ns[2] = (int64_t)(ms * 1e6);
    GB_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
    m->last_insert_ns = (int64_t)(ms * 1e6);
    m->size += (int64_t)c[0];
    m->windows += (int64_t)c[3];
    m->fixed_stride = 1;
    if (n_windows) *n_windows = (int64_t)c[3];
    *handled = true;
    return GB_OK;
}

// host scan of the record chain (PairedEndData.getPairs.read, PairedEndData.scala:24-32)
int scan_records(const uint8_t *bin, size_t n_bytes, int64_t n_reads, int k,
                        std::vector<unsigned long long> &off, std::vector<int64_t> &winp)
{
    off.resize((size_t)n_reads + 1);
    winp.resize((size_t)n_reads + 1);
    size_t pos = 0;
    int64_t w = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        if (pos >= n_bytes) { set_error("truncated .bin stream at read %lld", (long long)r); return GB_E_ARG; }
        unsigned int len = bin[pos];
        size_t nxt = pos + 1 + (len + 3) / 4;
        if (nxt > n_bytes) { set_error("truncated .bin stream at read %lld", (long long)r); return GB_E_ARG; }
        off[(size_t)r] = pos;
        winp[(size_t)r] = w;
        if ((int)len >= k) w += len - k + 1;
        pos = nxt;
    }
    off[(size_t)n_reads] = pos;
    winp[(size_t)n_reads] = w;
    return GB_OK;
}

int check_keys(const Map *m, const uint64_t *keys, int64_t n)
{
    const unsigned long long kmask = (1ull << (2 * m->k)) - 1;
    for (int64_t i = 0; i < n; i++)
        if (keys[i] & ~kmask) { set_error("key %lld is longer than k = %d", (long long)i, m->k); return GB_E_K_RANGE; }
    return GB_OK;
}

int check_map(gb_map *h, Map **m)
{
    if (!h) { set_error("null map handle"); return GB_E_ARG; }
    *m = reinterpret_cast<Map *>(h);
    GB_CUDA(cudaSetDevice((*m)->device));
    return GB_OK;
}

} // namespace gb

namespace gb {
// mapReduce export into DEVICE arrays of m->size entries (synchronises the map's stream)
int map_export_device(Map *m, unsigned long long *d_keys, int *d_vals)
{
    if (m->size == 0) return GB_OK;
    unsigned long long n = m->cap;
    GB_TRY(map_zero_counters(m));
    export_kernel<<<grid_for(n, 256, 16), 256, 0, m->stream>>>(m->view(), d_keys, d_vals, (unsigned long long)m->size, m->d_counters);
    GB_LAUNCHED();
    unsigned long long c[4];
    GB_TRY(map_read_counters(m, c));
    if ((int64_t)c[1] != m->size) { set_error("internal: exported %llu keys, size is %lld", c[1], (long long)m->size); return GB_E_INVARIANT; }
    return GB_OK;
}
} // namespace gb

// ================================================================ C ABI

extern "C" {

const char *gb_last_error(void) { return g_err; }
int gb_version(void) { return 100; }

int gb_device_count(int *n)
{
    if (!n) { set_error("null argument"); return GB_E_ARG; }
    GB_CUDA(cudaGetDeviceCount(n));
    return GB_OK;
}

int gb_host_alloc(size_t n_bytes, void **ptr)
{
    if (!ptr) { set_error("null argument"); return GB_E_ARG; }
    GB_CUDA(cudaHostAlloc(ptr, n_bytes ? n_bytes : 16, cudaHostAllocDefault));
    return GB_OK;
}

int gb_host_free(void *ptr)
{
    if (ptr) GB_CUDA(cudaFreeHost(ptr));
    return GB_OK;
}

int gb_map_create(int k, int64_t min_capacity, int device, uint32_t flags, gb_map **out)
{
    if (!out) { set_error("null out pointer"); return GB_E_ARG; }
    *out = nullptr;
    if (k < 1 || k > 31) { set_error("k = %d outside 1..31", k); return GB_E_K_RANGE; }
    if (min_capacity < 0) { set_error("negative capacity"); return GB_E_ARG; }
    int ndev = 0;
    GB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("device %d not present (%d devices)", device, ndev); return GB_E_CUDA; }
    GB_CUDA(cudaSetDevice(device));
    GB_TRY(pool_setup(device));
    Map *m = new Map();
    memset(m->pup, 0, sizeof m->pup);
    m->k = k;
    m->device = device;
    m->v210 = (flags & GB_FLAG_HASH_SCALA_210) != 0;
    const unsigned long long cap0 = cap_for(min_capacity);
    int r = GB_OK;
    do {
        if ((r = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) == cudaSuccess ? GB_OK : GB_E_CUDA)) break;
        {   // the bucket pass of the pipelined insert runs here: its CTAs go first when SM slots free up under the upsert grid
            int prio_lo = 0, prio_hi = 0;
            cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
            if ((r = cudaStreamCreateWithPriority(&m->copy_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess ? GB_OK : GB_E_CUDA)) break;
        }
        if ((r = cudaEventCreate(&m->ev0) == cudaSuccess ? GB_OK : GB_E_CUDA)) break;
        if ((r = cudaEventCreate(&m->ev1) == cudaSuccess ? GB_OK : GB_E_CUDA)) break;
        if ((r = cudaEventCreate(&m->t0) == cudaSuccess ? GB_OK : GB_E_CUDA)) break;
        if ((r = cudaEventCreate(&m->t1) == cudaSuccess ? GB_OK : GB_E_CUDA)) break;
        for (int i = 0; i < 4 && r == GB_OK; i++) r = cudaEventCreate(&m->fev[i]) == cudaSuccess ? GB_OK : GB_E_CUDA;
        if (r) break;
        if ((r = cudaMalloc((void **)&m->d_counters, 8 * sizeof(unsigned long long)) == cudaSuccess ? GB_OK : GB_E_OOM)) break;
        {
            void *none = nullptr;
            unsigned long long none_cap = 0;
            if ((r = map_swap_table(m, cap0, &none, &none_cap))) break;
        }
        if ((r = cudaStreamSynchronize(m->stream) == cudaSuccess ? GB_OK : GB_E_CUDA)) break;
    } while (0);
    if (r != GB_OK) {
        if (!g_err[0]) set_error("map creation failed: %s", cudaGetErrorString(cudaGetLastError()));
        gb_map_destroy(reinterpret_cast<gb_map *>(m));
        return r;
    }
    *out = reinterpret_cast<gb_map *>(m);
    return GB_OK;
}

int gb_map_destroy(gb_map *h)
{
    if (!h) return GB_OK;
    Map *m = reinterpret_cast<Map *>(h);
    cudaSetDevice(m->device);
    if (m->replica) gb_map_destroy(reinterpret_cast<gb_map *>(m->replica));
    m->replica = nullptr;
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->table) cudaFree(m->table);
    if (m->spare) cudaFree(m->spare);
    if (m->stage) cudaFree(m->stage);
    if (m->fp) cudaFree(m->fp);
    m->arena.destroy();
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->d_counters) cudaFree(m->d_counters);
    if (m->d_spread) cudaFree(m->d_spread);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    for (int i = 0; i < 4; i++)
        if (m->fev[i]) cudaEventDestroy(m->fev[i]);
    for (int i = 0; i < 4; i++)
        if (m->pe[i]) cudaEventDestroy(m->pe[i]);
    if (m->part) { m->part->release(); delete m->part; }
    if (m->part2) { m->part2->release(); delete m->part2; }
    for (int i = 0; i < 2; i++) {
        if (m->pready[i]) cudaEventDestroy(m->pready[i]);
        if (m->pfree[i]) cudaEventDestroy(m->pfree[i]);
    }
    for (int i = 0; i < 16; i++)
        if (m->pup[i]) cudaEventDestroy(m->pup[i]);
    if (m->t0) cudaEventDestroy(m->t0);
    if (m->t1) cudaEventDestroy(m->t1);
    if (m->stream) cudaStreamDestroy(m->stream);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    delete m;
    return GB_OK;
}

int gb_map_insert_reads_device(gb_map *h, const uint8_t *d_bin, size_t n_bytes, const uint64_t *d_offsets,
                               int64_t n_reads, int64_t *n_windows)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    ArenaScope scope(&m->arena);
    if (n_windows) *n_windows = 0;
    if (n_reads < 0 || (!d_bin && n_reads > 0)) { set_error("bad arguments"); return GB_E_ARG; }
    if (n_reads == 0) return GB_OK;
    if (d_offsets)
        return insert_device(m, d_bin, n_bytes, (const unsigned long long *)d_offsets, false, 0, 0, n_reads, nullptr, n_windows);
    // fixed-stride stream: take the length from the first record, verify all of them on the device
    uint8_t len0 = 0;
    GB_CUDA(cudaMemcpyAsync(&len0, d_bin, 1, cudaMemcpyDeviceToHost, m->stream));
    GB_CUDA(cudaStreamSynchronize(m->stream));
    unsigned int rec = 1 + (len0 + 3) / 4;
    if ((unsigned long long)n_reads * rec > n_bytes) { set_error("truncated .bin stream"); return GB_E_ARG; }
    GB_TRY(map_zero_counters(m));
    verify_fixed_kernel<<<grid_for((unsigned long long)n_reads, 256), 256, 0, m->stream>>>(d_bin, rec, len0, n_reads, m->d_counters);
    GB_LAUNCHED();
    unsigned long long c[4];
    GB_TRY(map_read_counters(m, c));
    if (c[2]) { set_error("records are not fixed-length: pass d_offsets"); return GB_E_ARG; }
    m->fixed_stride = 1;
    return insert_device(m, d_bin, n_bytes, nullptr, true, rec, len0, n_reads, nullptr, n_windows);
}

int gb_map_insert_reads(gb_map *h, const uint8_t *bin, size_t n_bytes, int64_t n_reads, int64_t *n_windows)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    ArenaScope scope(&m->arena);
    if (n_windows) *n_windows = 0;
    if (n_reads < 0 || (!bin && n_reads > 0)) { set_error("bad arguments"); return GB_E_ARG; }
    if (n_reads == 0) return GB_OK;
    if (n_bytes == 0) { set_error("truncated .bin stream at read 0"); return GB_E_ARG; }

    // fixed-stride fast path: if the stream is long enough for n_reads records of the first record's size,
    // copy exactly those bytes and let the device check every length byte (sound: equal length bytes at
    // i*rec imply the record chain is i*rec).  Otherwise scan the chain on the host.
    unsigned int len0 = bin[0], rec = 1 + (len0 + 3) / 4;
    bool try_fixed = (unsigned long long)n_reads * rec <= n_bytes;
    if (try_fixed) { // chunked copy overlapped with the single-pass bucket pass (when the batch qualifies: *handled)
        bool handled = false;
        GB_TRY(insert_host_pipelined(m, bin, n_reads, rec, len0, n_windows, &handled));
        if (handled) return GB_OK;
    }
    DeviceBuf d_bin, d_off;
    if (try_fixed) {
        size_t used = (size_t)n_reads * rec;
        GB_TRY(d_bin.alloc(used + 16, m->stream));
        GB_CUDA(cudaMemcpyAsync(d_bin.p, bin, used, cudaMemcpyHostToDevice, m->stream));
        GB_TRY(map_zero_counters(m));
        verify_fixed_kernel<<<grid_for((unsigned long long)n_reads, 256), 256, 0, m->stream>>>(
            (const uint8_t *)d_bin.p, rec, len0, n_reads, m->d_counters);
        GB_LAUNCHED();
        unsigned long long c[4];
        GB_TRY(map_read_counters(m, c));
        if (!c[2]) {
            m->fixed_stride = 1;
            return insert_device(m, (const uint8_t *)d_bin.p, used, nullptr, true, rec, len0, n_reads, nullptr, n_windows);
        }
    }
    m->fixed_stride = 0;
    std::vector<unsigned long long> off;
    std::vector<int64_t> winp;
    GB_TRY(scan_records(bin, n_bytes, n_reads, m->k, off, winp));
    size_t used = (size_t)off[(size_t)n_reads];
    if (!d_bin.p || !try_fixed || used > (size_t)n_reads * rec) {
        d_bin.release();
        GB_TRY(d_bin.alloc(used + 16, m->stream));
        GB_CUDA(cudaMemcpyAsync(d_bin.p, bin, used, cudaMemcpyHostToDevice, m->stream));
    }
    GB_TRY(d_off.alloc(off.size() * 8, m->stream));
    GB_CUDA(cudaMemcpyAsync(d_off.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, m->stream));
    return insert_device(m, (const uint8_t *)d_bin.p, used, (const unsigned long long *)d_off.p, false, 0, 0, n_reads,
                         winp.data(), n_windows);
}

int gb_map_insert_records_device(gb_map *h, const uint8_t *d_bin, size_t n_bytes, uint32_t rec_bytes, int64_t n_records, uint32_t max_len,
                                 int64_t *n_windows)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    ArenaScope scope(&m->arena);
    if (n_windows) *n_windows = 0;
    if (n_records < 0 || (!d_bin && n_records > 0) || rec_bytes < 1 || rec_bytes > (uint32_t)MAX_REC_BYTES || max_len > 255) { set_error("bad arguments"); return GB_E_ARG; }
    if (n_records == 0) return GB_OK;
    if ((unsigned long long)n_records * rec_bytes > n_bytes) { set_error("truncated record stream"); return GB_E_ARG; }
    GB_TRY(map_zero_counters(m));
    verify_records_kernel<<<grid_for((unsigned long long)n_records, 256), 256, 0, m->stream>>>(d_bin, rec_bytes, max_len, n_records, m->d_counters);
    GB_LAUNCHED();
    unsigned long long c[4];
    GB_TRY(map_read_counters(m, c));
    if (c[2]) { set_error("a record is longer than max_len = %u or than its %u-byte stride", max_len, rec_bytes); return GB_E_ARG; }
    m->fixed_stride = 1;
    return insert_device(m, d_bin, n_bytes, nullptr, true, rec_bytes, max_len, n_records, nullptr, n_windows, true);
}
} // extern "C"

namespace gb {
int map_insert_records(Map *m, const uint8_t *d_bin, size_t n_bytes, unsigned int rec_bytes, int64_t n_records, unsigned int max_len,
                       int64_t *n_windows)
{
    if (n_windows) *n_windows = 0;
    if (n_records <= 0) return GB_OK;
    return insert_device(m, d_bin, n_bytes, nullptr, true, rec_bytes, max_len, n_records, nullptr, n_windows, true);
}
} // namespace gb

extern "C" {

static int update_common(gb_map *h, const uint64_t *keys, const int32_t *vals, int64_t n, bool set)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    ArenaScope scope(&m->arena);
    if (n < 0 || (n > 0 && (!keys || (set && !vals)))) { set_error("bad arguments"); return GB_E_ARG; }
    if (n == 0) return GB_OK;
    GB_TRY(check_keys(m, keys, n));
    m->noncanonical = true;
    m->kept_valid = false;
    DeviceBuf dk, dv;
    GB_TRY(dk.alloc((size_t)n * 8, m->stream));
    GB_CUDA(cudaMemcpyAsync(dk.p, keys, (size_t)n * 8, cudaMemcpyHostToDevice, m->stream));
    if (set) {
        GB_TRY(dv.alloc((size_t)n * 4, m->stream));
        GB_CUDA(cudaMemcpyAsync(dv.p, vals, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    }
    int64_t done = 0;
    while (done < n) {
        int64_t budget = 0;
        GB_TRY(map_budget(m, n - done, &budget));
        int64_t take = std::min(n - done, std::max<int64_t>(budget, 1));
        GB_TRY(map_zero_counters(m));
        if (set) GB_TRY(map_launch_update_set(m, (const unsigned long long *)dk.p + done, (const int *)dv.p + done, take, m->stream, false));
        else GB_TRY(map_launch_update_counts(m, (const unsigned long long *)dk.p + done, take, m->stream));
        unsigned long long c[4];
        GB_TRY(map_read_counters(m, c));
        m->size += (int64_t)c[0];
        done += take;
    }
    return GB_OK;
}

int gb_map_update_counts(gb_map *h, const uint64_t *keys, int64_t n) { return update_common(h, keys, nullptr, n, false); }
int gb_map_update(gb_map *h, const uint64_t *keys, const int32_t *vals, int64_t n) { return update_common(h, keys, vals, n, true); }

int gb_map_size(gb_map *h, int64_t *size)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    if (!size) { set_error("null argument"); return GB_E_ARG; }
    *size = m->size;
    return GB_OK;
}

int gb_map_lookup(gb_map *h, const uint64_t *keys, int64_t n, int32_t *counts, uint8_t *found)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    ArenaScope scope(&m->arena);
    if (n < 0 || (n > 0 && !keys)) { set_error("bad arguments"); return GB_E_ARG; }
    if (n == 0) return GB_OK;
    GB_TRY(check_keys(m, keys, n));
    DeviceBuf dk, dc, df;
    GB_TRY(dk.alloc((size_t)n * 8, m->stream));
    GB_TRY(dc.alloc((size_t)n * 4, m->stream));
    GB_TRY(df.alloc((size_t)n, m->stream));
    GB_CUDA(cudaMemcpyAsync(dk.p, keys, (size_t)n * 8, cudaMemcpyHostToDevice, m->stream));
    lookup_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, m->stream>>>((const unsigned long long *)dk.p, n, m->view(), (int *)dc.p, (uint8_t *)df.p);
    GB_LAUNCHED();
    if (counts) GB_CUDA(cudaMemcpyAsync(counts, dc.p, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    if (found) GB_CUDA(cudaMemcpyAsync(found, df.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    GB_CUDA(cudaStreamSynchronize(m->stream));
    return GB_OK;
}

int gb_map_delete_below(gb_map *h, int32_t min_count)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    unsigned long long n = m->cap;
    if (m->size == 0) return GB_OK;
    GB_TRY(map_zero_counters(m));
    // survivors go to the staging buffer (which a previous deleteAll's vertex array may occupy: it dies here)
    m->kept_valid = false;
    GB_TRY(map_stage(m, (size_t)m->size + (size_t)(m->size + 1) / 2));
    unsigned long long *sk = m->stage;
    int *sv = reinterpret_cast<int *>(m->stage + m->size);
    GB_CUDA(cudaEventRecord(m->fev[0], m->stream));
    compact_survivors_kernel<<<grid_for(n, 256, 32), 256, 0, m->stream>>>(m->view(), min_count, sk, sv, m->d_counters);
    GB_LAUNCHED();
    GB_CUDA(cudaEventRecord(m->fev[1], m->stream));
    unsigned long long c[4];
    GB_TRY(map_read_counters(m, c));
    {
        float ms = 0;
        GB_CUDA(cudaEventElapsedTime(&ms, m->fev[0], m->fev[1]));
        m->filter_ns[0] = (int64_t)(ms * 1e6);
        m->filter_ns[1] = 0;
        m->slots_swept = (int64_t)n;
    }
    int64_t keep = (int64_t)c[1];
    if (keep == m->size) return GB_OK;
    // a fresh table sized for the survivors
    void *old = nullptr;
    unsigned long long old_alloc = 0;
    GB_TRY(map_swap_table(m, cap_for(keep), &old, &old_alloc));
    map_retire_table(m, old, old_alloc);
    m->grows++;
    GB_TRY(map_zero_counters(m));
    // survivors are distinct: their index in the compacted array becomes their vertex id (Graph.buildGraph skips numbering)
    const bool as_vertices = keep < (1ll << 30);
    GB_CUDA(cudaEventRecord(m->fev[2], m->stream));
    GB_TRY(map_launch_update_set(m, sk, sv, keep, m->stream, as_vertices));
    GB_CUDA(cudaEventRecord(m->fev[3], m->stream));
    m->filter_ns[1] = -1; // resolved lazily by gb_map_phase_ns (the call stays asynchronous)
    m->size = keep;
    m->kept_keys = sk;
    m->kept_n = keep;
    m->kept_valid = as_vertices;
    return GB_OK;
}

int gb_map_export(gb_map *h, uint64_t *keys, int32_t *vals, int64_t cap, int64_t *n_out)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    ArenaScope scope(&m->arena);
    if (n_out) *n_out = m->size;
    if (!keys && !vals) return GB_OK;
    if (cap < m->size) { set_error("export buffer too small: %lld < %lld", (long long)cap, (long long)m->size); return GB_E_CAPACITY; }
    if (m->size == 0) return GB_OK;
    DeviceBuf dk, dv;
    GB_TRY(dk.alloc((size_t)m->size * 8, m->stream));
    GB_TRY(dv.alloc((size_t)m->size * 4, m->stream));
    unsigned long long n = m->cap;
    GB_TRY(map_zero_counters(m));
    export_kernel<<<grid_for(n, 256, 16), 256, 0, m->stream>>>(m->view(), (unsigned long long *)dk.p, (int *)dv.p,
                                                             (unsigned long long)m->size, m->d_counters);
    GB_LAUNCHED();
    if (keys) GB_CUDA(cudaMemcpyAsync(keys, dk.p, (size_t)m->size * 8, cudaMemcpyDeviceToHost, m->stream));
    if (vals) GB_CUDA(cudaMemcpyAsync(vals, dv.p, (size_t)m->size * 4, cudaMemcpyDeviceToHost, m->stream));
    unsigned long long c[4];
    GB_TRY(map_read_counters(m, c));
    if ((int64_t)c[1] != m->size) { set_error("internal: exported %llu keys, size is %lld", c[1], (long long)m->size); return GB_E_INVARIANT; }
    return GB_OK;
}

int gb_map_clear(gb_map *h, int64_t min_capacity)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    if (min_capacity < 0) { set_error("negative capacity"); return GB_E_ARG; }
    unsigned long long nb = cap_for(min_capacity);
    if (nb != m->cap) {
        void *old = nullptr;
        unsigned long long old_alloc = 0;
        GB_TRY(map_swap_table(m, nb, &old, &old_alloc));
        map_retire_table(m, old, old_alloc);
    } else {
        GB_TRY(init_table(m->table, m->cap, m->stream, false));
    }
    m->size = 0;
    m->kept_valid = false;
    m->noncanonical = false;
    m->windows = 0;
    return GB_OK;
}

long long gb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

static long long *tune_field(const char *name)
{
    static const struct { const char *name; long long Tuning::*field; } table[] = {
        { "insert_path", &Tuning::insert_path }, { "single_pass", &Tuning::single_pass }, { "single_pass_min", &Tuning::single_pass_min },
        { "slice_bits", &Tuning::slice_bits }, { "batches", &Tuning::batches }, { "h2d_chunks", &Tuning::h2d_chunks },
        { "route", &Tuning::route }, { "a2a", &Tuning::a2a },
        { "pgraph_sharded", &Tuning::pgraph_sharded }, { "trace", &Tuning::trace },
    };
    if (name)
        for (const auto &e : table)
            if (!strcmp(name, e.name)) return &(g_tune.*(e.field));
    return nullptr;
}

int gb_tune(const char *name, int64_t value, int64_t *previous)
{
    long long *f = tune_field(name);
    if (!f) { set_error("unknown tuning key '%s'", name ? name : "(null)"); return GB_E_ARG; }
    if (previous) *previous = *f;
    *f = value;
    return GB_OK;
}

int gb_tune_get(const char *name, int64_t *value)
{
    long long *f = tune_field(name);
    if (!f || !value) { set_error("unknown tuning key '%s'", name ? name : "(null)"); return GB_E_ARG; }
    *value = *f;
    return GB_OK;
}

int gb_timer_start(gb_map *h)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    GB_CUDA(cudaEventRecord(m->t0, m->stream));
    return GB_OK;
}

int gb_timer_stop(gb_map *h, int64_t *ns)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    if (!ns) { set_error("null argument"); return GB_E_ARG; }
    GB_CUDA(cudaEventRecord(m->t1, m->stream));
    GB_CUDA(cudaEventSynchronize(m->t1));
    float ms = 0;
    GB_CUDA(cudaEventElapsedTime(&ms, m->t0, m->t1));
    *ns = (int64_t)((double)ms * 1e6);
    return GB_OK;
}

int gb_sync(gb_map *h)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    GB_CUDA(cudaStreamSynchronize(m->stream));
    return GB_OK;
}

int gb_bench_random_atomics(int device, size_t table_bytes, int64_t n_updates, int iters, int64_t *ns_per_iter)
{
    if (!ns_per_iter || table_bytes < 1024 || n_updates <= 0 || iters <= 0) { set_error("bad arguments"); return GB_E_ARG; }
    GB_CUDA(cudaSetDevice(device));
    const unsigned long long words = table_bytes / 8; // the table's actual size, not rounded to a power of two
    unsigned long long *t = nullptr;
    GB_CUDA(cudaMalloc((void **)&t, words * 8));
    GB_CUDA(cudaMemset(t, 0, words * 8));
    cudaEvent_t e0, e1;
    GB_CUDA(cudaEventCreate(&e0));
    GB_CUDA(cudaEventCreate(&e1));
    unsigned int grid = (unsigned int)((n_updates + 256 * 8 - 1) / (256 * 8));
    random_atomics_kernel<<<grid, 256>>>(t, words, n_updates, 1);
    GB_LAUNCHED();
    GB_CUDA(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) {
        random_atomics_kernel<<<grid, 256>>>(t, words, n_updates, 2 + i);
        GB_LAUNCHED();
    }
    GB_CUDA(cudaEventRecord(e1));
    GB_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    GB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ns_per_iter = (int64_t)(ms * 1e6 / iters);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    GB_CUDA(cudaFree(t));
    return GB_OK;
}

int gb_map_phase_ns(gb_map *h, int64_t ns[8])
{
    Map *m;
    GB_TRY(check_map(h, &m));
    if (!ns) { set_error("null argument"); return GB_E_ARG; }
    if (m->filter_ns[1] < 0) {
        GB_CUDA(cudaEventSynchronize(m->fev[3]));
        float ms = 0;
        GB_CUDA(cudaEventElapsedTime(&ms, m->fev[2], m->fev[3]));
        m->filter_ns[1] = (int64_t)(ms * 1e6);
    }
    memset(ns, 0, 8 * sizeof(int64_t));
    ns[0] = m->phase_ns[0] + m->phase_ns[1]; // L2-blocked insert: bucket pass
    ns[1] = m->phase_ns[2];                  // L2-blocked insert: slice-ordered upsert
    ns[2] = m->filter_ns[0];                 // deleteAll: table sweep
    ns[3] = m->filter_ns[1];                 // deleteAll: survivors into their new table
    ns[4] = m->slots_swept;                  // slots the sweep streamed (16 B each)
    ns[5] = m->graph_ns[0];                  // Graph.buildGraph: membership probes (masks_kernel)
    ns[6] = m->graph_ns[1];                  // Graph.buildGraph: list ranking (all jump_kernel launches)
    ns[7] = m->graph_ns[2];                  // ... and how many launches that took
    return GB_OK;
}

int gb_map_stats(gb_map *h, int64_t stats[8])
{
    Map *m;
    GB_TRY(check_map(h, &m));
    if (!stats) { set_error("null argument"); return GB_E_ARG; }
    memset(stats, 0, 8 * sizeof(int64_t));
    stats[0] = (int64_t)m->cap;
    stats[1] = stats[0] * (int64_t)SLOT_BYTES;
    stats[2] = m->grows;
    stats[3] = m->windows;
    stats[4] = m->last_insert_ns;
    stats[5] = m->fixed_stride;
    stats[6] = m->phase_ns[0] + m->phase_ns[1]; // partitioned insert: count + scatter
    stats[7] = m->phase_ns[2];                  // partitioned insert: slice-ordered upsert
    return GB_OK;
}

} // extern "C"
