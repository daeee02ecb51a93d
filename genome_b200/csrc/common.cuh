// common.cuh -- shared device/host helpers of libgenome_b200 (sm_100a only).
// Citations: paths relative to /root/reference, S/ = src/main/scala/ru/ifmo/genome/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/genome_b200.h"

namespace gb {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define GB_CUDA(expr)                                                         \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return gb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

// every kernel launch of the library goes through this: counts it (gb_launch_count) and checks the launch
void note_launch();
#define GB_LAUNCHED()                      \
    do {                                   \
        gb::note_launch();                 \
        GB_CUDA(cudaGetLastError());       \
    } while (0)

#define GB_TRY(expr)                \
    do {                            \
        int _r = (expr);            \
        if (_r != GB_OK) return _r; \
    } while (0)

// ---------------------------------------------------------------- tuning and test hooks (gb_tune, include/genome_b200.h)
// The library reads no environment variables on its data path.  Every choice below has a measured default;
// the others exist so that the parity tests can force a path that small inputs would not take by themselves.
struct Tuning {
    long long insert_path = 0;          // 0 = by table size, 1 = direct (fused extract + upsert), 2 = L2-blocked (bucket pass + slice-ordered upsert)
    long long single_pass = 1;          // L2-blocked insert on one GPU: bucket pass without a count pass (per-(bucket, CTA) slabs)
    long long single_pass_min = 1 << 20; // ... for batches of at least this many k-windows (smaller ones: the slabs would dwarf the batch)
    long long slice_bits = -1;          // table slices of the L2-blocked insert = 2^slice_bits; -1 = from the table size (64 MiB slices)
    long long batches = 0;              // sub-batches of one insert call; 0 = default (1 on one GPU, 2 sharded)
    long long h2d_chunks = 4;           // host insert: chunks of the host-to-device copy overlapped with the bucket pass
    long long route = 0;                // sharded insert: 0 = by shard size, 1 = one level (owner, slice) on the wire, 2 = two levels
    long long a2a = 0;                  // sharded insert: 0 = the bucket pass stores into the peers' inboxes (NVLink stores), 1 = staged
                                        // ncclSend/ncclRecv (what a box without peer access takes by itself)
    long long pgraph_sharded = 1;       // Graph.buildGraph over shards without a replica (sgraph.cuh); 0 = all-gather the shards, build replicated
    long long trace = 0;                // phase timings on stderr
};
extern Tuning g_tune;

#define GB_HD __host__ __device__ __forceinline__
// ---------------------------------------------------------------- table layout
// Structure of arrays over ONE allocation of 16 bytes per slot: keys[cap] (u64) | counts[cap] (i32) | vids[cap] (u32).
// Why not one 16-byte record per slot (rounds 1 and 2a): on B200 a 128-byte L2 line that is both READ (key compare) and WRITTEN
// (count red) costs twice the L2 time of the same requests on separate lines -- measured with the access pattern alone,
// profiles/r2c_l2_request_modes*.jsonl: load + red on one slot 4.2 cycles per update per SM, on separate arrays 2.1; different
// sectors of one line do not help (4.25).  With the arrays apart the key lines are read-mostly (written once per distinct key by
// the claiming CAS), the count lines are written only, and the vertex ids are not touched before Graph.buildGraph.
struct Table {
    unsigned long long *key = nullptr; // EMPTY_KEY when free; k <= 31 keys use at most 62 bits
    int *count = nullptr;              // DNAMap[Int] value; the word of a FREE slot holds 1, the count its claimer starts with
    unsigned int *vid = nullptr;       // dense vertex id, assigned by gb_graph_build / deleteAll
    unsigned long long cap = 0;
};
constexpr size_t SLOT_BYTES = 16;
GB_HD Table table_view(void *base, unsigned long long cap)
{
    Table t;
    t.key = static_cast<unsigned long long *>(base);
    t.count = reinterpret_cast<int *>(t.key + cap);
    t.vid = reinterpret_cast<unsigned int *>(t.count + cap);
    t.cap = cap;
    return t;
}
// one slot's contents in registers
struct Slot {
    unsigned long long key;
    int count;
    unsigned int vid;
};

constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned int NONE32 = 0xFFFFFFFFu;

constexpr int SM_COUNT = 148; // B200

// ---------------------------------------------------------------- k-mer arithmetic (device + host)

// Long1DNASeq.hashCode = long.## (S/dna/DNASeq.scala:103), scala-library 2.9.1: iv = (int)v;
// if (iv == v) iv else (int)(v ^ (v >>> 32)).  For 0 <= v < 2^62 both branches equal (int)(v ^ (v >>> 32)).
// V210 = scala >= 2.10: low ^ (high + (low >>> 31)).
template <bool V210>
GB_HD int scala_hash(unsigned long long v)
{
    unsigned int low = (unsigned int)v, high = (unsigned int)(v >> 32);
    if (V210) return (int)(low ^ (high + (low >> 31)));
    return (int)(low ^ high);
}

// Long1DNASeq.complement (165-168) then .reverse (155-163): 2-bit groups reversed, then aligned down.
GB_HD unsigned long long revcomp(unsigned long long x, int k)
{
    unsigned long long v = ~x; // complement = x ^ mask; the bits above 2k are shifted out below
#ifdef __CUDA_ARCH__
    v = __brevll(v);
#else
    v = ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
    v = ((v >> 2) & 0x3333333333333333ull) | ((v & 0x3333333333333333ull) << 2);
    v = ((v >> 4) & 0x0f0f0f0f0f0f0f0full) | ((v & 0x0f0f0f0f0f0f0f0full) << 4);
    v = ((v >> 8) & 0x00ff00ff00ff00ffull) | ((v & 0x00ff00ff00ff00ffull) << 8);
    v = ((v >> 16) & 0x0000ffff0000ffffull) | ((v & 0x0000ffff0000ffffull) << 16);
    v = (v >> 32) | (v << 32);
#endif
    // full bit reversal also swapped the two bits of every base: swap them back
    v = ((v >> 1) & 0x5555555555555555ull) | ((v & 0x5555555555555555ull) << 1);
    return v >> (64 - 2 * k);
}

// FreqFilter.add (S/data/FreqFilter.scala:31-32): signed 32-bit compare, tie => reverse complement
template <bool V210>
GB_HD unsigned long long canonical(unsigned long long x, unsigned long long rc)
{
    return scala_hash<V210>(x) < scala_hash<V210>(rc) ? x : rc;
}

// x.drop(1) :+ base (Graph.scala:279) and base +: x.take(k-1) (Graph.scala:273)
GB_HD unsigned long long kmer_append(unsigned long long x, int k, unsigned int b)
{
    return (x >> 2) | ((unsigned long long)b << (2 * (k - 1)));
}
GB_HD unsigned long long kmer_prepend(unsigned long long x, int k, unsigned int b)
{
    return ((x << 2) & ((1ull << (2 * k)) - 1)) | b;
}

// slot / owner hash: free choice, unobservable through DNAMap
GB_HD unsigned long long mix64(unsigned long long x)
{
    x *= 0x9E3779B97F4A7C15ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 29;
    return x;
}
// slot = floor(h * cap / 2^64): any capacity (not only powers of two), monotonic in h, so the top bits of h still
// select contiguous slices of the table
GB_HD unsigned long long slot_of(unsigned long long h, unsigned long long cap)
{
#ifdef __CUDA_ARCH__
    return __umul64hi(h, cap);
#else
    return (unsigned long long)(((unsigned __int128)h * cap) >> 64);
#endif
}
GB_HD unsigned long long next_slot(unsigned long long i, unsigned long long cap) { return i + 1 == cap ? 0 : i + 1; }
GB_HD unsigned int owner_of(unsigned long long h, unsigned int parts)
{
    return (unsigned int)(((h & 0xFFFFFFFFull) * parts) >> 32);
}

// one-byte fingerprint of a key (never 0: 0 marks an empty slot in the fingerprint array)
GB_HD unsigned int fp_tag(unsigned long long h) { return (unsigned int)((h >> 20) & 0xFF) | 1u; }

#ifdef __CUDACC__
// ---------------------------------------------------------------- device memory access helpers
// table arrays are read with L1 bypass (random access, no reuse inside an SM; L1 is not coherent)
__device__ __forceinline__ unsigned long long load_key(const Table &t, unsigned long long i)
{
    unsigned long long k;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(k) : "l"(t.key + i));
    return k;
}
__device__ __forceinline__ int load_count(const Table &t, unsigned long long i)
{
    int c;
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(c) : "l"(t.count + i));
    return c;
}
__device__ __forceinline__ unsigned int load_vid(const Table &t, unsigned long long i)
{
    unsigned int v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(t.vid + i));
    return v;
}
__device__ __forceinline__ Slot load_slot(const Table &t, unsigned long long i)
{
    Slot s;
    s.key = load_key(t, i);
    s.count = load_count(t, i);
    s.vid = load_vid(t, i);
    return s;
}

__device__ __forceinline__ void red_add_s32(int *p, int v)
{
    asm volatile("red.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}


// probe for `key`; returns slot index or -1 (the caller loads the count or the vertex id it needs: they live in other arrays).
// With `fp` (one byte per slot, 0 = empty, else fp_tag of the resident key; 1/16 of the table, L2-resident) the probe walks the
// fingerprints and touches the table only on a tag match: a negative probe -- 3 of 4 in Graph.buildGraph -- costs no DRAM access.
__device__ __forceinline__ long long probe_find(const Table &table, unsigned long long key, const uint8_t *fp = nullptr)
{
    const unsigned long long h = mix64(key), cap = table.cap;
    unsigned long long i = slot_of(h, cap);
    if (fp) {
        const unsigned int tag = fp_tag(h);
        for (;;) {
            const unsigned int t = fp[i];
            if (t == 0) return -1;
            if (t == tag && load_key(table, i) == key) return (long long)i;
            i = next_slot(i, cap);
        }
    }
    for (;;) { // the table is never full (map_budget), so an EMPTY slot always ends the probe
        const unsigned long long cur = load_key(table, i);
        if (cur == key) return (long long)i;
        if (cur == EMPTY_KEY) return -1;
        i = next_slot(i, cap);
    }
}

// Graph.buildGraph.contains (S/data/graph/Graph.scala:270): q is present if q or rc(q) is a stored key.
// Keys inserted by FreqFilter.add are stored in canonical orientation, so one probe of canonical(q)
// decides; both orientations are probed only when the two hashes tie (FreqFilter's rule then stores
// either orientation) or when keys were inserted as-is through update (dual).  If both
// orientations are stored, the numerically smaller key is the PRIMARY one: it alone carries a vertex id.
// On success *slot is the index of the primary stored slot and *strand = 1 when that stored key is rc(q) != q.
template <bool V210>
__device__ __forceinline__ bool find_oriented(const Table &table, int k, bool dual, unsigned long long q, unsigned long long *slot,
                                              unsigned int *strand, const uint8_t *fp = nullptr)
{
    unsigned long long r = revcomp(q, k);
    int hq = scala_hash<V210>(q), hr = scala_hash<V210>(r);
    if (!dual && hq != hr) {
        unsigned long long c = hq < hr ? q : r;
        const long long i = probe_find(table, c, fp);
        if (i < 0) return false;
        *slot = (unsigned long long)i;
        *strand = c != q;
        return true;
    }
    const long long iq = probe_find(table, q, fp);
    const long long ir = r != q ? probe_find(table, r, fp) : -1;
    if (iq < 0 && ir < 0) return false;
    bool use_r = ir >= 0 && (iq < 0 || r < q);
    *slot = (unsigned long long)(use_r ? ir : iq);
    *strand = use_r;
    return true;
}

// a stored key is SECONDARY (no vertex of its own) when rc(key) is stored too and is numerically smaller
template <bool V210>
__device__ __forceinline__ bool is_secondary(const Table &table, int k, bool dual, unsigned long long key,
                                             const uint8_t *fp = nullptr)
{
    unsigned long long r = revcomp(key, k);
    if (r >= key) return false;
    if (!dual && scala_hash<V210>(key) != scala_hash<V210>(r)) return false;
    return probe_find(table, r, fp) >= 0;
}
#endif

// ---------------------------------------------------------------- transient device memory
// Every top-level call that needs scratch space takes it from the handle's ARENA: one cudaMalloc'ed block, bump
// allocation, reset when the next top-level call on the handle begins (every such call ends synchronised).  The
// stream-ordered pool (cudaMallocAsync) is not used: with GB-sized blocks coming and going it remaps physical memory
// and stalls calls for 100s of ms (measured: graph builds of 3 ms taking 20..400 ms).
// freed arena blocks are kept (per process) and handed to the next arena that grows: handles come and go (a Graph per
// build, a replica map per sharded build) but their scratch memory is recycled without touching the driver
void *arena_cache_get(size_t bytes, size_t *got);
void arena_cache_put(void *p, size_t bytes);

struct Arena {
    char *base = nullptr;
    size_t cap = 0, off = 0;
    int depth = 0;
    void *overflow[64];
    size_t overflow_size[64];
    int n_overflow = 0;
    size_t overflow_bytes = 0;
    int alloc(void **p, size_t n)
    {
        n = (n + 255) & ~(size_t)255;
        if (!n) n = 256;
        if (off + n <= cap) { *p = base + off; off += n; return GB_OK; }
        if (n_overflow == 64) { set_error("scratch arena exhausted"); return GB_E_OOM; }
        size_t got = 0;
        *p = arena_cache_get(n, &got);
        if (!*p) {
            GB_CUDA(cudaMalloc(p, n));
            got = n;
        }
        overflow[n_overflow] = *p;
        overflow_size[n_overflow++] = got;
        overflow_bytes += n;
        return GB_OK;
    }
    void reset()
    {
        if (n_overflow) { // grow: one block large enough for what the last call needed
            cudaDeviceSynchronize();
            for (int i = 0; i < n_overflow; i++) arena_cache_put(overflow[i], overflow_size[i]);
            if (base) arena_cache_put(base, cap);
            size_t want = cap + overflow_bytes + (cap + overflow_bytes) / 4, got = 0;
            base = (char *)arena_cache_get(want, &got);
            if (base) cap = got;
            else {
                cap = cudaMalloc((void **)&base, want) == cudaSuccess ? want : 0;
                if (!cap) base = nullptr;
                cudaGetLastError();
            }
            n_overflow = 0;
            overflow_bytes = 0;
        }
        off = 0;
    }
    // make the base block at least `want` bytes before a call that will make many allocations (each allocation beyond the
    // base costs one of the 64 overflow entries).  Only at the start of a scope, before anything was handed out.
    int reserve(size_t want)
    {
        if (off != 0 || n_overflow != 0 || want <= cap) return GB_OK;
        cudaDeviceSynchronize(); // the old block may still be in use by work queued earlier
        if (base) arena_cache_put(base, cap);
        size_t got = 0;
        base = (char *)arena_cache_get(want, &got);
        if (base) { cap = got; return GB_OK; }
        cap = 0;
        if (cudaMalloc((void **)&base, want) != cudaSuccess) {
            base = nullptr;
            cudaGetLastError();
            set_error("out of device memory for %zu bytes of scratch", want);
            return GB_E_OOM;
        }
        cap = want;
        return GB_OK;
    }
    void destroy()
    {
        for (int i = 0; i < n_overflow; i++) arena_cache_put(overflow[i], overflow_size[i]);
        if (base) arena_cache_put(base, cap);
        base = nullptr;
        cap = off = overflow_bytes = 0;
        n_overflow = 0;
    }
};
extern thread_local Arena *tl_arena;
struct ArenaScope {
    Arena *a, *prev;
    explicit ArenaScope(Arena *arena) : a(arena), prev(tl_arena)
    {
        if (a->depth++ == 0) a->reset();
        tl_arena = a;
    }
    ~ArenaScope()
    {
        a->depth--;
        tl_arena = prev;
    }
};

struct DeviceBuf {
    void *p = nullptr;
    bool owned = false;
    DeviceBuf() = default;
    DeviceBuf(const DeviceBuf &) = delete;
    DeviceBuf &operator=(const DeviceBuf &) = delete;
    ~DeviceBuf() { release(); }
    void release()
    {
        if (p && owned) cudaFree(p);
        p = nullptr;
    }
    // from the current arena when a top-level call opened one, else a plain cudaMalloc
    int alloc(size_t n, cudaStream_t = nullptr)
    {
        release();
        if (tl_arena) { owned = false; return tl_arena->alloc(&p, n ? n : 16); }
        owned = true;
        GB_CUDA(cudaMalloc(&p, n ? n : 16));
        return GB_OK;
    }
};

// ---------------------------------------------------------------- host-side handle state
struct Comm;

struct Map {
    int k = 0;
    int device = 0;
    bool v210 = false;
    bool noncanonical = false; // keys were inserted through update/update_counts as-is
    unsigned long long cap = 0; // capacity in slots (any multiple of 1024)
    void *table = nullptr;     // keys | counts | vids of `cap` slots, packed at the front of the allocation (table_view)
    unsigned long long alloc_cap = 0; // the allocation behind `table` holds this many slots (>= cap)
    void *spare = nullptr;     // the other table allocation of the clear / filter cycle, kept for reuse
    Table view() const { return table_view(table, cap); }
    unsigned long long spare_cap = 0;
    unsigned long long *stage = nullptr; // key staging of the partitioned insert and of the filter (grow-only)
    size_t stage_cap = 0;
    int64_t size = 0;          // live keys (host mirror, exact after every call)
    int64_t grows = 0, windows = 0, last_insert_ns = 0, fixed_stride = 0;
    int64_t phase_ns[3] = { 0, 0, 0 }; // last partitioned insert: count, scatter, upsert (0 = direct path used)
    struct PartWork *part = nullptr;   // workspaces of the partitioned insert (partition.cuh), one per staging half
    struct PartWork *part2 = nullptr;
    cudaEvent_t pe[4] = { nullptr, nullptr, nullptr, nullptr };   // bucket-stream begin/end, scratch
    cudaEvent_t pready[2] = { nullptr, nullptr }, pfree[2] = { nullptr, nullptr }; // staging half filled / consumed
    cudaEvent_t pup[16];               // upsert begin/end per sub-batch (timing)
    int n_pup = 0;
    cudaStream_t stream = nullptr;      // every mutating call is asynchronous on this stream
    cudaStream_t copy_stream = nullptr; // high priority: the bucket pass of an overlapped (sub-batched) insert
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, t0 = nullptr, t1 = nullptr;
    cudaEvent_t fev[4] = { nullptr, nullptr, nullptr, nullptr }; // phase boundaries of deleteAll / Graph.buildGraph (gb_map_phase_ns)
    int64_t filter_ns[2] = { 0, 0 };   // last deleteAll: table sweep (compact_survivors_kernel), re-insert of the survivors
    int64_t slots_swept = 0;           // slots the last deleteAll streamed
    int64_t graph_ns[4] = { 0, 0, 0, 0 }; // last Graph.buildGraph on this map: membership probes, list ranking (ns), jump launches, vertices
    // scratch
    unsigned long long *d_counters = nullptr; // [0] new keys [1] survivors / exported [2] stream flags [3] k-windows
    unsigned long long *d_spread = nullptr;   // spread new-key tallies of insert_keys_kernel (partition.cu)
    Comm *comm = nullptr;
    // after deleteAll (or in a replica): the stored keys as one device array whose index IS the vertex id written in
    // the slots, so Graph.buildGraph needs no numbering pass.  Any mutation invalidates it.
    const unsigned long long *kept_keys = nullptr;
    int64_t kept_n = 0;
    bool kept_valid = false;
    uint8_t *fp = nullptr;     // fingerprint per slot, valid together with kept_valid (probe_find)
    unsigned long long fp_cap = 0;
    Map *replica = nullptr;    // sharded maps: the all-gathered copy Graph.buildGraph runs on (comm.cu)
    Arena arena;
};


// sharded Graph.buildGraph: this rank computes the membership masks of vertices [lo, hi) only; gather(ctx, base, elem)
// makes every rank's range of the device array `base` (elem bytes per vertex) visible on all ranks
struct ShardPlan {
    unsigned long long lo = 0, hi = 0;
    int (*gather)(void *ctx, void *base, size_t elem_bytes) = nullptr;
    void *ctx = nullptr;
};
int graph_build_sharded(gb_map *h, gb_graph **out, const ShardPlan *sp);
int check_map(gb_map *h, Map **m);
// assert(key.length == k) of apply / contains (S/ds/ArrayDNAMap.scala:182-206): a query with bits above 2k is an error
int check_keys(const Map *m, const uint64_t *keys, int64_t n);

inline unsigned int grid_for(unsigned long long n, int threads, int per_sm = 16)
{
    unsigned long long g = (n + threads - 1) / threads;
    unsigned long long cap = (unsigned long long)SM_COUNT * per_sm;
    if (g > cap) g = cap;
    return (unsigned int)(g ? g : 1);
}
int map_swap_table(Map *m, unsigned long long new_cap, void **old_table, unsigned long long *old_alloc_cap);
void map_retire_table(Map *m, void *t, unsigned long long alloc_cap);
int map_stage(Map *m, size_t n_u64);
int pool_setup(int device);
inline unsigned long long cap_for(int64_t keys)
{
    // load <= 1/3 at `keys`, a multiple of 1024 slots, at least 1024.  Measured on C2 (profiles/r2i_*: counting table at 3 / 2.5 / 2 /
    // 1.5 slots per expected key): insert 1.95 / 2.16 / 2.25 / 2.63 ms -- the probing a fuller table needs costs more than its
    // smaller clear and filter sweep give back (0.34 -> 0.16 ms); Graph.buildGraph's membership probes want the low load anyway
    unsigned long long c = ((unsigned long long)(keys > 0 ? keys : 0) * 3 + 1023) / 1024 * 1024;
    return c < 1024 ? 1024 : c;
}

} // namespace gb
