// walk.cu -- paired-end path support and node splitting on the device-resident graph.
// Replaces the pair loop, the WalkingActor round trips and the per-node matrix sweep of GraphSimplifier.startup
// (S/scripts/GraphSimplifier.scala:188-317, relative to /root/reference); the per-item logic is in walk.cuh.
//
//   gb_graph_pair_support   getGraphMap -> device multimap; one thread per read pair looks up the four first-k-mer position
//                           lists and applies annotate; the surviving orientation cases are compacted into a work list; one
//                           thread per case runs the bitset walk on a private local edge table in global scratch.  Cases whose
//                           neighbourhood outgrows the table are retried by the same kernel with a larger table (3 tiers).
//   gb_graph_split_nodes    one thread per node: thresholded in x out matrix, bipartite components, new node copies, edge ends
//                           rewired in place, unsupported edges dropped through the graph rewrite of graph.cu.
// Latency-bound graph search (pointer chasing through out4 / edge_off), not a bandwidth kernel; unmeasured so far.
#include <vector>

#include "common.cuh"
#include "extract.cuh"
#include "graph_types.cuh"
#include "scan.cuh"
#include "walk.cuh"

namespace gb {

__global__ void walk_out_table_kernel(GraphView g, unsigned int *out4)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.n_edges) return;
    out4[4ull * g.edge_start[e] + base_at(g.bases, g.edge_off[e])] = (unsigned int)e;
}

// in-edges of every node by preceding base (walk.cuh: in_slot_base); a clash means the graph is not a de Bruijn graph
__global__ void walk_in_table_kernel(GraphView g, unsigned int *in4, unsigned int *bad)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= g.n_edges) return;
    if (atomicExch(&in4[4ull * g.edge_end[e] + in_slot_base(g, (unsigned int)e)], (unsigned int)e) != NONE32) atomicOr(bad, 1u);
}

// putNew (S/ds/ArrayDNAMap.scala:152-162) of every getGraphMap entry: first free slot from the key's home
__global__ void posmap_insert_kernel(const unsigned long long *kmer, unsigned long long n, unsigned int *slot, unsigned long long cap)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long s = slot_of(mix64(kmer[i]), cap);
    while (atomicCAS(&slot[s], NONE32, (unsigned int)i) != NONE32) s = next_slot(s, cap);
}

// fixed-stride check: equal length bytes at i * rec imply that the record chain is i * rec
__global__ void walk_verify_fixed_kernel(const uint8_t *bin, unsigned int rec, unsigned int len0, unsigned long long n_reads,
                                         unsigned int *ragged)
{
    unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads && bin[r * rec] != len0) atomicOr(ragged, 1u);
}

struct PairStream {
    const uint8_t *bin;
    const unsigned long long *off; // record offsets, or nullptr when every record has rec_bytes bytes
    unsigned int rec_bytes;
    unsigned long long n_pairs;
};

// counters: [0] cases in the work list [1] cases in the overflow list [2] badPairs [3] walked cases [4] error flags
constexpr int WC_LIST = 0, WC_OVERFLOW = 1, WC_BAD = 2, WC_WALKED = 3, WC_ERROR = 4;

// one thread per pair (GraphSimplifier.scala:213-219): both reads at least k long, then the two orientation cases
// (p1, p2) and (p2, p1); a case that annotate keeps and whose position lists are non-empty goes to the work list as (x, y)
__global__ void walk_filter_kernel(GraphView g, PosMap m, PairStream ps, int lo, int hi, unsigned long long *cases,
                                   unsigned long long *counters)
{
    unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ps.n_pairs) return;
    const unsigned long long o1 = ps.off ? ps.off[2 * p] : 2 * p * ps.rec_bytes;
    const unsigned long long o2 = ps.off ? ps.off[2 * p + 1] : (2 * p + 1) * ps.rec_bytes;
    if ((int)ps.bin[o1] < g.k || (int)ps.bin[o2] < g.k) return;
    const unsigned long long a = record_first_kmer(ps.bin, o1, g.k), b = record_first_kmer(ps.bin, o2, g.k);
    for (int c = 0; c < 2; c++) {
        const unsigned long long x = c ? b : a, y = c ? a : b;
        Pos p1[WALK_MAXPOS], p2[WALK_MAXPOS];
        int n1, n2;
        const int r = case_positions(g, m, x, y, lo, hi, p1, &n1, p2, &n2);
        if (r == CASE_TOO_MANY_POSITIONS) atomicOr(&counters[WC_ERROR], 1ull);
        if (r != CASE_WALKED) continue;
        const unsigned long long at = atomicAdd(&counters[WC_LIST], 1ull);
        cases[2 * at] = x;
        cases[2 * at + 1] = y;
    }
}

// one thread per surviving case; worker w owns entries [w * lmax, (w + 1) * lmax) of `scratch`
__global__ void __launch_bounds__(128)
walk_cases_kernel(GraphView g, PosMap m, const unsigned long long *cases, unsigned long long n_cases, int lo, int hi,
                  WalkEntry *scratch, int lmax, unsigned long long n_workers, unsigned int *support,
                  unsigned long long *overflow_cases, unsigned long long *counters)
{
    const unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_workers) return;
    WalkTable t;
    t.e = scratch + w * (unsigned long long)lmax;
    t.cap = lmax;
    t.n = 0;
    for (unsigned long long c = w; c < n_cases; c += n_workers) {
        const unsigned long long x = cases[2 * c], y = cases[2 * c + 1];
        const int r = process_case(g, m, x, y, lo, hi, t, support, &counters[WC_BAD]);
        if (r == CASE_WALKED) atomicAdd(&counters[WC_WALKED], 1ull);
        else if (r == CASE_OVERFLOW) {
            const unsigned long long at = atomicAdd(&counters[WC_OVERFLOW], 1ull);
            overflow_cases[2 * at] = x;
            overflow_cases[2 * at + 1] = y;
        }
    }
}

__global__ void split_count_kernel(GraphView g, const unsigned int *in4, const unsigned int *support, int cutoff,
                                   unsigned long long *n_new)
{
    unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.n_nodes) return;
    n_new[v] = (unsigned long long)split_node(in4 + 4 * v, g.out4 + 4 * v, support, cutoff).n_new;
}

// replaceEnd / replaceStart (S/data/graph/Graph.scala:197-209) onto the new copies; in-edges alone in their component and
// out-edges no component reached are flagged for removal (GraphSimplifier.scala:301-302,309)
__global__ void split_apply_kernel(GraphView g, const unsigned int *in4, const unsigned int *support, int cutoff,
                                   const unsigned long long *new_base, unsigned long long *node_kmer2, unsigned int *edge_start,
                                   unsigned int *edge_end, unsigned int *kill, unsigned long long *n_killed)
{
    unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.n_nodes) return;
    const SplitPlan p = split_node(in4 + 4 * v, g.out4 + 4 * v, support, cutoff);
    const unsigned long long first = g.n_nodes + new_base[v];
    for (int c = 0; c < p.n_new; c++) node_kmer2[first + c] = g.node_kmer[v]; // addNode(node.seq) (303)
    for (int s = 0; s < 4; s++) {
        if (p.in_comp[s] >= 0) edge_end[in4[4 * v + s]] = (unsigned int)(first + p.in_comp[s]);
        else if (p.in_comp[s] == -1 && atomicExch(&kill[in4[4 * v + s]], 1u) == 0) atomicAdd(n_killed, 1ull);
        if (p.out_comp[s] >= 0) edge_start[g.out4[4 * v + s]] = (unsigned int)(first + p.out_comp[s]);
        else if (p.out_comp[s] == -1 && atomicExch(&kill[g.out4[4 * v + s]], 1u) == 0) atomicAdd(n_killed, 1ull);
    }
}

// getAll / contains (S/ds/ArrayDNAMap.scala:103-113,232) for n keys, one key per thread (walk.cuh: posmap_lookup)
__global__ void posmap_lookup_kernel(PosMap m, const unsigned long long *keys, long long n, int max_per, unsigned int *ids,
                                     unsigned int *dists, unsigned int *counts)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    counts[i] = posmap_lookup(m, keys[i], max_per, ids ? ids + i * max_per : nullptr, dists ? dists + i * max_per : nullptr);
}

// the DNAMap[GraphPosition] of Graph.getGraphMap as its own handle: a snapshot, independent of the graph it came from
struct GraphMap {
    int k = 0, device = 0;
    cudaStream_t stream = nullptr;
    int64_t n = 0;
    unsigned long long cap = 0;
    unsigned long long *kmer = nullptr;
    unsigned int *id = nullptr, *dist = nullptr, *slot = nullptr;
    Arena arena; // scratch of the lookups
    PosMap view() const
    {
        PosMap m;
        m.slot = slot; m.cap = cap; m.kmer = kmer; m.id = id; m.dist = dist;
        return m;
    }
};

static int graph_map_free(GraphMap *gm)
{
    if (gm->stream) { cudaStreamSynchronize(gm->stream); cudaStreamDestroy(gm->stream); }
    cudaFree(gm->kmer); cudaFree(gm->id); cudaFree(gm->dist); cudaFree(gm->slot);
    gm->arena.destroy();
    delete gm;
    return GB_OK;
}

static GraphView view_of(const Graph *g, const unsigned int *out4)
{
    GraphView v;
    v.k = g->k;
    v.n_nodes = (unsigned long long)g->n_nodes;
    v.n_edges = (unsigned long long)g->n_edges;
    v.node_kmer = g->node_kmer;
    v.edge_start = g->edge_start;
    v.edge_end = g->edge_end;
    v.edge_off = g->edge_off;
    v.bases = g->bases;
    v.out4 = out4;
    return v;
}

#define WLAUNCH(kernel, n, threads, ...)                                                                      \
    do {                                                                                                      \
        unsigned long long _n = (unsigned long long)(n);                                                      \
        if (_n) {                                                                                             \
            kernel<<<(unsigned int)((_n + (threads) - 1) / (threads)), (threads), 0, st>>>(__VA_ARGS__);      \
            GB_LAUNCHED();                                                                                    \
        }                                                                                                     \
    } while (0)

// CUDA events of one call (timings reported through gb_graph_stats)
struct Events {
    cudaEvent_t e[4] = { nullptr, nullptr, nullptr, nullptr };
    ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    int init()
    {
        for (cudaEvent_t &x : e) GB_CUDA(cudaEventCreate(&x));
        return GB_OK;
    }
    int64_t ns(int a, int b) const
    {
        float ms = 0;
        return cudaEventElapsedTime(&ms, e[a], e[b]) == cudaSuccess ? (int64_t)(ms * 1e6) : 0;
    }
};

static int read_counters(const unsigned long long *d, unsigned long long *h, int n, cudaStream_t st)
{
    GB_CUDA(cudaMemcpyAsync(h, d, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    return GB_OK;
}

} // namespace gb

using namespace gb;

extern "C" int gb_graph_pair_support(gb_graph *h, const uint8_t *bin, size_t n_bytes, int64_t n_pairs, int range_lo, int range_hi,
                                     uint32_t *support, int64_t *bad_pairs, int64_t *walked_cases)
{
    if (!h) { set_error("null graph handle"); return GB_E_ARG; }
    Graph *g = reinterpret_cast<Graph *>(h);
    if (bad_pairs) *bad_pairs = 0;
    if (walked_cases) *walked_cases = 0;
    if (n_pairs < 0 || (n_pairs > 0 && !bin) || !support) { set_error("bad arguments"); return GB_E_ARG; }
    if (range_lo < 0 || range_hi < range_lo || range_hi > WALK_MAX_RANGE) {
        set_error("range %d..%d not supported (0 <= first <= last <= %d)", range_lo, range_hi, WALK_MAX_RANGE);
        return GB_E_ARG;
    }
    GB_CUDA(cudaSetDevice(g->device));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
    memset(support, 0, (size_t)E * 4 * sizeof(uint32_t));
    if (n_pairs == 0 || E == 0) return GB_OK;

    const int64_t n_pos = g->n_nodes + g->n_bases - g->n_edges; // Graph.scala:97
    if (n_pos >= (int64_t)NONE32) { set_error("graph map of %lld entries exceeds 32-bit entry indices", (long long)n_pos); return GB_E_CAPACITY; }
    if (n_bytes == 0) { set_error("truncated .bin stream at read 0"); return GB_E_ARG; }

    // PairedEndData.getPairs framing (S/data/PairedEndData.scala:20-36).  Fixed-stride fast path as in gb_map_insert_reads: if
    // the stream is long enough for 2 * n_pairs records of the first record's size, copy those bytes and let the device
    // compare every length byte; otherwise (or if they differ) the record chain is scanned on the host and travels along.
    const int64_t n_reads = 2 * n_pairs;
    const unsigned int len0 = bin[0], rec = 1 + (len0 + 3) / 4;
    bool fixed = (unsigned long long)n_reads * rec <= n_bytes;
    DeviceBuf d_bin, d_off;
    size_t used = 0;
    if (fixed) {
        used = (size_t)n_reads * rec;
        GB_TRY(d_bin.alloc(used + 16));
        GB_CUDA(cudaMemcpyAsync(d_bin.p, bin, used, cudaMemcpyHostToDevice, st));
        Tmp<unsigned int> ragged;
        GB_TRY(ragged.alloc(1, st));
        GB_TRY(ragged.zero(1));
        WLAUNCH(walk_verify_fixed_kernel, n_reads, 256, (const uint8_t *)d_bin.p, rec, len0, (unsigned long long)n_reads, ragged.p);
        unsigned int r = 0;
        GB_CUDA(cudaMemcpyAsync(&r, ragged.p, 4, cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaStreamSynchronize(st));
        fixed = r == 0;
    }
    if (!fixed) {
        std::vector<unsigned long long> off;
        std::vector<int64_t> winp;
        GB_TRY(scan_records(bin, n_bytes, n_reads, g->k, off, winp));
        const size_t need = (size_t)off[(size_t)n_reads];
        if (need > used) {
            GB_TRY(d_bin.alloc(need + 16));
            GB_CUDA(cudaMemcpyAsync(d_bin.p, bin, need, cudaMemcpyHostToDevice, st));
        }
        GB_TRY(d_off.alloc(off.size() * 8));
        GB_CUDA(cudaMemcpyAsync(d_off.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, st));
        GB_CUDA(cudaStreamSynchronize(st)); // `off` goes out of scope
    }

    Tmp<unsigned int> out4, slot, pid, pdist, d_support;
    Tmp<unsigned long long> pk, counters, cases[2];
    Events ev;
    GB_TRY(ev.init());
    GB_CUDA(cudaEventRecord(ev.e[0], st));
    GB_TRY(out4.alloc(4 * N, st));
    GB_TRY(out4.fill_ff(4 * N));
    GraphView gv = view_of(g, out4.p);
    WLAUNCH(walk_out_table_kernel, E, 256, gv, out4.p);

    // getGraphMap: the same entries gb_graph_positions exports (graphmap.cu), indexed for getAll
    GB_TRY(pk.alloc((size_t)n_pos, st));
    GB_TRY(pid.alloc((size_t)n_pos, st));
    GB_TRY(pdist.alloc((size_t)n_pos, st));
    GB_TRY(graph_positions_device(g, pk.p, pid.p, pdist.p));
    const unsigned long long cap = ((unsigned long long)n_pos * 2 + 1024) / 1024 * 1024; // load <= 1/2
    GB_TRY(slot.alloc(cap, st));
    GB_TRY(slot.fill_ff(cap));
    WLAUNCH(posmap_insert_kernel, n_pos, 256, pk.p, (unsigned long long)n_pos, slot.p, cap);
    PosMap pm;
    pm.slot = slot.p; pm.cap = cap; pm.kmer = pk.p; pm.id = pid.p; pm.dist = pdist.p;

    GB_TRY(d_support.alloc(4 * E, st));
    GB_TRY(d_support.zero(4 * E));
    GB_TRY(counters.alloc(8, st));
    GB_TRY(counters.zero(8));
    GB_TRY(cases[0].alloc(4 * (size_t)n_pairs, st)); // two cases of two k-mers per pair
    GB_CUDA(cudaEventRecord(ev.e[1], st));

    PairStream ps;
    ps.bin = (const uint8_t *)d_bin.p;
    ps.off = fixed ? nullptr : (const unsigned long long *)d_off.p;
    ps.rec_bytes = rec;
    ps.n_pairs = (unsigned long long)n_pairs;
    WLAUNCH(walk_filter_kernel, n_pairs, 256, gv, pm, ps, range_lo, range_hi, cases[0].p, counters.p);
    GB_CUDA(cudaEventRecord(ev.e[2], st));
    unsigned long long c[8];
    GB_TRY(read_counters(counters.p, c, 8, st));
    if (c[WC_ERROR]) { set_error("a k-mer occupies more than %d graph positions", WALK_MAXPOS); return GB_E_CAPACITY; }

    // tiers of local-table size: every surviving case first, then only those that outgrew the previous tier
    static const int tier_lmax[3] = { 32, 512, 8192 };
    static const unsigned long long tier_workers[3] = { 148ull * 128, 2048, 128 };
    unsigned long long n_cases = c[WC_LIST];
    g->stats[7] = (int64_t)n_cases;
    int cur = 0;
    for (int tier = 0; tier < 3 && n_cases; tier++) {
        const unsigned long long workers = n_cases < tier_workers[tier] ? n_cases : tier_workers[tier];
        Tmp<WalkEntry> scratch;
        GB_TRY(scratch.alloc((size_t)workers * tier_lmax[tier], st));
        GB_TRY(cases[cur ^ 1].alloc(2 * (size_t)n_cases, st));
        GB_CUDA(cudaMemsetAsync(counters.p + WC_OVERFLOW, 0, 8, st));
        WLAUNCH(walk_cases_kernel, workers, 128, gv, pm, cases[cur].p, n_cases, range_lo, range_hi, scratch.p, tier_lmax[tier],
                workers, d_support.p, cases[cur ^ 1].p, counters.p);
        GB_TRY(read_counters(counters.p, c, 8, st));
        n_cases = c[WC_OVERFLOW];
        cur ^= 1;
    }
    if (n_cases) {
        set_error("%llu read pairs see more than %d edges within %d bases: walk table exhausted", n_cases, tier_lmax[2], range_hi);
        return GB_E_CAPACITY;
    }
    GB_CUDA(cudaEventRecord(ev.e[3], st));
    GB_CUDA(cudaMemcpyAsync(support, d_support.p, (size_t)E * 16, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    g->stats[4] = ev.ns(0, 1);
    g->stats[5] = ev.ns(1, 2);
    g->stats[6] = ev.ns(2, 3);
    if (bad_pairs) *bad_pairs = (int64_t)c[WC_BAD];
    if (walked_cases) *walked_cases = (int64_t)c[WC_WALKED];
    return GB_OK;
}

extern "C" int gb_graph_split_nodes(gb_graph *h, const uint32_t *support, int32_t cutoff, int64_t *edges_removed, int64_t *nodes_added)
{
    if (!h) { set_error("null graph handle"); return GB_E_ARG; }
    Graph *g = reinterpret_cast<Graph *>(h);
    if (edges_removed) *edges_removed = 0;
    if (nodes_added) *nodes_added = 0;
    if (!support && g->n_edges) { set_error("bad arguments"); return GB_E_ARG; }
    GB_CUDA(cudaSetDevice(g->device));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
    if (!E) return GB_OK;

    Tmp<unsigned int> out4, in4, d_support, kill, bad;
    Tmp<unsigned long long> n_new, totals;
    GB_TRY(out4.alloc(4 * N, st));
    GB_TRY(out4.fill_ff(4 * N));
    GB_TRY(in4.alloc(4 * N, st));
    GB_TRY(in4.fill_ff(4 * N));
    GB_TRY(bad.alloc(1, st));
    GB_TRY(bad.zero(1));
    GraphView gv = view_of(g, out4.p);
    WLAUNCH(walk_out_table_kernel, E, 256, gv, out4.p);
    WLAUNCH(walk_in_table_kernel, E, 256, gv, in4.p, bad.p);
    GB_TRY(d_support.alloc(4 * E, st));
    GB_CUDA(cudaMemcpyAsync(d_support.p, support, (size_t)E * 16, cudaMemcpyHostToDevice, st));
    GB_TRY(n_new.alloc(N, st));
    GB_TRY(totals.alloc(2, st));
    GB_TRY(totals.zero(2));
    WLAUNCH(split_count_kernel, N, 256, gv, in4.p, d_support.p, cutoff, n_new.p);
    GB_TRY(exclusive_scan_u64(n_new.p, N, totals.p + 0, st));
    unsigned int b = 0;
    GB_CUDA(cudaMemcpyAsync(&b, bad.p, 4, cudaMemcpyDeviceToHost, st));
    unsigned long long tot[2];
    GB_TRY(read_counters(totals.p, tot, 2, st));
    if (b) { set_error("two in-edges of a node share their preceding base: not a de Bruijn graph"); return GB_E_INVARIANT; }
    const unsigned long long added = tot[0];
    if (N + added >= (unsigned long long)NONE32) { set_error("node count exceeds 32-bit indices"); return GB_E_CAPACITY; }

    unsigned long long *node_kmer2 = nullptr;
    GB_TRY(g->store[g->cur].alloc((void **)&node_kmer2, (size_t)(N + added) * 8));
    GB_CUDA(cudaMemcpyAsync(node_kmer2, g->node_kmer, (size_t)N * 8, cudaMemcpyDeviceToDevice, st));
    GB_TRY(kill.alloc(E, st));
    GB_TRY(kill.zero(E));
    WLAUNCH(split_apply_kernel, N, 256, gv, in4.p, d_support.p, cutoff, n_new.p, node_kmer2, g->edge_start, g->edge_end, kill.p,
            totals.p + 1);
    GB_TRY(read_counters(totals.p, tot, 2, st));
    g->node_kmer = node_kmer2;
    g->n_nodes = (int64_t)(N + added);
    if (nodes_added) *nodes_added = (int64_t)added;
    if (edges_removed) *edges_removed = (int64_t)tot[1];
    if (tot[1]) GB_TRY(graph_remove_flagged(g, kill.p)); // toRemove.foreach(removeEdge) (316)
    return GB_OK;
}

// ---------------------------------------------------------------- DNAMap[GraphPosition]
extern "C" int gb_graph_map_create(gb_graph *h, gb_graph_map **out)
{
    if (!h || !out) { set_error("null argument"); return GB_E_ARG; }
    *out = nullptr;
    Graph *g = reinterpret_cast<Graph *>(h);
    GB_CUDA(cudaSetDevice(g->device));
    const int64_t n = g->n_nodes + g->n_bases - g->n_edges; // Graph.scala:97
    if (n >= (int64_t)NONE32) { set_error("graph map of %lld entries exceeds 32-bit entry indices", (long long)n); return GB_E_CAPACITY; }
    GraphMap *gm = new GraphMap();
    gm->k = g->k; gm->device = g->device; gm->n = n;
    gm->cap = ((unsigned long long)n * 2 + 1024) / 1024 * 1024; // load <= 1/2
    int rc = GB_OK;
    auto fail = [&](int code) { graph_map_free(gm); return code; };
    if (cudaStreamCreateWithFlags(&gm->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(GB_E_CUDA); }
    if (cudaMalloc((void **)&gm->kmer, (size_t)(n ? n : 1) * 8) != cudaSuccess || cudaMalloc((void **)&gm->id, (size_t)(n ? n : 1) * 4) != cudaSuccess ||
        cudaMalloc((void **)&gm->dist, (size_t)(n ? n : 1) * 4) != cudaSuccess || cudaMalloc((void **)&gm->slot, (size_t)gm->cap * 4) != cudaSuccess) {
        cudaGetLastError();
        set_error("out of device memory for a graph map of %lld entries", (long long)n);
        return fail(GB_E_OOM);
    }
    {
        // the entries are produced on the graph's stream (they read the graph's arrays), the index on it too; the handle's own
        // stream takes over once everything is in place
        cudaStream_t st = g->stream;
        if ((rc = graph_positions_device(g, gm->kmer, gm->id, gm->dist)) != GB_OK) return fail(rc);
        if (cudaMemsetAsync(gm->slot, 0xFF, (size_t)gm->cap * 4, st) != cudaSuccess) { set_error("cudaMemset failed"); return fail(GB_E_CUDA); }
        if (n) {
            posmap_insert_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(gm->kmer, (unsigned long long)n, gm->slot, gm->cap);
            note_launch();
            if (cudaGetLastError() != cudaSuccess) { set_error("posmap_insert_kernel launch failed"); return fail(GB_E_CUDA); }
        }
        if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("graph map build failed"); return fail(GB_E_CUDA); }
    }
    *out = reinterpret_cast<gb_graph_map *>(gm);
    return GB_OK;
}

extern "C" int gb_graph_map_destroy(gb_graph_map *h)
{
    if (!h) return GB_OK;
    GraphMap *gm = reinterpret_cast<GraphMap *>(h);
    cudaSetDevice(gm->device);
    return graph_map_free(gm);
}

extern "C" int gb_graph_map_size(gb_graph_map *h, int64_t *size)
{
    if (!h || !size) { set_error("null argument"); return GB_E_ARG; }
    *size = reinterpret_cast<GraphMap *>(h)->n;
    return GB_OK;
}

extern "C" int gb_graph_map_get_all(gb_graph_map *h, const uint64_t *keys, int64_t n, int max_per_key, uint32_t *ids, uint32_t *dists,
                                    uint32_t *counts)
{
    if (!h) { set_error("null graph map handle"); return GB_E_ARG; }
    GraphMap *gm = reinterpret_cast<GraphMap *>(h);
    if (n < 0 || max_per_key < 0 || (n > 0 && (!keys || !counts))) { set_error("bad arguments"); return GB_E_ARG; }
    if (n == 0) return GB_OK;
    const unsigned long long kmask = (1ull << (2 * gm->k)) - 1;
    for (int64_t i = 0; i < n; i++)
        if (keys[i] & ~kmask) { set_error("key %lld is longer than k = %d", (long long)i, gm->k); return GB_E_K_RANGE; }
    GB_CUDA(cudaSetDevice(gm->device));
    ArenaScope scope(&gm->arena);
    cudaStream_t st = gm->stream;
    const bool want = max_per_key > 0 && (ids || dists);
    Tmp<unsigned long long> dk;
    Tmp<unsigned int> di, dd, dc;
    GB_TRY(dk.alloc((size_t)n, st));
    GB_TRY(dc.alloc((size_t)n, st));
    if (want && ids) GB_TRY(di.alloc((size_t)n * max_per_key, st));
    if (want && dists) GB_TRY(dd.alloc((size_t)n * max_per_key, st));
    GB_CUDA(cudaMemcpyAsync(dk.p, keys, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    if (di.p) GB_CUDA(cudaMemsetAsync(di.p, 0xFF, (size_t)n * max_per_key * 4, st));
    if (dd.p) GB_CUDA(cudaMemsetAsync(dd.p, 0xFF, (size_t)n * max_per_key * 4, st));
    WLAUNCH(posmap_lookup_kernel, n, 256, gm->view(), dk.p, (long long)n, want ? max_per_key : 0, di.p, dd.p, dc.p);
    GB_CUDA(cudaMemcpyAsync(counts, dc.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (di.p) GB_CUDA(cudaMemcpyAsync(ids, di.p, (size_t)n * max_per_key * 4, cudaMemcpyDeviceToHost, st));
    if (dd.p) GB_CUDA(cudaMemcpyAsync(dists, dd.p, (size_t)n * max_per_key * 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    return GB_OK;
}
