// sgraph.cu -- the CUDA backend of the sharded Graph.buildGraph (sgraph.cuh; S/data/graph/Graph.scala:269-382 relative to
// /root/reference): one thread per item on the rank's stream, scratch from the map's arena, the resulting graph in the
// graph's store arena.  Entry points:
//   graph_build_on_fabric            called by gb_pmap_graph_build (comm.cu) with the NCCL + CUDA-IPC fabric, one rank per GPU
//   gb_graph_build_virtual_shards    the same build over P VIRTUAL ranks on one device (LocalFabric): the kept k-mers of a
//                                    single-GPU map are dealt to P ranks that live in one address space.  This is how the
//                                    device code of the sharded build is validated (and profiled) on a single GPU.
// The algorithm (functors + orchestration, the code compiled here) is also checked against the oracle on the CPU by
// tests/test_sgraph_emul_cpu.py through a g++ backend; on the device by tests/test_sgraph_gpu.py (virtual ranks) and
// tests/test_parity_multigpu.py (1, 2 and 8 ranks).  gb_pmap_graph_build runs this build by default (gb_tune pgraph_sharded;
// 0 = the replicated build).
#include <vector>

#include "common.cuh"
#include "extract.cuh"
#include "graph_types.cuh"
#include "scan.cuh"
#include "sgraph.cuh"

namespace gb {
namespace sg {

// Scratch of one build: a few large chunks (8 MB, doubling) taken from the calling handle's arena and handed out by bumping.
// A build makes ~15 allocations per rank, a virtual-shard build 16 times that; the arena itself keeps at most 64 blocks
// beyond its base, which such a build would exhaust on its first call (the arena regrows to one block for the next call).
struct Pool {
    Arena *arena = nullptr; // an ArenaScope is open on it
    char *cur = nullptr;
    size_t left = 0, next = (size_t)8 << 20;
    int alloc(void **p, size_t bytes)
    {
        bytes = (bytes + 255) & ~(size_t)255;
        if (!bytes) bytes = 256;
        if (bytes > left) {
            const size_t want = bytes > next ? bytes : next;
            if (next < ((size_t)1 << 30)) next *= 2;
            void *c = nullptr;
            GB_TRY(arena->alloc(&c, want));
            cur = (char *)c;
            left = want;
        }
        *p = cur;
        cur += bytes;
        left -= bytes;
        return GB_OK;
    }
};

struct Exec {
    cudaStream_t st = nullptr;
    Pool *scratch = nullptr;
    Arena *store = nullptr;   // the resulting graph's store arena
};

int sg_alloc(Exec &ex, void **p, size_t bytes) { return ex.scratch->alloc(p, bytes); }
int sg_graph_alloc(Exec &ex, void **p, size_t bytes) { return ex.store->alloc(p, bytes ? bytes : 16); }
int sg_zero(Exec &ex, void *p, size_t bytes)
{
    if (bytes) GB_CUDA(cudaMemsetAsync(p, 0, bytes, ex.st));
    return GB_OK;
}
int sg_fill_ff(Exec &ex, void *p, size_t bytes)
{
    if (bytes) GB_CUDA(cudaMemsetAsync(p, 0xFF, bytes, ex.st));
    return GB_OK;
}
int sg_copy(Exec &ex, void *dst, const void *src, size_t bytes)
{
    if (bytes) GB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ex.st));
    return GB_OK;
}
int sg_read(Exec &ex, void *host, const void *dev, size_t bytes)
{
    if (bytes) GB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ex.st));
    GB_CUDA(cudaStreamSynchronize(ex.st));
    return GB_OK;
}
int sg_write_host(Exec &ex, void *dev, const void *host, size_t bytes)
{
    if (bytes) GB_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ex.st));
    return GB_OK;
}
int sg_sync(Exec &ex)
{
    GB_CUDA(cudaStreamSynchronize(ex.st));
    return GB_OK;
}
int sg_scan(Exec &ex, u64 *data, u64 n, u64 *total_host)
{
    u64 *d_total;
    GB_TRY(sg_new(ex, &d_total, 1));
    GB_TRY(exclusive_scan_u64(data, n, d_total, ex.st));
    return sg_read(ex, total_host, d_total, 8);
}

template <class Op>
__global__ void __launch_bounds__(256) items_kernel(u64 n, Op op)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) op(i);
}
template <class Op>
int sg_launch(Exec &ex, u64 n, const Op &op)
{
    if (n == 0) return GB_OK;
    if ((n + 255) / 256 > 0x7FFFFFFFull) { set_error("sharded build: %llu items exceed one grid", n); return GB_E_CAPACITY; }
    items_kernel<Op><<<(unsigned int)((n + 255) / 256), 256, 0, ex.st>>>(n, op);
    GB_LAUNCHED();
    return GB_OK;
}

} // namespace sg

// the kept k-mers of each local rank come out of its map (all live keys; counts are not needed), the result goes into a
// fresh Graph handle
static int build_with(sg::Fabric &fab, Map *const *maps, const std::vector<std::pair<const unsigned long long *, unsigned long long>> *given,
                      cudaStream_t stream, sg::Pool *scratch, int device, int k, bool dual, bool v210, gb_graph **out)
{
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    GB_CUDA(cudaEventCreate(&ev0));
    GB_CUDA(cudaEventCreate(&ev1));
    struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } guard{ ev0, ev1 };
    const int nl = (int)fab.mine.size();
    GB_CUDA(cudaEventRecord(ev0, stream));
    Graph *g = new Graph();
    g->k = k;
    g->device = device;
    auto fail = [&](int rc) {
        cudaStreamSynchronize(stream);
        gb_graph_destroy(reinterpret_cast<gb_graph *>(g));
        return rc;
    };
    if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(GB_E_CUDA); }
    g->cur = 0;
    g->store[0].reset();
    std::vector<sg::Exec> ex((size_t)nl);
    std::vector<sg::RankInput> in((size_t)nl);
    for (int l = 0; l < nl; l++) {
        ex[(size_t)l].st = stream;
        ex[(size_t)l].scratch = scratch;
        ex[(size_t)l].store = &g->store[0];
        in[(size_t)l].ex = &ex[(size_t)l];
        if (given) {
            in[(size_t)l].keys = (*given)[(size_t)l].first;
            in[(size_t)l].n = (*given)[(size_t)l].second;
        } else {
            Map *m = maps[l];
            unsigned long long *keys = nullptr;
            int *vals = nullptr;
            int rc = sg::sg_new(ex[(size_t)l], &keys, (size_t)m->size);
            if (rc == GB_OK) rc = sg::sg_new(ex[(size_t)l], &vals, (size_t)m->size);
            if (rc == GB_OK) rc = map_export_device(m, keys, vals); // synchronises the map's stream
            if (rc != GB_OK) return fail(rc);
            in[(size_t)l].keys = keys;
            in[(size_t)l].n = (unsigned long long)m->size;
        }
    }
    sg::Result res;
    const int rc = sg::build(fab, in, k, dual, v210, &res);
    if (rc != GB_OK) return fail(rc);
    if (cudaEventRecord(ev1, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) { set_error("sharded build failed on the device"); return fail(GB_E_CUDA); }
    float ms = 0;
    cudaEventElapsedTime(&ms, ev0, ev1);
    g->node_kmer = res.node_kmer; g->edge_start = res.edge_start; g->edge_end = res.edge_end;
    g->edge_off = res.edge_off; g->bases = res.bases;
    g->n_nodes = (int64_t)res.n_nodes; g->n_edges = (int64_t)res.n_edges; g->n_bases = (int64_t)res.n_bases;
    g->stats[0] = (int64_t)res.kept;
    g->stats[1] = res.jump_rounds + res.seg_rounds;
    g->stats[2] = (int64_t)res.cycle_vertices;
    g->stats[3] = (int64_t)(ms * 1e6);
    g->stats[4] = (int64_t)res.segments; // shared with the pair-support timings, which a fresh graph does not have yet
    *out = reinterpret_cast<gb_graph *>(g);
    return GB_OK;
}

int graph_build_on_fabric(sg::Fabric &fab, Map *const *maps, cudaStream_t stream, bool dual, gb_graph **out)
{
    Map *m = maps[0];
    sg::Pool pool;
    pool.arena = &m->arena;
    return build_with(fab, maps, nullptr, stream, &pool, m->device, m->k, dual, m->v210, out);
}

} // namespace gb

using namespace gb;

extern "C" int gb_graph_build_virtual_shards(gb_map *h, int n_shards, gb_graph **out)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    ArenaScope scope(&m->arena);
    if (!out) { set_error("null out pointer"); return GB_E_ARG; }
    *out = nullptr;
    if (n_shards < 1 || n_shards > sg::MAXR) { set_error("n_shards must be in 1..%d", sg::MAXR); return GB_E_ARG; }
    // every live key of the map, dealt to the virtual ranks in table order
    const unsigned long long n = (unsigned long long)m->size;
    // ~120 B of scratch per key (export, staging, window, neighbours, three scan arrays) + a replicated segment list per rank
    GB_TRY(m->arena.reserve((size_t)n * (140 + 2 * (size_t)n_shards) + ((size_t)64 << 20)));
    DeviceBuf keys, vals;
    GB_TRY(keys.alloc((size_t)n * 8));
    GB_TRY(vals.alloc((size_t)n * 4));
    GB_TRY(map_export_device(m, (unsigned long long *)keys.p, (int *)vals.p));
    std::vector<std::pair<const unsigned long long *, unsigned long long>> given;
    for (int r = 0; r < n_shards; r++) {
        const unsigned long long lo = n * (unsigned long long)r / n_shards, hi = n * (unsigned long long)(r + 1) / n_shards;
        given.push_back({ (const unsigned long long *)keys.p + lo, hi - lo });
    }
    std::vector<sg::Exec *> none((size_t)n_shards, nullptr);
    // the fabric copies through the ranks' Execs, which build_with creates: a LocalFabric over one shared Exec does the same
    sg::Pool pool;
    pool.arena = &m->arena;
    sg::Exec shared;
    shared.st = m->stream;
    shared.scratch = &pool;
    shared.store = nullptr;
    for (auto &e : none) e = &shared;
    sg::LocalFabric fab(n_shards, none);
    return build_with(fab, nullptr, &given, m->stream, &pool, m->device, m->k, m->noncanonical, m->v210, out);
}
