// sgraph_emul.cpp -- TEST INFRASTRUCTURE: compiles genome_b200/csrc/sgraph.cuh (the sharded Graph.buildGraph: per-item ops
// AND the orchestration) with g++, with serial loops for the kernels, malloc for device memory and the in-process fabric for
// the collectives, so that the whole algorithm is checked against the oracle on a box without a GPU.  Not part of the
// product: nothing in genome_b200/ builds or loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../genome_b200/csrc/sgraph.cuh"
#include "superkmer.cuh"

namespace gb {
void set_error(const char *, ...) {}
thread_local Arena *tl_arena = nullptr;
namespace sg {

struct Exec {
    std::vector<void *> owned, graph;
    ~Exec()
    {
        for (void *p : owned) free(p);
        for (void *p : graph) free(p);
    }
};
int sg_alloc(Exec &ex, void **p, size_t bytes)
{
    // poison: a read of memory the algorithm never wrote should not look like a plausible value
    *p = malloc(bytes ? bytes : 1); // exact size: an address-sanitizer build of this harness sees every overrun
    if (!*p) return GB_E_OOM;
    memset(*p, 0xA5, bytes);
    ex.owned.push_back(*p);
    return GB_OK;
}
int sg_graph_alloc(Exec &ex, void **p, size_t bytes)
{
    *p = malloc(bytes ? bytes : 1);
    if (!*p) return GB_E_OOM;
    memset(*p, 0xA5, bytes);
    ex.graph.push_back(*p);
    return GB_OK;
}
int sg_zero(Exec &, void *p, size_t bytes) { memset(p, 0, bytes); return GB_OK; }
int sg_fill_ff(Exec &, void *p, size_t bytes) { memset(p, 0xFF, bytes); return GB_OK; }
int sg_copy(Exec &, void *dst, const void *src, size_t bytes) { memmove(dst, src, bytes); return GB_OK; }
int sg_read(Exec &, void *host, const void *dev, size_t bytes) { memcpy(host, dev, bytes); return GB_OK; }
int sg_write_host(Exec &, void *dev, const void *host, size_t bytes) { memcpy(dev, host, bytes); return GB_OK; }
int sg_sync(Exec &) { return GB_OK; }
int sg_scan(Exec &, u64 *data, u64 n, u64 *total_host)
{
    u64 t = 0;
    for (u64 i = 0; i < n; i++) { const u64 v = data[i]; data[i] = t; t += v; }
    *total_host = t;
    return GB_OK;
}
template <class Op> int sg_launch(Exec &, u64 n, const Op &op)
{
    for (u64 i = 0; i < n; i++) op(i);
    return GB_OK;
}

} // namespace sg
} // namespace gb

using namespace gb;
using namespace gb::sg;

// ---- one THREAD per rank, each with its own Fabric object that drives exactly one rank: the control flow of the
// one-process-per-GPU form (own copy of the global arrays, real sums, real exchanges), minus NCCL and CUDA IPC
struct Hub {
    int P;
    std::mutex mu;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long generation = 0;
    std::vector<const void *> ptr;
    std::vector<Row> row_a, row_b;
    explicit Hub(int n) : P(n), ptr((size_t)n), row_a((size_t)n), row_b((size_t)n) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long gen = generation;
        if (++waiting == P) { waiting = 0; generation++; cv.notify_all(); }
        else cv.wait(lk, [&] { return generation != gen; });
    }
};
struct ThreadFabric : Fabric {
    Hub &hub;
    Exec &ex;
    int me;
    ThreadFabric(Hub &h, Exec &e, int rank) : hub(h), ex(e), me(rank)
    {
        P = h.P;
        mine.push_back(rank);
    }
    int allgather_host(const void *const *contrib, void *const *all, size_t bytes) override
    {
        hub.ptr[(size_t)me] = contrib[0];
        hub.wait();
        for (int r = 0; r < P; r++) memcpy((char *)all[0] + (size_t)r * bytes, hub.ptr[(size_t)r], bytes);
        hub.wait();
        return GB_OK;
    }
    int windows(const size_t *bytes_of_rank, void **window, PeerPtrs *peers) override
    {
        GB_TRY(sg_alloc(ex, &window[0], bytes_of_rank[me]));
        hub.ptr[(size_t)me] = window[0];
        hub.wait();
        for (int r = 0; r < P; r++) peers[0].p[r] = (void *)hub.ptr[(size_t)r];
        hub.wait();
        return GB_OK;
    }
    int alltoallv_u64(const u64 *const *send, const Row *soff, const Row *scnt, u64 *const *recv, const Row *roff, const Row *rcnt) override
    {
        hub.ptr[(size_t)me] = send[0];
        hub.row_a[(size_t)me] = soff[0];
        hub.row_b[(size_t)me] = scnt[0];
        hub.wait();
        int rc = GB_OK;
        for (int p = 0; p < P; p++) {
            if (hub.row_b[(size_t)p].v[me] != rcnt[0].v[p]) rc = GB_E_INVARIANT;
            else memcpy(recv[0] + roff[0].v[p], (const u64 *)hub.ptr[(size_t)p] + hub.row_a[(size_t)p].v[me], (size_t)rcnt[0].v[p] * 8);
        }
        hub.wait();
        return rc;
    }
    int barrier() override { hub.wait(); return GB_OK; }
    int allgatherv_u64(u64 *const *buf, const u64 *off, const u64 *cnt) override
    {
        hub.ptr[(size_t)me] = buf[0];
        hub.wait();
        for (int p = 0; p < P; p++)
            if (p != me) memcpy(buf[0] + off[p], (const u64 *)hub.ptr[(size_t)p] + off[p], (size_t)cnt[p] * 8);
        hub.wait();
        return GB_OK;
    }
    int allreduce_sum(void *buf, size_t count, int elem_bytes) override
    {
        hub.ptr[(size_t)me] = buf;
        hub.wait();
        std::vector<u64> sum(count, 0);
        for (int p = 0; p < P; p++)
            for (size_t i = 0; i < count; i++)
                sum[i] += elem_bytes == 8 ? ((const u64 *)hub.ptr[(size_t)p])[i] : (u64)((const u32 *)hub.ptr[(size_t)p])[i];
        hub.wait();
        for (size_t i = 0; i < count; i++) {
            if (elem_bytes == 8) ((u64 *)buf)[i] = sum[i];
            else ((u32 *)buf)[i] = (u32)sum[i];
        }
        hub.wait();
        return GB_OK;
    }
};

// ---- one PROCESS per rank: the collectives are callbacks into the host program (tests/gloo_sgraph_worker.py implements them
// with torch.distributed over gloo and POSIX shared memory for the peer windows), so peers live in other address spaces and
// a rank sees their windows at addresses of its own -- the situation of CUDA IPC mappings + NCCL
extern "C" {
struct EmulFabricCallbacks {
    int (*allgather)(const void *mine, void *all, uint64_t bytes);
    int (*window)(uint64_t my_bytes, void **mine, void **peers /* [P] */);
    int (*alltoallv)(const uint64_t *send, const uint64_t *soff, const uint64_t *scnt, uint64_t *recv, const uint64_t *roff,
                     const uint64_t *rcnt);
    int (*barrier)(void);
    int (*allgatherv)(uint64_t *buf, const uint64_t *off, const uint64_t *cnt);
    int (*allreduce_sum)(void *buf, uint64_t count, int elem_bytes);
};
}
struct CallbackFabric : Fabric {
    const EmulFabricCallbacks &cb;
    CallbackFabric(const EmulFabricCallbacks &c, int n_ranks, int rank) : cb(c)
    {
        P = n_ranks;
        mine.push_back(rank);
    }
    int allgather_host(const void *const *contrib, void *const *all, size_t bytes) override { return cb.allgather(contrib[0], all[0], bytes); }
    int windows(const size_t *bytes_of_rank, void **window, PeerPtrs *peers) override
    {
        for (int r = 0; r < MAXR; r++) peers[0].p[r] = nullptr;
        return cb.window(bytes_of_rank[mine[0]], &window[0], peers[0].p);
    }
    int alltoallv_u64(const u64 *const *send, const Row *soff, const Row *scnt, u64 *const *recv, const Row *roff, const Row *rcnt) override
    {
        return cb.alltoallv((const uint64_t *)send[0], (const uint64_t *)soff[0].v, (const uint64_t *)scnt[0].v, (uint64_t *)recv[0],
                            (const uint64_t *)roff[0].v, (const uint64_t *)rcnt[0].v);
    }
    int barrier() override { return cb.barrier(); }
    int allgatherv_u64(u64 *const *buf, const u64 *off, const u64 *cnt) override
    {
        return cb.allgatherv((uint64_t *)buf[0], (const uint64_t *)off, (const uint64_t *)cnt);
    }
    int allreduce_sum(void *buf, size_t count, int elem_bytes) override { return cb.allreduce_sum(buf, count, elem_bytes); }
};

extern "C" {

// this process's rank of a build whose fabric is `cb`; two-call pattern like emul_sharded_build (sizes first)
int emul_sharded_build_rank(int k, int dual, int v210, int P, int rank, const uint64_t *keys, uint64_t n, const EmulFabricCallbacks *cb,
                            uint64_t *out, uint64_t *node_kmer, uint32_t *edge_start, uint32_t *edge_end, uint64_t *edge_off, uint32_t *bases)
{
    Exec ex;
    CallbackFabric fab(*cb, P, rank);
    std::vector<RankInput> in{ RankInput{ &ex, (const u64 *)keys, n } };
    Result res;
    const int rc = build(fab, in, k, dual != 0, v210 != 0, &res);
    if (rc != GB_OK) return rc;
    out[0] = res.n_nodes; out[1] = res.n_edges; out[2] = res.n_bases;
    out[3] = res.kept; out[4] = res.segments; out[5] = res.cycle_vertices; out[6] = (uint64_t)res.jump_rounds; out[7] = (uint64_t)res.seg_rounds;
    if (node_kmer) {
        memcpy(node_kmer, res.node_kmer, res.n_nodes * 8);
        memcpy(edge_start, res.edge_start, res.n_edges * 4);
        memcpy(edge_end, res.edge_end, res.n_edges * 4);
        memcpy(edge_off, res.edge_off, (res.n_edges + 1) * 8);
        memcpy(bases, res.bases, ((res.n_bases + 15) / 16) * 4);
    }
    return 0;
}

// the build with one thread per rank (ThreadFabric): every rank's copy of the result must be the same graph.  Returns 0, a
// GB_E_* code, or 1000 + r when rank r's copy differs from rank 0's.  Output = rank 0's copy (same layout as below).
int emul_sharded_build_threads(int k, int dual, int v210, int P, const uint64_t *keys, const uint64_t *off, uint64_t *out,
                               uint64_t *node_kmer, uint32_t *edge_start, uint32_t *edge_end, uint64_t *edge_off, uint32_t *bases)
{
    Hub hub(P);
    std::vector<Exec> ex((size_t)P);
    std::vector<Result> res((size_t)P);
    std::vector<int> rcs((size_t)P, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < P; r++)
        th.emplace_back([&, r] {
            ThreadFabric fab(hub, ex[(size_t)r], r);
            std::vector<RankInput> in{ RankInput{ &ex[(size_t)r], (const u64 *)keys + off[r], off[r + 1] - off[r] } };
            rcs[(size_t)r] = build(fab, in, k, dual != 0, v210 != 0, &res[(size_t)r]);
        });
    for (auto &t : th) t.join();
    for (int r = 0; r < P; r++)
        if (rcs[(size_t)r] != GB_OK) return rcs[(size_t)r];
    const Result &a = res[0];
    for (int r = 1; r < P; r++) {
        const Result &b = res[(size_t)r];
        if (a.n_nodes != b.n_nodes || a.n_edges != b.n_edges || a.n_bases != b.n_bases || a.segments != b.segments ||
            a.cycle_vertices != b.cycle_vertices || memcmp(a.node_kmer, b.node_kmer, a.n_nodes * 8) ||
            memcmp(a.edge_start, b.edge_start, a.n_edges * 4) || memcmp(a.edge_end, b.edge_end, a.n_edges * 4) ||
            memcmp(a.edge_off, b.edge_off, (a.n_edges + 1) * 8) || memcmp(a.bases, b.bases, ((a.n_bases + 15) / 16) * 4))
            return 1000 + r;
    }
    out[0] = a.n_nodes; out[1] = a.n_edges; out[2] = a.n_bases;
    out[3] = a.kept; out[4] = a.segments; out[5] = a.cycle_vertices; out[6] = (uint64_t)a.jump_rounds; out[7] = (uint64_t)a.seg_rounds;
    if (node_kmer) {
        memcpy(node_kmer, a.node_kmer, a.n_nodes * 8);
        memcpy(edge_start, a.edge_start, a.n_edges * 4);
        memcpy(edge_end, a.edge_end, a.n_edges * 4);
        memcpy(edge_off, a.edge_off, (a.n_edges + 1) * 8);
        memcpy(bases, a.bases, ((a.n_bases + 15) / 16) * 4);
    }
    return 0;
}

// the sharded build over P in-process ranks; rank r starts with keys[off[r] .. off[r + 1]).  Two-call pattern: with
// node_kmer == NULL only the sizes (out[0..2] = nodes, edges, bases) and stats (out[3..7] = kept, segments, cycle vertices,
// jump rounds, segment rounds) are returned.
int emul_sharded_build(int k, int dual, int v210, int P, const uint64_t *keys, const uint64_t *off, uint64_t *out,
                       uint64_t *node_kmer, uint32_t *edge_start, uint32_t *edge_end, uint64_t *edge_off, uint32_t *bases)
{
    std::vector<Exec> ex((size_t)P);
    std::vector<Exec *> pex;
    std::vector<RankInput> in;
    for (int r = 0; r < P; r++) {
        pex.push_back(&ex[(size_t)r]);
        in.push_back(RankInput{ &ex[(size_t)r], (const u64 *)keys + off[r], off[r + 1] - off[r] });
    }
    LocalFabric fab(P, pex);
    Result res;
    const int rc = build(fab, in, k, dual != 0, v210 != 0, &res);
    if (rc != GB_OK) return rc;
    out[0] = res.n_nodes; out[1] = res.n_edges; out[2] = res.n_bases;
    out[3] = res.kept; out[4] = res.segments; out[5] = res.cycle_vertices; out[6] = (uint64_t)res.jump_rounds; out[7] = (uint64_t)res.seg_rounds;
    if (node_kmer) {
        memcpy(node_kmer, res.node_kmer, res.n_nodes * 8);
        memcpy(edge_start, res.edge_start, res.n_edges * 4);
        memcpy(edge_end, res.edge_end, res.n_edges * 4);
        memcpy(edge_off, res.edge_off, (res.n_edges + 1) * 8);
        memcpy(bases, res.bases, ((res.n_bases + 15) / 16) * 4);
    }
    return 0;
}

// the super-k-mer splitting (superkmer.cuh) of a fixed-stride `.bin` stream: two passes like the device code would run them.
// per_owner[P] receives the record counts; with `records` != NULL owner o's records are written, 16 bytes each, one owner
// after the other (o's block starts at record sum(per_owner[0..o))).  Returns the number of k-windows.
uint64_t emul_superkmers(const uint8_t *bin, uint32_t rec_bytes, uint64_t n_reads, int k, int P, uint64_t *per_owner, uint64_t *records)
{
    Exec ex;
    const int m = minimizer_len(k);
    uint64_t windows = 0;
    for (int o = 0; o < P; o++) per_owner[o] = 0;
    sg_launch(ex, n_reads, SkCountOp{ bin, nullptr, rec_bytes, k, m, P, (u64 *)per_owner, (u64 *)&windows });
    if (records) {
        std::vector<u64> cursor((size_t)P, 0);
        std::vector<u64 *> out((size_t)P);
        uint64_t at = 0;
        for (int o = 0; o < P; o++) { out[(size_t)o] = (u64 *)records + 2 * at; at += per_owner[o]; }
        sg_launch(ex, n_reads, SkEmitOp{ bin, nullptr, rec_bytes, k, m, P, cursor.data(), out.data() });
        for (int o = 0; o < P; o++)
            if (cursor[(size_t)o] != per_owner[o]) return ~0ull;
    }
    return windows;
}

// owner of a k-mer and of its 8 neighbours (4 successors, then 4 predecessors): incremental form vs full recomputation
void emul_owners(int k, int P, uint64_t x, uint32_t *full9, uint32_t *incr8)
{
    const int m = minimizer_len(k);
    const u64 rcx = revcomp(x, k);
    const MinParts mp = min_parts(x, rcx, k, m);
    full9[0] = owner_of_kmer(x, k, m, P);
    for (u32 b = 0; b < 4; b++) {
        full9[1 + b] = owner_of_kmer(kmer_append(x, k, b), k, m, P);
        full9[5 + b] = owner_of_kmer(kmer_prepend(x, k, b), k, m, P);
        const Neighbour s = neighbour_of(mp, x, rcx, k, m, P, true, b), p = neighbour_of(mp, x, rcx, k, m, P, false, b);
        // the shifted reverse complements must be the reverse complements: a wrong one is reported as an impossible owner
        incr8[b] = s.rq == revcomp(s.q, k) && s.q == kmer_append(x, k, b) ? s.owner : 0xFFFFFFFFu;
        incr8[4 + b] = p.rq == revcomp(p.q, k) && p.q == kmer_prepend(x, k, b) ? p.owner : 0xFFFFFFFFu;
    }
}

} // extern "C"
