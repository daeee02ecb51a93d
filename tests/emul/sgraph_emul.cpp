// sgraph_emul.cpp -- TEST INFRASTRUCTURE: compiles genome_b200/csrc/sgraph.cuh (the sharded Graph.buildGraph: per-item ops
// AND the orchestration) with g++, with serial loops for the kernels, malloc for device memory and the in-process fabric for
// the collectives, so that the whole algorithm is checked against the oracle on a box without a GPU.  Not part of the
// product: nothing in genome_b200/ builds or loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../genome_b200/csrc/sgraph.cuh"

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
    *p = malloc(bytes + 64);
    if (!*p) return GB_E_OOM;
    memset(*p, 0xA5, bytes + 64);
    ex.owned.push_back(*p);
    return GB_OK;
}
int sg_graph_alloc(Exec &ex, void **p, size_t bytes)
{
    *p = malloc(bytes + 64);
    if (!*p) return GB_E_OOM;
    memset(*p, 0xA5, bytes + 64);
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

extern "C" {

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
        incr8[b] = neighbour_owner(mp, x, rcx, k, m, P, true, b);
        incr8[4 + b] = neighbour_owner(mp, x, rcx, k, m, P, false, b);
    }
}

} // extern "C"
