// walk_emul.cpp -- TEST INFRASTRUCTURE: compiles genome_b200/csrc/walk.cuh with g++ and runs the per-item functions the
// CUDA kernels of walk.cu call (one thread per item there, a serial loop here), so that the walk / split logic can be
// checked against the oracle on a box without a GPU.  Not part of the product: nothing in genome_b200/ builds or loads it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../genome_b200/csrc/walk.cuh"

using namespace gb;

namespace gb {
void set_error(const char *, ...) {}
thread_local Arena *tl_arena = nullptr;
} // namespace gb

static GraphView make_view(int k, uint64_t n_nodes, uint64_t n_edges, const uint64_t *node_kmer, const uint32_t *edge_start,
                           const uint32_t *edge_end, const uint64_t *edge_off, const uint32_t *bases, const uint32_t *out4)
{
    GraphView g;
    g.k = k; g.n_nodes = n_nodes; g.n_edges = n_edges;
    g.node_kmer = (const unsigned long long *)node_kmer;
    g.edge_start = edge_start; g.edge_end = edge_end;
    g.edge_off = (const unsigned long long *)edge_off;
    g.bases = bases; g.out4 = out4;
    return g;
}

static std::vector<uint32_t> out_table(const GraphView &g)
{
    std::vector<uint32_t> out4(4 * g.n_nodes + 4, NONE32);
    for (uint64_t e = 0; e < g.n_edges; e++) out4[4 * g.edge_start[e] + base_at(g.bases, g.edge_off[e])] = (uint32_t)e; // walk_out_table_kernel
    return out4;
}

extern "C" {

// gb_graph_pair_support's device work, serially: returns 0, or -2 when a k-mer has too many positions; *overflowed = cases that
// did not fit a table of lmax entries (the library would retry them in the next tier; here they are reported)
int emul_pair_support(int k, uint64_t n_nodes, uint64_t n_edges, const uint64_t *node_kmer, const uint32_t *edge_start,
                      const uint32_t *edge_end, const uint64_t *edge_off, const uint32_t *bases, uint64_t n_pos,
                      const uint64_t *pos_kmer, const uint32_t *pos_id, const uint32_t *pos_dist, const uint8_t *bin,
                      const uint64_t *rec_off, uint64_t n_pairs, int lo, int hi, int lmax, uint32_t *support, uint64_t *bad,
                      uint64_t *walked, uint64_t *overflowed)
{
    GraphView g0 = make_view(k, n_nodes, n_edges, node_kmer, edge_start, edge_end, edge_off, bases, nullptr);
    std::vector<uint32_t> out4 = out_table(g0);
    GraphView g = make_view(k, n_nodes, n_edges, node_kmer, edge_start, edge_end, edge_off, bases, out4.data());

    const unsigned long long cap = (n_pos * 2 + 1024) / 1024 * 1024;
    std::vector<uint32_t> slot(cap, NONE32);
    for (uint64_t i = 0; i < n_pos; i++) { // posmap_insert_kernel
        unsigned long long s = slot_of(mix64(pos_kmer[i]), cap);
        while (slot[s] != NONE32) s = next_slot(s, cap);
        slot[s] = (uint32_t)i;
    }
    PosMap m;
    m.slot = slot.data(); m.cap = cap; m.kmer = (const unsigned long long *)pos_kmer; m.id = pos_id; m.dist = pos_dist;

    std::vector<unsigned long long> cases;
    for (uint64_t p = 0; p < n_pairs; p++) { // walk_filter_kernel
        const unsigned long long o1 = rec_off[2 * p], o2 = rec_off[2 * p + 1];
        if ((int)bin[o1] < k || (int)bin[o2] < k) continue;
        const unsigned long long a = record_first_kmer(bin, o1, k), b = record_first_kmer(bin, o2, k);
        for (int c = 0; c < 2; c++) {
            const unsigned long long x = c ? b : a, y = c ? a : b;
            Pos p1[WALK_MAXPOS], p2[WALK_MAXPOS];
            int n1, n2;
            const int r = case_positions(g, m, x, y, lo, hi, p1, &n1, p2, &n2);
            if (r == CASE_TOO_MANY_POSITIONS) return -2;
            if (r != CASE_WALKED) continue;
            cases.push_back(x);
            cases.push_back(y);
        }
    }
    std::vector<WalkEntry> scratch((size_t)lmax);
    WalkTable t;
    t.e = scratch.data(); t.cap = lmax; t.n = 0;
    unsigned long long nbad = 0;
    *walked = 0; *overflowed = 0;
    memset(support, 0, (size_t)n_edges * 16);
    for (size_t c = 0; c < cases.size() / 2; c++) { // walk_cases_kernel
        const int r = process_case(g, m, cases[2 * c], cases[2 * c + 1], lo, hi, t, support, &nbad);
        if (r == CASE_WALKED) (*walked)++;
        else if (r == CASE_OVERFLOW) (*overflowed)++;
    }
    *bad = nbad;
    return 0;
}

// one walk (walk_one) on a fresh table: returns good (0/1) or -1 on overflow; emit[4 * n_edges] receives the pathEdges flags
int emul_walk(int k, uint64_t n_nodes, uint64_t n_edges, const uint64_t *node_kmer, const uint32_t *edge_start,
              const uint32_t *edge_end, const uint64_t *edge_off, const uint32_t *bases, uint32_t id1, uint32_t dist1,
              uint32_t id2, uint32_t dist2, int lo, int hi, int lmax, uint8_t *emit)
{
    GraphView g0 = make_view(k, n_nodes, n_edges, node_kmer, edge_start, edge_end, edge_off, bases, nullptr);
    std::vector<uint32_t> out4 = out_table(g0);
    GraphView g = make_view(k, n_nodes, n_edges, node_kmer, edge_start, edge_end, edge_off, bases, out4.data());
    std::vector<WalkEntry> scratch((size_t)lmax);
    WalkTable t;
    t.e = scratch.data(); t.cap = lmax;
    table_reset(t);
    Pos p1 = { id1, dist1 }, p2 = { id2, dist2 };
    const int r = walk_one(g, t, p1, p2, lo, hi);
    memset(emit, 0, (size_t)n_edges * 4);
    if (r < 0) return r;
    for (int i = 1; i < t.n; i++)
        for (int b = 0; b < 4; b++)
            if (t.e[i].flags & (1u << b)) emit[4ull * t.e[i].eid + b] = 1;
    return r;
}

// gb_graph_split_nodes' device work, serially.  edge_start / edge_end are rewired in place, node_kmer2 must hold n_nodes +
// (upper bound 4 * n_nodes) entries; returns the number of nodes added, -1 on an in-slot clash; kill[n_edges] flags removals
int64_t emul_split(int k, uint64_t n_nodes, uint64_t n_edges, const uint64_t *node_kmer, uint32_t *edge_start, uint32_t *edge_end,
                   const uint64_t *edge_off, const uint32_t *bases, const uint32_t *support, int cutoff, uint64_t *node_kmer2,
                   uint32_t *kill)
{
    GraphView g0 = make_view(k, n_nodes, n_edges, node_kmer, edge_start, edge_end, edge_off, bases, nullptr);
    std::vector<uint32_t> out4 = out_table(g0);
    GraphView g = make_view(k, n_nodes, n_edges, node_kmer, edge_start, edge_end, edge_off, bases, out4.data());
    std::vector<uint32_t> in4(4 * n_nodes + 4, NONE32);
    for (uint64_t e = 0; e < n_edges; e++) { // walk_in_table_kernel
        uint32_t &s = in4[4 * g.edge_end[e] + in_slot_base(g, (unsigned int)e)];
        if (s != NONE32) return -1;
        s = (uint32_t)e;
    }
    std::vector<unsigned long long> base(n_nodes + 1, 0);
    for (uint64_t v = 0; v < n_nodes; v++) // split_count_kernel + exclusive scan
        base[v + 1] = base[v] + (unsigned long long)split_node(&in4[4 * v], &out4[4 * v], support, cutoff).n_new;
    memcpy(node_kmer2, node_kmer, n_nodes * 8);
    memset(kill, 0, n_edges * 4);
    for (uint64_t v = 0; v < n_nodes; v++) { // split_apply_kernel
        const SplitPlan p = split_node(&in4[4 * v], &out4[4 * v], support, cutoff);
        const unsigned long long first = n_nodes + base[v];
        for (int c = 0; c < p.n_new; c++) node_kmer2[first + c] = node_kmer[v];
        for (int s = 0; s < 4; s++) {
            if (p.in_comp[s] >= 0) edge_end[in4[4 * v + s]] = (uint32_t)(first + p.in_comp[s]);
            else if (p.in_comp[s] == -1) kill[in4[4 * v + s]] = 1;
            if (p.out_comp[s] >= 0) edge_start[out4[4 * v + s]] = (uint32_t)(first + p.out_comp[s]);
            else if (p.out_comp[s] == -1) kill[out4[4 * v + s]] = 1;
        }
    }
    return (int64_t)base[n_nodes];
}

// gb_graph_map_create + gb_graph_map_get_all, serially: putNew of every entry (posmap_insert_kernel), then posmap_lookup per key
int emul_graph_map_get_all(uint64_t n_pos, const uint64_t *pos_kmer, const uint32_t *pos_id, const uint32_t *pos_dist, uint64_t n_keys,
                           const uint64_t *keys, int max_per, uint32_t *ids, uint32_t *dists, uint32_t *counts)
{
    const unsigned long long cap = (n_pos * 2 + 1024) / 1024 * 1024;
    std::vector<uint32_t> slot(cap, NONE32);
    for (uint64_t i = 0; i < n_pos; i++) {
        unsigned long long s = slot_of(mix64(pos_kmer[i]), cap);
        while (slot[s] != NONE32) s = next_slot(s, cap);
        slot[s] = (uint32_t)i;
    }
    PosMap m;
    m.slot = slot.data(); m.cap = cap; m.kmer = (const unsigned long long *)pos_kmer; m.id = pos_id; m.dist = pos_dist;
    for (uint64_t i = 0; i < n_keys; i++)
        counts[i] = posmap_lookup(m, keys[i], max_per, ids ? ids + i * max_per : nullptr, dists ? dists + i * max_per : nullptr);
    return 0;
}

} // extern "C"
