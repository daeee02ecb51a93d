// walk.cuh -- paired-end path support on the compacted graph: the per-item logic of
// S/scripts/GraphSimplifier.scala:33-127 (WalkingActor), 192-206 (annotate), 213-248 (the pair loop) and 268-313 (the in x out
// matrix of a node and its split), paths relative to /root/reference.  Everything here is a __host__ __device__ function of
// plain arrays: the kernels of walk.cu call them one item per thread, and tests/emul/ compiles the same header with g++ to
// run the same code serially against the oracle where no GPU is available.
//
// The reference explores walks with a memoised DFS over states (prevEdge, dist1), pruned by a backward Dijkstra from the
// target (43-72, 91-113).  Here the same state space is a bitset per edge: F[e] bit d = "dfs(e.end, d, e) is called", G[e] bit
// d = "... and returns true".  F grows forwards from the start position (shift by the next edge's length), G grows backwards
// from the target, both to their fixed points; an edge pair (e, e') is in pathEdges iff F[e] & (G[e'] >> len(e')) != 0.
// The Dijkstra prune and the memo only save work in the reference (a pruned call returns false either way), so the sets
// are identical.  State lives in a per-case LOCAL edge table (the edges within range.last bases of the start position).
#pragma once
#include "common.cuh"

namespace gb {

#define GB_HDN __host__ __device__

constexpr int WALK_W = 8;          // 64-bit words per distance bitset: range.last <= 64 * WALK_W - 1
constexpr int WALK_MAX_RANGE = 64 * WALK_W - 1;
constexpr int WALK_MAXPOS = 8;     // positions per getAll (1 on a freshly built graph: every oriented k-mer occurs once)

struct GraphView {
    int k;
    unsigned long long n_nodes, n_edges;
    const unsigned long long *node_kmer;
    const unsigned int *edge_start, *edge_end;
    const unsigned long long *edge_off;
    const unsigned int *bases;
    const unsigned int *out4; // [4 * n_nodes] out-edge by first base, NONE32 = none
};

// Graph.getGraphMap (S/data/graph/Graph.scala:90-119) as a device multimap: the entries are the (k-mer, id, dist) arrays of
// gb_graph_positions, `slot` is an open-addressing index over them.  putNew (S/ds/ArrayDNAMap.scala:152-162) takes the first
// free slot after the key's home, getAll (103-113) walks the run up to the first free slot.
struct PosMap {
    const unsigned int *slot; // [cap] entry index, NONE32 = free
    unsigned long long cap;
    const unsigned long long *kmer;
    const unsigned int *id, *dist; // dist 0: NodeGraphPosition(id), dist >= 1: EdgeGraphPosition(id, dist)
};

struct Pos {
    unsigned int id, dist;
};

GB_HD unsigned int base_at(const unsigned int *bases, unsigned long long pos)
{
    return (bases[pos >> 4] >> (2 * (unsigned int)(pos & 15))) & 3u;
}
GB_HD unsigned long long edge_len(const GraphView &g, unsigned int e) { return g.edge_off[e + 1] - g.edge_off[e]; }

// p.take(k) of the record at byte `off` (1 length byte, then 4 bases per byte, first base in the low bits)
GB_HD unsigned long long record_first_kmer(const uint8_t *bin, unsigned long long off, int k)
{
    unsigned long long x = 0;
    for (int i = 0; i < (k + 3) / 4; i++) x |= (unsigned long long)bin[off + 1 + i] << (8 * i);
    return x & ((1ull << (2 * k)) - 1);
}

GB_HD int posmap_get_all(const PosMap &m, unsigned long long key, Pos *out, int cap)
{
    int n = 0;
    unsigned long long i = slot_of(mix64(key), m.cap);
    for (;;) {
        const unsigned int e = m.slot[i];
        if (e == NONE32) return n;
        if (m.kmer[e] == key) {
            if (n < cap) { out[n].id = m.id[e]; out[n].dist = m.dist[e]; }
            n++;
        }
        i = next_slot(i, m.cap);
    }
}

// getAll (ArrayDNAMap.scala:103-113) into plain arrays: returns the number of positions under `key` (contains = count > 0,
// 232); the first max_per of them go to ids / dists (either may be null)
GB_HD unsigned int posmap_lookup(const PosMap &m, unsigned long long key, int max_per, unsigned int *ids, unsigned int *dists)
{
    unsigned int n = 0;
    unsigned long long i = slot_of(mix64(key), m.cap);
    for (;;) {
        const unsigned int e = m.slot[i];
        if (e == NONE32) return n;
        if (m.kmer[e] == key) {
            if ((int)n < max_per) {
                if (ids) ids[n] = m.id[e];
                if (dists) dists[n] = m.dist[e];
            }
            n++;
        }
        i = next_slot(i, m.cap);
    }
}

// annotate (GraphSimplifier.scala:192-206)
GB_HD bool annotate_drops(const Pos *p1, int n1, const Pos *p2, int n2, int k, int lo, int hi)
{
    for (int i = 0; i < n1; i++) {
        if (p1[i].dist == 0) continue;
        for (int j = 0; j < n2; j++) {
            if (p2[j].dist == 0 || p2[j].id != p1[i].id) continue;
            const long long d = (long long)p2[j].dist - (long long)p1[i].dist + k;
            if (lo <= d && d <= hi) return true;
        }
    }
    return false;
}

// ---------------------------------------------------------------- distance bitsets (bit d <-> dist1 == d)
GB_HD void bs_zero(unsigned long long *a) { for (int w = 0; w < WALK_W; w++) a[w] = 0; }
GB_HD bool bs_any(const unsigned long long *a)
{
    unsigned long long o = 0;
    for (int w = 0; w < WALK_W; w++) o |= a[w];
    return o != 0;
}
GB_HD bool bs_test(const unsigned long long *a, long long d)
{
    return d >= 0 && d < 64 * WALK_W && ((a[d >> 6] >> (d & 63)) & 1);
}
// bits lo..hi inclusive, clipped to the bitset
GB_HD void bs_range(unsigned long long *a, long long lo, long long hi)
{
    if (lo < 0) lo = 0;
    if (hi > 64 * WALK_W - 1) hi = 64 * WALK_W - 1;
    for (int w = 0; w < WALK_W; w++) {
        const long long b0 = 64ll * w, b1 = b0 + 63;
        unsigned long long v = 0;
        if (lo <= b1 && hi >= b0) {
            const long long l = lo > b0 ? lo - b0 : 0, h = hi < b1 ? hi - b0 : 63;
            v = (h == 63 ? ~0ull : ((1ull << (h + 1)) - 1)) & ~((1ull << l) - 1);
        }
        a[w] = v;
    }
}
// dst = src << s (towards larger distances); bits beyond the bitset are dropped
GB_HD void bs_shl(unsigned long long *dst, const unsigned long long *src, unsigned long long s)
{
    if (s >= 64ull * WALK_W) { bs_zero(dst); return; }
    const int ws = (int)(s >> 6), bsft = (int)(s & 63);
    for (int w = WALK_W - 1; w >= 0; w--) {
        unsigned long long v = 0;
        if (w - ws >= 0) {
            v = src[w - ws] << bsft;
            if (bsft && w - ws - 1 >= 0) v |= src[w - ws - 1] >> (64 - bsft);
        }
        dst[w] = v;
    }
}
// dst = src >> s
GB_HD void bs_shr(unsigned long long *dst, const unsigned long long *src, unsigned long long s)
{
    if (s >= 64ull * WALK_W) { bs_zero(dst); return; }
    const int ws = (int)(s >> 6), bsft = (int)(s & 63);
    for (int w = 0; w < WALK_W; w++) {
        unsigned long long v = 0;
        if (w + ws < WALK_W) {
            v = src[w + ws] >> bsft;
            if (bsft && w + ws + 1 < WALK_W) v |= src[w + ws + 1] << (64 - bsft);
        }
        dst[w] = v;
    }
}

// ---------------------------------------------------------------- the local edge table of one orientation case
constexpr unsigned int WF_EMIT = 15u, WF_DIRTY = 16u;
constexpr int NXT_UNKNOWN = -1, NXT_NONE = -2;

struct WalkEntry {
    unsigned int eid;   // graph edge, NONE32 for entry 0 (the null prevEdge of a walk that starts on a node)
    unsigned int endn;  // node the state stands on (the edge's end)
    unsigned int len;   // edge length in bases (only lengths <= range.last ever enter the table, bar a start edge)
    unsigned int flags; // bits 0..3: (this edge, out-edge with first base b) is in the case's pathEdges; WF_DIRTY
    int nxt[4];         // local index of the out-edge with first base b of endn; NXT_UNKNOWN / NXT_NONE
    unsigned long long F[WALK_W], G[WALK_W];
};

struct WalkTable {
    WalkEntry *e;
    int n, cap;
};

GB_HD void table_reset(WalkTable &t)
{
    t.n = 1;
    WalkEntry &r = t.e[0];
    r.eid = NONE32; r.endn = 0; r.len = 0; r.flags = 0;
    for (int b = 0; b < 4; b++) r.nxt[b] = NXT_UNKNOWN;
    bs_zero(r.F); bs_zero(r.G);
}

// local index of graph edge `eid`, adding it; -1 when the table is full
GB_HD int table_find_or_add(WalkTable &t, const GraphView &g, unsigned int eid)
{
    for (int i = 1; i < t.n; i++)
        if (t.e[i].eid == eid) return i;
    if (t.n == t.cap) return -1;
    WalkEntry &x = t.e[t.n];
    x.eid = eid;
    x.endn = g.edge_end[eid];
    const unsigned long long len = edge_len(g, eid);
    x.len = len > 0xFFFFFFFFull ? 0xFFFFFFFFu : (unsigned int)len;
    x.flags = 0;
    for (int b = 0; b < 4; b++) x.nxt[b] = NXT_UNKNOWN;
    bs_zero(x.F); bs_zero(x.G);
    return t.n++;
}

constexpr int WALK_OVERFLOW = -1;

// WalkingActor.receive (GraphSimplifier.scala:77-126) for one (pos1, pos2): returns 1 = good, 0 = not good, WALK_OVERFLOW when
// the local table is too small (nothing has been committed: the caller retries the whole case with a larger table).
// The pathEdges of the walk are OR-ed into the entries' emit bits.
GB_HDN inline int walk_one(const GraphView &g, WalkTable &t, Pos pos1, Pos pos2, int lo, int hi)
{
    // 80-90, 114-118: target (node2, dist2, endEdge), start (node0, dist0, startEdge)
    unsigned int node2, end_edge = NONE32, node0, start_edge = NONE32;
    long long dist2 = 0, dist0 = 0;
    if (pos2.dist == 0) node2 = pos2.id;
    else { node2 = g.edge_start[pos2.id]; dist2 = pos2.dist; end_edge = pos2.id; }
    if (pos1.dist == 0) node0 = pos1.id;
    else { node0 = g.edge_end[pos1.id]; dist0 = (long long)edge_len(g, pos1.id) - (long long)pos1.dist; start_edge = pos1.id; }

    for (int i = 0; i < t.n; i++) { // a fresh walk on the case's table: states go, the emitted pairs and the topology stay
        bs_zero(t.e[i].F); bs_zero(t.e[i].G);
        t.e[i].flags &= WF_EMIT;
    }
    const long long dmax = (long long)hi - dist2; // a state with dist1 + dist2 > range.last returns false (94)
    if (dmax < 0 || dist0 > dmax) return 0;

    int root = 0;
    if (start_edge == NONE32) t.e[0].endn = node0;
    else {
        root = table_find_or_add(t, g, start_edge);
        if (root < 0) return WALK_OVERFLOW;
    }
    t.e[root].F[dist0 >> 6] |= 1ull << (dist0 & 63);
    t.e[root].flags |= WF_DIRTY;

    unsigned long long maskd[WALK_W], tgt[WALK_W], s[WALK_W];
    bs_range(maskd, 0, dmax);
    bs_range(tgt, (long long)lo - dist2, dmax); // node1 == node2 && range.contains(dist1 + dist2) (97)

    // forwards: F[e'] |= F[e] << len(e') for every out-edge e' of e's end node (104-105), to the fixed point
    for (bool any = true; any;) {
        any = false;
        for (int i = 0; i < t.n; i++) {
            if (!(t.e[i].flags & WF_DIRTY)) continue;
            t.e[i].flags &= ~WF_DIRTY;
            const unsigned int v = t.e[i].endn;
            for (int b = 0; b < 4; b++) {
                int j = t.e[i].nxt[b];
                if (j == NXT_NONE) continue;
                unsigned int e2 = NONE32;
                unsigned long long len2;
                if (j >= 0) len2 = t.e[j].len;
                else {
                    e2 = g.out4[4ull * v + b];
                    if (e2 == NONE32) { t.e[i].nxt[b] = NXT_NONE; continue; }
                    len2 = edge_len(g, e2);
                }
                if (len2 > (unsigned long long)dmax) continue;
                bs_shl(s, t.e[i].F, len2);
                bool fresh = false, nonempty = false;
                for (int w = 0; w < WALK_W; w++) { s[w] &= maskd[w]; nonempty |= s[w] != 0; }
                if (!nonempty) continue;
                if (j < 0) {
                    j = table_find_or_add(t, g, e2);
                    if (j < 0) return WALK_OVERFLOW;
                    t.e[i].nxt[b] = j;
                }
                for (int w = 0; w < WALK_W; w++) {
                    fresh |= (s[w] & ~t.e[j].F[w]) != 0;
                    t.e[j].F[w] |= s[w];
                }
                if (fresh) { t.e[j].flags |= WF_DIRTY; any = true; }
            }
        }
    }

    // backwards: G[e] = F[e] & (target(e) | OR over out-edges e' of (G[e'] >> len(e'))) (97-107), to the fixed point
    for (int i = 0; i < t.n; i++)
        if (t.e[i].endn == node2)
            for (int w = 0; w < WALK_W; w++) t.e[i].G[w] = t.e[i].F[w] & tgt[w];
    for (bool any = true; any;) {
        any = false;
        for (int i = 0; i < t.n; i++) {
            if (!bs_any(t.e[i].F)) continue;
            for (int b = 0; b < 4; b++) {
                const int j = t.e[i].nxt[b];
                if (j < 0 || !bs_any(t.e[j].G)) continue;
                bs_shr(s, t.e[j].G, t.e[j].len);
                for (int w = 0; w < WALK_W; w++) {
                    const unsigned long long add = t.e[i].F[w] & s[w] & ~t.e[i].G[w];
                    if (add) { t.e[i].G[w] |= add; any = true; }
                }
            }
        }
    }

    // pathEdges: (prevEdge, endEdge) where the walk may stop (98-100), (prevEdge, e) where it goes on successfully (105-107)
    for (int i = 1; i < t.n; i++) {
        if (!bs_any(t.e[i].F)) continue;
        for (int b = 0; b < 4; b++) {
            const int j = t.e[i].nxt[b];
            if (j < 0) continue;
            bs_shr(s, t.e[j].G, t.e[j].len);
            bool hit = false;
            for (int w = 0; w < WALK_W; w++) hit |= (t.e[i].F[w] & s[w]) != 0;
            if (hit) t.e[i].flags |= 1u << b;
        }
        if (end_edge != NONE32 && t.e[i].endn == node2) {
            bool hit = false;
            for (int w = 0; w < WALK_W; w++) hit |= (t.e[i].F[w] & tgt[w]) != 0;
            if (hit) t.e[i].flags |= 1u << base_at(g.bases, g.edge_off[end_edge]);
        }
    }
    return bs_test(t.e[root].G, dist0) ? 1 : 0;
}

GB_HD void walk_add_u32(unsigned int *p, unsigned int v)
{
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
GB_HD void walk_add_u64(unsigned long long *p, unsigned long long v)
{
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}

constexpr int CASE_SKIPPED = 0, CASE_WALKED = 1, CASE_OVERFLOW = -1, CASE_TOO_MANY_POSITIONS = -2;

// does the orientation case (x, y) reach the walkers?  lists f = getAll(x), getAll(rc(y)) (214-217); dropped by annotate (219)
// or with an empty list (no futures, 238)
GB_HDN inline int case_positions(const GraphView &g, const PosMap &m, unsigned long long x, unsigned long long y, int lo, int hi,
                                 Pos *p1, int *n1, Pos *p2, int *n2)
{
    *n1 = posmap_get_all(m, x, p1, WALK_MAXPOS);
    *n2 = posmap_get_all(m, revcomp(y, g.k), p2, WALK_MAXPOS);
    if (*n1 > WALK_MAXPOS || *n2 > WALK_MAXPOS) return CASE_TOO_MANY_POSITIONS;
    if (annotate_drops(p1, *n1, p2, *n2, g.k, lo, hi)) return CASE_SKIPPED;
    if (*n1 == 0 || *n2 == 0) return CASE_SKIPPED;
    return CASE_WALKED;
}

// one orientation case of one read pair, all of 219-248: every (pos1, pos2) is walked, the union of their pathEdges adds 1 to
// support[4 * e1 + b] (= pathsMap((e1, e2)), e2 = the out-edge with first base b of e1's end node), a case without a good walk
// adds 1 to *bad_pairs.  Nothing is committed on CASE_OVERFLOW.
GB_HDN inline int process_case(const GraphView &g, const PosMap &m, unsigned long long x, unsigned long long y, int lo, int hi,
                               WalkTable &t, unsigned int *support, unsigned long long *bad_pairs)
{
    Pos p1[WALK_MAXPOS], p2[WALK_MAXPOS];
    int n1, n2;
    const int r = case_positions(g, m, x, y, lo, hi, p1, &n1, p2, &n2);
    if (r != CASE_WALKED) return r;
    table_reset(t);
    bool good = false;
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n2; j++) {
            const int w = walk_one(g, t, p1[i], p2[j], lo, hi);
            if (w == WALK_OVERFLOW) return CASE_OVERFLOW;
            good |= w == 1;
        }
    for (int i = 1; i < t.n; i++)
        for (unsigned int b = 0; b < 4; b++)
            if (t.e[i].flags & (1u << b)) walk_add_u32(&support[4ull * t.e[i].eid + b], 1u);
    if (!good) walk_add_u64(bad_pairs, 1ull);
    return CASE_WALKED;
}

// ---------------------------------------------------------------- node splitting (GraphSimplifier.scala:268-313)
// in-slot of an edge at its end node = the base that precedes the node's k-mer on the edge: the spelled string is
// start k-mer ++ seq, the end k-mer is its last k characters, so that base is character len - 1.  Two in-edges of one node
// never share it (each oriented k-mer lies on one edge), which also bounds the in-degree by 4.
GB_HD unsigned int in_slot_base(const GraphView &g, unsigned int e)
{
    const unsigned long long len = edge_len(g, e);
    if (len - 1 < (unsigned long long)g.k) return (unsigned int)(g.node_kmer[g.edge_start[e]] >> (2 * (len - 1))) & 3u;
    return base_at(g.bases, g.edge_off[e] + (len - 1 - g.k));
}

struct SplitPlan {
    int n_new;          // components with both in- and out-edges: each becomes a copy of the node (303-307)
    int in_comp[4];     // per in-slot: index of its new node (0..n_new-1), -1 = edge removed (301-302), -2 = no edge
    int out_comp[4];    // per out-slot: index of its new node, -1 = edge removed (309), -2 = no edge
};

// the matrix of 271-273 thresholded at `cutoff`, and the bipartite components of dfsLeft / dfsRight (279-300).  New nodes are
// numbered by the smallest in-slot of their component.  Nodes without in- or without out-edges are left alone (268-269).
GB_HDN inline SplitPlan split_node(const unsigned int in4[4], const unsigned int out4[4], const unsigned int *support, int cutoff)
{
    SplitPlan p;
    p.n_new = 0;
    int nin = 0, nout = 0;
    for (int s = 0; s < 4; s++) {
        p.in_comp[s] = in4[s] == NONE32 ? -2 : -1;
        p.out_comp[s] = out4[s] == NONE32 ? -2 : -1;
        nin += in4[s] != NONE32;
        nout += out4[s] != NONE32;
    }
    if (nin == 0 || nout == 0) { // untouched: every present edge keeps its node
        for (int s = 0; s < 4; s++) { if (p.in_comp[s] == -1) p.in_comp[s] = -3; if (p.out_comp[s] == -1) p.out_comp[s] = -3; }
        return p;
    }
    bool adj[4][4];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            adj[i][j] = in4[i] != NONE32 && out4[j] != NONE32 && (long long)support[4ull * in4[i] + j] >= (long long)cutoff;
    bool seen_in[4] = { false, false, false, false };
    for (int i0 = 0; i0 < 4; i0++) {
        if (in4[i0] == NONE32 || seen_in[i0]) continue;
        unsigned int l = 1u << i0, r = 0;
        for (bool grow = true; grow;) { // closure of {i0} under adj
            grow = false;
            for (int i = 0; i < 4; i++)
                if (l >> i & 1)
                    for (int j = 0; j < 4; j++)
                        if (adj[i][j] && !(r >> j & 1)) { r |= 1u << j; grow = true; }
            for (int j = 0; j < 4; j++)
                if (r >> j & 1)
                    for (int i = 0; i < 4; i++)
                        if (adj[i][j] && !(l >> i & 1)) { l |= 1u << i; grow = true; }
        }
        for (int i = 0; i < 4; i++) if (l >> i & 1) seen_in[i] = true;
        if (r == 0) continue; // in-edge alone: removed (in_comp stays -1)
        for (int i = 0; i < 4; i++) if (l >> i & 1) p.in_comp[i] = p.n_new;
        for (int j = 0; j < 4; j++) if (r >> j & 1) p.out_comp[j] = p.n_new;
        p.n_new++;
    }
    return p;
}

} // namespace gb
