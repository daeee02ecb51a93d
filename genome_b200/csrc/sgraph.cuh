// sgraph.cuh -- Graph.buildGraph (S/data/graph/Graph.scala:269-382, relative to /root/reference) over P shards WITHOUT a
// replica of the filtered table: SURVEY 8(e) "graph build: near-linear".  The replicated build of comm.cu gathers every kept
// k-mer on every rank, so its time grows with the total; here every rank works on its own keys and only chain boundaries,
// the compacted graph (2 bits per base) and a segment list (one entry per chain piece) cross the NVLink fabric.
//
//   1. re-route    the kept k-mers go to owner = hash of their MINIMIZER (smallest hashed canonical m-mer): a k-mer and its
//                  reverse complement share the owner, and so do ~90 % of the (k-1)-overlap neighbours -- the chains of the
//                  de Bruijn graph become long rank-local runs (the count table keeps its hash-prefix shards; ownership is
//                  unobservable through DNAMap).
//   2. index       open-addressing index (u32 entry per slot) over the received keys; entry index = local vertex id.
//   3. masks       incoming / outcoming (Graph.scala:272-282): 8 membership probes per key, in the owner's index -- through
//                  the peer mapping (NVLink load) when the neighbour's minimizer differs.
//   4. classify    terminal rule (323), node / edge numbering (local scans + rank offsets), SEGMENT heads = interior
//                  vertices whose predecessor lives on another rank.
//   5. rank        list ranking by pointer jumping over LOCAL predecessor links only: every interior vertex learns
//                  (edge, rank) if its chain's head is local, else (segment, rank inside the segment).
//   6. segments    one entry per segment: the resolved entry of its remote predecessor, read through the peer mapping; the
//                  entries are all-gathered and ranked (a list ~10x shorter than the vertex list), then folded back.
//   7. close       edge ends / lengths, offsets, 2-bit bases: written into the (small) global graph arrays of each rank and
//                  combined with one sum per array (every element has exactly one writer).
//
// The code below is backend-agnostic: every "kernel" is a functor run once per item by sg_launch, memory and collectives
// go through the Exec / Fabric interfaces.  sgraph.cu provides the CUDA Exec (one thread per item on the rank's stream) and
// comm.cu the NCCL + CUDA-IPC Fabric; tests/emul/sgraph_emul.cpp provides serial loops and an in-process Fabric, so the
// whole algorithm is checked against the oracle on a box without a GPU.  The same in-process Fabric also runs P VIRTUAL
// shards on one device (gb_graph_build_virtual_shards), which is how the device code is validated on a single GPU.
#pragma once
#include <vector>

#include "common.cuh"
#include "sgraph_fabric.cuh"

namespace gb {
namespace sg {

typedef uint8_t u8;

// ---------------------------------------------------------------- backend interface
struct Exec; // one per rank driven by this process: CUDA stream + arena, or nothing at all (emulation)
int sg_alloc(Exec &ex, void **p, size_t bytes); // scratch that lives until the build returns
int sg_graph_alloc(Exec &ex, void **p, size_t bytes); // memory of the resulting graph
int sg_zero(Exec &ex, void *p, size_t bytes);
int sg_fill_ff(Exec &ex, void *p, size_t bytes);
int sg_copy(Exec &ex, void *dst, const void *src, size_t bytes);   // device -> device
int sg_read(Exec &ex, void *host, const void *dev, size_t bytes);  // synchronises
int sg_write_host(Exec &ex, void *dev, const void *host, size_t bytes); // host memory stays valid until the next synchronising call
int sg_sync(Exec &ex);
int sg_scan(Exec &ex, u64 *data, u64 n, u64 *total_host);          // in-place exclusive prefix sums; synchronises
template <class Op> int sg_launch(Exec &ex, u64 n, const Op &op);  // op(i) for i in [0, n)

template <typename T> static int sg_new(Exec &ex, T **p, size_t n) { return sg_alloc(ex, (void **)p, (n ? n : 1) * sizeof(T)); }

// ---------------------------------------------------------------- per-item helpers
#define SG_HD __host__ __device__ __forceinline__

SG_HD u64 ld64(const u64 *p)
{
#ifdef __CUDA_ARCH__
    u64 v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p)); // entries change under our feet: never from L1
    return v;
#else
    return *p;
#endif
}
SG_HD void at_or32(u32 *p, u32 v)
{
#ifdef __CUDA_ARCH__
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
SG_HD u64 at_add64(u64 *p, u64 v)
{
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    u64 o = *p;
    *p += v;
    return o;
#endif
}
SG_HD u32 at_cas32(u32 *p, u32 cmp, u32 v)
{
#ifdef __CUDA_ARCH__
    return atomicCAS(p, cmp, v);
#else
    u32 o = *p;
    if (o == cmp) *p = v;
    return o;
#endif
}
// `*(base + which) += 1` for every calling lane, returning the value before this lane's increment.  On the device the lanes
// of a warp that hit the same counter are combined into one atomic (P counters take millions of hits: the owner histogram
// and the scatter cursors); the order inside a group follows the lane number.
SG_HD u64 at_inc64_grouped(u64 *base, u32 which)
{
#ifdef __CUDA_ARCH__
    const u32 active = __activemask();
    const u32 peers = __match_any_sync(active, which);
    const u32 lane = threadIdx.x & 31, leader = (u32)__ffs((int)peers) - 1;
    u64 first = 0;
    if (lane == leader) first = atomicAdd(base + which, (u64)__popc(peers));
    first = __shfl_sync(peers, first, (int)leader);
    return first + (u64)__popc(peers & ((1u << lane) - 1));
#else
    return at_add64(base + which, 1);
#endif
}
SG_HD int popc4(u32 m) { return (int)((m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1)); }
SG_HD u32 first_bit4(u32 m) { return (m & 1) ? 0u : (m & 2) ? 1u : (m & 4) ? 2u : 3u; }

// vertex entry A[u], the layout of graph.cu: tag (2) | ptr (31) | dist (31)
//   UNRES: ptr = LOCAL oriented index of an interior ancestor, dist = steps to it
//   RES:   ptr = global edge id, or SEG_FLAG | local segment id; dist = rank from the head of the edge / segment
//   TERM:  low 32 bits = global node index;  NONE: isolated k-mer, palindrome alias, secondary key
constexpr u64 TAG_UNRES = 0, TAG_RES = 1, TAG_TERM = 2, TAG_NONE = 3;
constexpr u32 SEG_FLAG = 1u << 30;
SG_HD u64 a_make(u64 tag, u64 ptr, u64 dist) { return (tag << 62) | (ptr << 31) | (dist & 0x7FFFFFFFull); }
SG_HD u32 a_tag(u64 a) { return (u32)(a >> 62); }
SG_HD u32 a_ptr(u64 a) { return (u32)((a >> 31) & 0x7FFFFFFFu); }
SG_HD u32 a_dist(u64 a) { return (u32)(a & 0x7FFFFFFFu); }

struct Peer {
    const u64 *keys; // [n] the rank's k-mers; index = local vertex id
    const u32 *slot; // [cap] open-addressing index over keys, NONE32 = free
    const u8 *tag;   // [cap] fingerprint of the key behind slot[i] (fp_tag of its hash, never 0), 0 = free: a negative probe -- 3 of
                     // 4 in Graph.buildGraph -- ends on one byte, and a tag mismatch skips the two dependent loads (slot, key)
    const u8 *mask8; // [n] out | in << 4 of the stored orientation
    u64 *A;          // [2n] vertex entries
    u64 n, cap;
};

struct Ctx {
    int k, m, P, me, idx_bits;
    bool dual, v210;
    Peer peer[MAXR];
};

// global oriented vertex reference: rank in the top bits, 2 * vid + strand below; NONE32 = none
SG_HD u32 g_rank(const Ctx &c, u32 g) { return (u32)((u64)g >> c.idx_bits); }
SG_HD u32 g_idx(const Ctx &c, u32 g) { return g & (u32)((1ull << c.idx_bits) - 1); }
SG_HD u32 g_make(const Ctx &c, u32 rank, u32 idx) { return (u32)(((u64)rank << c.idx_bits) | idx); }

SG_HD int khash(const Ctx &c, u64 v) { return c.v210 ? scala_hash<true>(v) : scala_hash<false>(v); }

// ---- minimizer ownership.  H(p) = hash of the canonical m-mer at position p (0 <= p < w = k - m + 1); the owner of a k-mer
// is a function of min_p H(p), which x and rc(x) share (same multiset of canonical m-mers).  (first, mid, last) = H(0),
// min H(1 .. w-2), H(w-1): the successors of x keep positions 1 .. w-1 and add one m-mer at the end, the predecessors keep
// 0 .. w-2 and add one at the front, so the 8 neighbour owners cost one m-mer hash each.
constexpr u32 H_NONE = 0xFFFFFFFFu;
// 32-bit finalizer (murmur3 fmix32): the m-mer hashes are the inner loop of the owner function (21 per k-mer, 8 more for its
// neighbours), and an m-mer has at most 22 bits -- 64-bit multiplies here made the probes three times as long as the single-GPU kernel
SG_HD u32 fmix32(u32 h)
{
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
SG_HD u32 mmer_hash(u64 fwd, u64 rc) { return fmix32((u32)(fwd < rc ? fwd : rc) + 0x9E3779B9u); }
SG_HD int minimizer_len(int k)
{
    int m = (k + 1) / 2;
    if (m < 2) m = 2;
    if (m > 11) m = 11;
    return m > k ? k : m;
}
struct MinParts {
    u32 first, mid, last;
};
SG_HD MinParts min_parts(u64 x, u64 rcx, int k, int m)
{
    const u64 mm = (1ull << (2 * m)) - 1;
    const int w = k - m + 1;
    MinParts r;
    r.first = mmer_hash(x & mm, (rcx >> (2 * (k - m))) & mm);
    r.mid = H_NONE;
    for (int p = 1; p < w - 1; p++) {
        const u32 h = mmer_hash((x >> (2 * p)) & mm, (rcx >> (2 * (k - m - p))) & mm);
        r.mid = h < r.mid ? h : r.mid;
    }
    r.last = w > 1 ? mmer_hash((x >> (2 * (w - 1))) & mm, rcx & mm) : r.first;
    return r;
}
SG_HD u32 min3(u32 a, u32 b, u32 c) { a = a < b ? a : b; return a < c ? a : c; }
// the minimum of w hashes crowds towards 0: it is hashed once more before it picks the rank
SG_HD u32 owner_from_hash(u32 h, int P) { return (u32)(((u64)fmix32(h ^ 0x7F4A7C15u) * (u64)P) >> 32); }
SG_HD u32 owner_of_kmer(u64 x, int k, int m, int P)
{
    const MinParts mp = min_parts(x, revcomp(x, k), k, m);
    return owner_from_hash(min3(mp.first, mp.mid, mp.last), P);
}
// the neighbour x.drop(1) :+ b (succ = true) or b +: x.take(k-1) (succ = false): the k-mer, its reverse complement (one shift
// of rc(x), no bit reversal) and its owner from the parts of x
struct Neighbour {
    u64 q, rq;
    u32 owner;
};
SG_HD Neighbour neighbour_of(const MinParts &mp, u64 x, u64 rcx, int k, int m, int P, bool succ, u32 b)
{
    const u64 mm = (1ull << (2 * m)) - 1;
    const int w = k - m + 1;
    Neighbour nb;
    u32 h;
    if (succ) {
        nb.q = kmer_append(x, k, b);
        nb.rq = kmer_prepend(rcx, k, 3u - b);
        const u32 hn = mmer_hash((nb.q >> (2 * (w - 1))) & mm, nb.rq & mm);
        h = w > 1 ? min3(mp.mid, mp.last, hn) : hn;
    } else {
        nb.q = kmer_prepend(x, k, b);
        nb.rq = kmer_append(rcx, k, 3u - b);
        const u32 hn = mmer_hash(nb.q & mm, (nb.rq >> (2 * (k - m))) & mm);
        h = w > 1 ? min3(mp.first, mp.mid, hn) : hn;
    }
    nb.owner = owner_from_hash(h, P);
    return nb;
}
SG_HD u32 neighbour_owner(const MinParts &mp, u64 x, u64 rcx, int k, int m, int P, bool succ, u32 b)
{
    return neighbour_of(mp, x, rcx, k, m, P, succ, b).owner;
}

// ---- membership in one rank's index
// the probe from slot i on, with the tag of slot i already loaded (t8)
SG_HD bool probe_from(const Peer &t, u64 key, u32 tg, u64 i, u32 t8, u32 *vid)
{
    for (;;) {
        if (t8 == 0) return false;
        if (t8 == tg) {
            const u32 e = t.slot[i];
            if (t.keys[e] == key) { *vid = e; return true; }
        }
        i = next_slot(i, t.cap);
        t8 = t.tag[i];
    }
}
SG_HD bool probe_idx(const Peer &t, u64 key, u32 *vid)
{
    if (t.n == 0) return false;
    const u64 h = mix64(key), i = slot_of(h, t.cap);
    return probe_from(t, key, fp_tag(h), i, t.tag[i], vid);
}
// Graph.buildGraph.contains (Graph.scala:270) with the orientation rules of common.cuh find_oriented: one probe of the
// canonical orientation unless the two hashes tie or keys were inserted as-is (dual); if both orientations are stored the
// numerically smaller key is the primary one.  *g = the primary vertex, strand bit set when the stored key is rc(q) != q.
SG_HD bool find_g(const Ctx &c, u32 owner, u64 q, u64 r /* = revcomp(q) */, u32 *g)
{
    const Peer &t = c.peer[owner];
    const int hq = khash(c, q), hr = khash(c, r);
    u32 v;
    if (!c.dual && hq != hr) {
        const u64 cq = hq < hr ? q : r;
        if (!probe_idx(t, cq, &v)) return false;
        *g = g_make(c, owner, 2 * v + (cq != q));
        return true;
    }
    u32 vq = 0, vr = 0;
    const bool fq = probe_idx(t, q, &vq);
    const bool fr = r != q && probe_idx(t, r, &vr);
    if (!fq && !fr) return false;
    const bool use_r = fr && (!fq || r < q);
    *g = g_make(c, owner, 2 * (use_r ? vr : vq) + (use_r ? 1u : 0u));
    return true;
}
SG_HD bool is_secondary(const Ctx &c, const Peer &t, u64 key)
{
    const u64 r = revcomp(key, c.k);
    if (r >= key) return false;
    if (!c.dual && khash(c, key) != khash(c, r)) return false;
    u32 v;
    return probe_idx(t, r, &v);
}

SG_HD u32 comp_mask4(u32 m) { return ((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3); }
SG_HD void oriented_masks(u32 m8, u32 s, u32 *out, u32 *in)
{
    const u32 o = m8 & 15, i = m8 >> 4;
    *out = s ? comp_mask4(i) : o;
    *in = s ? comp_mask4(o) : i;
}
// Graph.scala:323: terminal iff (in != 1 || out != 1) && (in != 0 || out != 0)
SG_HD u32 classify(u32 out, u32 in)
{
    const int no = popc4(out), ni = popc4(in);
    if (no == 0 && ni == 0) return (u32)TAG_NONE;
    if (no == 1 && ni == 1) return (u32)TAG_UNRES;
    return (u32)TAG_TERM;
}
SG_HD bool is_alias(const Ctx &c, u32 g)
{
    // (vid, 1) of a palindromic key is the same oriented k-mer as (vid, 0); only even k has palindromes
    if (!(g & 1) || (c.k & 1)) return false;
    const u64 x = c.peer[g_rank(c, g)].keys[g_idx(c, g) >> 1];
    return revcomp(x, c.k) == x;
}
SG_HD u32 normalise(const Ctx &c, u32 g) { return is_alias(c, g) ? g ^ 1u : g; }
SG_HD u32 vertex_type(const Ctx &c, u32 g, u32 *out, u32 *in)
{
    const u32 i = g_idx(c, g);
    oriented_masks(c.peer[g_rank(c, g)].mask8[i >> 1], i & 1, out, in);
    return classify(*out, *in);
}
SG_HD u64 oriented_kmer(u64 key, u32 s, int k) { return s ? revcomp(key, k) : key; }

// arrays of the rank that runs an op (u = LOCAL oriented index everywhere below)
struct Local {
    const u32 *nbr_out, *nbr_in; // [n] the unique oriented successor / predecessor of the stored orientation (global refs)
    u64 *node_idx, *edge_idx, *seg_idx; // [2n] counts, then exclusive prefix sums
    u32 *seg_pred;                // [segments] global ref of the segment head's predecessor
    u32 *edge_first;              // [edges of this rank] global ref of the first vertex after the start node
    u64 node_base, edge_base, seg_base;
};
// the global graph arrays (this process's copy)
struct Global {
    u64 *node_kmer;
    u32 *edge_start, *edge_end;
    u64 *edge_len; // becomes edge_off after the scan
    const u64 *edge_off;
    u32 *bases;
};

SG_HD u32 pred_of(const Ctx &c, const Local &L, u32 u)
{
    const u32 p = (u & 1) ? (L.nbr_out[u >> 1] ^ 1u) : L.nbr_in[u >> 1];
    return normalise(c, p);
}
SG_HD u32 succ_of(const Ctx &c, const Local &L, u32 u)
{
    const u32 p = (u & 1) ? (L.nbr_in[u >> 1] ^ 1u) : L.nbr_out[u >> 1];
    return normalise(c, p);
}

// ---------------------------------------------------------------- ops (one item per call)
struct OwnerCountOp { // histogram of the owners of this rank's kept keys
    const u64 *keys; int k, m, P; u64 *cnt;
    SG_HD void operator()(u64 i) const { at_inc64_grouped(cnt, owner_of_kmer(keys[i], k, m, P)); }
};
struct OwnerScatterOp { // keys grouped by owner (cursor = exclusive offsets of the histogram)
    const u64 *keys; int k, m, P; u64 *cursor; u64 *out;
    SG_HD void operator()(u64 i) const { out[at_inc64_grouped(cursor, owner_of_kmer(keys[i], k, m, P))] = keys[i]; }
};
struct IndexInsertOp { // putNew of entry i: first free slot from the key's home; the slot's tag follows (read only after a barrier)
    const u64 *keys; u32 *slot; u8 *tag; u64 cap;
    SG_HD void operator()(u64 i) const
    {
        const u64 h = mix64(keys[i]);
        u64 s = slot_of(h, cap);
        while (at_cas32(slot + s, NONE32, (u32)i) != NONE32) s = next_slot(s, cap);
        tag[s] = (u8)fp_tag(h);
    }
};
// incoming / outcoming (Graph.scala:272-282) of every stored key with ONE ITEM PER (stored key, neighbour query).  The first form
// (MasksOp, rounds 2a-2q) kept the 8 queries of a key in the registers of one thread: 76 registers, a third of the warps resident
// (profiles/masks_r2r_ncu_full_summary.csv), which hides little of the latency of a probe that goes to a peer's index over
// NVLink.  Here a key's minimizer parts are computed once (PartsOp), every query is an item of its own (ProbeOp: 8 consecutive
// items = 8 consecutive lanes share a key, 26 registers) and a third sweep folds the 8 answers of a key into its mask byte and
// unique neighbours (CombineOp: one 32-byte read per key).  Per rank of 8 virtual ranks on C2: 0.229 -> 0.018 + 0.094 + 0.009 ms.
struct PartsOp { // per stored key: is it a vertex at all (a secondary orientation is not), and the parts of its minimizer
    Ctx c; u32 *parts; u8 *skip;
    SG_HD void operator()(u64 v) const
    {
        const Peer &me = c.peer[c.me];
        const u64 x = me.keys[v];
        skip[v] = is_secondary(c, me, x) ? 1 : 0;
        const MinParts mp = min_parts(x, revcomp(x, c.k), c.k, c.m);
        parts[3 * v] = mp.first;
        parts[3 * v + 1] = mp.mid;
        parts[3 * v + 2] = mp.last;
    }
};
struct ProbeOp { // item i = 8 v + j: the successor of key v by base j / 2 (j even) or its predecessor (j odd), NONE32 = absent
    Ctx c; const u32 *parts; const u8 *skip; u32 *found;
    SG_HD void operator()(u64 i) const
    {
        const u64 v = i >> 3;
        const u32 j = (u32)i & 7u;
        u32 g = NONE32;
        if (!skip[v]) {
            const u64 x = c.peer[c.me].keys[v], rcx = revcomp(x, c.k);
            const MinParts mp{ parts[3 * v], parts[3 * v + 1], parts[3 * v + 2] };
            const Neighbour nb = neighbour_of(mp, x, rcx, c.k, c.m, c.P, (j & 1) == 0, j >> 1);
            u32 f = 0;
            if (find_g(c, nb.owner, nb.q, nb.rq, &f)) g = f;
        }
        found[i] = g;
    }
};
struct CombineOp { // the 8 answers of a key -> incoming / outcoming masks and the unique neighbour of either side
    const u32 *found; u8 *mask8; u32 *nbr_out, *nbr_in;
    SG_HD void operator()(u64 v) const
    {
        u32 out = 0, in = 0, so = NONE32, si = NONE32;
        for (u32 j = 0; j < 8; j++) {
            const u32 g = found[8 * v + j];
            if (g == NONE32) continue;
            if ((j & 1) == 0) { out |= 1u << (j >> 1); so = g; } else { in |= 1u << (j >> 1); si = g; }
        }
        mask8[v] = (u8)(out | (in << 4));
        nbr_out[v] = so;
        nbr_in[v] = si;
    }
};
struct ClassifyOp { // per oriented vertex: is it a node, how many edges start there, does it head a segment
    Ctx c; Local L;
    SG_HD void operator()(u64 uu) const
    {
        const u32 u = (u32)uu, g = g_make(c, c.me, u);
        u32 out = 0, in = 0;
        const u32 t = is_alias(c, g) ? (u32)TAG_NONE : vertex_type(c, g, &out, &in);
        L.node_idx[u] = t == TAG_TERM;
        L.edge_idx[u] = t == TAG_TERM ? (u64)popc4(out) : 0;
        u64 seg = 0;
        if (t == TAG_UNRES) {
            const u32 p = pred_of(c, L, u);
            u32 po, pi;
            seg = g_rank(c, p) != (u32)c.me && vertex_type(c, p, &po, &pi) != TAG_TERM;
        }
        L.seg_idx[u] = seg;
    }
};
struct InitVerticesOp { // entries of terminals / none / interiors that do not follow a node; node k-mers (Graph.scala:343-347)
    Ctx c; Local L; Global G;
    SG_HD void operator()(u64 uu) const
    {
        const u32 u = (u32)uu, g = g_make(c, c.me, u);
        u64 *A = c.peer[c.me].A;
        if (is_alias(c, g)) { A[u] = a_make(TAG_NONE, 0, 0); return; }
        u32 out, in;
        const u32 t = vertex_type(c, g, &out, &in);
        if (t == TAG_TERM) {
            const u64 ni = L.node_base + L.node_idx[u];
            A[u] = a_make(TAG_TERM, 0, 0) | ni;
            G.node_kmer[ni] = oriented_kmer(c.peer[c.me].keys[u >> 1], u & 1, c.k);
        } else if (t == TAG_NONE) {
            A[u] = a_make(TAG_NONE, 0, 0);
        } else {
            const u32 p = pred_of(c, L, u);
            u32 po, pi;
            if (vertex_type(c, p, &po, &pi) == TAG_TERM) return; // head of an edge: written by the node's owner (StartEdgesOp)
            if (g_rank(c, p) == (u32)c.me) {
                A[u] = a_make(TAG_UNRES, g_idx(c, p), 1);
            } else {
                const u64 s = L.seg_idx[u];
                A[u] = a_make(TAG_RES, SEG_FLAG | s, 0);
                L.seg_pred[s] = p;
            }
        }
    }
};
struct StartEdgesOp { // buildEdges (Graph.scala:349-365), first step of every edge, out-bases in Base.fromInt order (351)
    Ctx c; Local L; Global G;
    SG_HD void operator()(u64 uu) const
    {
        const u32 u = (u32)uu, g = g_make(c, c.me, u);
        u32 out, in;
        if (is_alias(c, g) || vertex_type(c, g, &out, &in) != TAG_TERM) return;
        const u64 x = oriented_kmer(c.peer[c.me].keys[u >> 1], u & 1, c.k);
        u64 e = L.edge_base + L.edge_idx[u];
        const u32 my_node = (u32)(L.node_base + L.node_idx[u]);
        for (u32 b = 0; b < 4; b++) {
            if (!(out & (1u << b))) continue;
            const u64 q = kmer_append(x, c.k, b);
            u32 w = NONE32;
            find_g(c, owner_of_kmer(q, c.k, c.m, c.P), q, revcomp(q, c.k), &w); // present by construction
            w = normalise(c, w);
            G.edge_start[e] = my_node;
            L.edge_first[e - L.edge_base] = w;
            u32 wo, wi;
            if (vertex_type(c, w, &wo, &wi) != TAG_TERM) c.peer[g_rank(c, w)].A[g_idx(c, w)] = a_make(TAG_RES, e, 0); // NVLink store
            e++;
        }
    }
};
// in-place pointer jumping over local links (graph.cu jump_kernel): every entry is read and written whole
constexpr int JUMPS = 4;
struct JumpOp {
    u64 *A; u32 *pending;
    SG_HD void operator()(u64 u) const
    {
        u64 a = ld64(A + u);
        if (a_tag(a) != TAG_UNRES) return;
        for (int j = 0; j < JUMPS; j++) {
            const u64 ap = ld64(A + a_ptr(a));
            a = a_make(a_tag(ap), a_ptr(ap), (u64)a_dist(a) + a_dist(ap));
            if (a_tag(a) != TAG_UNRES) break;
        }
        A[u] = a;
        if (a_tag(a) == TAG_UNRES) *pending = 1;
    }
};
struct SegFillOp { // entry of segment s = what its remote predecessor resolved to, one step further
    Ctx c; Local L; const u64 *seg_base_of_rank; u64 *S;
    SG_HD void operator()(u64 s) const
    {
        const u32 p = L.seg_pred[s], r = g_rank(c, p);
        const u64 a = ld64(c.peer[r].A + g_idx(c, p)); // NVLink load
        u64 e = a_make(TAG_NONE, 0, 0);               // a predecessor on an unresolved local loop cannot happen; dead if it does
        if (a_tag(a) == TAG_RES) {
            if (a_ptr(a) & SEG_FLAG) e = a_make(TAG_UNRES, seg_base_of_rank[r] + (a_ptr(a) & ~SEG_FLAG), (u64)a_dist(a) + 1);
            else e = a_make(TAG_RES, a_ptr(a), (u64)a_dist(a) + 1);
        }
        S[L.seg_base + s] = e;
    }
};
struct SegJumpOp { // list ranking of the segment list; chains of segments that close on themselves stay UNRES (perfect cycles)
    u64 *S; u32 *pending;
    SG_HD void operator()(u64 i) const
    {
        u64 a = ld64(S + i);
        if (a_tag(a) != TAG_UNRES) return;
        for (int j = 0; j < JUMPS; j++) {
            const u64 ap = ld64(S + a_ptr(a));
            a = a_make(a_tag(ap), a_ptr(ap), (u64)a_dist(a) + a_dist(ap));
            if (a_tag(a) != TAG_UNRES) break;
        }
        S[i] = a;
        if (a_tag(a) == TAG_UNRES) *pending = 1;
    }
};
struct FinalizeOp { // (segment, rank in segment) -> (edge, rank in edge)
    Ctx c; Local L; const u64 *S;
    SG_HD void operator()(u64 u) const
    {
        u64 *A = c.peer[c.me].A;
        const u64 a = A[u];
        if (a_tag(a) != TAG_RES || !(a_ptr(a) & SEG_FLAG)) return;
        const u64 sgm = S[L.seg_base + (a_ptr(a) & ~SEG_FLAG)];
        A[u] = a_tag(sgm) == TAG_RES ? a_make(TAG_RES, a_ptr(sgm), (u64)a_dist(sgm) + a_dist(a)) : a_make(TAG_UNRES, 0, 0);
    }
};
struct CloseOp { // the last interior vertex of every chain closes its edge: end node and length (Graph.scala:363-364)
    Ctx c; Local L; Global G; u64 *cycle_vertices;
    SG_HD void operator()(u64 uu) const
    {
        const u32 u = (u32)uu;
        const u64 a = c.peer[c.me].A[u];
        if (a_tag(a) == TAG_RES) {
            // terminal or not is read off the (immutable) mask byte: an interior successor's entry may be changing right now
            // (its owner's FinalizeOp), a terminal's entry is final since InitVerticesOp
            const u32 s = succ_of(c, L, u);
            u32 so, si;
            if (vertex_type(c, s, &so, &si) == TAG_TERM) {
                G.edge_end[a_ptr(a)] = (u32)ld64(c.peer[g_rank(c, s)].A + g_idx(c, s));
                G.edge_len[a_ptr(a)] = (u64)a_dist(a) + 2;
            }
        } else if (a_tag(a) == TAG_UNRES) {
            at_add64(cycle_vertices, 1); // perfect cycle: never reached from a node, ignored like Graph.scala:375
        }
    }
};
struct CloseNodeEdgesOp { // node -> node edges of length 1
    Ctx c; Local L; Global G;
    SG_HD void operator()(u64 j) const
    {
        const u32 w = L.edge_first[j];
        u32 wo, wi;
        if (vertex_type(c, w, &wo, &wi) == TAG_TERM) {
            G.edge_end[L.edge_base + j] = (u32)ld64(c.peer[g_rank(c, w)].A + g_idx(c, w));
            G.edge_len[L.edge_base + j] = 1;
        }
    }
};
SG_HD void put_base(u32 *bases, u64 pos, u32 b) { at_or32(bases + (pos >> 4), b << (2 * (u32)(pos & 15))); }
struct WriteBasesOp { // Edge.seq: the base appended at every step of the walk (Graph.scala:352,358)
    Ctx c; Local L; Global G; const u64 *edge_len;
    SG_HD void operator()(u64 uu) const
    {
        const u32 u = (u32)uu;
        const u64 a = c.peer[c.me].A[u];
        u32 out, in;
        if (a_tag(a) == TAG_RES) {
            const u64 x = c.peer[c.me].keys[u >> 1];
            const u32 last = (u & 1) ? 3u - (u32)(x & 3) : (u32)(x >> (2 * (c.k - 1))) & 3u;
            const u64 pos = G.edge_off[a_ptr(a)] + a_dist(a);
            put_base(G.bases, pos, last);
            if ((u64)a_dist(a) + 2 == edge_len[a_ptr(a)]) { // tail: the step into the end node appends its single out-base
                oriented_masks(c.peer[c.me].mask8[u >> 1], u & 1, &out, &in);
                put_base(G.bases, pos + 1, first_bit4(out));
            }
        } else if (a_tag(a) == TAG_TERM) {
            oriented_masks(c.peer[c.me].mask8[u >> 1], u & 1, &out, &in);
            u64 e = L.edge_base + L.edge_idx[u];
            for (u32 b = 0; b < 4; b++) {
                if (!(out & (1u << b))) continue;
                if (edge_len[e] == 1) put_base(G.bases, G.edge_off[e], b);
                e++;
            }
        }
    }
};

// ---------------------------------------------------------------- orchestration
struct RankInput {
    Exec *ex;
    const u64 *keys; // device: this rank's kept k-mers (any order)
    u64 n;
};
struct Result {
    u64 *node_kmer = nullptr;
    u32 *edge_start = nullptr, *edge_end = nullptr;
    u64 *edge_off = nullptr; // [E + 1]
    u32 *bases = nullptr;
    u64 n_nodes = 0, n_edges = 0, n_bases = 0;
    u64 kept = 0, segments = 0, cycle_vertices = 0, cross_links = 0; // totals over the ranks
    int jump_rounds = 0, seg_rounds = 0;
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline u64 index_cap(u64 n) { return (n * 2 + 1024) / 1024 * 1024; } // load <= 1/2
// peer window of a rank with n keys: keys | A | slot | mask8 | tag
struct WindowLayout {
    size_t keys, A, slot, mask8, tag, bytes;
    explicit WindowLayout(u64 n)
    {
        keys = 0;
        A = align256(keys + (size_t)n * 8);
        slot = align256(A + (size_t)n * 16);
        mask8 = align256(slot + (size_t)index_cap(n) * 4);
        tag = align256(mask8 + (size_t)n);
        bytes = align256(tag + (size_t)index_cap(n));
    }
};
inline size_t base_words(u64 n_bases) { return (size_t)((n_bases + 15) / 16) + 2; }

// One sharded build.  in[l] belongs to rank fab.mine[l]; the resulting graph (identical on every process) is allocated
// through in[0].ex.  Collective: every process of the fabric calls it with the same k / dual / v210.
inline int build(Fabric &fab, const std::vector<RankInput> &in, int k, bool dual, bool v210, Result *res)
{
    const int P = fab.P, nl = (int)fab.mine.size();
    if (P < 1 || P > MAXR || nl < 1 || (int)in.size() != nl) { set_error("sharded build: bad rank set"); return GB_E_ARG; }
    const int m = minimizer_len(k);
    int rank_bits = 0;
    while ((1 << rank_bits) < P) rank_bits++;
    const int idx_bits = 32 - rank_bits;

    // ---- 1. owners of this rank's keys, counts to everybody
    std::vector<u64 *> d_cnt(nl);
    std::vector<u64> cnt_mat((size_t)nl * P * P), scnt_flat((size_t)nl * P);
    for (int l = 0; l < nl; l++) {
        Exec &ex = *in[l].ex;
        GB_TRY(sg_new(ex, &d_cnt[l], 2 * (size_t)P));
        GB_TRY(sg_zero(ex, d_cnt[l], 2 * (size_t)P * 8));
        GB_TRY(sg_launch(ex, in[l].n, OwnerCountOp{ in[l].keys, k, m, P, d_cnt[l] }));
        GB_TRY(sg_read(ex, &scnt_flat[(size_t)l * P], d_cnt[l], (size_t)P * 8));
    }
    {
        std::vector<const void *> contrib(nl);
        std::vector<void *> all(nl);
        for (int l = 0; l < nl; l++) { contrib[l] = &scnt_flat[(size_t)l * P]; all[l] = &cnt_mat[(size_t)l * P * P]; }
        GB_TRY(fab.allgather_host(contrib.data(), all.data(), (size_t)P * 8));
    }
    // cnt_mat[l][src][dst] is the same matrix for every l
    std::vector<u64> n_of(P, 0);
    for (int src = 0; src < P; src++)
        for (int dst = 0; dst < P; dst++) n_of[dst] += cnt_mat[(size_t)src * P + dst];
    u64 kept = 0;
    for (int r = 0; r < P; r++) {
        kept += n_of[r];
        if (2 * n_of[r] + 2 > (1ull << idx_bits)) {
            set_error("sharded build: %llu k-mers on rank %d exceed the %d-bit vertex index", n_of[r], r, idx_bits);
            return GB_E_CAPACITY;
        }
    }

    fab.tick("owners counted");
    // ---- 2. peer windows, keys to their owners, index
    std::vector<size_t> wbytes(P);
    for (int r = 0; r < P; r++) wbytes[r] = WindowLayout(n_of[r]).bytes;
    std::vector<void *> window(nl);
    std::vector<PeerPtrs> peers(nl);
    GB_TRY(fab.windows(wbytes.data(), window.data(), peers.data()));
    std::vector<Ctx> ctx(nl);
    for (int l = 0; l < nl; l++) {
        Ctx &c = ctx[l];
        c.k = k; c.m = m; c.P = P; c.me = fab.mine[l]; c.idx_bits = idx_bits; c.dual = dual; c.v210 = v210;
        for (int r = 0; r < P; r++) {
            const WindowLayout wl(n_of[r]);
            char *b = (char *)peers[l].p[r];
            c.peer[r].keys = (const u64 *)(b + wl.keys);
            c.peer[r].A = (u64 *)(b + wl.A);
            c.peer[r].slot = (const u32 *)(b + wl.slot);
            c.peer[r].mask8 = (const u8 *)(b + wl.mask8);
            c.peer[r].tag = (const u8 *)(b + wl.tag);
            c.peer[r].n = n_of[r];
            c.peer[r].cap = index_cap(n_of[r]);
        }
        for (int r = P; r < MAXR; r++) c.peer[r] = Peer{ nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0 };
    }
    {
        std::vector<u64 *> send(nl), recv(nl);
        std::vector<Row> soff(nl), scnt(nl), roff(nl), rcnt(nl);
        for (int l = 0; l < nl; l++) {
            Exec &ex = *in[l].ex;
            const int me = fab.mine[l];
            u64 t = 0, cursor[MAXR];
            for (int p = 0; p < P; p++) { soff[l].v[p] = cursor[p] = t; scnt[l].v[p] = cnt_mat[(size_t)me * P + p]; t += scnt[l].v[p]; }
            t = 0;
            for (int p = 0; p < P; p++) { roff[l].v[p] = t; rcnt[l].v[p] = cnt_mat[(size_t)p * P + me]; t += rcnt[l].v[p]; }
            GB_TRY(sg_new(ex, &send[l], (size_t)in[l].n));
            // cursors live behind the counts; filled from the host copy
            for (int p = 0; p < P; p++) scnt_flat[(size_t)l * P + p] = cursor[p];
            u64 *d_cursor = d_cnt[l] + P;
            GB_TRY(sg_write_host(ex, d_cursor, &scnt_flat[(size_t)l * P], (size_t)P * 8));
            GB_TRY(sg_launch(ex, in[l].n, OwnerScatterOp{ in[l].keys, k, m, P, d_cursor, send[l] }));
            recv[l] = (u64 *)((char *)window[l] + WindowLayout(n_of[me]).keys);
        }
        std::vector<const u64 *> csend(send.begin(), send.end());
        GB_TRY(fab.alltoallv_u64(csend.data(), soff.data(), scnt.data(), recv.data(), roff.data(), rcnt.data()));
    }
    for (int l = 0; l < nl; l++) {
        Exec &ex = *in[l].ex;
        const Peer &me = ctx[l].peer[ctx[l].me];
        GB_TRY(sg_fill_ff(ex, (void *)me.slot, (size_t)me.cap * 4));
        GB_TRY(sg_zero(ex, (void *)me.tag, (size_t)me.cap));
        GB_TRY(sg_launch(ex, me.n, IndexInsertOp{ me.keys, (u32 *)me.slot, (u8 *)me.tag, me.cap }));
    }
    GB_TRY(fab.barrier());

    fab.tick("windows + re-routing + index");
    // ---- 3. masks
    std::vector<Local> loc(nl);
    std::vector<u32 *> nbr_out(nl), nbr_in(nl);
    for (int l = 0; l < nl; l++) {
        Exec &ex = *in[l].ex;
        const u64 n = ctx[l].peer[ctx[l].me].n;
        GB_TRY(sg_new(ex, &nbr_out[l], (size_t)n));
        GB_TRY(sg_new(ex, &nbr_in[l], (size_t)n));
        u8 *mask8 = (u8 *)ctx[l].peer[ctx[l].me].mask8;
        u32 *parts = nullptr, *found = nullptr;
        u8 *skip = nullptr;
        GB_TRY(sg_new(ex, &parts, 3 * (size_t)n));
        GB_TRY(sg_new(ex, &skip, (size_t)n));
        GB_TRY(sg_new(ex, &found, 8 * (size_t)n));
        GB_TRY(sg_launch(ex, n, PartsOp{ ctx[l], parts, skip }));
        GB_TRY(sg_launch(ex, 8 * n, ProbeOp{ ctx[l], parts, skip, found }));
        GB_TRY(sg_launch(ex, n, CombineOp{ found, mask8, nbr_out[l], nbr_in[l] }));
        loc[l] = Local{ nbr_out[l], nbr_in[l], nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0 };
    }
    GB_TRY(fab.barrier());

    fab.tick("masks");
    // ---- 4. classify, numbering
    std::vector<u64> tot((size_t)nl * 3), tot_all((size_t)nl * 3 * P);
    for (int l = 0; l < nl; l++) {
        Exec &ex = *in[l].ex;
        const u64 n2 = 2 * ctx[l].peer[ctx[l].me].n;
        GB_TRY(sg_new(ex, &loc[l].node_idx, (size_t)n2));
        GB_TRY(sg_new(ex, &loc[l].edge_idx, (size_t)n2));
        GB_TRY(sg_new(ex, &loc[l].seg_idx, (size_t)n2));
        GB_TRY(sg_launch(ex, n2, ClassifyOp{ ctx[l], loc[l] }));
        GB_TRY(sg_scan(ex, loc[l].node_idx, n2, &tot[(size_t)l * 3 + 0]));
        GB_TRY(sg_scan(ex, loc[l].edge_idx, n2, &tot[(size_t)l * 3 + 1]));
        GB_TRY(sg_scan(ex, loc[l].seg_idx, n2, &tot[(size_t)l * 3 + 2]));
    }
    {
        std::vector<const void *> contrib(nl);
        std::vector<void *> all(nl);
        for (int l = 0; l < nl; l++) { contrib[l] = &tot[(size_t)l * 3]; all[l] = &tot_all[(size_t)l * 3 * P]; }
        GB_TRY(fab.allgather_host(contrib.data(), all.data(), 3 * 8));
    }
    std::vector<u64> node_base(P), edge_base(P), seg_base(P), seg_cnt(P);
    u64 N = 0, E = 0, S = 0;
    for (int r = 0; r < P; r++) {
        node_base[r] = N; edge_base[r] = E; seg_base[r] = S;
        N += tot_all[(size_t)r * 3 + 0];
        E += tot_all[(size_t)r * 3 + 1];
        seg_cnt[r] = tot_all[(size_t)r * 3 + 2];
        S += seg_cnt[r];
    }
    if (N >= 0xFFFFFFFFull || E >= (1ull << 30) || S >= (1ull << 30)) {
        set_error("sharded build: graph too large: %llu nodes, %llu edges, %llu segments", N, E, S);
        return GB_E_CAPACITY;
    }

    fab.tick("classify + scans");
    // ---- 5. the global arrays of this process, vertex entries, first step of every edge
    Exec &ex0 = *in[0].ex;
    Global G;
    GB_TRY(sg_graph_alloc(ex0, (void **)&G.node_kmer, (N ? N : 1) * 8));
    GB_TRY(sg_graph_alloc(ex0, (void **)&G.edge_start, (E ? E : 1) * 4));
    GB_TRY(sg_graph_alloc(ex0, (void **)&G.edge_end, (E ? E : 1) * 4));
    GB_TRY(sg_graph_alloc(ex0, (void **)&G.edge_len, (E + 1) * 8));
    G.edge_off = G.edge_len;
    G.bases = nullptr;
    GB_TRY(sg_zero(ex0, G.node_kmer, (N ? N : 1) * 8));
    GB_TRY(sg_zero(ex0, G.edge_start, (E ? E : 1) * 4));
    GB_TRY(sg_zero(ex0, G.edge_end, (E ? E : 1) * 4));
    GB_TRY(sg_zero(ex0, G.edge_len, (E + 1) * 8));
    if (nl > 1) GB_TRY(sg_sync(ex0)); // the ranks of this process write into the shared copy from their own streams
    for (int l = 0; l < nl; l++) {
        Exec &ex = *in[l].ex;
        const int me = ctx[l].me;
        const u64 n2 = 2 * ctx[l].peer[me].n;
        loc[l].node_base = node_base[me]; loc[l].edge_base = edge_base[me]; loc[l].seg_base = seg_base[me];
        GB_TRY(sg_new(ex, &loc[l].seg_pred, (size_t)seg_cnt[me]));
        GB_TRY(sg_new(ex, &loc[l].edge_first, (size_t)tot_all[(size_t)me * 3 + 1]));
        GB_TRY(sg_launch(ex, n2, InitVerticesOp{ ctx[l], loc[l], G }));
        GB_TRY(sg_launch(ex, n2, StartEdgesOp{ ctx[l], loc[l], G }));
    }
    GB_TRY(fab.barrier());

    fab.tick("vertices + edge starts");
    // ---- 6. list ranking over local links
    int jump_rounds = 0;
    for (int l = 0; l < nl; l++) {
        Exec &ex = *in[l].ex;
        const u64 n2 = 2 * ctx[l].peer[ctx[l].me].n;
        u32 *d_pending;
        GB_TRY(sg_new(ex, &d_pending, 1));
        int bound = 2, rounds = 0;
        while ((1ull << (bound - 2)) < n2 + 1) bound++; // each launch makes at least one jump
        u32 pending = n2 ? 1 : 0;
        while (pending && rounds < bound) {
            GB_TRY(sg_zero(ex, d_pending, 4));
            GB_TRY(sg_launch(ex, n2, JumpOp{ ctx[l].peer[ctx[l].me].A, d_pending }));
            GB_TRY(sg_read(ex, &pending, d_pending, 4));
            rounds++;
        }
        jump_rounds = rounds > jump_rounds ? rounds : jump_rounds;
    }
    GB_TRY(fab.barrier());

    fab.tick("local list ranking");
    // ---- 7. segment list: fill, gather, rank, fold back
    int seg_rounds = 0;
    {
        std::vector<u64 *> Sg(nl);
        for (int l = 0; l < nl; l++) {
            Exec &ex = *in[l].ex;
            u64 *d_seg_base;
            GB_TRY(sg_new(ex, &Sg[l], (size_t)S));
            GB_TRY(sg_new(ex, &d_seg_base, (size_t)P));
            GB_TRY(sg_write_host(ex, d_seg_base, seg_base.data(), (size_t)P * 8));
            GB_TRY(sg_launch(ex, seg_cnt[ctx[l].me], SegFillOp{ ctx[l], loc[l], d_seg_base, Sg[l] }));
        }
        GB_TRY(fab.allgatherv_u64(Sg.data(), seg_base.data(), seg_cnt.data()));
        for (int l = 0; l < nl; l++) {
            Exec &ex = *in[l].ex;
            u32 *d_pending;
            GB_TRY(sg_new(ex, &d_pending, 1));
            int bound = 2, rounds = 0;
            while ((1ull << (bound - 2)) < S + 1) bound++;
            u32 pending = S ? 1 : 0;
            while (pending && rounds < bound) {
                GB_TRY(sg_zero(ex, d_pending, 4));
                GB_TRY(sg_launch(ex, S, SegJumpOp{ Sg[l], d_pending }));
                GB_TRY(sg_read(ex, &pending, d_pending, 4));
                rounds++;
            }
            seg_rounds = rounds > seg_rounds ? rounds : seg_rounds;
            GB_TRY(sg_launch(ex, 2 * ctx[l].peer[ctx[l].me].n, FinalizeOp{ ctx[l], loc[l], Sg[l] }));
        }
    }

    fab.tick("segments");
    // ---- 8. edge ends and lengths
    std::vector<u64> cyc(nl, 0), cyc_all((size_t)nl * P);
    for (int l = 0; l < nl; l++) {
        Exec &ex = *in[l].ex;
        const int me = ctx[l].me;
        u64 *d_cyc;
        GB_TRY(sg_new(ex, &d_cyc, 1));
        GB_TRY(sg_zero(ex, d_cyc, 8));
        GB_TRY(sg_launch(ex, 2 * ctx[l].peer[me].n, CloseOp{ ctx[l], loc[l], G, d_cyc }));
        GB_TRY(sg_launch(ex, tot_all[(size_t)me * 3 + 1], CloseNodeEdgesOp{ ctx[l], loc[l], G }));
        GB_TRY(sg_read(ex, &cyc[l], d_cyc, 8));
    }
    {
        std::vector<const void *> contrib(nl);
        std::vector<void *> all(nl);
        for (int l = 0; l < nl; l++) { contrib[l] = &cyc[l]; all[l] = &cyc_all[(size_t)l * P]; }
        GB_TRY(fab.allgather_host(contrib.data(), all.data(), 8));
    }
    GB_TRY(fab.allreduce_sum(G.node_kmer, (size_t)N, 8));
    GB_TRY(fab.allreduce_sum(G.edge_start, (size_t)E, 4));
    GB_TRY(fab.allreduce_sum(G.edge_end, (size_t)E, 4));
    GB_TRY(fab.allreduce_sum(G.edge_len, (size_t)E, 8));

    fab.tick("edges closed + array sums");
    // ---- 9. offsets and bases
    u64 *edge_len;
    GB_TRY(sg_new(ex0, &edge_len, (size_t)E + 1));
    GB_TRY(sg_copy(ex0, edge_len, G.edge_len, (E + 1) * 8));
    u64 n_bases = 0;
    GB_TRY(sg_scan(ex0, G.edge_len, E + 1, &n_bases));
    GB_TRY(sg_graph_alloc(ex0, (void **)&G.bases, base_words(n_bases) * 4));
    GB_TRY(sg_zero(ex0, G.bases, base_words(n_bases) * 4));
    if (nl > 1) GB_TRY(sg_sync(ex0));
    for (int l = 0; l < nl; l++)
        GB_TRY(sg_launch(*in[l].ex, 2 * ctx[l].peer[ctx[l].me].n, WriteBasesOp{ ctx[l], loc[l], G, edge_len }));
    for (int l = 0; l < nl; l++) GB_TRY(sg_sync(*in[l].ex));
    GB_TRY(fab.allreduce_sum(G.bases, base_words(n_bases), 4));
    GB_TRY(fab.barrier()); // nobody releases its window while a peer may still read it
    fab.tick("bases written + summed");

    res->node_kmer = G.node_kmer; res->edge_start = G.edge_start; res->edge_end = G.edge_end;
    res->edge_off = G.edge_len; res->bases = G.bases;
    res->n_nodes = N; res->n_edges = E; res->n_bases = n_bases;
    res->kept = kept; res->segments = S; res->jump_rounds = jump_rounds; res->seg_rounds = seg_rounds;
    res->cycle_vertices = 0;
    for (int r = 0; r < P; r++) res->cycle_vertices += cyc_all[r];
    return GB_OK;
}

// All P ranks in one process (virtual shards on one device, or the g++ emulation): peers are plain pointers, transfers are
// copies, barriers are the order of issue (the ranks share one stream).
struct LocalFabric : Fabric {
    std::vector<Exec *> ex;
    LocalFabric(int n_ranks, const std::vector<Exec *> &execs) : ex(execs)
    {
        P = n_ranks;
        for (int r = 0; r < n_ranks; r++) mine.push_back(r);
    }
    int allgather_host(const void *const *contrib, void *const *all, size_t bytes) override
    {
        for (int l = 0; l < P; l++)
            for (int r = 0; r < P; r++) memcpy((char *)all[l] + (size_t)r * bytes, contrib[r], bytes);
        return GB_OK;
    }
    int windows(const size_t *bytes_of_rank, void **window, PeerPtrs *peers) override
    {
        for (int r = 0; r < P; r++) GB_TRY(sg_alloc(*ex[r], &window[r], bytes_of_rank[r]));
        for (int l = 0; l < P; l++)
            for (int r = 0; r < P; r++) peers[l].p[r] = window[r];
        return GB_OK;
    }
    int alltoallv_u64(const u64 *const *send, const Row *soff, const Row *scnt, u64 *const *recv, const Row *roff,
                      const Row *rcnt) override
    {
        for (int s = 0; s < P; s++)
            for (int d = 0; d < P; d++) {
                if (scnt[s].v[d] != rcnt[d].v[s]) { set_error("sharded build: count mismatch %d -> %d", s, d); return GB_E_INVARIANT; }
                if (scnt[s].v[d]) GB_TRY(sg_copy(*ex[s], recv[d] + roff[d].v[s], send[s] + soff[s].v[d], (size_t)scnt[s].v[d] * 8));
            }
        return GB_OK;
    }
    int barrier() override
    {
        for (int r = 0; r < P; r++) GB_TRY(sg_sync(*ex[r]));
        return GB_OK;
    }
    int allgatherv_u64(u64 *const *buf, const u64 *off, const u64 *cnt) override
    {
        GB_TRY(barrier());
        for (int s = 0; s < P; s++)
            for (int d = 0; d < P; d++)
                if (s != d && cnt[s]) GB_TRY(sg_copy(*ex[d], buf[d] + off[s], buf[s] + off[s], (size_t)cnt[s] * 8));
        return GB_OK;
    }
    int allreduce_sum(void *, size_t, int) override { return barrier(); }
};

} // namespace sg
} // namespace gb
