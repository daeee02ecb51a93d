// graph.cu -- Graph.buildGraph and the MapGraph operators on one B200, as data-parallel sweeps plus
// pointer-jumping list ranking.  Replaces S/data/graph/Graph.scala:54-72,121-149,161-165,211-230,269-382
// (paths relative to /root/reference, S/ = src/main/scala/ru/ifmo/genome/).
//
// Vertex model.  Every stored (primary) k-mer gets a dense id vid; the ORIENTED vertex u = 2*vid + s is the
// stored key (s = 0) or its reverse complement (s = 1).  The reference materialises both strands
// (termKmers = set ++ set.map(revComplement), Graph.scala:330-333), so every sweep below runs over 2n
// oriented vertices.  For a palindromic key (even k) the s = 1 alias does not exist.
#include <time.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "scan.cuh"
#include "graph_types.cuh"

namespace gb {

static inline size_t base_words(int64_t n_bases) { return (size_t)((n_bases + 15) / 16) + 2; }

// ------------------------------------------------------------------------------------------------
// vertex table A: one u64 per oriented vertex
//   tag (bits 63..62): 0 interior, unresolved: ptr = an interior ancestor, dist = steps to it
//                      1 interior, resolved:   ptr = edge id, dist = rank from the head of the chain
//                      2 terminal:             low 32 bits = node index
//                      3 none (isolated (0,0) k-mer, palindrome alias)
//   ptr  (bits 61..31), dist (bits 30..0)
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long TAG_UNRES = 0ull, TAG_RES = 1ull, TAG_TERM = 2ull, TAG_NONE = 3ull;
__host__ __device__ __forceinline__ unsigned long long a_make(unsigned long long tag, unsigned long long ptr, unsigned long long dist)
{
    return (tag << 62) | (ptr << 31) | (dist & 0x7FFFFFFFull); // dist wraps only on perfect cycles, where it is unused
}
__host__ __device__ __forceinline__ unsigned int a_tag(unsigned long long a) { return (unsigned int)(a >> 62); }
__host__ __device__ __forceinline__ unsigned int a_ptr(unsigned long long a) { return (unsigned int)((a >> 31) & 0x7FFFFFFFu); }
__host__ __device__ __forceinline__ unsigned int a_dist(unsigned long long a) { return (unsigned int)(a & 0x7FFFFFFFu); }

__device__ __forceinline__ unsigned int comp_mask4(unsigned int m) // bit b <-> bit 3-b
{
    return ((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3);
}

// masks of an oriented vertex from the stored key's byte (low nibble out, high nibble in)
__device__ __forceinline__ void oriented_masks(unsigned int m8, unsigned int s, unsigned int *out, unsigned int *in)
{
    unsigned int o = m8 & 15, i = m8 >> 4;
    *out = s ? comp_mask4(i) : o;
    *in = s ? comp_mask4(o) : i;
}

// Graph.scala:323: terminal iff (in != 1 || out != 1) && (in != 0 || out != 0)
__device__ __forceinline__ unsigned int classify(unsigned int out, unsigned int in)
{
    int no = __popc(out), ni = __popc(in);
    if (no == 0 && ni == 0) return TAG_NONE;
    if (no == 1 && ni == 1) return TAG_UNRES;
    return TAG_TERM;
}

__device__ __forceinline__ unsigned long long oriented_kmer(unsigned long long key, unsigned int s, int k)
{
    return s ? revcomp(key, k) : key;
}

// ================================================================ build kernels

constexpr int TILE = 1024; // slots per CTA tile in the compaction passes (256 threads x 4)

template <bool V210>
__device__ __forceinline__ bool slot_is_vertex(const Table &table, int k, bool dual, unsigned long long key)
{
    return key != EMPTY_KEY && !is_secondary<V210>(table, k, dual, key);
}

template <bool V210>
__global__ void __launch_bounds__(256)
count_vertices_kernel(Table table, int k, bool dual, unsigned long long *tile_count)
{
    const unsigned long long n = table.cap;
    unsigned int c = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        unsigned long long i = (unsigned long long)blockIdx.x * TILE + j * 256 + threadIdx.x;
        if (i < n) c += slot_is_vertex<V210>(table, k, dual, load_key(table, i));
    }
    unsigned long long base = block_alloc(c, nullptr); // exclusive prefix inside the CTA, no counter
    __shared__ unsigned int s_total;
    if (threadIdx.x == 255) s_total = (unsigned int)base + c;
    __syncthreads();
    if (threadIdx.x == 0) tile_count[blockIdx.x] = s_total;
}

// assigns vid in slot order (deterministic), writes keys[vid] and the slot's vid field
template <bool V210>
__global__ void __launch_bounds__(256)
assign_vertices_kernel(Table table, int k, bool dual, const unsigned long long *tile_base, unsigned long long *keys)
{
    const unsigned long long n = table.cap;
    // thread t owns 4 CONSECUTIVE slots so that vids follow slot order
    unsigned long long i0 = (unsigned long long)blockIdx.x * TILE + threadIdx.x * 4;
    unsigned long long s[4];
    bool live[4];
    unsigned int c = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        live[j] = false;
        if (i0 + j < n) {
            s[j] = load_key(table, i0 + j);
            live[j] = slot_is_vertex<V210>(table, k, dual, s[j]);
        }
        c += live[j];
    }
    unsigned long long v = tile_base[blockIdx.x] + block_alloc(c, nullptr);
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (live[j]) {
            keys[v] = s[j];
            table.vid[i0 + j] = (unsigned int)v;
            v++;
        }
}

// incoming / outcoming (Graph.scala:272-282) of every stored key: 8 membership probes, either orientation.
// incoming / outcoming (Graph.scala:272-282) of every stored k-mer: 8 membership probes (contains, Graph.scala:270).
// mask8 = out | in << 4; nbr_out / nbr_in = the oriented neighbour when there is exactly one.
// `check_secondary`: keys[] came from a compacted array, not from the numbering pass, so an entry may be the SECONDARY
// orientation of a k-mer stored twice (hash tie / as-is inserts): it is no vertex (mask 0 = isolated, never referenced).
// ONE LANE PER (stored k-mer, neighbour query): 8 consecutive lanes share a key (lane j < 4: successor by base j, j >= 4:
// predecessor by base j - 4), so the warp's ballot of `found` holds the mask bytes of its 4 keys, and the unique neighbour of a
// side comes from the lane that found it by one shuffle.  Rounds 1-2 ran one THREAD per k-mer with its 8 probes unrolled (4 500
// SASS instructions, 64 registers, 40 % of the warps resident, 1.36 ms on C2); this form: 26 registers, 83 % resident, 8 x the
// independent probes in flight per resident thread, 0.60 ms (profiles/r2s_masks_timing.json, masks_r2r / masks_r2s ncu
// summaries).  With the fingerprint array (`fp`, one byte per slot, L2-resident) a negative probe -- 3 of 4 -- ends on one byte.
// rc of a neighbour is one shift of rc(x):
// rc(x.drop(1) :+ b) = comp(b) +: rc(x).take(k-1), rc(b +: x.take(k-1)) = rc(x).drop(1) :+ comp(b) -- so successor and
// predecessor lanes run the SAME instructions on swapped operands (A = (s.drop(1) :+ c), B = (comp(c) +: t.take(k-1)) with
// (s, t, c) = (x, rc x, b) or (rc x, x, comp b): {A, B} = {q, rc q} either way), and the warp stays converged up to the probe
// loop: the first per-lane form branched on `rc(x) < x` and on `j < 4` and ran its body twice per warp (ncu source page; 0.77 ms).
template <bool V210>
__global__ void __launch_bounds__(256)
masks_kernel(Table table, int k, bool dual, const unsigned long long *keys, unsigned long long lo,
                  unsigned long long n, bool check_secondary, const uint8_t *fp, uint8_t *mask8, unsigned int *nbr_out, unsigned int *nbr_in)
{
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int lane = threadIdx.x & 31, j = lane & 7, group = lane & 24;
    const bool live = (t >> 3) < n; // no early exit: every lane of the warp takes part in the ballot and the shuffles;
    const unsigned long long v = lo + (live ? (t >> 3) : 0); // the lanes behind the last key redo key `lo` and drop the result
    const unsigned long long x = keys[v];
    const unsigned long long rcx = revcomp(x, k);
    // a SECONDARY orientation (rc stored too and numerically smaller) is no vertex: mask 0, never referenced.  Only hash ties
    // (even k) and as-is inserts can store both orientations
    bool skip = !live;
    const bool suspect = check_secondary & (rcx < x) & (dual | (scala_hash<V210>(x) == scala_hash<V210>(rcx)));
    if (suspect) skip |= probe_find(table, rcx, fp) >= 0;
    __syncwarp();
    const bool succ = j < 4;
    const unsigned int c = succ ? (j & 3u) : ((j & 3u) ^ 3u);
    const unsigned long long A = kmer_append(succ ? x : rcx, k, c), B = kmer_prepend(succ ? rcx : x, k, c ^ 3u);
    const int hA = scala_hash<V210>(A), hB = scala_hash<V210>(B);
    bool found = false;
    unsigned int w = NONE32;
    if (!dual && hA != hB) { // one stored orientation possible: the canonical one (always, for odd k)
        const bool pickA = hA < hB;
        const long long i = skip ? -1 : probe_find(table, pickA ? A : B, fp);
        found = i >= 0;
        if (found) w = 2 * load_vid(table, (unsigned long long)i) + (unsigned int)(pickA != succ); // strand: the stored key is rc(q)
    } else if (!skip) {
        unsigned long long at = 0;
        unsigned int strand = 0;
        found = find_oriented<V210>(table, k, dual, succ ? A : B, &at, &strand, fp);
        if (found) w = 2 * load_vid(table, at) + strand;
    }
    const unsigned int byte = (__ballot_sync(0xFFFFFFFFu, found) >> group) & 0xFFu;
    const unsigned int out = byte & 0xFu, in = byte >> 4;
    // the neighbour reference matters only when its side has exactly one bit; any finder will do: take the highest
    const unsigned int so = __shfl_sync(0xFFFFFFFFu, w, out ? group + (31 - __clz(out)) : lane);
    const unsigned int si = __shfl_sync(0xFFFFFFFFu, w, in ? group + 4 + (31 - __clz(in)) : lane);
    if (live && j == 0) {
        mask8[v] = (uint8_t)byte;
        nbr_out[v] = out ? so : NONE32;
        nbr_in[v] = in ? si : NONE32;
    }
}

// gb_map_neighbour_masks: the same probes for arbitrary query k-mers
template <bool V210>
__global__ void query_masks_kernel(Table table, int k, bool dual, const unsigned long long *q, long long n, uint8_t *masks)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long x = q[i];
    unsigned int out = 0, in = 0;
#pragma unroll
    for (unsigned int b = 0; b < 4; b++) {
        unsigned long long at;
        unsigned int strand;
        if (find_oriented<V210>(table, k, dual, kmer_append(x, k, b), &at, &strand)) out |= 1u << b;
        if (find_oriented<V210>(table, k, dual, kmer_prepend(x, k, b), &at, &strand)) in |= 1u << b;
    }
    masks[i] = (uint8_t)(out | (in << 4));
}

struct BuildArrays {
    const unsigned long long *keys;
    const uint8_t *mask8;
    const unsigned int *nbr_out, *nbr_in;
    unsigned long long n; // stored (primary) keys; 2n oriented vertices
    int k;
};

__device__ __forceinline__ bool is_alias(const BuildArrays &B, unsigned int u)
{
    // (vid, 1) of a palindromic key is the same oriented k-mer as (vid, 0); only even k has palindromes
    if (!(u & 1) || (B.k & 1)) return false;
    unsigned long long x = B.keys[u >> 1];
    return revcomp(x, B.k) == x;
}
__device__ __forceinline__ unsigned int normalise(const BuildArrays &B, unsigned int u) { return is_alias(B, u) ? u ^ 1u : u; }

__device__ __forceinline__ unsigned int vertex_type(const BuildArrays &B, unsigned int u, unsigned int *out, unsigned int *in)
{
    oriented_masks(B.mask8[u >> 1], u & 1, out, in);
    return classify(*out, *in);
}
// unique predecessor / successor of an interior vertex
__device__ __forceinline__ unsigned int pred_of(const BuildArrays &B, unsigned int u)
{
    unsigned int p = (u & 1) ? (B.nbr_out[u >> 1] ^ 1u) : B.nbr_in[u >> 1];
    return normalise(B, p);
}
__device__ __forceinline__ unsigned int succ_of(const BuildArrays &B, unsigned int u)
{
    unsigned int p = (u & 1) ? (B.nbr_in[u >> 1] ^ 1u) : B.nbr_out[u >> 1];
    return normalise(B, p);
}

// per oriented vertex: is it a node, and how many edges start there
__global__ void classify_kernel(BuildArrays B, unsigned long long *node_cnt, unsigned long long *edge_cnt)
{
    unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= 2 * B.n) return;
    unsigned int out, in;
    unsigned int t = is_alias(B, (unsigned int)u) ? (unsigned int)TAG_NONE : vertex_type(B, (unsigned int)u, &out, &in);
    node_cnt[u] = t == TAG_TERM;
    edge_cnt[u] = t == TAG_TERM ? __popc(out) : 0;
}

// A for terminals / none / non-head interiors; node k-mers (addNode, Graph.scala:343-347)
__global__ void init_vertices_kernel(BuildArrays B, const unsigned long long *node_idx, unsigned long long *A,
                                     unsigned long long *node_kmer)
{
    unsigned long long uu = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (uu >= 2 * B.n) return;
    unsigned int u = (unsigned int)uu, out, in;
    if (is_alias(B, u)) { A[u] = a_make(TAG_NONE, 0, 0); return; }
    unsigned int t = vertex_type(B, u, &out, &in);
    if (t == TAG_TERM) {
        unsigned long long ni = node_idx[u];
        A[u] = a_make(TAG_TERM, 0, 0) | ni;
        node_kmer[ni] = oriented_kmer(B.keys[u >> 1], u & 1, B.k);
    } else if (t == TAG_NONE) {
        A[u] = a_make(TAG_NONE, 0, 0);
    } else {
        unsigned int p = pred_of(B, u), po, pi;
        if (vertex_type(B, p, &po, &pi) != TAG_TERM) A[u] = a_make(TAG_UNRES, p, 1);
        // heads (predecessor is a node) are written by start_edges_kernel, which knows the edge id
    }
}

// buildEdges (Graph.scala:349-365), first step of every edge: one thread per oriented vertex that is a node.
// Edge ids follow (node index, base) order; out-bases are visited in Base.fromInt order (351).
template <bool V210>
__global__ void start_edges_kernel(BuildArrays B, Table table, bool dual,
                                   const unsigned long long *node_idx, const unsigned long long *edge_idx,
                                   unsigned long long *A, unsigned int *edge_start, unsigned int *edge_end,
                                   unsigned long long *edge_len)
{
    unsigned long long uu = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (uu >= 2 * B.n) return;
    unsigned int u = (unsigned int)uu, out, in;
    if (is_alias(B, u) || vertex_type(B, u, &out, &in) != TAG_TERM) return;
    const unsigned long long x = oriented_kmer(B.keys[u >> 1], u & 1, B.k);
    unsigned long long e = edge_idx[u];
    const unsigned int me = (unsigned int)node_idx[u];
    for (unsigned int b = 0; b < 4; b++) {
        if (!(out & (1u << b))) continue;
        unsigned long long at = 0;
        unsigned int strand = 0;
        find_oriented<V210>(table, B.k, dual, kmer_append(x, B.k, b), &at, &strand); // present by construction
        unsigned int w = normalise(B, 2 * load_vid(table, at) + strand), wo, wi;
        edge_start[e] = me;
        if (vertex_type(B, w, &wo, &wi) == TAG_TERM) {
            edge_end[e] = (unsigned int)node_idx[w];
            edge_len[e] = 1;
        } else {
            A[w] = a_make(TAG_RES, e, 0); // w is the head of edge e's chain
            edge_end[e] = NONE32;
            edge_len[e] = 0;
        }
        e++;
    }
}

__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// in-place pointer jumping: A[u] = (ancestor, distance) is an invariant under any interleaving because every
// 64-bit entry is read and written whole.  Each thread jumps up to JUMPS times per launch.
constexpr int JUMPS = 4;
__global__ void jump_kernel(unsigned long long *A, unsigned long long n2, unsigned long long *unresolved)
{
    unsigned long long u = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int pending = 0;
    if (u < n2) {
        unsigned long long a = ld_cg_u64(A + u);
        if (a_tag(a) == TAG_UNRES) {
#pragma unroll 1
            for (int j = 0; j < JUMPS; j++) {
                unsigned long long ap = ld_cg_u64(A + a_ptr(a));
                a = a_make(a_tag(ap), a_ptr(ap), (unsigned long long)a_dist(a) + a_dist(ap));
                if (a_tag(a) != TAG_UNRES) break;
            }
            A[u] = a;
            pending = a_tag(a) == TAG_UNRES;
        }
    }
    pending = __reduce_add_sync(0xFFFFFFFFu, pending);
    if ((threadIdx.x & 31) == 0 && pending) atomicAdd(unresolved, (unsigned long long)pending);
}

// the last interior vertex of every chain closes its edge: end node and length (Graph.scala:363-364)
__global__ void close_edges_kernel(BuildArrays B, const unsigned long long *A, unsigned int *edge_end,
                                   unsigned long long *edge_len, unsigned long long *cycle_vertices)
{
    unsigned long long uu = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int cyc = 0;
    if (uu < 2 * B.n) {
        unsigned int u = (unsigned int)uu;
        unsigned long long a = A[u];
        if (a_tag(a) == TAG_RES) {
            unsigned long long as = A[succ_of(B, u)];
            if (a_tag(as) == TAG_TERM) {
                edge_end[a_ptr(a)] = (unsigned int)as;
                edge_len[a_ptr(a)] = (unsigned long long)a_dist(a) + 2;
            }
        } else if (a_tag(a) == TAG_UNRES) {
            cyc = 1; // perfect cycle: never reached from a node, ignored like Graph.scala:375
        }
    }
    cyc = __reduce_add_sync(0xFFFFFFFFu, cyc);
    if ((threadIdx.x & 31) == 0 && cyc) atomicAdd(cycle_vertices, (unsigned long long)cyc);
}

__device__ __forceinline__ void put_base(unsigned int *bases, unsigned long long pos, unsigned int b)
{
    atomicOr(&bases[pos >> 4], b << (2 * (unsigned int)(pos & 15)));
}

// Edge.seq: the base appended at every step of the walk (builder += base, Graph.scala:352,358)
__global__ void write_bases_kernel(BuildArrays B, const unsigned long long *A, const unsigned long long *edge_idx,
                                   const unsigned long long *edge_off, const unsigned long long *edge_len,
                                   unsigned int *bases)
{
    unsigned long long uu = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (uu >= 2 * B.n) return;
    unsigned int u = (unsigned int)uu;
    unsigned long long a = A[u];
    unsigned int out, in;
    if (a_tag(a) == TAG_RES) {
        // the vertex reached after dist+1 steps: its last base is seq[dist]
        unsigned long long x = B.keys[u >> 1];
        unsigned int last = (u & 1) ? 3u - (unsigned int)(x & 3) : (unsigned int)(x >> (2 * (B.k - 1))) & 3u;
        unsigned long long pos = edge_off[a_ptr(a)] + a_dist(a);
        put_base(bases, pos, last);
        if ((unsigned long long)a_dist(a) + 2 == edge_len[a_ptr(a)]) {
            // tail: the step into the end node appends the tail's single out-base
            oriented_masks(B.mask8[u >> 1], u & 1, &out, &in);
            put_base(bases, pos + 1, (unsigned int)__ffs((int)out) - 1);
        }
    } else if (a_tag(a) == TAG_TERM) {
        // node -> node edges of length 1
        oriented_masks(B.mask8[u >> 1], u & 1, &out, &in);
        unsigned long long e = edge_idx[u];
        for (unsigned int b = 0; b < 4; b++) {
            if (!(out & (1u << b))) continue;
            if (edge_len[e] == 1) put_base(bases, edge_off[e], b);
            e++;
        }
    }
}

// ================================================================ operators on the compacted graph

__global__ void degrees_kernel(const unsigned int *edge_start, const unsigned int *edge_end, unsigned long long n_edges,
                               unsigned int *indeg, unsigned int *outdeg, unsigned int *in_edge, unsigned int *out_edge)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    atomicAdd(&outdeg[edge_start[e]], 1u);
    atomicAdd(&indeg[edge_end[e]], 1u);
    if (out_edge) out_edge[edge_start[e]] = (unsigned int)e; // unique when the degree is 1
    if (in_edge) in_edge[edge_end[e]] = (unsigned int)e;
}

// components (Graph.scala:54-72): hook the larger root under the smaller, then compress
__global__ void cc_init_kernel(unsigned int *parent, unsigned long long n)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = (unsigned int)i;
}
__global__ void cc_hook_kernel(const unsigned int *edge_start, const unsigned int *edge_end, unsigned long long n_edges,
                               unsigned int *parent, unsigned int *changed)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    unsigned int a = parent[edge_start[e]], b = parent[edge_end[e]];
    if (a == b) return;
    unsigned int hi = max(a, b), lo = min(a, b);
    atomicMin(&parent[hi], lo);
    *changed = 1;
}
__global__ void cc_compress_kernel(unsigned int *parent, unsigned long long n)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned int p = parent[i];
    while (true) {
        unsigned int pp = ((volatile unsigned int *)parent)[p];
        if (pp == p) break;
        p = pp;
    }
    parent[i] = p;
}
__global__ void cc_flag_roots_kernel(const unsigned int *parent, unsigned long long n, unsigned long long *is_root)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_root[i] = parent[i] == i;
}
__global__ void cc_label_kernel(const unsigned int *parent, const unsigned long long *root_rank, unsigned long long n,
                                unsigned int *label)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) label[i] = (unsigned int)root_rank[parent[i]];
}
// retain(components.maxBy(_.size)) (GraphBuilder.scala:52-54): sizes, then the smallest node k-mer per component
__global__ void cc_size_kernel(const unsigned int *parent, const unsigned long long *node_kmer, unsigned long long n,
                               unsigned int *size, unsigned long long *min_kmer)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    atomicAdd(&size[parent[i]], 1u);
    atomicMin(&min_kmer[parent[i]], node_kmer[i]);
}
__global__ void cc_best_kernel(const unsigned int *parent, const unsigned int *size, const unsigned long long *min_kmer,
                               unsigned long long n, int pass, unsigned long long *best /* [0] size [1] kmer */)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || parent[i] != i) return;
    if (pass == 0) atomicMax(&best[0], (unsigned long long)size[i]);
    else if (size[i] == best[0]) atomicMin(&best[1], min_kmer[i]);
}
__global__ void cc_keep_kernel(const unsigned int *parent, const unsigned int *size, const unsigned long long *min_kmer,
                               unsigned long long n, const unsigned long long *best, unsigned long long *node_keep)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned int r = parent[i];
    node_keep[i] = size[r] == best[0] && min_kmer[r] == best[1];
}

// ---- the generic rewrite: every source edge goes to a destination edge (identified by the source edge that
// heads its chain) at a base offset `dist`, or is dropped; nodes are kept or dropped.
struct Rewrite {
    unsigned long long *node_keep = nullptr; // [N] 0/1 -> (after scan) new node index
    unsigned int *head = nullptr;            // [E] source edge id heading the chain, NONE32 = dropped
    unsigned long long *dist = nullptr;      // [E] base offset inside the destination edge
    uint8_t *tail = nullptr;                 // [E] 1 when this source edge is the last of its chain
};

__global__ void rw_flag_heads_kernel(const unsigned int *head, unsigned long long n_edges, unsigned long long *is_head)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges) is_head[e] = head[e] == e;
}
__global__ void rw_edges_kernel(const unsigned int *head, const unsigned long long *dist, const uint8_t *tail,
                                const unsigned long long *new_edge, const unsigned long long *new_node,
                                const unsigned int *edge_start, const unsigned int *edge_end, const unsigned long long *edge_off,
                                unsigned long long n_edges, unsigned int *ns, unsigned int *ne, unsigned long long *nlen)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges || head[e] == NONE32) return;
    unsigned long long d = new_edge[head[e]];
    if (head[e] == e) ns[d] = (unsigned int)new_node[edge_start[e]];
    if (tail[e]) {
        ne[d] = (unsigned int)new_node[edge_end[e]];
        nlen[d] = dist[e] + (edge_off[e + 1] - edge_off[e]);
    }
}
__global__ void rw_nodes_kernel(const unsigned long long *keep_flag, const unsigned long long *new_node,
                                const unsigned long long *node_kmer, unsigned long long n, unsigned long long *out)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && keep_flag[i]) out[new_node[i]] = node_kmer[i];
}
// one thread per 16-base source word: every piece of it that belongs to a surviving edge is OR-ed into place
__global__ void rw_bases_kernel(const unsigned int *src, unsigned long long n_bases, const unsigned long long *edge_off,
                                unsigned long long n_edges, const unsigned int *head, const unsigned long long *dist,
                                const unsigned long long *new_edge, const unsigned long long *new_off, unsigned int *dst)
{
    unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long p0 = w * 16;
    if (p0 >= n_bases) return;
    unsigned long long p1 = min(p0 + 16, n_bases);
    const unsigned int word = src[w];
    // last edge with edge_off <= p0 (edge_off ascending; zero-length edges do not exist)
    unsigned long long lo = 0, hi = n_edges;
    while (hi - lo > 1) {
        unsigned long long mid = (lo + hi) >> 1;
        if (edge_off[mid] <= p0) lo = mid; else hi = mid;
    }
    unsigned long long e = lo, p = p0;
    while (p < p1) {
        unsigned long long eend = edge_off[e + 1];
        unsigned long long q = min(eend, p1);
        if (head[e] != NONE32) {
            unsigned int cnt = (unsigned int)(q - p);
            unsigned long long bits = ((unsigned long long)word >> (2 * (unsigned int)(p - p0))) & ((1ull << (2 * cnt)) - 1);
            unsigned long long dpos = new_off[new_edge[head[e]]] + dist[e] + (p - edge_off[e]);
            unsigned int sh = 2 * (unsigned int)(dpos & 15);
            unsigned long long placed = bits << sh; // up to 32 + 30 bits
            atomicOr(&dst[dpos >> 4], (unsigned int)placed);
            if (placed >> 32) atomicOr(&dst[(dpos >> 4) + 1], (unsigned int)(placed >> 32));
        }
        p = q;
        e++;
    }
}

// simplifyGraph (Graph.scala:211-230).  D = nodes with exactly one in-edge and one out-edge.  A chain of edges
// through D nodes becomes one edge (seq = e1.seq ++ e2.seq ...); chains that never leave D are cycles and
// vanish with their nodes (the self-loop case e1 == e2 is the cycle of length one); (0,0) nodes are removed.
__global__ void simp_init_kernel(const unsigned int *edge_start, const unsigned long long *edge_off, unsigned long long n_edges,
                                 const unsigned int *indeg, const unsigned int *outdeg, const unsigned int *in_edge,
                                 unsigned int *ptr, unsigned long long *dist, uint8_t *resolved)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    unsigned int s = edge_start[e];
    if (indeg[s] == 1 && outdeg[s] == 1) {
        unsigned int p = in_edge[s]; // predecessor edge in the chain
        ptr[e] = p;
        dist[e] = edge_off[p + 1] - edge_off[p];
        resolved[e] = 0;
    } else {
        ptr[e] = (unsigned int)e; // head of its own chain
        dist[e] = 0;
        resolved[e] = 1;
    }
}
// synchronous (double-buffered) jump: ptr/dist/resolved are three arrays, so entries are not read atomically
__global__ void simp_jump_kernel(const unsigned int *ptr, const unsigned long long *dist, const uint8_t *resolved,
                                 unsigned int *ptr2, unsigned long long *dist2, uint8_t *resolved2,
                                 unsigned long long n_edges, unsigned int *pending)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    unsigned int p = ptr[e];
    unsigned long long d = dist[e];
    uint8_t r = resolved[e];
    if (!r) {
        d += dist[p];
        r = resolved[p];
        p = ptr[p];
        if (!r) *pending = 1;
    }
    ptr2[e] = p;
    dist2[e] = d;
    resolved2[e] = r;
}
__global__ void simp_finish_kernel(const unsigned int *edge_end, unsigned long long n_edges, const unsigned int *indeg,
                                   const unsigned int *outdeg, const unsigned int *ptr, const uint8_t *resolved,
                                   unsigned int *head, uint8_t *tail)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    head[e] = resolved[e] ? ptr[e] : NONE32; // unresolved after the bound = on a cycle inside D
    unsigned int t = edge_end[e];
    tail[e] = !(indeg[t] == 1 && outdeg[t] == 1);
}
__global__ void simp_nodes_kernel(const unsigned int *indeg, const unsigned int *outdeg, unsigned long long n,
                                  unsigned long long *node_keep)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool d = indeg[i] == 1 && outdeg[i] == 1, z = indeg[i] == 0 && outdeg[i] == 0;
    node_keep[i] = !(d || z);
}

__global__ void identity_rewrite_kernel(unsigned long long n_edges, unsigned int *head, unsigned long long *dist, uint8_t *tail)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    head[e] = (unsigned int)e;
    dist[e] = 0;
    tail[e] = 1;
}
__global__ void fill_u64_kernel(unsigned long long *p, unsigned long long n, unsigned long long v)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void drop_edges_kernel(const unsigned int *idx, long long n, unsigned long long n_edges, unsigned int *head)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && idx[i] < n_edges) head[idx[i]] = NONE32;
}
__global__ void retain_edges_kernel(const unsigned int *edge_start, const unsigned int *edge_end, unsigned long long n_edges,
                                    const unsigned long long *node_keep, unsigned int *head)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    if (!node_keep[edge_start[e]] || !node_keep[edge_end[e]]) head[e] = NONE32;
}

// out-edges of every node ordered by first base (= outEdgeIds key order of a freshly built graph)
__device__ __forceinline__ unsigned int first_base(const unsigned int *bases, unsigned long long off)
{
    return (bases[off >> 4] >> (2 * (unsigned int)(off & 15))) & 3u;
}
__global__ void out_table_kernel(const unsigned int *edge_start, const unsigned long long *edge_off, const unsigned int *bases,
                                 unsigned long long n_edges, unsigned int *out4 /* [N][4] */)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    out4[4ull * edge_start[e] + first_base(bases, edge_off[e])] = (unsigned int)e;
}
// removeBubbles (Graph.scala:125-149) + similar (121-123)
__global__ void bubbles_kernel(const unsigned int *out4, unsigned long long n_nodes, const unsigned int *edge_end,
                               const unsigned long long *edge_off, unsigned int *head)
{
    unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    unsigned int out[4], no = 0;
    for (int b = 0; b < 4; b++)
        if (out4[4 * v + b] != NONE32) out[no++] = out4[4 * v + b];
    bool rm[4] = { false, false, false, false };
    for (unsigned int i = 0; i < no; i++) {
        if (rm[i]) continue;
        for (unsigned int j = i + 1; j < no; j++) {
            long long la = (long long)(edge_off[out[i] + 1] - edge_off[out[i]]);
            long long lb = (long long)(edge_off[out[j] + 1] - edge_off[out[j]]);
            long long d = la > lb ? la - lb : lb - la;
            if (edge_end[out[i]] == edge_end[out[j]] && d * 5 < max(la, lb)) rm[j] = true;
        }
    }
    for (unsigned int i = 0; i < no; i++)
        if (rm[i]) head[out[i]] = NONE32;
}

// EXTENSION (no reference counterpart): one simultaneous sweep of dead-end tip removal.
// OUT candidate: u != v, len < max_len, v has in-degree 1 and out-degree 0; IN candidate: u has in-degree 0 and
// out-degree 1.  An OUT candidate dies iff u has another out-edge that is not an OUT candidate or is strictly
// longer; an IN candidate dies iff v has another in-edge that is not an IN candidate or is strictly longer.
__global__ void tips_mark_kernel(const unsigned int *edge_start, const unsigned int *edge_end, const unsigned long long *edge_off,
                                 unsigned long long n_edges, const unsigned int *indeg, const unsigned int *outdeg,
                                 long long max_len, uint8_t *oc, uint8_t *ic)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    unsigned int u = edge_start[e], v = edge_end[e];
    long long len = (long long)(edge_off[e + 1] - edge_off[e]);
    bool ok = u != v && len < max_len;
    oc[e] = ok && indeg[v] == 1 && outdeg[v] == 0;
    ic[e] = ok && indeg[u] == 0 && outdeg[u] == 1;
}
// per node: among out-edges, is there a non-candidate, and the two largest candidate lengths' maximum
__global__ void tips_node_kernel(const unsigned int *edge_start, const unsigned int *edge_end, const unsigned long long *edge_off,
                                 unsigned long long n_edges, const uint8_t *oc, const uint8_t *ic,
                                 unsigned int *out_noncand, unsigned long long *out_maxlen, unsigned int *out_maxcnt_dummy,
                                 unsigned int *in_noncand, unsigned long long *in_maxlen)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    unsigned long long len = edge_off[e + 1] - edge_off[e];
    unsigned int u = edge_start[e], v = edge_end[e];
    if (!oc[e]) atomicAdd(&out_noncand[u], 1u); else atomicMax(&out_maxlen[u], len);
    if (!ic[e]) atomicAdd(&in_noncand[v], 1u); else atomicMax(&in_maxlen[v], len);
}
__global__ void tips_kill_kernel(const unsigned int *edge_start, const unsigned int *edge_end, const unsigned long long *edge_off,
                                 unsigned long long n_edges, const uint8_t *oc, const uint8_t *ic,
                                 const unsigned int *out_noncand, const unsigned long long *out_maxlen,
                                 const unsigned int *in_noncand, const unsigned long long *in_maxlen,
                                 unsigned int *head, unsigned long long *removed)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    unsigned long long len = edge_off[e + 1] - edge_off[e];
    unsigned int u = edge_start[e], v = edge_end[e];
    bool kill = false;
    // "another out-edge that is not a candidate" (e itself is a candidate, so any non-candidate is another edge)
    if (oc[e] && (out_noncand[u] > 0 || out_maxlen[u] > len)) kill = true;
    if (ic[e] && (in_noncand[v] > 0 || in_maxlen[v] > len)) kill = true;
    if (kill) {
        head[e] = NONE32;
        atomicAdd(removed, 1ull);
    }
}

// invariants of GraphSimplifier.scala:159-170 that the array form can violate: edge ends in range, at most one
// out-edge per (node, first base), no empty edge
__global__ void check_kernel(const unsigned int *edge_start, const unsigned int *edge_end, const unsigned long long *edge_off,
                             const unsigned int *bases, unsigned long long n_edges, unsigned long long n_nodes,
                             unsigned int *seen4, unsigned int *bad)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    if (edge_start[e] >= n_nodes || edge_end[e] >= n_nodes || edge_off[e + 1] <= edge_off[e]) { atomicOr(bad, 1u); return; }
    if (atomicAdd(&seen4[4ull * edge_start[e] + first_base(bases, edge_off[e])], 1u) != 0) atomicOr(bad, 2u);
}

// ================================================================ host side

static int check_graph(gb_graph *h, Graph **g)
{
    if (!h) { set_error("null graph handle"); return GB_E_ARG; }
    *g = reinterpret_cast<Graph *>(h);
    GB_CUDA(cudaSetDevice((*g)->device));
    return GB_OK;
}

static void graph_free_arrays(Graph *g)
{
    g->node_kmer = nullptr; g->edge_start = g->edge_end = nullptr; g->edge_off = nullptr; g->bases = nullptr;
}
// graph arrays come from a store arena (no cudaMalloc in the steady state)
template <typename T>
static int store_alloc(Arena &a, T **p, size_t n)
{
    return a.alloc((void **)p, (n ? n : 1) * sizeof(T));
}

#define LAUNCH(kernel, n, ...)                                                                        \
    do {                                                                                              \
        unsigned long long _n = (unsigned long long)(n);                                              \
        if (_n) {                                                                                     \
            kernel<<<(unsigned int)((_n + 255) / 256), 256, 0, st>>>(__VA_ARGS__);                   \
            GB_LAUNCHED();                                                                            \
        }                                                                                             \
    } while (0)

static int read_u64(const unsigned long long *d, unsigned long long *h, int n, cudaStream_t st)
{
    GB_CUDA(cudaMemcpyAsync(h, d, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    return GB_OK;
}

template <bool V210>
static int build_graph(Map *m, Graph *g, const ShardPlan *sp)
{
    cudaStream_t st = m->stream;
    const int k = m->k;
    const unsigned long long bits = m->cap; // table capacity in slots (passed to the kernels as `cap`)
    const bool dual = m->noncanonical;
    const unsigned long long slots = bits;
    cudaEvent_t ev0 = m->ev0, ev1 = m->ev1;
    GB_CUDA(cudaEventRecord(ev0, st));
    const bool trace = g_tune.trace != 0;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_begin = now_ms();
    auto tick = [&](const char *what) {
        if (!trace) return;
        cudaStreamSynchronize(st);
        fprintf(stderr, "[graph] %-28s %9.3f ms\n", what, now_ms() - t_begin);
    };

    unsigned long long n = 0;
    Tmp<unsigned long long> total, keys_tmp;
    GB_TRY(total.alloc(4, st));
    const unsigned long long *keys_p = nullptr;
    const bool given = m->kept_valid;
    if (given) {
        // the filter (or the replica insert) left the stored keys as one array whose index is the slot's vid
        n = (unsigned long long)m->kept_n;
        keys_p = m->kept_keys;
    } else {
        // ---- dense vertex ids in slot order
        const unsigned long long tiles = (slots + TILE - 1) / TILE;
        Tmp<unsigned long long> tile_cnt;
        GB_TRY(tile_cnt.alloc(tiles, st));
        count_vertices_kernel<V210><<<(unsigned int)tiles, 256, 0, st>>>(m->view(), k, dual, tile_cnt.p);
        GB_LAUNCHED();
        GB_TRY(exclusive_scan_u64(tile_cnt.p, tiles, total.p, st));
        GB_TRY(read_u64(total.p, &n, 1, st));
        if (n >= (1ull << 30)) { set_error("%llu stored k-mers on one GPU: the graph build addresses at most 2^30", n); return GB_E_CAPACITY; }
        tick("counted vertices");
        GB_TRY(keys_tmp.alloc(n, st));
        assign_vertices_kernel<V210><<<(unsigned int)tiles, 256, 0, st>>>(m->view(), k, dual, tile_cnt.p, keys_tmp.p);
        GB_LAUNCHED();
        keys_p = keys_tmp.p;
    }
    if (n >= (1ull << 30)) { set_error("%llu stored k-mers on one GPU: the graph build addresses at most 2^30", n); return GB_E_CAPACITY; }
    g->stats[0] = (int64_t)n;
    struct { const unsigned long long *p; } keys{ keys_p };
    tick("assigned vertices");
    GB_CUDA(cudaEventRecord(m->fev[0], st));
    // ---- in/out masks and unique neighbours
    Tmp<uint8_t> mask8;
    Tmp<unsigned int> nbr_out, nbr_in;
    GB_TRY(mask8.alloc(n, st));
    GB_TRY(nbr_out.alloc(n, st));
    GB_TRY(nbr_in.alloc(n, st));
    {
        // sharded build: this rank probes only its own range of the (identical) key array, then the ranks exchange ranges
        const unsigned long long lo = sp ? sp->lo : 0, cnt = sp ? sp->hi - sp->lo : n;
        // the fingerprint array is built together with the vertex array (deleteAll / replica insert)
        const uint8_t *fp = given ? m->fp : nullptr;
        LAUNCH(masks_kernel<V210>, 8 * cnt, m->view(), k, dual, keys.p, lo, cnt, given, fp, mask8.p, nbr_out.p, nbr_in.p); // a lane per probe
        GB_CUDA(cudaEventRecord(m->fev[1], st));
        if (sp) {
            GB_CUDA(cudaStreamSynchronize(st));
            GB_TRY(sp->gather(sp->ctx, mask8.p, 1));
            GB_TRY(sp->gather(sp->ctx, nbr_out.p, 4));
            GB_TRY(sp->gather(sp->ctx, nbr_in.p, 4));
        }
    }
    BuildArrays B{ keys.p, mask8.p, nbr_out.p, nbr_in.p, n, k };
    const unsigned long long n2 = 2 * n;

    tick("masks");
    // ---- nodes and edge slots
    Tmp<unsigned long long> node_idx, edge_idx;
    GB_TRY(node_idx.alloc(n2, st));
    GB_TRY(edge_idx.alloc(n2, st));
    LAUNCH(classify_kernel, n2, B, node_idx.p, edge_idx.p);
    GB_TRY(exclusive_scan_u64(node_idx.p, n2, total.p + 0, st));
    GB_TRY(exclusive_scan_u64(edge_idx.p, n2, total.p + 1, st));
    unsigned long long tot[2] = { 0, 0 };
    GB_TRY(read_u64(total.p, tot, 2, st));
    const unsigned long long N = tot[0], E = tot[1];
    if (N >= 0xFFFFFFFFull || E >= (1ull << 31)) { set_error("graph too large: %llu nodes, %llu edges", N, E); return GB_E_CAPACITY; }

    tick("classified + scanned");
    g->cur = 0;
    g->store[0].reset();
    GB_TRY(store_alloc(g->store[0], &g->node_kmer, N));
    GB_TRY(store_alloc(g->store[0], &g->edge_start, E));
    GB_TRY(store_alloc(g->store[0], &g->edge_end, E));
    GB_TRY(store_alloc(g->store[0], &g->edge_off, E + 1));
    g->n_nodes = (int64_t)N;
    g->n_edges = (int64_t)E;

    tick("graph arrays allocated");
    Tmp<unsigned long long> A;
    GB_TRY(A.alloc(n2, st));
    LAUNCH(init_vertices_kernel, n2, B, node_idx.p, A.p, g->node_kmer);
    LAUNCH(start_edges_kernel<V210>, n2, B, m->view(), dual, node_idx.p, edge_idx.p, A.p, g->edge_start, g->edge_end,
           g->edge_off);
    node_idx.release();

    tick("vertices + edge starts");
    // ---- list ranking: rank of every interior vertex from the head of its chain + the chain's edge id
    int rounds = 0, bound = 2;
    while ((1ull << (bound - 2)) < n2 + 1) bound++; // ceil(log2) + slack; each launch makes >= 1 jump
    GB_CUDA(cudaEventRecord(m->fev[2], st));
    unsigned long long pending = 1;
    while (pending && rounds < bound) {
        GB_CUDA(cudaMemsetAsync(total.p + 2, 0, 8, st));
        LAUNCH(jump_kernel, n2, A.p, n2, total.p + 2);
        GB_TRY(read_u64(total.p + 2, &pending, 1, st));
        rounds++;
    }
    g->stats[1] = rounds;
    GB_CUDA(cudaEventRecord(m->fev[3], st));

    tick("list ranking");
    // ---- edge ends and lengths, base offsets, bases
    GB_CUDA(cudaMemsetAsync(total.p + 3, 0, 8, st));
    LAUNCH(close_edges_kernel, n2, B, A.p, g->edge_end, g->edge_off, total.p + 3);
    Tmp<unsigned long long> edge_len;
    GB_TRY(edge_len.alloc(E + 1, st));
    GB_CUDA(cudaMemcpyAsync(edge_len.p, g->edge_off, E * 8, cudaMemcpyDeviceToDevice, st));
    GB_CUDA(cudaMemsetAsync(edge_len.p + E, 0, 8, st));
    GB_CUDA(cudaMemsetAsync(g->edge_off + E, 0, 8, st));
    GB_TRY(exclusive_scan_u64(g->edge_off, E + 1, total.p + 0, st));
    unsigned long long fin[4];
    GB_TRY(read_u64(total.p, fin, 4, st));
    g->n_bases = (int64_t)fin[0];
    g->stats[2] = (int64_t)fin[3];
    tick("edges closed + scanned");
    GB_TRY(store_alloc(g->store[0], &g->bases, base_words(g->n_bases)));
    GB_CUDA(cudaMemsetAsync(g->bases, 0, base_words(g->n_bases) * 4, st));
    LAUNCH(write_bases_kernel, n2, B, A.p, edge_idx.p, g->edge_off, edge_len.p, g->bases);
    tick("bases written");
    GB_CUDA(cudaEventRecord(ev1, st));
    GB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    GB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    g->stats[3] = (int64_t)(ms * 1e6);
    GB_CUDA(cudaEventElapsedTime(&ms, m->fev[0], m->fev[1]));
    m->graph_ns[0] = (int64_t)(ms * 1e6); // masks_kernel (membership probes)
    GB_CUDA(cudaEventElapsedTime(&ms, m->fev[2], m->fev[3]));
    m->graph_ns[1] = (int64_t)(ms * 1e6); // list ranking: all jump_kernel launches
    m->graph_ns[2] = rounds;
    m->graph_ns[3] = (int64_t)n;
    return GB_OK;
}

// apply a Rewrite: builds the new arrays and swaps them into g.  rw.node_keep holds 0/1 flags on entry.
static int apply_rewrite(Graph *g, Rewrite &rw)
{
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
    Tmp<unsigned long long> new_node, new_edge, total;
    GB_TRY(new_node.alloc(N, st));
    GB_TRY(new_edge.alloc(E, st));
    GB_TRY(total.alloc(4, st));
    GB_CUDA(cudaMemcpyAsync(new_node.p, rw.node_keep, N * 8, cudaMemcpyDeviceToDevice, st));
    GB_TRY(exclusive_scan_u64(new_node.p, N, total.p + 0, st));
    LAUNCH(rw_flag_heads_kernel, E, rw.head, E, new_edge.p);
    GB_TRY(exclusive_scan_u64(new_edge.p, E, total.p + 1, st));
    unsigned long long tot[2];
    GB_TRY(read_u64(total.p, tot, 2, st));
    const unsigned long long N2 = tot[0], E2 = tot[1];

    struct { unsigned long long *node_kmer; unsigned int *edge_start, *edge_end; unsigned long long *edge_off; unsigned int *bases; } ng;
    Arena &dst = g->store[g->cur ^ 1];
    dst.reset();
    GB_TRY(store_alloc(dst, &ng.node_kmer, N2));
    GB_TRY(store_alloc(dst, &ng.edge_start, E2));
    GB_TRY(store_alloc(dst, &ng.edge_end, E2));
    GB_TRY(store_alloc(dst, &ng.edge_off, E2 + 1));
    GB_CUDA(cudaMemsetAsync(ng.edge_off, 0, (E2 + 1) * 8, st));
    LAUNCH(rw_nodes_kernel, N, rw.node_keep, new_node.p, g->node_kmer, N, ng.node_kmer);
    LAUNCH(rw_edges_kernel, E, rw.head, rw.dist, rw.tail, new_edge.p, new_node.p, g->edge_start, g->edge_end, g->edge_off, E,
           ng.edge_start, ng.edge_end, ng.edge_off);
    GB_TRY(exclusive_scan_u64(ng.edge_off, E2 + 1, total.p + 2, st));
    unsigned long long nb = 0;
    GB_TRY(read_u64(total.p + 2, &nb, 1, st));
    GB_TRY(store_alloc(dst, &ng.bases, base_words((int64_t)nb)));
    GB_CUDA(cudaMemsetAsync(ng.bases, 0, base_words((int64_t)nb) * 4, st));
    LAUNCH(rw_bases_kernel, (g->n_bases + 15) / 16, g->bases, (unsigned long long)g->n_bases, g->edge_off, E, rw.head, rw.dist,
           new_edge.p, ng.edge_off, ng.bases);
    GB_CUDA(cudaStreamSynchronize(st));
    g->cur ^= 1;
    g->node_kmer = ng.node_kmer; g->edge_start = ng.edge_start; g->edge_end = ng.edge_end;
    g->edge_off = ng.edge_off; g->bases = ng.bases;
    g->n_nodes = (int64_t)N2; g->n_edges = (int64_t)E2; g->n_bases = (int64_t)nb;
    return GB_OK;
}

struct RewriteBufs {
    Tmp<unsigned long long> node_keep, dist;
    Tmp<unsigned int> head;
    Tmp<uint8_t> tail;
    Rewrite rw;
    // identity: keep everything
    int init(Graph *g)
    {
        cudaStream_t st = g->stream;
        const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
        GB_TRY(node_keep.alloc(N, st));
        GB_TRY(dist.alloc(E, st));
        GB_TRY(head.alloc(E, st));
        GB_TRY(tail.alloc(E, st));
        LAUNCH(fill_u64_kernel, N, node_keep.p, N, 1ull);
        LAUNCH(identity_rewrite_kernel, E, E, head.p, dist.p, tail.p);
        rw.node_keep = node_keep.p; rw.head = head.p; rw.dist = dist.p; rw.tail = tail.p;
        return GB_OK;
    }
};

struct Degrees {
    Tmp<unsigned int> indeg, outdeg, in_edge, out_edge;
    int compute(Graph *g)
    {
        cudaStream_t st = g->stream;
        const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
        GB_TRY(indeg.alloc(N, st)); GB_TRY(outdeg.alloc(N, st)); GB_TRY(in_edge.alloc(N, st)); GB_TRY(out_edge.alloc(N, st));
        GB_TRY(indeg.zero(N)); GB_TRY(outdeg.zero(N)); GB_TRY(in_edge.fill_ff(N)); GB_TRY(out_edge.fill_ff(N));
        LAUNCH(degrees_kernel, E, g->edge_start, g->edge_end, E, indeg.p, outdeg.p, in_edge.p, out_edge.p);
        return GB_OK;
    }
};

// weakly connected components: parent[i] = root (smallest node index of the component)
static int components(Graph *g, Tmp<unsigned int> &parent)
{
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
    GB_TRY(parent.alloc(N, st));
    LAUNCH(cc_init_kernel, N, parent.p, N);
    Tmp<unsigned int> changed;
    GB_TRY(changed.alloc(1, st));
    for (int it = 0; it < 64 && E; it++) {
        GB_TRY(changed.zero(1));
        LAUNCH(cc_hook_kernel, E, g->edge_start, g->edge_end, E, parent.p, changed.p);
        LAUNCH(cc_compress_kernel, N, parent.p, N);
        unsigned int c = 0;
        GB_CUDA(cudaMemcpyAsync(&c, changed.p, 4, cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaStreamSynchronize(st));
        if (!c) break;
    }
    return GB_OK;
}

__global__ void drop_flagged_kernel(const unsigned int *flag, unsigned long long n_edges, unsigned int *head)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges && flag[e]) head[e] = NONE32;
}

int graph_remove_flagged(Graph *g, const unsigned int *d_flag)
{
    cudaStream_t st = g->stream;
    RewriteBufs rb;
    GB_TRY(rb.init(g));
    LAUNCH(drop_flagged_kernel, g->n_edges, d_flag, (unsigned long long)g->n_edges, rb.rw.head);
    return apply_rewrite(g, rb.rw);
}

int graph_build_sharded(gb_map *h, gb_graph **out, const ShardPlan *sp)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    ArenaScope scope(&m->arena);
    if (!out) { set_error("null out pointer"); return GB_E_ARG; }
    *out = nullptr;
    Graph *g = new Graph();
    g->k = m->k;
    g->device = m->device;
    int r = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) == cudaSuccess ? GB_OK : GB_E_CUDA;
    if (r == GB_OK) r = m->v210 ? build_graph<true>(m, g, sp) : build_graph<false>(m, g, sp);
    if (r != GB_OK) {
        cudaStreamSynchronize(m->stream);
        gb_graph_destroy(reinterpret_cast<gb_graph *>(g));
        return r;
    }
    *out = reinterpret_cast<gb_graph *>(g);
    return GB_OK;
}

} // namespace gb

using namespace gb;

extern "C" {

int gb_map_neighbour_masks(gb_map *h, const uint64_t *keys, int64_t n, uint8_t *masks)
{
    Map *m;
    GB_TRY(check_map(h, &m));
    ArenaScope scope(&m->arena);
    if (n < 0 || (n > 0 && (!keys || !masks))) { set_error("bad arguments"); return GB_E_ARG; }
    if (n == 0) return GB_OK;
    GB_TRY(check_keys(m, keys, n));
    cudaStream_t st = m->stream;
    DeviceBuf dk, dm;
    GB_TRY(dk.alloc((size_t)n * 8));
    GB_TRY(dm.alloc((size_t)n));
    GB_CUDA(cudaMemcpyAsync(dk.p, keys, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    if (m->v210) LAUNCH(query_masks_kernel<true>, n, m->view(), m->k, m->noncanonical, (const unsigned long long *)dk.p, n, (uint8_t *)dm.p);
    else LAUNCH(query_masks_kernel<false>, n, m->view(), m->k, m->noncanonical, (const unsigned long long *)dk.p, n, (uint8_t *)dm.p);
    GB_CUDA(cudaMemcpyAsync(masks, dm.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    return GB_OK;
}

int gb_graph_build(gb_map *h, gb_graph **out) { return gb::graph_build_sharded(h, out, nullptr); }

int gb_graph_destroy(gb_graph *h)
{
    if (!h) return GB_OK;
    Graph *g = reinterpret_cast<Graph *>(h);
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    graph_free_arrays(g);
    g->arena.destroy();
    g->store[0].destroy();
    g->store[1].destroy();
    if (g->stream) cudaStreamDestroy(g->stream);
    delete g;
    return GB_OK;
}

int gb_graph_counts(gb_graph *h, int64_t *n_nodes, int64_t *n_edges, int64_t *n_edge_bases)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    if (n_nodes) *n_nodes = g->n_nodes;
    if (n_edges) *n_edges = g->n_edges;
    if (n_edge_bases) *n_edge_bases = g->n_bases;
    return GB_OK;
}

int gb_graph_stats(gb_graph *h, int64_t stats[8])
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    if (!stats) { set_error("null argument"); return GB_E_ARG; }
    memcpy(stats, g->stats, sizeof g->stats);
    return GB_OK;
}

int gb_graph_export(gb_graph *h, uint64_t *node_kmers, uint32_t *edge_start, uint32_t *edge_end, uint64_t *edge_off,
                    uint8_t *edge_bases_2bit)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    const size_t N = (size_t)g->n_nodes, E = (size_t)g->n_edges;
    if (node_kmers && N) GB_CUDA(cudaMemcpyAsync(node_kmers, g->node_kmer, N * 8, cudaMemcpyDeviceToHost, st));
    if (edge_start && E) GB_CUDA(cudaMemcpyAsync(edge_start, g->edge_start, E * 4, cudaMemcpyDeviceToHost, st));
    if (edge_end && E) GB_CUDA(cudaMemcpyAsync(edge_end, g->edge_end, E * 4, cudaMemcpyDeviceToHost, st));
    if (edge_off) GB_CUDA(cudaMemcpyAsync(edge_off, g->edge_off, (E + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (edge_bases_2bit && g->n_bases)
        GB_CUDA(cudaMemcpyAsync(edge_bases_2bit, g->bases, (size_t)(g->n_bases + 3) / 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    return GB_OK;
}

int gb_graph_components(gb_graph *h, uint32_t *node_label, int64_t *n_components)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes;
    if (n_components) *n_components = 0;
    if (!N) return GB_OK;
    Tmp<unsigned int> parent, label;
    GB_TRY(components(g, parent));
    Tmp<unsigned long long> rank, total;
    GB_TRY(rank.alloc(N, st));
    GB_TRY(total.alloc(1, st));
    GB_TRY(label.alloc(N, st));
    LAUNCH(cc_flag_roots_kernel, N, parent.p, N, rank.p);
    GB_TRY(exclusive_scan_u64(rank.p, N, total.p, st));
    LAUNCH(cc_label_kernel, N, parent.p, rank.p, N, label.p);
    unsigned long long nc = 0;
    GB_TRY(read_u64(total.p, &nc, 1, st));
    if (n_components) *n_components = (int64_t)nc;
    if (node_label) {
        GB_CUDA(cudaMemcpyAsync(node_label, label.p, N * 4, cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaStreamSynchronize(st));
    }
    return GB_OK;
}

int gb_graph_retain_largest(gb_graph *h)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
    if (!N) return GB_OK;
    Tmp<unsigned int> parent, size;
    Tmp<unsigned long long> min_kmer, best;
    GB_TRY(components(g, parent));
    GB_TRY(size.alloc(N, st));
    GB_TRY(size.zero(N));
    GB_TRY(min_kmer.alloc(N, st));
    GB_TRY(min_kmer.fill_ff(N));
    GB_TRY(best.alloc(2, st));
    unsigned long long init[2] = { 0ull, ~0ull };
    GB_CUDA(cudaMemcpyAsync(best.p, init, 16, cudaMemcpyHostToDevice, st));
    LAUNCH(cc_size_kernel, N, parent.p, g->node_kmer, N, size.p, min_kmer.p);
    LAUNCH(cc_best_kernel, N, parent.p, size.p, min_kmer.p, N, 0, best.p);
    LAUNCH(cc_best_kernel, N, parent.p, size.p, min_kmer.p, N, 1, best.p);
    RewriteBufs rb;
    GB_TRY(rb.init(g));
    LAUNCH(cc_keep_kernel, N, parent.p, size.p, min_kmer.p, N, best.p, rb.rw.node_keep);
    LAUNCH(retain_edges_kernel, E, g->edge_start, g->edge_end, E, rb.rw.node_keep, rb.rw.head);
    return apply_rewrite(g, rb.rw);
}

__global__ void keep_from_bytes_kernel(const uint8_t *keep8, unsigned long long n, unsigned long long *node_keep)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) node_keep[i] = keep8[i] != 0;
}

// MapGraph.retain(nodesSet) (Graph.scala:161-165) for an arbitrary node set: nodes with node_keep[i] != 0 stay, and so do the
// edges whose start AND end stay
int gb_graph_retain(gb_graph *h, const uint8_t *node_keep)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
    if (!N) return GB_OK;
    if (!node_keep) { set_error("null argument"); return GB_E_ARG; }
    Tmp<uint8_t> d_keep;
    GB_TRY(d_keep.alloc(N, st));
    GB_CUDA(cudaMemcpyAsync(d_keep.p, node_keep, N, cudaMemcpyHostToDevice, st));
    RewriteBufs rb;
    GB_TRY(rb.init(g));
    LAUNCH(keep_from_bytes_kernel, N, d_keep.p, N, rb.rw.node_keep);
    LAUNCH(retain_edges_kernel, E, g->edge_start, g->edge_end, E, rb.rw.node_keep, rb.rw.head);
    return apply_rewrite(g, rb.rw);
}

int gb_graph_simplify(gb_graph *h)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
    if (!N) return GB_OK;
    Degrees dg;
    GB_TRY(dg.compute(g));
    RewriteBufs rb;
    GB_TRY(rb.init(g));
    Tmp<unsigned int> ptr[2], pending;
    Tmp<unsigned long long> dist[2];
    Tmp<uint8_t> res[2];
    for (int i = 0; i < 2; i++) { GB_TRY(ptr[i].alloc(E, st)); GB_TRY(dist[i].alloc(E, st)); GB_TRY(res[i].alloc(E, st)); }
    GB_TRY(pending.alloc(1, st));
    LAUNCH(simp_init_kernel, E, g->edge_start, g->edge_off, E, dg.indeg.p, dg.outdeg.p, dg.in_edge.p, ptr[0].p, dist[0].p, res[0].p);
    int cur = 0, bound = 2;
    while ((1ull << (bound - 2)) < E + 1) bound++;
    for (int it = 0; it < bound && E; it++) {
        GB_TRY(pending.zero(1));
        LAUNCH(simp_jump_kernel, E, ptr[cur].p, dist[cur].p, res[cur].p, ptr[cur ^ 1].p, dist[cur ^ 1].p, res[cur ^ 1].p, E, pending.p);
        cur ^= 1;
        unsigned int c = 0;
        GB_CUDA(cudaMemcpyAsync(&c, pending.p, 4, cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaStreamSynchronize(st));
        if (!c) break;
    }
    LAUNCH(simp_finish_kernel, E, g->edge_end, E, dg.indeg.p, dg.outdeg.p, ptr[cur].p, res[cur].p, rb.rw.head, rb.rw.tail);
    GB_CUDA(cudaMemcpyAsync(rb.rw.dist, dist[cur].p, E * 8, cudaMemcpyDeviceToDevice, st));
    LAUNCH(simp_nodes_kernel, N, dg.indeg.p, dg.outdeg.p, N, rb.rw.node_keep);
    return apply_rewrite(g, rb.rw);
}

int gb_graph_remove_bubbles(gb_graph *h)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
    if (!E) return GB_OK;
    Tmp<unsigned int> out4;
    GB_TRY(out4.alloc(4 * N, st));
    GB_TRY(out4.fill_ff(4 * N));
    LAUNCH(out_table_kernel, E, g->edge_start, g->edge_off, g->bases, E, out4.p);
    RewriteBufs rb;
    GB_TRY(rb.init(g));
    LAUNCH(bubbles_kernel, N, out4.p, N, g->edge_end, g->edge_off, rb.rw.head);
    return apply_rewrite(g, rb.rw);
}

int gb_graph_remove_edges(gb_graph *h, const uint32_t *edge_idx, int64_t n)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    if (n < 0 || (n > 0 && !edge_idx)) { set_error("bad arguments"); return GB_E_ARG; }
    if (n == 0) return GB_OK;
    for (int64_t i = 0; i < n; i++)
        if ((int64_t)edge_idx[i] >= g->n_edges) { set_error("edge index %u out of range", edge_idx[i]); return GB_E_ARG; }
    DeviceBuf di;
    GB_TRY(di.alloc((size_t)n * 4));
    GB_CUDA(cudaMemcpyAsync(di.p, edge_idx, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    RewriteBufs rb;
    GB_TRY(rb.init(g));
    LAUNCH(drop_edges_kernel, n, (const unsigned int *)di.p, (long long)n, (unsigned long long)g->n_edges, rb.rw.head);
    return apply_rewrite(g, rb.rw);
}

// ---- the fine-grained mutators of trait Graph (Graph.scala:31-36) in bulk
__global__ void edit_replace_kernel(const unsigned int *idx, const unsigned int *ns, const unsigned int *ne, long long n, unsigned int *edge_start,
                                    unsigned int *edge_end)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (ns[i] != NONE32) edge_start[idx[i]] = ns[i]; // replaceStart (197-202)
    if (ne[i] != NONE32) edge_end[idx[i]] = ne[i];   // replaceEnd (204-209)
}
__global__ void edit_append_offsets_kernel(const unsigned long long *add_off, long long n_add, unsigned long long base, unsigned long long *edge_off)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_add) edge_off[i] = base + add_off[i]; // edge_off points at the old end: entry 0 rewrites the old total with itself
}
__global__ void edit_append_bases_kernel(const uint8_t *codes, unsigned long long n, unsigned long long at, unsigned int *bases)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) put_base(bases, at + i, codes[i] & 3u);
}
__global__ void edit_drop_nodes_kernel(const unsigned int *idx, long long n, unsigned long long *node_keep)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) node_keep[idx[i]] = 0;
}
__global__ void edit_check_ends_kernel(const unsigned int *edge_start, const unsigned int *edge_end, unsigned long long n_edges,
                                       const unsigned long long *node_keep, unsigned int *bad)
{
    unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n_edges && (!node_keep[edge_start[e]] || !node_keep[edge_end[e]])) *bad = 1;
}

// replaceStart / replaceEnd, addNode / addEdge, removeNode (Graph.scala:172-209) in bulk, applied in that order.
// New nodes get the indices n_nodes, n_nodes + 1, ...; new edges n_edges, ... (they may start or end at new nodes, and
// replacements may name new nodes too).  removeNode only drops a node, like the reference's (185-187): removing one that an
// edge still starts or ends at is refused with GB_E_INVARIANT (the array form cannot hold a dangling end); node and edge
// indices are compacted by the removal like by every other operator.
int gb_graph_edit(gb_graph *h, int64_t n_replace, const uint32_t *edge_idx, const uint32_t *new_start, const uint32_t *new_end,
                  int64_t n_add_nodes, const uint64_t *add_node_kmers, int64_t n_add_edges, const uint32_t *add_start,
                  const uint32_t *add_end, const uint64_t *add_off, const uint8_t *add_bases, int64_t n_remove_nodes,
                  const uint32_t *remove_nodes)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    if (n_replace < 0 || n_add_nodes < 0 || n_add_edges < 0 || n_remove_nodes < 0 || (n_replace && (!edge_idx || !new_start || !new_end)) ||
        (n_add_nodes && !add_node_kmers) || (n_add_edges && (!add_start || !add_end || !add_off || !add_bases)) ||
        (n_remove_nodes && !remove_nodes)) { set_error("bad arguments"); return GB_E_ARG; }
    const int64_t N0 = g->n_nodes, E0 = g->n_edges, N1 = N0 + n_add_nodes, E1 = E0 + n_add_edges;
    const unsigned long long kmask = (1ull << (2 * g->k)) - 1;
    for (int64_t i = 0; i < n_add_nodes; i++)
        if (add_node_kmers[i] & ~kmask) { set_error("new node %lld: k-mer longer than k = %d", (long long)i, g->k); return GB_E_K_RANGE; }
    for (int64_t i = 0; i < n_replace; i++)
        if ((int64_t)edge_idx[i] >= E0 || (new_start[i] != NONE32 && (int64_t)new_start[i] >= N1) || (new_end[i] != NONE32 && (int64_t)new_end[i] >= N1)) {
            set_error("replacement %lld out of range", (long long)i);
            return GB_E_ARG;
        }
    unsigned long long add_bases_n = 0;
    for (int64_t i = 0; i < n_add_edges; i++) {
        if ((int64_t)add_start[i] >= N1 || (int64_t)add_end[i] >= N1 || add_off[i + 1] <= add_off[i] || (i == 0 && add_off[0] != 0)) {
            set_error("new edge %lld: end out of range or empty sequence", (long long)i);
            return GB_E_ARG;
        }
        add_bases_n = add_off[i + 1];
    }
    for (int64_t i = 0; i < n_remove_nodes; i++)
        if ((int64_t)remove_nodes[i] >= N1) { set_error("node %u out of range", remove_nodes[i]); return GB_E_ARG; }
    if (N1 >= 0xFFFFFFFFll || E1 >= (1ll << 31)) { set_error("graph too large"); return GB_E_CAPACITY; }

    // 1. replaceStart / replaceEnd in place (new node indices are valid once step 2 has run; nothing reads them before)
    if (n_replace) {
        Tmp<unsigned int> di, ds, de;
        GB_TRY(di.alloc((size_t)n_replace, st)); GB_TRY(ds.alloc((size_t)n_replace, st)); GB_TRY(de.alloc((size_t)n_replace, st));
        GB_CUDA(cudaMemcpyAsync(di.p, edge_idx, (size_t)n_replace * 4, cudaMemcpyHostToDevice, st));
        GB_CUDA(cudaMemcpyAsync(ds.p, new_start, (size_t)n_replace * 4, cudaMemcpyHostToDevice, st));
        GB_CUDA(cudaMemcpyAsync(de.p, new_end, (size_t)n_replace * 4, cudaMemcpyHostToDevice, st));
        LAUNCH(edit_replace_kernel, n_replace, di.p, ds.p, de.p, (long long)n_replace, g->edge_start, g->edge_end);
    }
    // 2. addNode / addEdge: the arrays grow into the other store
    if (n_add_nodes || n_add_edges) {
        const unsigned long long nb0 = (unsigned long long)g->n_bases, nb1 = nb0 + add_bases_n;
        struct { unsigned long long *node_kmer; unsigned int *edge_start, *edge_end; unsigned long long *edge_off; unsigned int *bases; } ng;
        Arena &dst = g->store[g->cur ^ 1];
        dst.reset();
        GB_TRY(store_alloc(dst, &ng.node_kmer, (size_t)N1));
        GB_TRY(store_alloc(dst, &ng.edge_start, (size_t)E1));
        GB_TRY(store_alloc(dst, &ng.edge_end, (size_t)E1));
        GB_TRY(store_alloc(dst, &ng.edge_off, (size_t)E1 + 1));
        GB_TRY(store_alloc(dst, &ng.bases, base_words((int64_t)nb1)));
        GB_CUDA(cudaMemsetAsync(ng.bases, 0, base_words((int64_t)nb1) * 4, st));
        if (N0) GB_CUDA(cudaMemcpyAsync(ng.node_kmer, g->node_kmer, (size_t)N0 * 8, cudaMemcpyDeviceToDevice, st));
        if (E0) {
            GB_CUDA(cudaMemcpyAsync(ng.edge_start, g->edge_start, (size_t)E0 * 4, cudaMemcpyDeviceToDevice, st));
            GB_CUDA(cudaMemcpyAsync(ng.edge_end, g->edge_end, (size_t)E0 * 4, cudaMemcpyDeviceToDevice, st));
        }
        GB_CUDA(cudaMemcpyAsync(ng.edge_off, g->edge_off, (size_t)(E0 + 1) * 8, cudaMemcpyDeviceToDevice, st));
        if (nb0) GB_CUDA(cudaMemcpyAsync(ng.bases, g->bases, (size_t)((nb0 + 15) / 16) * 4, cudaMemcpyDeviceToDevice, st));
        if (n_add_nodes) GB_CUDA(cudaMemcpyAsync(ng.node_kmer + N0, add_node_kmers, (size_t)n_add_nodes * 8, cudaMemcpyHostToDevice, st));
        if (n_add_edges) {
            GB_CUDA(cudaMemcpyAsync(ng.edge_start + E0, add_start, (size_t)n_add_edges * 4, cudaMemcpyHostToDevice, st));
            GB_CUDA(cudaMemcpyAsync(ng.edge_end + E0, add_end, (size_t)n_add_edges * 4, cudaMemcpyHostToDevice, st));
            Tmp<unsigned long long> d_off;
            Tmp<uint8_t> d_codes;
            GB_TRY(d_off.alloc((size_t)n_add_edges + 1, st));
            GB_TRY(d_codes.alloc((size_t)add_bases_n, st));
            GB_CUDA(cudaMemcpyAsync(d_off.p, add_off, (size_t)(n_add_edges + 1) * 8, cudaMemcpyHostToDevice, st));
            GB_CUDA(cudaMemcpyAsync(d_codes.p, add_bases, (size_t)add_bases_n, cudaMemcpyHostToDevice, st));
            LAUNCH(edit_append_offsets_kernel, n_add_edges + 1, d_off.p, (long long)n_add_edges, nb0, ng.edge_off + E0);
            LAUNCH(edit_append_bases_kernel, add_bases_n, d_codes.p, add_bases_n, nb0, ng.bases);
        }
        GB_CUDA(cudaStreamSynchronize(st));
        g->cur ^= 1;
        g->node_kmer = ng.node_kmer; g->edge_start = ng.edge_start; g->edge_end = ng.edge_end;
        g->edge_off = ng.edge_off; g->bases = ng.bases;
        g->n_nodes = N1; g->n_edges = E1; g->n_bases = (int64_t)nb1;
    }
    // 3. removeNode
    if (n_remove_nodes) {
        Tmp<unsigned int> dr, bad;
        GB_TRY(dr.alloc((size_t)n_remove_nodes, st));
        GB_TRY(bad.alloc(1, st));
        GB_TRY(bad.zero(1));
        GB_CUDA(cudaMemcpyAsync(dr.p, remove_nodes, (size_t)n_remove_nodes * 4, cudaMemcpyHostToDevice, st));
        RewriteBufs rb;
        GB_TRY(rb.init(g));
        LAUNCH(edit_drop_nodes_kernel, n_remove_nodes, dr.p, (long long)n_remove_nodes, rb.rw.node_keep);
        LAUNCH(edit_check_ends_kernel, g->n_edges, g->edge_start, g->edge_end, (unsigned long long)g->n_edges, rb.rw.node_keep, bad.p);
        unsigned int b = 0;
        GB_CUDA(cudaMemcpyAsync(&b, bad.p, 4, cudaMemcpyDeviceToHost, st));
        GB_CUDA(cudaStreamSynchronize(st));
        if (b) { set_error("removeNode: an edge still starts or ends at a removed node"); return GB_E_INVARIANT; }
        return apply_rewrite(g, rb.rw);
    }
    GB_CUDA(cudaStreamSynchronize(st));
    return GB_OK;
}

int gb_graph_clip_tips(gb_graph *h, int64_t max_len, int64_t *removed)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
    if (removed) *removed = 0;
    if (!E) return GB_OK;
    Degrees dg;
    GB_TRY(dg.compute(g));
    Tmp<uint8_t> oc, ic;
    Tmp<unsigned int> out_nc, in_nc;
    Tmp<unsigned long long> out_ml, in_ml, cnt;
    GB_TRY(oc.alloc(E, st)); GB_TRY(ic.alloc(E, st));
    GB_TRY(out_nc.alloc(N, st)); GB_TRY(in_nc.alloc(N, st)); GB_TRY(out_ml.alloc(N, st)); GB_TRY(in_ml.alloc(N, st));
    GB_TRY(out_nc.zero(N)); GB_TRY(in_nc.zero(N)); GB_TRY(out_ml.zero(N)); GB_TRY(in_ml.zero(N));
    GB_TRY(cnt.alloc(1, st)); GB_TRY(cnt.zero(1));
    LAUNCH(tips_mark_kernel, E, g->edge_start, g->edge_end, g->edge_off, E, dg.indeg.p, dg.outdeg.p, (long long)max_len, oc.p, ic.p);
    LAUNCH(tips_node_kernel, E, g->edge_start, g->edge_end, g->edge_off, E, oc.p, ic.p, out_nc.p, out_ml.p, nullptr, in_nc.p, in_ml.p);
    RewriteBufs rb;
    GB_TRY(rb.init(g));
    LAUNCH(tips_kill_kernel, E, g->edge_start, g->edge_end, g->edge_off, E, oc.p, ic.p, out_nc.p, out_ml.p, in_nc.p, in_ml.p,
           rb.rw.head, cnt.p);
    unsigned long long c = 0;
    GB_TRY(read_u64(cnt.p, &c, 1, st));
    if (removed) *removed = (int64_t)c;
    return apply_rewrite(g, rb.rw);
}

int gb_graph_check(gb_graph *h)
{
    Graph *g;
    GB_TRY(check_graph(h, &g));
    ArenaScope scope(&g->arena);
    cudaStream_t st = g->stream;
    const unsigned long long N = (unsigned long long)g->n_nodes, E = (unsigned long long)g->n_edges;
    if (!E) return GB_OK;
    Tmp<unsigned int> seen4, bad;
    GB_TRY(seen4.alloc(4 * N, st));
    GB_TRY(seen4.zero(4 * N));
    GB_TRY(bad.alloc(1, st));
    GB_TRY(bad.zero(1));
    LAUNCH(check_kernel, E, g->edge_start, g->edge_end, g->edge_off, g->bases, E, N, seen4.p, bad.p);
    unsigned int b = 0;
    GB_CUDA(cudaMemcpyAsync(&b, bad.p, 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    if (b) { set_error("graph invariant broken (mask %u): GraphSimplifier.scala:159-170", b); return GB_E_INVARIANT; }
    return GB_OK;
}

} // extern "C"
