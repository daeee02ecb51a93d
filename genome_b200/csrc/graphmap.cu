// graphmap.cu -- Graph.getGraphMap (S/data/graph/Graph.scala:90-119, relative to /root/reference) as a bulk export:
// the (k-mer, GraphPosition) pairs the reference feeds to putNew, nodes first, then the interior k-mers of every edge.
// SURVEY 8(f) row 3.  Parity with the oracle's go_graph_map: tests/test_graphmap_gpu.py.  The DNAMap[GraphPosition] the
// reference builds from these pairs (putNew / getAll) is the caller's: every oriented k-mer occurs once, so an
// ArrayDNAMap with value = entry index serves as that multimap.
#include "common.cuh"
#include "graph_types.cuh"

namespace gb {

// bases [pos, pos + cnt) of the 2-bit stream (cnt <= 31) as a k-mer fragment: base j at bits 2j
__device__ __forceinline__ unsigned long long stream_bases(const unsigned int *bases, unsigned long long pos, unsigned int cnt)
{
    if (cnt == 0) return 0;
    const unsigned long long w = pos >> 4;
    const unsigned int sh = 2 * (unsigned int)(pos & 15);
    const unsigned int a = bases[w], b = bases[w + 1], c = bases[w + 2]; // the stream has two words of slack
    const unsigned int lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return v & ((1ull << (2 * cnt)) - 1);
}

__global__ void node_positions_kernel(const unsigned long long *node_kmer, unsigned long long n_nodes, unsigned long long *kmer,
                                      unsigned int *id, unsigned int *dist)
{
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    kmer[i] = node_kmer[i];
    id[i] = (unsigned int)i; // NodeGraphPosition(node)
    dist[i] = 0;
}

// one thread per edge base q: d = q - off[e] is the distance from the start node; d >= 1 yields EdgeGraphPosition(e, d) with
// the k-mer reached after d steps = bases [d, d + k) of start_kmer ++ seq (Graph.scala:105-112)
__global__ void edge_positions_kernel(const unsigned long long *node_kmer, const unsigned int *edge_start, const unsigned long long *edge_off,
                                      const unsigned int *bases, unsigned long long n_edges, unsigned long long n_bases,
                                      unsigned long long n_nodes, int k, unsigned long long *kmer, unsigned int *id, unsigned int *dist)
{
    unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_bases) return;
    unsigned long long lo = 0, hi = n_edges; // last edge with edge_off <= q
    while (hi - lo > 1) {
        unsigned long long mid = (lo + hi) >> 1;
        if (edge_off[mid] <= q) lo = mid; else hi = mid;
    }
    const unsigned long long e = lo, d = q - edge_off[e];
    if (d == 0) return;
    unsigned long long x;
    if (d < (unsigned long long)k) {
        const unsigned long long head = node_kmer[edge_start[e]] >> (2 * d);                  // k - d bases of the start node
        const unsigned long long tail = stream_bases(bases, edge_off[e], (unsigned int)d);    // the first d appended bases
        x = head | (tail << (2 * (k - (unsigned int)d)));
    } else {
        x = stream_bases(bases, edge_off[e] + d - k, (unsigned int)k);
    }
    const unsigned long long out = n_nodes + (edge_off[e] - e) + (d - 1); // edges before e contribute (len - 1) entries each
    kmer[out] = x;
    id[out] = (unsigned int)e;
    dist[out] = (unsigned int)d;
}

int graph_positions_device(Graph *g, unsigned long long *d_kmer, unsigned int *d_id, unsigned int *d_dist)
{
    cudaStream_t st = g->stream;
    if (g->n_nodes) {
        node_positions_kernel<<<(unsigned int)((g->n_nodes + 255) / 256), 256, 0, st>>>(g->node_kmer, (unsigned long long)g->n_nodes,
                                                                                      d_kmer, d_id, d_dist);
        GB_LAUNCHED();
    }
    if (g->n_bases) {
        edge_positions_kernel<<<(unsigned int)((g->n_bases + 255) / 256), 256, 0, st>>>(
            g->node_kmer, g->edge_start, g->edge_off, g->bases, (unsigned long long)g->n_edges, (unsigned long long)g->n_bases,
            (unsigned long long)g->n_nodes, g->k, d_kmer, d_id, d_dist);
        GB_LAUNCHED();
    }
    return GB_OK;
}

} // namespace gb

using namespace gb;

extern "C" int gb_graph_positions(gb_graph *h, uint64_t *kmers, uint32_t *ids, uint32_t *dists, int64_t cap, int64_t *n_out)
{
    if (!h) { set_error("null graph handle"); return GB_E_ARG; }
    Graph *g = reinterpret_cast<Graph *>(h);
    GB_CUDA(cudaSetDevice(g->device));
    ArenaScope scope(&g->arena);
    const int64_t n = g->n_nodes + g->n_bases - g->n_edges; // `total`, Graph.scala:97
    if (n_out) *n_out = n;
    if (!kmers && !ids && !dists) return GB_OK;
    if (cap < n) { set_error("positions buffer too small: %lld < %lld", (long long)cap, (long long)n); return GB_E_CAPACITY; }
    if (n == 0) return GB_OK;
    cudaStream_t st = g->stream;
    DeviceBuf dk, di, dd;
    GB_TRY(dk.alloc((size_t)n * 8));
    GB_TRY(di.alloc((size_t)n * 4));
    GB_TRY(dd.alloc((size_t)n * 4));
    GB_TRY(graph_positions_device(g, (unsigned long long *)dk.p, (unsigned int *)di.p, (unsigned int *)dd.p));
    if (kmers) GB_CUDA(cudaMemcpyAsync(kmers, dk.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    if (ids) GB_CUDA(cudaMemcpyAsync(ids, di.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (dists) GB_CUDA(cudaMemcpyAsync(dists, dd.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    GB_CUDA(cudaStreamSynchronize(st));
    return GB_OK;
}
