// graph_types.cuh -- the device-resident MapGraph (S/data/graph/Graph.scala:152-262, relative to /root/reference)
#pragma once
#include "common.cuh"

namespace gb {

struct Graph {
    int k = 0, device = 0;
    cudaStream_t stream = nullptr;
    int64_t n_nodes = 0, n_edges = 0, n_bases = 0;
    unsigned long long *node_kmer = nullptr; // [n_nodes]
    unsigned int *edge_start = nullptr;      // [n_edges] node index
    unsigned int *edge_end = nullptr;        // [n_edges]
    unsigned long long *edge_off = nullptr;  // [n_edges + 1], in bases, ascending
    unsigned int *bases = nullptr;           // 2-bit stream, 16 bases per word, base j at bits 2(j%16) of word j/16
    Arena arena;    // scratch of the operators
    Arena store[2]; // the arrays above live in store[cur]; a rewrite builds the new ones in store[cur ^ 1]
    int cur = 0;
    int64_t stats[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; // [0] kept k-mers [1] jump rounds [2] cycle vertices dropped [3] build ns
};

// Graph.getGraphMap entries into DEVICE arrays of n_nodes + n_bases - n_edges elements, on g->stream (graphmap.cu)
int graph_positions_device(Graph *g, unsigned long long *d_kmer, unsigned int *d_id, unsigned int *d_dist);
// removeEdge for every edge e with d_flag[e] != 0 (graph.cu); the caller holds an ArenaScope on g->arena
int graph_remove_flagged(Graph *g, const unsigned int *d_flag);

} // namespace gb
