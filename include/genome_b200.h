/*
 * genome_b200.h -- C ABI of libgenome_b200.so: the B200 (sm_100a) replacement for winger/genome's
 * k-mer -> de Bruijn graph hot path.  Plain pointers and sizes only; every entry point names the
 * reference interface it replaces (paths relative to /root/reference, S/ = src/main/scala/ru/ifmo/genome/).
 *
 * Conventions
 *   - every function returns 0 (GB_OK) or a negative GB_E_* code; gb_last_error() gives the thread-local text.
 *   - handles own all device memory; host buffers are caller-allocated; exports use size-then-fill.
 *   - calls on one handle must be serialised by the caller; different handles are independent.
 *   - a k-mer is one uint64: base i at bits 2i..2i+1, code A0 G1 C2 T3 (S/dna/DNASeq.scala:80-85,
 *     S/dna/Base.scala:13-16).  Supported k: 1..31 (k = 32 is broken in the reference itself).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with GB_E_CUDA.
 */
#ifndef GENOME_B200_H
#define GENOME_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_OK            0
#define GB_E_ARG        -1   /* null pointer, negative size, truncated .bin stream */
#define GB_E_K_RANGE    -2   /* k outside 1..31 (assert key.length == k, S/ds/ArrayDNAMap.scala:182-206) */
#define GB_E_OOM        -3
#define GB_E_CUDA       -4
#define GB_E_NCCL       -5
#define GB_E_CAPACITY   -6   /* more than 2^31-1 kept k-mers on one GPU, or a caller buffer too small */
#define GB_E_INVARIANT  -7   /* a graph invariant failed (assert at S/data/graph/Graph.scala:357) */
#define GB_E_STATE      -8   /* call not valid in the handle's current state */

/* gb_map_create flags */
#define GB_FLAG_HASH_SCALA_210  1u  /* Long.## of scala >= 2.10 instead of the pinned 2.9.1 */

typedef struct gb_map gb_map;     /* DNAMap[Int]: one shard (one GPU) of the k-mer count table */
typedef struct gb_graph gb_graph; /* Graph / MapGraph */
typedef struct gb_comm gb_comm;   /* the set of shards of a PartitionedDNAMap: one rank per GPU */
typedef struct gb_graph_map gb_graph_map; /* DNAMap[GraphPosition] as built by Graph.getGraphMap */

const char *gb_last_error(void);
/* diagnostics: kernels launched by this library in this process so far */
long long gb_launch_count(void);
/* diagnostics: empirical random-access ceiling R_gups of SURVEY 8(d): a bare kernel doing ONE random 64-bit atomicAdd per
 * element over an array of exactly table_bytes (no hashing, no probing, no preceding load): ns per pass of n_updates,
 * averaged over iters passes */
int gb_bench_random_atomics(int device, size_t table_bytes, int64_t n_updates, int iters, int64_t *ns_per_iter);
/* diagnostics: update(key, 1, _ + 1) with the table slice in SHARED memory (one CTA per bucket of keys_per_bucket synthetic keys with
 * C2's multiplicity profile: slice of 2^slots_log2 slots initialised on chip, shared-memory atomics, slice streamed out once): ns per
 * pass over n_buckets buckets; *distinct_keys (optional) = keys claimed per pass.  The design question of DESIGN.md 3.1. */
int gb_bench_smem_upsert(int device, int slots_log2, int64_t keys_per_bucket, int64_t n_buckets, int ctas_per_sm, int iters,
                         int64_t *ns_per_iter, int64_t *distinct_keys);
/* diagnostics: the request rate of the upsert's access pattern with hashing and probing stripped away: n_updates random 16-byte
 * slots of a region of region_bytes (choose it to fit L2 or not); mode = base + 100 * log2(updates in flight per thread; 0 = 4)
 * + 1000 * P (P > 0: persistent grid of P CTAs per SM, software-pipelined); base 1 = 8-byte load, 2 = 4-byte red.add, 3 = load
 * then a dependent red (an existing key), 4 = load, 64-bit CAS, red (a new key), 5 = load and an independent red, 6 = load then a
 * dependent red on another slot, 7 = atom.add with a return value, 8 = 16-byte load then red, 9 = load then a plain store,
 * 10 = load from one half / red into the other, 11 / 12 = loads and reds issued by different SMs / warps, 13 = separate key and
 * count arrays, 14 = keys and counts in different sectors of one 128-byte line, 15 / 16 = 13 / 14 with a CAS on the key.
 * ns per pass. */
int gb_bench_l2_requests(int device, size_t region_bytes, int64_t n_updates, int mode, int iters, int64_t *ns_per_iter);
/* tuning and test hooks (process-wide; not part of the reference surface).  The library reads no environment variable on
 * its data path: every default is a measured choice, and the parity tests use these keys to force a path that
 * small inputs would not take by themselves.  Keys: insert_path (0 auto, 1 direct, 2 L2-blocked), single_pass,
 * single_pass_min, slice_bits, batches, h2d_chunks, route (0 auto, 1 one level, 2 two levels), a2a (0 peer stores, 1 NCCL staged),
 * pgraph_sharded (1 = Graph.buildGraph over shards without a replica, the default; 0 = replicated after an all-gather), trace.
 * *previous (optional) receives the old value. */
int gb_tune(const char *name, int64_t value, int64_t *previous);
int gb_tune_get(const char *name, int64_t *value);
int gb_version(void);
int gb_device_count(int *n);

/* pinned host staging memory for the .bin stream (optional; any host pointer is accepted by the calls below) */
int gb_host_alloc(size_t n_bytes, void **ptr);
int gb_host_free(void *ptr);

/* ------------------------------------------------------------------------------------------------
 * DNAMap[Int]  (trait DNAMap, S/ds/ArrayDNAMap.scala:49-60; ArrayDNAMap 62-243)
 * ---------------------------------------------------------------------------------------------- */

/* new ArrayDNAMap[Int](k) (ArrayDNAMap.scala:62-72).  min_capacity = expected number of distinct keys
 * (0 = unknown: the table starts small and grows by rehashing, like rescale() 217-230). */
int gb_map_create(int k, int64_t min_capacity, int device, uint32_t flags, gb_map **out);
int gb_map_destroy(gb_map *m);

/* FreqFilter.add over a `.bin` stream (S/data/FreqFilter.scala:28-36,44-49 + PairedEndData.getPairs,
 * S/data/PairedEndData.scala:20-36): for each of the first n_reads reads with len >= k, every k-window x is
 * canonicalised (y = x if x.hashCode < rc(x).hashCode else rc(x)) and update(y, 1, _ + 1) is applied.
 * `bin` is a HOST buffer; *n_windows (optional) receives the number of updates. */
int gb_map_insert_reads(gb_map *m, const uint8_t *bin, size_t n_bytes, int64_t n_reads, int64_t *n_windows);

/* same, with the stream already resident in device memory (d_bin: cudaMalloc'ed on the map's device).
 * d_offsets: n_reads+1 record offsets on the device, or NULL when every record has the same length. */
int gb_map_insert_reads_device(gb_map *m, const uint8_t *d_bin, size_t n_bytes, const uint64_t *d_offsets,
                               int64_t n_reads, int64_t *n_windows);

/* same for n_records records laid out at a FIXED STRIDE of rec_bytes, each with its own length byte (<= max_len, and
 * 1 + ceil(len / 4) <= rec_bytes; the rest of a record is padding): ragged reads -- or pieces of reads, e.g. 16-byte super-k-mer
 * records -- without an offset array.  Records shorter than k contribute nothing (FreqFilter.scala:29). */
int gb_map_insert_records_device(gb_map *m, const uint8_t *d_bin, size_t n_bytes, uint32_t rec_bytes, int64_t n_records, uint32_t max_len,
                                 int64_t *n_windows);

/* update(key, 1, _ + 1) for n keys taken as they are (DNAMap.update(key, v0, f), ArrayDNAMap.scala:198-203
 * with the closure of FreqFilter.scala:33).  Keys need not be canonical. */
int gb_map_update_counts(gb_map *m, const uint64_t *keys, int64_t n);
/* update(key, v) (ArrayDNAMap.scala:191-196) for n (key, value) pairs; later pairs win on duplicate keys
 * only if the caller orders the calls (within one call duplicates are a caller error). */
int gb_map_update(gb_map *m, const uint64_t *keys, const int32_t *vals, int64_t n);

/* size (ArrayDNAMap.scala:71) */
int gb_map_size(gb_map *m, int64_t *size);
/* apply / contains (ArrayDNAMap.scala:181-184,232): counts[i] = value or 0, found[i] = 0/1; either may be NULL */
int gb_map_lookup(gb_map *m, const uint64_t *keys, int64_t n, int32_t *counts, uint8_t *found);
/* deleteAll((k, v) => v < min_count) (ArrayDNAMap.scala:212-215 with the predicate of FreqFilter.scala:55) */
int gb_map_delete_below(gb_map *m, int32_t min_count);
/* mapReduce / foreach (ArrayDNAMap.scala:234-241): the host closure runs over the exported arrays.
 * Call with keys == NULL to get *n only.  Order is unspecified (hash order in the reference too). */
int gb_map_export(gb_map *m, uint64_t *keys, int32_t *vals, int64_t cap, int64_t *n);
/* incoming / outcoming (S/data/graph/Graph.scala:270-282) for n query k-mers against the current contents:
 * masks[i] bit b (0..3) = outcoming has base b, bit 4+b = incoming has base b (Base.fromInt order). */
int gb_map_neighbour_masks(gb_map *m, const uint64_t *keys, int64_t n, uint8_t *masks);

/* empty the map and size it for min_capacity distinct keys (a fresh `new ArrayDNAMap[Int](k)` that reuses the
 * handle, its streams and the pooled device memory) */
int gb_map_clear(gb_map *m, int64_t min_capacity);
/* device-side stopwatch on the handle's stream (CUDA events): start, issue calls, stop (synchronises) */
int gb_timer_start(gb_map *m);
int gb_timer_stop(gb_map *m, int64_t *ns);
/* wait for every asynchronous operation issued on the handle */
int gb_sync(gb_map *m);

/* counters: [0] capacity (slots) [1] table bytes [2] rehash/grow count [3] windows inserted so far
 * [4] last insert time in ns (CUDA events) [5] fixed-stride fast path used (0/1)
 * [6] [7] when the last insert took the L2-blocked path: ns spent bucketing k-mers by table slice / upserting them
 * (0 0 = the fused random-access kernel was used; gb_tune("insert_path") forces a path) */
int gb_map_stats(gb_map *m, int64_t stats[8]);
/* CUDA-event durations (ns) of the handle's last calls, for the roofline report: [0] bucket pass and [1] slice-ordered upsert of
 * the last L2-blocked insert, [2] table sweep and [3] re-insert of the survivors of the last deleteAll, [4] slots that sweep
 * streamed (16 bytes each), [5] membership probes and [6] list ranking of the last Graph.buildGraph on this map, [7] the number of
 * pointer-jumping launches of that ranking */
int gb_map_phase_ns(gb_map *m, int64_t ns[8]);

/* ------------------------------------------------------------------------------------------------
 * Graph  (trait Graph / MapGraph / object Graph, S/data/graph/Graph.scala)
 * Nodes and edges are addressed by their index in the most recent gb_graph_counts/export (the
 * reference's ids are not reproducible even reference-vs-reference).
 * ---------------------------------------------------------------------------------------------- */

/* Graph.buildGraph(k, kmersFreq) (Graph.scala:269-382) on the map's current contents */
int gb_graph_build(gb_map *m, gb_graph **out);
/* The SHARDED form of Graph.buildGraph (csrc/sgraph.cuh: the build gb_pmap_graph_build runs over the GPUs of a box when
 * gb_tune pgraph_sharded = 1, the default) over n_shards VIRTUAL ranks on this map's one device: same result as gb_graph_build up
 * to node / edge numbering.  A diagnostic entry point: it exists so that the multi-GPU algorithm can be checked and profiled
 * on a single GPU (tests/test_sgraph_gpu.py).  1 <= n_shards <= 16. */
int gb_graph_build_virtual_shards(gb_map *m, int n_shards, gb_graph **out);
int gb_graph_destroy(gb_graph *g);
/* getNodes.size, getEdges.size, getEdges.map(_.seq.length).sum (GraphBuilder.scala:39) */
int gb_graph_counts(gb_graph *g, int64_t *n_nodes, int64_t *n_edges, int64_t *n_edge_bases);
/* node_kmers[n_nodes]; edge_start/edge_end[n_edges] = node indices; edge_off[n_edges+1] in BASES into one
 * 2-bit stream edge_bases_2bit[(n_edge_bases+3)/4] (base j at byte j/4, bits 2(j%4)).  Any pointer may be NULL. */
int gb_graph_export(gb_graph *g, uint64_t *node_kmers, uint32_t *edge_start, uint32_t *edge_end,
                    uint64_t *edge_off, uint8_t *edge_bases_2bit);
/* Graph.components (Graph.scala:54-72): node_label[n_nodes] in 0..*n_components-1 (may be NULL) */
int gb_graph_components(gb_graph *g, uint32_t *node_label, int64_t *n_components);
/* graph.retain(components.maxBy(_.size)) (GraphBuilder.scala:52-54, MapGraph.retain Graph.scala:161-165);
 * ties (the norm: strand twins) go to the component holding the smallest node k-mer */
int gb_graph_retain_largest(gb_graph *g);
/* MapGraph.retain(nodesSet) (Graph.scala:161-165) for any node set: node_keep[n_nodes] (host), non-zero = keep; edges stay
 * when both ends stay. */
int gb_graph_retain(gb_graph *g, const uint8_t *node_keep);
/* MapGraph.simplifyGraph (Graph.scala:211-230) */
int gb_graph_simplify(gb_graph *g);
/* Graph.removeBubbles (Graph.scala:125-149), out-edges taken in Base.fromInt order */
int gb_graph_remove_bubbles(gb_graph *g);
/* removeEdge for a list of edge indices (Graph.scala:191-195; the call at GraphSimplifier.scala:316) */
int gb_graph_remove_edges(gb_graph *g, const uint32_t *edge_idx, int64_t n);
/* The remaining mutators of trait Graph (Graph.scala:31-36) in bulk, applied in this order:
 *   replaceStart / replaceEnd (197-209): for i < n_replace, edge edge_idx[i] gets start new_start[i] and end new_end[i]
 *                                        (0xFFFFFFFF = keep);
 *   addNode (172-176) / addEdge (178-184): n_add_nodes nodes with k-mers add_node_kmers[] get the indices n_nodes, n_nodes + 1,
 *                                        ...; n_add_edges edges (add_start[i] -> add_end[i], bases add_bases[add_off[i] ..
 *                                        add_off[i + 1]) as one code per byte, add_off[0] = 0) get the indices n_edges, ...;
 *                                        replacements and new edges may name the new nodes;
 *   removeNode (185-187): the listed nodes are dropped; like the reference's it does not touch edges, so a node that an edge
 *                                        still starts or ends at is refused (GB_E_INVARIANT; the earlier steps stay applied).
 * Every array is a host array; any count may be 0. */
int gb_graph_edit(gb_graph *g, int64_t n_replace, const uint32_t *edge_idx, const uint32_t *new_start, const uint32_t *new_end,
                  int64_t n_add_nodes, const uint64_t *add_node_kmers, int64_t n_add_edges, const uint32_t *add_start,
                  const uint32_t *add_end, const uint64_t *add_off, const uint8_t *add_bases, int64_t n_remove_nodes,
                  const uint32_t *remove_nodes);
/* EXTENSION, no reference counterpart: one sweep of dead-end tip removal */
int gb_graph_clip_tips(gb_graph *g, int64_t max_len, int64_t *removed);
/* invariants of S/scripts/GraphSimplifier.scala:159-170 evaluated on the device; GB_E_INVARIANT if broken */
int gb_graph_check(gb_graph *g);
/* Graph.getGraphMap (Graph.scala:90-119) as a bulk export: the (k-mer, GraphPosition) pairs the reference feeds to putNew.
 * Entry i: kmers[i] oriented as in the graph (not canonicalised); dists[i] == 0: NodeGraphPosition(node ids[i]);
 * dists[i] >= 1: EdgeGraphPosition(edge ids[i], dists[i]).  *n = nodes + edge bases - edges (line 97).  Call with all three
 * arrays NULL to get *n only. */
int gb_graph_positions(gb_graph *g, uint64_t *kmers, uint32_t *ids, uint32_t *dists, int64_t cap, int64_t *n);
/* Graph.getGraphMap (Graph.scala:90-119) as a DEVICE-resident DNAMap[GraphPosition]: putNew (S/ds/ArrayDNAMap.scala:152-162)
 * of every entry gb_graph_positions lists.  A snapshot: later changes of the graph do not show, the handle outlives the graph.
 * EXPERIMENTAL this round: covered by the emulation tests only, its GPU test is gated (tests/test_graphmap_gpu.py). */
int gb_graph_map_create(gb_graph *g, gb_graph_map **out);
int gb_graph_map_destroy(gb_graph_map *gm);
int gb_graph_map_size(gb_graph_map *gm, int64_t *size);               /* nodeMap.size (Graph.scala:117) */
/* getAll (ArrayDNAMap.scala:103-113) / contains (232) for n k-mers, oriented as given (the map holds both strands, no
 * canonicalisation): counts[i] = number of positions under keys[i]; the first max_per_key of them are written to
 * ids / dists [i * max_per_key ...] (dist 0: NodeGraphPosition(id), else EdgeGraphPosition(id, dist); unused = 0xFFFFFFFF).
 * ids and dists may be NULL (contains only). */
int gb_graph_map_get_all(gb_graph_map *gm, const uint64_t *keys, int64_t n, int max_per_key, uint32_t *ids, uint32_t *dists,
                         uint32_t *counts);
/* The pair loop of GraphSimplifier.startup (S/scripts/GraphSimplifier.scala:188-263) with the WalkingActor walks (33-127) and
 * annotate (192-206) over the first n_pairs read pairs of a HOST `.bin` stream: for both orientations of every pair whose
 * reads are at least k long, graphMap.getAll of the first k-mers, the bounded walk between every pair of positions with
 * range = range_lo..range_hi (the reference hard-codes 180 to 250, line 153; range_hi <= 511 here), and pathsMap.
 * support[4 * e1 + b] = pathsMap((e1, e2)), e2 being the out-edge of e1's end node whose first base is b (a pathsMap key
 * is always such a pair); support has 4 * n_edges entries.  *bad_pairs = badPairs (246-248), *walked_cases = orientation
 * cases that reached the walkers.  Edge indices are those of the current gb_graph_export. */
int gb_graph_pair_support(gb_graph *g, const uint8_t *bin, size_t n_bytes, int64_t n_pairs, int range_lo, int range_hi,
                          uint32_t *support, int64_t *bad_pairs, int64_t *walked_cases);
/* The node sweep of GraphSimplifier.startup (268-316): per node with in- and out-edges the matrix support >= cutoff, its
 * bipartite components; each component with out-edges moves to a new copy of the node (addNode / replaceEnd / replaceStart),
 * in-edges alone in their component and out-edges no component reached are removed.  The emptied originals stay until
 * gb_graph_simplify, which the reference calls next (318).  support as filled by gb_graph_pair_support on this graph. */
int gb_graph_split_nodes(gb_graph *g, const uint32_t *support, int32_t cutoff, int64_t *edges_removed, int64_t *nodes_added);
/* counters of the build: [0] stored k-mers seen [1] pointer-jumping launches [2] oriented k-mers on perfect
 * cycles, dropped like the reference does (Graph.scala:375) [3] build time in ns (CUDA events);
 * of the last gb_graph_pair_support on the handle, in ns (CUDA events): [4] out-edge table + graph map (positions, putNew)
 * [5] pair filter (getAll x 4, annotate) [6] walks (all tiers); [7] orientation cases that survived the filter */
int gb_graph_stats(gb_graph *g, int64_t stats[8]);

/* ------------------------------------------------------------------------------------------------
 * PartitionedDNAMap  (S/ds/PartitionedDNAMap.scala:15-63): one hash-prefix shard per GPU, one process
 * (rank) per GPU, k-mers routed to their owner by an all-to-all over NVLink.
 * Every gb_pmap_* call is collective: all ranks of the communicator call it in the same order.
 * ---------------------------------------------------------------------------------------------- */
#define GB_UNIQUE_ID_BYTES 128
int gb_comm_unique_id(uint8_t id[GB_UNIQUE_ID_BYTES]);            /* rank 0, then broadcast by the host */
int gb_comm_create(const uint8_t id[GB_UNIQUE_ID_BYTES], int rank, int n_ranks, int device, gb_comm **out);
int gb_comm_destroy(gb_comm *c);
/* Element-wise sum over the ranks of a HOST array, in place; collective, same n on every rank.  For host-side compositions of
 * per-rank partial results: with the graph identical on every rank (gb_pmap_graph_build), each rank runs gb_graph_pair_support
 * on its own slice of the pairs and the support counts, badPairs and walked cases are summed (GraphSimplifier.scala:239-265). */
int gb_comm_allreduce_sum_u32(gb_comm *c, uint32_t *host, int64_t n);
int gb_comm_allreduce_sum_i64(gb_comm *c, int64_t *host, int64_t n);

/* new PartitionedDNAMap[Int](k) (PartitionedDNAMap.scala:15-28): this rank's shard, bound to the communicator */
int gb_pmap_create(gb_comm *c, int k, int64_t min_capacity_per_shard, uint32_t flags, gb_map **out);
/* FreqFilter.add over THIS RANK's slice of the read stream; k-mers go to partition(key) (60-63) */
int gb_pmap_insert_reads(gb_map *m, const uint8_t *bin, size_t n_bytes, int64_t n_reads, int64_t *n_windows);
int gb_pmap_insert_reads_device(gb_map *m, const uint8_t *d_bin, size_t n_bytes, const uint64_t *d_offsets,
                                int64_t n_reads, int64_t *n_windows);
/* size = sum over partitions (PartitionedDNAMap.scala:31) */
int gb_pmap_size(gb_map *m, int64_t *size);
/* deleteAll broadcast (49-51): no communication, every shard filters itself */
int gb_pmap_delete_below(gb_map *m, int32_t min_count);
/* apply/contains routed to the owner shard and back (33,53); every rank passes its own queries */
int gb_pmap_lookup(gb_map *m, const uint64_t *keys, int64_t n, int32_t *counts, uint8_t *found);
/* Graph.buildGraph over all shards; the compacted graph is materialised on every rank */
int gb_pmap_graph_build(gb_map *m, gb_graph **out);
/* owner shard of a k-mer (any function is unobservable in the reference) */
int gb_pmap_owner(gb_map *m, const uint64_t *keys, int64_t n, int32_t *owner);
/* the same partition(key) (PartitionedDNAMap.scala:60-63) for n_parts shards: host arithmetic, needs no GPU */
int gb_owner_of(const uint64_t *keys, int64_t n, int n_parts, int32_t *owner);
/* the ownership rule of the sharded graph build (gb_pmap_graph_build re-routes the kept k-mers by it): the hash of the
 * k-mer's minimizer (smallest hashed canonical m-mer), shared by a k-mer, its reverse complement and ~90 % of its (k-1)-overlap
 * neighbours; host arithmetic, needs no GPU */
int gb_owner_of_minimizer(const uint64_t *keys, int64_t n, int k, int n_parts, int32_t *owner);

#ifdef __cplusplus
}
#endif
#endif
