/*
 * oracle.h -- CPU restatement of winger/genome's k-mer -> de Bruijn graph path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (genome_b200/, include/, hostcpp/) may
 * include, link or call this.  Allowed users: tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY UNPINNED: the reference (Scala 2.9.1 + Akka 2.1-SNAPSHOT + Kryo 2.14-SNAPSHOT) has no tests,
 * no golden vectors and cannot be built here (no JVM, dead snapshot repositories).  This file is a
 * line-by-line restatement of the cited Scala; the only reference-held datum it is checked against
 * is the Edge/Node print-out at application.conf:73 (tests/test_oracle_golden.py).
 *
 * Citations are relative to /root/reference; S/ = src/main/scala/ru/ifmo/genome/.
 */
#ifndef GENOME_ORACLE_H
#define GENOME_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- arithmetic (S/dna/Base.scala:13-19, S/dna/DNASeq.scala:74-169) ---- */
int32_t  go_hash(uint64_t v, int variant);            /* variant 291: scala 2.9.1 Long.##, 210: scala 2.10+ */
uint64_t go_revcomp(uint64_t x, int k);               /* complement (165-168) then reverse (155-163) */
uint64_t go_canonical(uint64_t x, int k, int variant);/* S/data/FreqFilter.scala:31-32 */
int32_t  go_improve(int32_t h);                       /* S/ds/ArrayDNAMap.scala:267-272 */
uint64_t go_prepend(uint64_t x, int k, int base);     /* base +: x.take(k-1)  (Graph.scala:273) */
uint64_t go_append(uint64_t x, int k, int base);      /* x.drop(1) :+ base    (Graph.scala:279) */

/* ---- PartitionedDNAMap[Int] of ArrayDNAMap (S/ds/ArrayDNAMap.scala:62-243, PartitionedDNAMap.scala) ---- */
typedef struct go_map go_map;
go_map  *go_map_new(int k, int partitions, int variant);
void     go_map_free(go_map *m);
int      go_map_k(const go_map *m);
void     go_map_update1(go_map *m, uint64_t key);               /* update(key, 1, _ + 1) */
void     go_map_update(go_map *m, uint64_t key, int32_t v);      /* update(key, v) */
int      go_map_apply(const go_map *m, uint64_t key, int32_t *v);/* 1 if present */
int64_t  go_map_size(const go_map *m);
void     go_map_delete_below(go_map *m, int32_t rounds);         /* deleteAll((k,v) => v < rounds) */
int64_t  go_map_export(const go_map *m, uint64_t *keys, int32_t *vals, int64_t cap); /* iterator order */
int64_t  go_map_bins(const go_map *m);                           /* sum of bins over partitions */

/* ---- reads (.bin layout, S/data/PairedEndData.scala:20-36) and FreqFilter (S/data/FreqFilter.scala:25-58) ---- */
/* number of k-windows in the first n_reads reads; -1 on a truncated stream */
int64_t  go_count_windows(const uint8_t *bin, size_t n_bytes, int64_t n_reads, int k);
/* FreqFilter.add over the first n_reads reads (pairs are just consecutive reads); returns windows inserted */
int64_t  go_insert_reads(go_map *m, const uint8_t *bin, size_t n_bytes, int64_t n_reads);
/* same result, timed CPU baseline: `threads` extractor threads route k-mers by PartitionedDNAMap.partition
 * through in-memory buckets to one single-threaded inserter per partition (one actor per partition). */
int64_t  go_insert_reads_mt(go_map *m, const uint8_t *bin, size_t n_bytes, int64_t n_reads, int threads);
/* window extraction only: writes canonical k-mers in stream order; returns count */
int64_t  go_extract_canonical(const uint8_t *bin, size_t n_bytes, int64_t n_reads, int k, int variant,
                              uint64_t *out, int64_t cap);

/* ---- graph (S/data/graph/Graph.scala) ---- */
typedef struct go_graph go_graph;
go_graph *go_build_graph(const go_map *m);                       /* Graph.buildGraph 269-382 */
void      go_graph_free(go_graph *g);
void      go_graph_counts(const go_graph *g, int64_t *n_nodes, int64_t *n_edges, int64_t *n_edge_bases);
/* live nodes in id order; live edges in id order; seq = one byte per base (code 0..3) */
void      go_graph_export(const go_graph *g, uint64_t *node_kmer, int64_t *node_id,
                          int64_t *edge_start_id, int64_t *edge_end_id, int64_t *edge_off, uint8_t *edge_bases);
void      go_graph_edge_ids(const go_graph *g, int64_t *edge_id); /* ids of the live edges, export order */
/* components 54-72: label[i] for live node i (export order), returns number of components */
int64_t   go_graph_components(const go_graph *g, int64_t *label);
/* GraphBuilder.scala:52-54; tie rule (reference order is JDK-dependent): largest, then the component
 * holding the smallest node k-mer */
void      go_graph_retain_largest(go_graph *g);
void      go_graph_simplify(go_graph *g);                        /* MapGraph.simplifyGraph 211-230 */
void      go_graph_remove_bubbles(go_graph *g);                  /* Graph.removeBubbles 125-149 */
int64_t   go_graph_remove_edges(go_graph *g, const int64_t *edge_ids, int64_t n); /* removeEdge 191-195 */
/* EXTENSION (no reference semantics): one sweep of dead-end tip removal, see DESIGN.md */
int64_t   go_graph_clip_tips(go_graph *g, int64_t max_len);
/* Graph.getGraphMap 90-119 as the list of putNew(kmer, position) calls; dist 0 = NodeGraphPosition(id) */
int64_t   go_graph_map(const go_graph *g, uint64_t *kmer, int64_t *id, int32_t *dist, int64_t cap);
/* invariants of S/scripts/GraphSimplifier.scala:159-170; 0 = ok */
int       go_graph_check(const go_graph *g);

/* ---- paired-end path support (S/scripts/GraphSimplifier.scala:33-127,188-317) ----
 * A graph position is (id, dist): dist == 0 -> NodeGraphPosition(id), dist >= 1 -> EdgeGraphPosition(id, dist). */
/* WalkingActor.receive (77-126) for one (pos1, pos2) and range lo..hi: returns `good`, writes the pathEdges set as
 * (prevEdge.id, edge.id) pairs (at most cap of them), *n_pairs = size of the set */
int       go_walk(const go_graph *g, int64_t id1, int32_t dist1, int64_t id2, int32_t dist2, int lo, int hi,
                  int64_t *pairs, int64_t cap, int64_t *n_pairs);
/* the pair loop (188-263) over the first n_pairs pairs of a `.bin` stream: pathsMap as (e1, e2, count) sorted by (e1, e2);
 * returns the number of triples (fills at most cap), -1 on a truncated stream */
int64_t   go_pair_support(const go_graph *g, const uint8_t *bin, size_t n_bytes, int64_t n_pairs, int lo, int hi,
                          int64_t *e1, int64_t *e2, int32_t *cnt, int64_t cap, int64_t *bad_pairs, int64_t *walked);
/* the in x out matrices, node splitting and edge removal of 266-317 (simplifyGraph excluded); returns edges removed */
int64_t   go_graph_split(go_graph *g, const int64_t *e1, const int64_t *e2, const int32_t *cnt, int64_t n, int32_t cutoff,
                         int64_t *nodes_added);

#ifdef __cplusplus
}
#endif
#endif
