/*
 * oracle.c -- CPU restatement of winger/genome's k-mer -> de Bruijn graph path (see oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no reference tests/golden vectors exist, no JVM here).
 * Every function cites the Scala it follows; S/ = /root/reference/src/main/scala/ru/ifmo/genome/.
 */
#include "oracle.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * Long1DNASeq arithmetic.  A k-mer (k <= 31) is one u64, base i at bits 2i..2i+1 (S/dna/DNASeq.scala:80-85),
 * base code A0 G1 C2 T3 (S/dna/Base.scala:13-16), complement = code ^ 3 (Base.scala:19).
 * ---------------------------------------------------------------------------------------------- */

/* Long1DNASeq.hashCode = long.## (DNASeq.scala:103).  scala-library 2.9.1 (project/Build.scala:21):
 * ScalaRunTime.hash(lv: Long) = { val iv = lv.toInt; if (iv == lv) iv else lv.hashCode } with
 * java.lang.Long.hashCode = (int)(v ^ (v >>> 32)).  Variant 210 is scala >= 2.10's formula, kept as a
 * switch in case a JVM run ever contradicts the 2.9.1 reading. */
int32_t go_hash(uint64_t v, int variant)
{
    if (variant == 210) {
        int32_t low = (int32_t)(uint32_t)v;
        int32_t low_sign = (int32_t)((uint32_t)low >> 31);
        int32_t high = (int32_t)(uint32_t)(v >> 32);
        return low ^ (int32_t)((uint32_t)high + (uint32_t)low_sign);
    }
    int32_t iv = (int32_t)(uint32_t)v;
    if ((int64_t)iv == (int64_t)v) return iv;
    return (int32_t)(uint32_t)(v ^ (v >> 32));
}

/* Long1DNASeq.reverse (DNASeq.scala:155-163) and complement (165-168); revComplement = complement.reverse (28) */
static uint64_t reverse_groups(uint64_t i, int k)
{
    i = (i & 0x3333333333333333ULL) << 2 | ((i >> 2) & 0x3333333333333333ULL);
    i = (i & 0x0f0f0f0f0f0f0f0fULL) << 4 | ((i >> 4) & 0x0f0f0f0f0f0f0f0fULL);
    i = (i & 0x00ff00ff00ff00ffULL) << 8 | ((i >> 8) & 0x00ff00ff00ff00ffULL);
    i = (i << 48) | ((i & 0xffff0000ULL) << 16) | ((i >> 16) & 0xffff0000ULL) | (i >> 48);
    return i >> (2 * (32 - k));
}

uint64_t go_revcomp(uint64_t x, int k)
{
    uint64_t mask = (1ULL << (2 * k)) - 1; /* k <= 31; the k == 32 branch of the reference is broken */
    return reverse_groups(x ^ mask, k);
}

/* FreqFilter.add (S/data/FreqFilter.scala:31-32): y = if (x.hashCode < rcx.hashCode) x else rcx */
uint64_t go_canonical(uint64_t x, int k, int variant)
{
    uint64_t rc = go_revcomp(x, k);
    return go_hash(x, variant) < go_hash(rc, variant) ? x : rc;
}

/* ArrayDNAMap.improve (S/ds/ArrayDNAMap.scala:267-272), 32-bit wrapping arithmetic, >>> logical */
int32_t go_improve(int32_t hcode)
{
    uint32_t h = (uint32_t)hcode + ~((uint32_t)hcode << 9);
    h = h ^ (h >> 14);
    h = h + (h << 4);
    return (int32_t)(h ^ (h >> 10));
}

/* base +: x.take(k-1)  (Graph.scala:273 -> DNASeq.scala:123,135-143) */
uint64_t go_prepend(uint64_t x, int k, int base)
{
    uint64_t m1 = (1ULL << (2 * (k - 1))) - 1;
    return ((x & m1) << 2) | (uint64_t)base;
}

/* x.drop(1) :+ base  (Graph.scala:279 -> DNASeq.scala:125,145-153) */
uint64_t go_append(uint64_t x, int k, int base)
{
    return (x >> 2) | ((uint64_t)base << (2 * (k - 1)));
}

/* ------------------------------------------------------------------------------------------------
 * ArrayDNAMap.Container (S/ds/ArrayDNAMap.scala:74-179) for T = Int and k <= 32 (8-byte keys, 249-257)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t bins, size;
    uint64_t *keys;
    int32_t *ar;
    uint64_t *set, *del; /* bitsets */
} container;

static inline int bit_get(const uint64_t *b, int64_t i) { return (int)((b[i >> 6] >> (i & 63)) & 1); }
static inline void bit_set(uint64_t *b, int64_t i) { b[i >> 6] |= 1ULL << (i & 63); }
static inline void bit_clr(uint64_t *b, int64_t i) { b[i >> 6] &= ~(1ULL << (i & 63)); }

static container *container_new(int64_t bins)
{
    container *c = (container *)calloc(1, sizeof *c);
    c->bins = bins;
    c->keys = (uint64_t *)calloc((size_t)bins, 8);
    c->ar = (int32_t *)calloc((size_t)bins, 4);
    c->set = (uint64_t *)calloc((size_t)(bins + 63) / 64, 8);
    c->del = (uint64_t *)calloc((size_t)(bins + 63) / 64, 8);
    return c;
}

static void container_free(container *c)
{
    if (!c) return;
    free(c->keys); free(c->ar); free(c->set); free(c->del); free(c);
}

/* Container.apply 90-101 */
static int container_apply(const container *c, uint64_t key, int32_t hash, int32_t *v)
{
    int64_t mask = c->bins - 1;
    int64_t i = (int64_t)(uint32_t)go_improve(hash) & mask;
    while (bit_get(c->set, i)) {
        if (!bit_get(c->del, i) && c->keys[i] == key) { if (v) *v = c->ar[i]; return 1; }
        i = (i + 1) & mask;
    }
    return 0;
}

/* Container.update(key, v) 115-127 */
static void container_update(container *c, uint64_t key, int32_t hash, int32_t v)
{
    int64_t mask = c->bins - 1;
    int64_t i = (int64_t)(uint32_t)go_improve(hash) & mask;
    while (!bit_get(c->del, i) && bit_get(c->set, i) && c->keys[i] != key) i = (i + 1) & mask;
    if (bit_get(c->del, i) || !bit_get(c->set, i)) {
        bit_set(c->set, i); bit_clr(c->del, i);
        c->keys[i] = key;
        c->size++;
    }
    c->ar[i] = v;
}

/* Container.update(key, v0, f) 129-150 with v0 = 1, f = _ + 1 (FreqFilter.scala:33) */
static void container_update1(container *c, uint64_t key, int32_t hash)
{
    int64_t mask = c->bins - 1;
    int64_t i = (int64_t)(uint32_t)go_improve(hash) & mask;
    int64_t first_pos = -1;
    while (bit_get(c->set, i) && (bit_get(c->del, i) || c->keys[i] != key)) {
        if (bit_get(c->del, i)) first_pos = i;
        i = (i + 1) & mask;
    }
    if (!bit_get(c->set, i)) {
        if (first_pos != -1) { i = first_pos; bit_clr(c->del, i); }
        bit_set(c->set, i);
        c->keys[i] = key;
        c->size++;
        c->ar[i] = 1;
    } else {
        c->ar[i] = (int32_t)((uint32_t)c->ar[i] + 1u);
    }
}

/* Container.putNew 152-162 */
static void container_put_new(container *c, uint64_t key, int32_t hash, int32_t v)
{
    int64_t mask = c->bins - 1;
    int64_t i = (int64_t)(uint32_t)go_improve(hash) & mask;
    while (!bit_get(c->del, i) && bit_get(c->set, i)) i = (i + 1) & mask;
    bit_set(c->set, i); bit_clr(c->del, i);
    c->size++;
    c->keys[i] = key;
    c->ar[i] = v;
}

/* ArrayDNAMap (62-243): container + rescale 217-230, load factors 246-247, initial 16 bins (72) */
typedef struct { container *c; int variant; } array_map;

static void array_map_rescale(array_map *a)
{
    container *c = a->c;
    if ((c->bins > 16 && (double)c->size < (double)c->bins * 0.3) || (double)c->bins * 0.7 < (double)c->size) {
        int64_t nb = 16;
        while ((double)nb * 0.7 < (double)c->size) nb *= 2;
        container *n = container_new(nb);
        for (int64_t i = 0; i < c->bins; i++)
            if (bit_get(c->set, i) && !bit_get(c->del, i))
                container_put_new(n, c->keys[i], go_hash(c->keys[i], a->variant), c->ar[i]);
        container_free(c);
        a->c = n;
    }
}

/* PartitionedDNAMap (S/ds/PartitionedDNAMap.scala:15-63) without the actors: P local ArrayDNAMaps */
struct go_map {
    int k, parts, variant;
    array_map *p;
};

go_map *go_map_new(int k, int partitions, int variant)
{
    if (k < 1 || k > 31 || partitions < 1) return NULL;
    go_map *m = (go_map *)calloc(1, sizeof *m);
    m->k = k; m->parts = partitions; m->variant = variant ? variant : 291;
    m->p = (array_map *)calloc((size_t)partitions, sizeof *m->p);
    for (int i = 0; i < partitions; i++) { m->p[i].c = container_new(16); m->p[i].variant = m->variant; }
    return m;
}

void go_map_free(go_map *m)
{
    if (!m) return;
    for (int i = 0; i < m->parts; i++) container_free(m->p[i].c);
    free(m->p); free(m);
}

int go_map_k(const go_map *m) { return m->k; }

/* PartitionedDNAMap.partition 60-63: fix(key.hashCode % P); Java % takes the dividend's sign */
static inline int partition_of(const go_map *m, int32_t hash)
{
    int i = hash % m->parts;
    return i < 0 ? i + m->parts : i;
}

void go_map_update1(go_map *m, uint64_t key)
{
    int32_t h = go_hash(key, m->variant);
    array_map *a = &m->p[partition_of(m, h)];
    container_update1(a->c, key, h);
    array_map_rescale(a);
}

void go_map_update(go_map *m, uint64_t key, int32_t v)
{
    int32_t h = go_hash(key, m->variant);
    array_map *a = &m->p[partition_of(m, h)];
    container_update(a->c, key, h, v);
    array_map_rescale(a);
}

int go_map_apply(const go_map *m, uint64_t key, int32_t *v)
{
    int32_t h = go_hash(key, m->variant);
    return container_apply(m->p[partition_of(m, h)].c, key, h, v);
}

int64_t go_map_size(const go_map *m)
{
    int64_t s = 0;
    for (int i = 0; i < m->parts; i++) s += m->p[i].c->size;
    return s;
}

int64_t go_map_bins(const go_map *m)
{
    int64_t s = 0;
    for (int i = 0; i < m->parts; i++) s += m->p[i].c->bins;
    return s;
}

/* ArrayDNAMap.deleteAll 212-215 -> Container.deleteAll 164-173 with p = (k, v) => v < rounds (FreqFilter.scala:55) */
void go_map_delete_below(go_map *m, int32_t rounds)
{
    for (int p = 0; p < m->parts; p++) {
        container *c = m->p[p].c;
        for (int64_t i = 0; i < c->bins; i++)
            if (bit_get(c->set, i) && !bit_get(c->del, i) && c->ar[i] < rounds) { bit_set(c->del, i); c->size--; }
        array_map_rescale(&m->p[p]);
    }
}

/* Container.iterator 175-178, partitions concatenated (PartitionedDNAMap.mapReduce 55-58) */
int64_t go_map_export(const go_map *m, uint64_t *keys, int32_t *vals, int64_t cap)
{
    int64_t n = 0;
    for (int p = 0; p < m->parts; p++) {
        const container *c = m->p[p].c;
        for (int64_t i = 0; i < c->bins; i++)
            if (bit_get(c->set, i) && !bit_get(c->del, i)) {
                if (n < cap) { if (keys) keys[n] = c->keys[i]; if (vals) vals[n] = c->ar[i]; }
                n++;
            }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * .bin decode + window extraction
 * PairedEndData.getPairs.read (S/data/PairedEndData.scala:24-32): 1 length byte, (len+3)/4 packed bytes,
 * base i in byte i/4 at bits 2(i%4) (DNASeq.scala:285-303 / ArrayDNASeq.apply 46-51).
 * seq.sliding(k) (FreqFilter.scala:29-30): all len-k+1 windows when len >= k, none otherwise.
 * ---------------------------------------------------------------------------------------------- */
static inline int read_base(const uint8_t *p, int i) { return (p[i >> 2] >> (2 * (i & 3))) & 3; }

int64_t go_count_windows(const uint8_t *bin, size_t n_bytes, int64_t n_reads, int k)
{
    size_t pos = 0;
    int64_t w = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        if (pos >= n_bytes) return -1;
        int len = bin[pos];
        size_t bl = (size_t)(len + 3) / 4;
        if (pos + 1 + bl > n_bytes) return -1;
        if (len >= k) w += len - k + 1;
        pos += 1 + bl;
    }
    return w;
}

typedef void (*kmer_sink)(void *ctx, uint64_t canonical);

static int64_t for_each_window(const uint8_t *bin, size_t n_bytes, int64_t r0, int64_t r1, size_t start_pos,
                               int k, int variant, kmer_sink sink, void *ctx, size_t *end_pos)
{
    size_t pos = start_pos;
    int64_t w = 0;
    uint64_t mask = (1ULL << (2 * k)) - 1;
    for (int64_t r = r0; r < r1; r++) {
        if (pos >= n_bytes) break;
        int len = bin[pos];
        size_t bl = (size_t)(len + 3) / 4;
        if (pos + 1 + bl > n_bytes) break;
        const uint8_t *p = bin + pos + 1;
        if (len >= k) {
            uint64_t x = 0;
            for (int i = 0; i < len; i++) {
                /* window [i-k+1, i]: base j of the window at bits 2j => newest base enters at the top */
                x = (x >> 2) | ((uint64_t)read_base(p, i) << (2 * (k - 1)));
                if (i >= k - 1) {
                    sink(ctx, go_canonical(x & mask, k, variant));
                    w++;
                }
            }
        }
        pos += 1 + bl;
    }
    if (end_pos) *end_pos = pos;
    return w;
}

static void sink_update1(void *ctx, uint64_t y) { go_map_update1((go_map *)ctx, y); }

int64_t go_insert_reads(go_map *m, const uint8_t *bin, size_t n_bytes, int64_t n_reads)
{
    return for_each_window(bin, n_bytes, 0, n_reads, 0, m->k, m->variant, sink_update1, m, NULL);
}

typedef struct { uint64_t *out; int64_t cap, n; } extract_ctx;
static void sink_extract(void *ctx, uint64_t y)
{
    extract_ctx *e = (extract_ctx *)ctx;
    if (e->n < e->cap) e->out[e->n] = y;
    e->n++;
}

int64_t go_extract_canonical(const uint8_t *bin, size_t n_bytes, int64_t n_reads, int k, int variant,
                             uint64_t *out, int64_t cap)
{
    extract_ctx e = { out, cap, 0 };
    for_each_window(bin, n_bytes, 0, n_reads, 0, k, variant ? variant : 291, sink_extract, &e, NULL);
    return e.n;
}

/* ---- multi-threaded timing variant: T extractor threads -> P x T buckets -> P single-threaded inserters ---- */
typedef struct { uint64_t *v; int64_t n, cap; } bucket;
typedef struct {
    go_map *m;
    const uint8_t *bin; size_t n_bytes;
    int64_t r0, r1; size_t pos0;
    bucket *row;      /* this thread's P buckets */
    int64_t windows;
} extract_job;

static void sink_bucket(void *ctx, uint64_t y)
{
    extract_job *j = (extract_job *)ctx;
    bucket *b = &j->row[partition_of(j->m, go_hash(y, j->m->variant))];
    if (b->n == b->cap) { b->cap = b->cap ? b->cap * 2 : 4096; b->v = (uint64_t *)realloc(b->v, (size_t)b->cap * 8); }
    b->v[b->n++] = y;
}

static void *extract_thread(void *arg)
{
    extract_job *j = (extract_job *)arg;
    j->windows = for_each_window(j->bin, j->n_bytes, j->r0, j->r1, j->pos0, j->m->k, j->m->variant, sink_bucket, j, NULL);
    return NULL;
}

typedef struct { go_map *m; int part; bucket *all; int threads; } insert_job;

static void *insert_thread(void *arg)
{
    insert_job *j = (insert_job *)arg;
    array_map *a = &j->m->p[j->part];
    for (int t = 0; t < j->threads; t++) {
        bucket *b = &j->all[(size_t)t * j->m->parts + j->part];
        for (int64_t i = 0; i < b->n; i++) {
            container_update1(a->c, b->v[i], go_hash(b->v[i], j->m->variant));
            array_map_rescale(a);
        }
        b->n = 0;
    }
    return NULL;
}

int64_t go_insert_reads_mt(go_map *m, const uint8_t *bin, size_t n_bytes, int64_t n_reads, int threads)
{
    if (threads < 1) threads = 1;
    const int64_t chunk = 1 << 18; /* reads per round, bounds bucket memory */
    int P = m->parts;
    bucket *all = (bucket *)calloc((size_t)threads * P, sizeof *all);
    extract_job *ej = (extract_job *)calloc((size_t)threads, sizeof *ej);
    insert_job *ij = (insert_job *)calloc((size_t)P, sizeof *ij);
    pthread_t *th = (pthread_t *)calloc((size_t)(threads > P ? threads : P), sizeof *th);
    int64_t total = 0;
    size_t pos = 0;
    for (int64_t r = 0; r < n_reads; r += chunk) {
        int64_t re = r + chunk < n_reads ? r + chunk : n_reads;
        /* split [r, re) into `threads` runs of reads; record boundaries need a sequential length-byte scan */
        int64_t per = (re - r + threads - 1) / threads;
        size_t p = pos;
        int64_t rr = r;
        for (int t = 0; t < threads; t++) {
            int64_t t1 = rr + per < re ? rr + per : re;
            ej[t] = (extract_job){ m, bin, n_bytes, rr, t1, p, &all[(size_t)t * P], 0 };
            for (; rr < t1 && p < n_bytes; rr++) p += 1 + (size_t)(bin[p] + 3) / 4;
        }
        pos = p;
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, extract_thread, &ej[t]);
        for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); total += ej[t].windows; }
        for (int q = 0; q < P; q++) { ij[q] = (insert_job){ m, q, all, threads }; pthread_create(&th[q], NULL, insert_thread, &ij[q]); }
        for (int q = 0; q < P; q++) pthread_join(th[q], NULL);
    }
    for (int i = 0; i < threads * P; i++) free(all[i].v);
    free(all); free(ej); free(ij); free(th);
    return total;
}

/* ------------------------------------------------------------------------------------------------
 * MapGraph (S/data/graph/Graph.scala:152-262), Node (Node.scala:41-52), Edge (Edge.scala:11-24)
 * ids are 1-based counters (nodeIdGen/edgeIdGen.incrementAndGet, 173,179); index = id - 1.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t kmer;
    int alive;
    int64_t *in; int nin, cin; /* inEdgeIds (a Set) */
    int64_t out[4];            /* outEdgeIds: Base -> edge id, 0 = absent */
} node_t;

typedef struct {
    int64_t start, end; /* node ids */
    uint8_t *seq; int64_t len;
    int alive;
} edge_t;

struct go_graph {
    int k;
    node_t *nodes; int64_t nn, cn;
    edge_t *edges; int64_t ne, ce;
};

static int64_t graph_add_node(go_graph *g, uint64_t kmer)
{
    if (g->nn == g->cn) { g->cn = g->cn ? g->cn * 2 : 64; g->nodes = (node_t *)realloc(g->nodes, (size_t)g->cn * sizeof(node_t)); }
    node_t *n = &g->nodes[g->nn++];
    memset(n, 0, sizeof *n);
    n->kmer = kmer; n->alive = 1;
    return g->nn; /* id */
}

static void node_in_add(node_t *n, int64_t eid)
{
    for (int i = 0; i < n->nin; i++) if (n->in[i] == eid) return;
    if (n->nin == n->cin) { n->cin = n->cin ? n->cin * 2 : 4; n->in = (int64_t *)realloc(n->in, (size_t)n->cin * 8); }
    n->in[n->nin++] = eid;
}

static void node_in_del(node_t *n, int64_t eid)
{
    for (int i = 0; i < n->nin; i++) if (n->in[i] == eid) { n->in[i] = n->in[--n->nin]; return; }
}

/* MapGraph.addEdge 178-184; takes ownership of seq */
static int64_t graph_add_edge(go_graph *g, int64_t start, int64_t end, uint8_t *seq, int64_t len)
{
    if (g->ne == g->ce) { g->ce = g->ce ? g->ce * 2 : 64; g->edges = (edge_t *)realloc(g->edges, (size_t)g->ce * sizeof(edge_t)); }
    edge_t *e = &g->edges[g->ne++];
    e->start = start; e->end = end; e->seq = seq; e->len = len; e->alive = 1;
    int64_t id = g->ne;
    g->nodes[start - 1].out[seq[0]] = id;
    node_in_add(&g->nodes[end - 1], id);
    return id;
}

/* MapGraph.removeEdge 191-195 */
static void graph_remove_edge(go_graph *g, int64_t id)
{
    edge_t *e = &g->edges[id - 1];
    if (!e->alive) return;
    if (g->nodes[e->start - 1].alive) g->nodes[e->start - 1].out[e->seq[0]] = 0;
    if (g->nodes[e->end - 1].alive) node_in_del(&g->nodes[e->end - 1], id);
    e->alive = 0;
}

void go_graph_free(go_graph *g)
{
    if (!g) return;
    for (int64_t i = 0; i < g->nn; i++) free(g->nodes[i].in);
    for (int64_t i = 0; i < g->ne; i++) free(g->edges[i].seq);
    free(g->nodes); free(g->edges); free(g);
}

/* Graph.buildGraph.contains 270 */
static int kept_contains(const go_map *m, uint64_t x)
{
    return go_map_apply(m, x, NULL) || go_map_apply(m, go_revcomp(x, m->k), NULL);
}

/* incoming 272-276 / outcoming 278-282: bases in Base.fromInt order A,G,C,T = 0,1,2,3 */
static int incoming(const go_map *m, uint64_t x, int *bases)
{
    int n = 0;
    for (int b = 0; b < 4; b++) if (kept_contains(m, go_prepend(x, m->k, b))) bases[n++] = b;
    return n;
}

static int outcoming(const go_map *m, uint64_t x, int *bases)
{
    int n = 0;
    for (int b = 0; b < 4; b++) if (kept_contains(m, go_append(x, m->k, b))) bases[n++] = b;
    return n;
}

/* Graph.buildGraph 269-382 */
go_graph *go_build_graph(const go_map *m)
{
    int k = m->k;
    go_graph *g = (go_graph *)calloc(1, sizeof *g);
    g->k = k;

    /* op1 320-329: classify every stored k-mer; termKmers 330-333 = set ++ set.map(revComplement) */
    int64_t n = go_map_size(m);
    uint64_t *keys = (uint64_t *)malloc((size_t)(n ? n : 1) * 8);
    go_map_export(m, keys, NULL, n);
    go_map *node_map = go_map_new(k, 1, m->variant); /* nodeMap 343-347: k-mer -> node id */
    int bs[4];
    for (int pass = 0; pass < 2; pass++)
        for (int64_t i = 0; i < n; i++) {
            uint64_t read = keys[i];
            int in = incoming(m, read, bs), out = outcoming(m, read, bs);
            if ((in != 1 || out != 1) && (in != 0 || out != 0)) {
                uint64_t t = pass == 0 ? read : go_revcomp(read, k);
                if (!go_map_apply(node_map, t, NULL)) {
                    int64_t id = graph_add_node(g, t);
                    go_map_update(node_map, t, (int32_t)id);
                }
            }
        }
    free(keys);

    /* buildEdges 349-365 for every terminal k-mer (367-374) */
    int64_t nn = g->nn;
    for (int64_t ni = 0; ni < nn; ni++) {
        uint64_t read = g->nodes[ni].kmer;
        int outs[4];
        int no = outcoming(m, read, outs);
        for (int oi = 0; oi < no; oi++) {
            int64_t cap = 64, len = 1;
            uint8_t *seq = (uint8_t *)malloc((size_t)cap);
            seq[0] = (uint8_t)outs[oi];
            uint64_t cur = go_append(read, k, outs[oi]);
            int32_t end_id;
            while (!go_map_apply(node_map, cur, &end_id)) {
                int o2[4];
                int c = outcoming(m, cur, o2);
                if (c != 1) { /* assert(out.size == 1) 357 */
                    fprintf(stderr, "oracle: non-terminal k-mer with %d successors\n", c);
                    abort();
                }
                if (len == cap) { cap *= 2; seq = (uint8_t *)realloc(seq, (size_t)cap); }
                seq[len++] = (uint8_t)o2[0];
                cur = go_append(cur, k, o2[0]);
            }
            graph_add_edge(g, ni + 1, end_id, seq, len);
        }
    }
    go_map_free(node_map);
    return g;
}

void go_graph_counts(const go_graph *g, int64_t *n_nodes, int64_t *n_edges, int64_t *n_edge_bases)
{
    int64_t a = 0, b = 0, c = 0;
    for (int64_t i = 0; i < g->nn; i++) a += g->nodes[i].alive;
    for (int64_t i = 0; i < g->ne; i++) if (g->edges[i].alive) { b++; c += g->edges[i].len; }
    if (n_nodes) *n_nodes = a;
    if (n_edges) *n_edges = b;
    if (n_edge_bases) *n_edge_bases = c;
}

void go_graph_export(const go_graph *g, uint64_t *node_kmer, int64_t *node_id, int64_t *edge_start_id,
                     int64_t *edge_end_id, int64_t *edge_off, uint8_t *edge_bases)
{
    int64_t a = 0;
    for (int64_t i = 0; i < g->nn; i++) if (g->nodes[i].alive) {
        if (node_kmer) node_kmer[a] = g->nodes[i].kmer;
        if (node_id) node_id[a] = i + 1;
        a++;
    }
    int64_t b = 0, off = 0;
    for (int64_t i = 0; i < g->ne; i++) if (g->edges[i].alive) {
        const edge_t *e = &g->edges[i];
        if (edge_start_id) edge_start_id[b] = e->start;
        if (edge_end_id) edge_end_id[b] = e->end;
        if (edge_off) edge_off[b] = off;
        if (edge_bases) memcpy(edge_bases + off, e->seq, (size_t)e->len);
        off += e->len;
        b++;
    }
    if (edge_off) edge_off[b] = off;
}

/* Graph.components 54-72: undirected reachability over in-edge starts and out-edge ends */
static int64_t components_by_index(const go_graph *g, int64_t *comp /* per node index, -1 for dead */)
{
    int64_t nc = 0;
    int64_t *stack = (int64_t *)malloc((size_t)(g->nn ? g->nn : 1) * 8);
    for (int64_t i = 0; i < g->nn; i++) comp[i] = -1;
    for (int64_t s = 0; s < g->nn; s++) {
        if (!g->nodes[s].alive || comp[s] >= 0) continue;
        int64_t sp = 0;
        stack[sp++] = s; comp[s] = nc;
        while (sp) {
            const node_t *nd = &g->nodes[stack[--sp]];
            for (int i = 0; i < nd->nin; i++) {
                int64_t o = g->edges[nd->in[i] - 1].start - 1;
                if (g->nodes[o].alive && comp[o] < 0) { comp[o] = nc; stack[sp++] = o; }
            }
            for (int b = 0; b < 4; b++) if (nd->out[b]) {
                int64_t o = g->edges[nd->out[b] - 1].end - 1;
                if (g->nodes[o].alive && comp[o] < 0) { comp[o] = nc; stack[sp++] = o; }
            }
        }
        nc++;
    }
    free(stack);
    return nc;
}

int64_t go_graph_components(const go_graph *g, int64_t *label)
{
    int64_t *comp = (int64_t *)malloc((size_t)(g->nn ? g->nn : 1) * 8);
    int64_t nc = components_by_index(g, comp);
    int64_t a = 0;
    for (int64_t i = 0; i < g->nn; i++) if (g->nodes[i].alive) label[a++] = comp[i];
    free(comp);
    return nc;
}

/* GraphBuilder.scala:52-54: retain(components.maxBy(_.size)); MapGraph.retain 161-165 */
void go_graph_retain_largest(go_graph *g)
{
    if (!g->nn) return;
    int64_t *comp = (int64_t *)malloc((size_t)g->nn * 8);
    int64_t nc = components_by_index(g, comp);
    if (!nc) { free(comp); return; }
    int64_t *size = (int64_t *)calloc((size_t)nc, 8);
    uint64_t *mink = (uint64_t *)malloc((size_t)nc * 8);
    memset(mink, 0xff, (size_t)nc * 8);
    for (int64_t i = 0; i < g->nn; i++) if (comp[i] >= 0) {
        size[comp[i]]++;
        if (g->nodes[i].kmer < mink[comp[i]]) mink[comp[i]] = g->nodes[i].kmer;
    }
    int64_t best = 0;
    for (int64_t c = 1; c < nc; c++)
        if (size[c] > size[best] || (size[c] == size[best] && mink[c] < mink[best])) best = c;
    for (int64_t i = 0; i < g->nn; i++) if (g->nodes[i].alive && comp[i] != best) g->nodes[i].alive = 0;
    for (int64_t i = 0; i < g->ne; i++) if (g->edges[i].alive) {
        edge_t *e = &g->edges[i];
        if (!g->nodes[e->start - 1].alive || !g->nodes[e->end - 1].alive) e->alive = 0;
    }
    free(comp); free(size); free(mink);
}

/* MapGraph.simplifyGraph 211-230; node visiting order is ConcurrentHashMap order in the reference, id order
 * here -- the resulting edge multiset does not depend on it ("simplify is confluent"). */
void go_graph_simplify(go_graph *g)
{
    int64_t nn = g->nn;
    for (int64_t i = 0; i < nn; i++) {
        node_t *nd = &g->nodes[i];
        if (!nd->alive) continue;
        int nout = 0; int64_t e2 = 0;
        for (int b = 0; b < 4; b++) if (nd->out[b]) { nout++; e2 = nd->out[b]; }
        if (nd->nin == 0 && nout == 0) {
            nd->alive = 0;
        } else if (nd->nin == 1 && nout == 1) {
            int64_t e1 = nd->in[0];
            if (e1 == e2) {
                graph_remove_edge(g, e1);
            } else {
                graph_remove_edge(g, e1);
                graph_remove_edge(g, e2);
                edge_t a = g->edges[e1 - 1], b = g->edges[e2 - 1];
                uint8_t *seq = (uint8_t *)malloc((size_t)(a.len + b.len));
                memcpy(seq, a.seq, (size_t)a.len);
                memcpy(seq + a.len, b.seq, (size_t)b.len);
                graph_add_edge(g, a.start, b.end, seq, a.len + b.len);
                nd = &g->nodes[i]; /* edges array may have moved, nodes did not; keep pointer fresh anyway */
            }
            nd->alive = 0;
        }
    }
}

/* Graph.similar 121-123 */
static int similar(int64_t la, int64_t lb)
{
    int64_t d = la > lb ? la - lb : lb - la;
    return d * 5 < (la > lb ? la : lb);
}

/* Graph.removeBubbles 125-149.  out = node.outEdges.values.toArray: for a freshly built graph the map's
 * insertion order is Base.fromInt order (buildEdges 351); that order is used here unconditionally. */
void go_graph_remove_bubbles(go_graph *g)
{
    for (int64_t i = 0; i < g->nn; i++) {
        node_t *nd = &g->nodes[i];
        if (!nd->alive) continue;
        int64_t out[4]; int no = 0;
        for (int b = 0; b < 4; b++) if (nd->out[b]) out[no++] = nd->out[b];
        int rm[4] = { 0, 0, 0, 0 };
        for (int a = 0; a < no; a++) {
            if (rm[a]) continue;
            for (int b = a + 1; b < no; b++) {
                const edge_t *ea = &g->edges[out[a] - 1], *eb = &g->edges[out[b] - 1];
                if (ea->end == eb->end && similar(ea->len, eb->len)) rm[b] = 1;
            }
        }
        for (int a = 0; a < no; a++) if (rm[a]) graph_remove_edge(g, out[a]);
    }
}

int64_t go_graph_remove_edges(go_graph *g, const int64_t *edge_ids, int64_t n)
{
    int64_t r = 0;
    for (int64_t i = 0; i < n; i++)
        if (edge_ids[i] >= 1 && edge_ids[i] <= g->ne && g->edges[edge_ids[i] - 1].alive) { graph_remove_edge(g, edge_ids[i]); r++; }
    return r;
}

/* EXTENSION: tip clipping (the reference has no such routine).  One simultaneous sweep:
 * an edge e = (u -> v), u != v, len(e) < max_len is an OUT-tip candidate if v has in-degree 1 and
 * out-degree 0, an IN-tip candidate if u has in-degree 0 and out-degree 1.  An OUT candidate is removed
 * iff u has another out-edge that is not an OUT candidate or is strictly longer; an IN candidate is
 * removed iff v has another in-edge that is not an IN candidate or is strictly longer. */
static int node_outdeg(const node_t *n) { return (n->out[0] != 0) + (n->out[1] != 0) + (n->out[2] != 0) + (n->out[3] != 0); }

int64_t go_graph_clip_tips(go_graph *g, int64_t max_len)
{
    int64_t ne = g->ne, removed = 0;
    uint8_t *oc = (uint8_t *)calloc((size_t)(ne ? ne : 1), 1), *ic = (uint8_t *)calloc((size_t)(ne ? ne : 1), 1);
    uint8_t *kill = (uint8_t *)calloc((size_t)(ne ? ne : 1), 1);
    for (int64_t i = 0; i < ne; i++) {
        const edge_t *e = &g->edges[i];
        if (!e->alive || e->start == e->end || e->len >= max_len) continue;
        const node_t *u = &g->nodes[e->start - 1], *v = &g->nodes[e->end - 1];
        if (v->nin == 1 && node_outdeg(v) == 0) oc[i] = 1;
        if (u->nin == 0 && node_outdeg(u) == 1) ic[i] = 1;
    }
    for (int64_t i = 0; i < ne; i++) {
        const edge_t *e = &g->edges[i];
        if (oc[i]) {
            const node_t *u = &g->nodes[e->start - 1];
            for (int b = 0; b < 4; b++) {
                int64_t o = u->out[b];
                if (o && o != i + 1 && (!oc[o - 1] || g->edges[o - 1].len > e->len)) kill[i] = 1;
            }
        }
        if (ic[i]) {
            const node_t *v = &g->nodes[e->end - 1];
            for (int j = 0; j < v->nin; j++) {
                int64_t o = v->in[j];
                if (o != i + 1 && (!ic[o - 1] || g->edges[o - 1].len > e->len)) kill[i] = 1;
            }
        }
    }
    for (int64_t i = 0; i < ne; i++) if (kill[i]) { graph_remove_edge(g, i + 1); removed++; }
    free(oc); free(ic); free(kill);
    return removed;
}

/* Graph.getGraphMap (S/data/graph/Graph.scala:90-119): the k-mer -> GraphPosition multimap as a list of putNew calls in
 * the reference's order (nodes, then edges; ids = this oracle's ids).  Entry i: kmer[i] (oriented, NOT canonicalised),
 * id[i] = node id with dist[i] = 0 (NodeGraphPosition) or edge id with dist[i] >= 1 (EdgeGraphPosition(edgeId, dist)).
 * Returns the number of entries (= sum of edge lengths + nodes - edges, line 97); fills at most cap of them. */
int64_t go_graph_map(const go_graph *g, uint64_t *kmer, int64_t *id, int32_t *dist, int64_t cap)
{
    int64_t n = 0;
    for (int64_t i = 0; i < g->nn; i++) {
        if (!g->nodes[i].alive) continue;
        if (n < cap) { kmer[n] = g->nodes[i].kmer; id[n] = i + 1; dist[n] = 0; }
        n++;
    }
    for (int64_t e = 0; e < g->ne; e++) {
        const edge_t *ed = &g->edges[e];
        if (!ed->alive) continue;
        uint64_t seq = go_append(g->nodes[ed->start - 1].kmer, g->k, ed->seq[0]); /* edge.start.seq.drop(1) :+ start */
        int32_t d = 1;
        for (int64_t j = 1; j < ed->len; j++) { /* for (base <- edge.seq.tail) */
            if (n < cap) { kmer[n] = seq; id[n] = e + 1; dist[n] = d; }
            n++;
            seq = go_append(seq, g->k, ed->seq[j]);
            d++;
        }
    }
    return n;
}

/* invariants asserted at S/scripts/GraphSimplifier.scala:159-170 */
int go_graph_check(const go_graph *g)
{
    for (int64_t i = 0; i < g->ne; i++) {
        const edge_t *e = &g->edges[i];
        if (!e->alive) continue;
        const node_t *s = &g->nodes[e->start - 1], *t = &g->nodes[e->end - 1];
        if (!s->alive || !t->alive) return 1;
        if (s->out[e->seq[0]] != i + 1) return 2;
        int found = 0;
        for (int j = 0; j < t->nin; j++) found |= t->in[j] == i + 1;
        if (!found) return 3;
    }
    for (int64_t i = 0; i < g->nn; i++) {
        const node_t *n = &g->nodes[i];
        if (!n->alive) continue;
        for (int j = 0; j < n->nin; j++) {
            const edge_t *e = &g->edges[n->in[j] - 1];
            if (!e->alive || e->end != i + 1) return 4;
        }
        for (int b = 0; b < 4; b++) if (n->out[b]) {
            const edge_t *e = &g->edges[n->out[b] - 1];
            if (!e->alive || e->start != i + 1 || e->seq[0] != b) return 5;
        }
    }
    return 0;
}

/* ids of the live edges in go_graph_export order */
void go_graph_edge_ids(const go_graph *g, int64_t *edge_id)
{
    int64_t n = 0;
    for (int64_t i = 0; i < g->ne; i++) if (g->edges[i].alive) edge_id[n++] = i + 1;
}

/* ================================================================================================
 * Paired-end path support: S/scripts/GraphSimplifier.scala
 *   WalkingActor.reachable 43-72, WalkingActor.receive 77-126, annotate 192-206, the pair loop 213-248,
 *   the per-node in x out matrix and node splitting 268-317.
 * ============================================================================================== */

/* int64 -> int64 open-addressing map (keys >= 0), the stand-in for the Scala mutable.Map / Set instances below */
typedef struct { int64_t *k, *v; int64_t cap, n; } imap;
static void imap_init(imap *m, int64_t cap) { m->cap = cap; m->n = 0; m->k = (int64_t *)malloc((size_t)cap * 8); m->v = (int64_t *)malloc((size_t)cap * 8); memset(m->k, 0xFF, (size_t)cap * 8); }
static void imap_free(imap *m) { free(m->k); free(m->v); m->k = m->v = NULL; }
static int64_t imap_slot(const imap *m, int64_t key)
{
    uint64_t h = (uint64_t)key * 0x9E3779B97F4A7C15ULL;
    int64_t i = (int64_t)((h >> 20) & (uint64_t)(m->cap - 1));
    while (m->k[i] != -1 && m->k[i] != key) i = (i + 1) & (m->cap - 1);
    return i;
}
static int64_t *imap_get(const imap *m, int64_t key) { int64_t i = imap_slot(m, key); return m->k[i] == key ? &m->v[i] : NULL; }
static void imap_put(imap *m, int64_t key, int64_t val)
{
    if ((m->n + 1) * 2 > m->cap) {
        imap o = *m;
        imap_init(m, o.cap * 2);
        for (int64_t i = 0; i < o.cap; i++) if (o.k[i] != -1) imap_put(m, o.k[i], o.v[i]);
        imap_free(&o);
    }
    int64_t i = imap_slot(m, key);
    if (m->k[i] != key) { m->k[i] = key; m->n++; }
    m->v[i] = val;
}

/* WalkingActor.reachable (43-72): Dijkstra from `node` BACKWARDS over in-edges, distances <= range.last; the result maps
 * node id -> smallest distance.  (The LinkedHashMap cache of 41,66-69 only saves recomputation and is not restated.) */
typedef struct { int64_t dist, node; } heap_item;
static void reachable(const go_graph *g, int64_t node, int hi, imap *set)
{
    int64_t hn = 0, hc = 64;
    heap_item *heap = (heap_item *)malloc((size_t)hc * sizeof *heap);
    heap[hn++] = (heap_item){ 0, node };
    while (hn) {
        heap_item top = heap[0]; /* PriorityQueue with the reversed ordering of 47-51: smallest distance first */
        heap[0] = heap[--hn];
        for (int64_t i = 0;;) {
            int64_t l = 2 * i + 1, r = l + 1, s = i;
            if (l < hn && heap[l].dist < heap[s].dist) s = l;
            if (r < hn && heap[r].dist < heap[s].dist) s = r;
            if (s == i) break;
            heap_item t = heap[i]; heap[i] = heap[s]; heap[s] = t; i = s;
        }
        if (imap_get(set, top.node)) continue;
        imap_put(set, top.node, top.dist);
        const node_t *u = &g->nodes[top.node - 1];
        for (int j = 0; j < u->nin; j++) {
            const edge_t *e = &g->edges[u->in[j] - 1];
            int64_t d2 = top.dist + e->len;
            if (d2 <= hi) {
                if (hn == hc) { hc *= 2; heap = (heap_item *)realloc(heap, (size_t)hc * sizeof *heap); }
                int64_t i = hn++;
                heap[i] = (heap_item){ d2, e->start };
                while (i && heap[(i - 1) / 2].dist > heap[i].dist) {
                    heap_item t = heap[i]; heap[i] = heap[(i - 1) / 2]; heap[(i - 1) / 2] = t; i = (i - 1) / 2;
                }
            }
        }
    }
    free(heap);
}

typedef struct {
    const go_graph *g;
    int lo, hi;
    int64_t node2, dist2, end_edge; /* end_edge 0 = null */
    imap reach, memo, *path_edges;
} walk_ctx;

static int64_t pair_key(const go_graph *g, int64_t e1, int64_t e2) { return e1 * (g->ne + 1) + e2; }

/* the nested dfs of WalkingActor.receive (91-113); prev_edge 0 = null */
static int walk_dfs(walk_ctx *c, int64_t node1, int64_t dist1, int64_t prev_edge)
{
    const go_graph *g = c->g;
    /* memo keys are only ever created below the prune test, i.e. with dist1 <= hi */
    int64_t mkey = dist1 <= c->hi ? prev_edge * (c->hi + 1) + dist1 : -1;
    if (mkey >= 0) { int64_t *m = imap_get(&c->memo, mkey); if (m) return (int)*m; }
    int64_t *r = imap_get(&c->reach, node1);
    if (dist1 + c->dist2 + (r ? *r : c->hi + 1) > c->hi) return 0;
    int cur = 0;
    if (node1 == c->node2 && c->lo <= dist1 + c->dist2 && dist1 + c->dist2 <= c->hi) {
        if (prev_edge && c->end_edge) imap_put(c->path_edges, pair_key(g, prev_edge, c->end_edge), 1);
        cur = 1;
    }
    const node_t *nd = &g->nodes[node1 - 1];
    for (int b = 0; b < 4; b++) {
        int64_t eid = nd->out[b];
        if (!eid) continue;
        const edge_t *e = &g->edges[eid - 1];
        int res = walk_dfs(c, e->end, dist1 + e->len, eid);
        if (res && prev_edge) imap_put(c->path_edges, pair_key(g, prev_edge, eid), 1);
        cur |= res;
    }
    imap_put(&c->memo, mkey, cur);
    return cur;
}

/* WalkingActor.receive for one (pos1, pos2) (77-126).  A position is (id, dist): dist == 0 -> NodeGraphPosition(id),
 * dist >= 1 -> EdgeGraphPosition(id, dist).  path_edges receives the (prevEdge.id, edge.id) pairs; returns `good`. */
static int walk_positions(const go_graph *g, int64_t id1, int32_t d1, int64_t id2, int32_t d2, int lo, int hi, imap *path_edges)
{
    walk_ctx c;
    c.g = g; c.lo = lo; c.hi = hi; c.path_edges = path_edges;
    if (d2 == 0) { c.node2 = id2; c.dist2 = 0; c.end_edge = 0; }
    else { c.node2 = g->edges[id2 - 1].start; c.dist2 = d2; c.end_edge = id2; }
    int64_t start_edge = d1 == 0 ? 0 : id1;
    imap_init(&c.reach, 64);
    imap_init(&c.memo, 64);
    reachable(g, c.node2, hi, &c.reach);
    int64_t node0, dist0;
    if (d1 == 0) { node0 = id1; dist0 = 0; }
    else { node0 = g->edges[id1 - 1].end; dist0 = g->edges[id1 - 1].len - d1; }
    int good = walk_dfs(&c, node0, dist0, start_edge);
    imap_free(&c.reach);
    imap_free(&c.memo);
    return good;
}

int go_walk(const go_graph *g, int64_t id1, int32_t dist1, int64_t id2, int32_t dist2, int lo, int hi,
            int64_t *pairs, int64_t cap, int64_t *n_pairs)
{
    imap pe;
    imap_init(&pe, 64);
    int good = walk_positions(g, id1, dist1, id2, dist2, lo, hi, &pe);
    int64_t n = 0;
    for (int64_t i = 0; i < pe.cap; i++) if (pe.k[i] != -1) {
        if (n < cap) { pairs[2 * n] = pe.k[i] / (g->ne + 1); pairs[2 * n + 1] = pe.k[i] % (g->ne + 1); }
        n++;
    }
    imap_free(&pe);
    if (n_pairs) *n_pairs = n;
    return good;
}

/* the graphMap of 188 as a sorted multimap: getAll(key) = every entry whose k-mer equals key (ArrayDNAMap.getAll 103-113) */
typedef struct { uint64_t kmer; int64_t id; int32_t dist; } gm_entry;
static int gm_cmp(const void *a, const void *b)
{
    const gm_entry *x = (const gm_entry *)a, *y = (const gm_entry *)b;
    return x->kmer < y->kmer ? -1 : x->kmer > y->kmer;
}
static int64_t gm_get_all(const gm_entry *gm, int64_t n, uint64_t key, const gm_entry **first)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) / 2; if (gm[mid].kmer < key) lo = mid + 1; else hi = mid; }
    int64_t c = 0;
    while (lo + c < n && gm[lo + c].kmer == key) c++;
    *first = gm + lo;
    return c;
}

/* annotate (192-206): the pair is dropped when some edge position of the first list and some edge position of the second
 * lie on the same edge with range.contains((dist2 - dist1) + k) */
static int annotate_drops(const gm_entry *p1, int64_t n1, const gm_entry *p2, int64_t n2, int k, int lo, int hi)
{
    for (int64_t i = 0; i < n1; i++) {
        if (p1[i].dist == 0) continue;
        for (int64_t j = 0; j < n2; j++) {
            if (p2[j].dist == 0) continue;
            int64_t d = (int64_t)p2[j].dist - p1[i].dist + k;
            if (p1[i].id == p2[j].id && lo <= d && d <= hi) return 1;
        }
    }
    return 0;
}

static int pe_cmp(const void *a, const void *b)
{
    const int64_t *x = (const int64_t *)a, *y = (const int64_t *)b;
    return x[0] < y[0] ? -1 : x[0] > y[0];
}

/* the pair loop of GraphSimplifier.startup (188-263) over the first n_pairs pairs of a `.bin` stream: pathsMap as (e1, e2, count)
 * triples sorted by (e1, e2); *bad_pairs = badPairs (259-261); *walked = orientation cases that survived annotate with both
 * position lists non-empty.  Returns the number of triples (fills at most cap), -1 on a truncated stream. */
int64_t go_pair_support(const go_graph *g, const uint8_t *bin, size_t n_bytes, int64_t n_pairs, int lo, int hi,
                        int64_t *e1, int64_t *e2, int32_t *cnt, int64_t cap, int64_t *bad_pairs, int64_t *walked)
{
    const int k = g->k;
    int64_t gn = go_graph_map(g, NULL, NULL, NULL, 0);
    uint64_t *gk = (uint64_t *)malloc((size_t)(gn ? gn : 1) * 8);
    int64_t *gi = (int64_t *)malloc((size_t)(gn ? gn : 1) * 8);
    int32_t *gd = (int32_t *)malloc((size_t)(gn ? gn : 1) * 4);
    go_graph_map(g, gk, gi, gd, gn);
    gm_entry *gm = (gm_entry *)malloc((size_t)(gn ? gn : 1) * sizeof *gm);
    for (int64_t i = 0; i < gn; i++) gm[i] = (gm_entry){ gk[i], gi[i], gd[i] };
    free(gk); free(gi); free(gd);
    qsort(gm, (size_t)gn, sizeof *gm, gm_cmp);

    imap paths; /* pathsMap: (e1, e2) -> count */
    imap_init(&paths, 1024);
    int64_t bad = 0, nwalked = 0;
    size_t pos = 0;
    for (int64_t p = 0; p < n_pairs; p++) {
        uint64_t first[2] = { 0, 0 };
        int len[2];
        for (int r = 0; r < 2; r++) { /* PairedEndData.getPairs.read (20-36) */
            if (pos >= n_bytes) { imap_free(&paths); free(gm); return -1; }
            len[r] = bin[pos];
            size_t nxt = pos + 1 + (size_t)(len[r] + 3) / 4;
            if (nxt > n_bytes) { imap_free(&paths); free(gm); return -1; }
            if (len[r] >= k)
                for (int i = 0; i < k; i++) first[r] |= (uint64_t)read_base(bin + pos + 1, i) << (2 * i); /* p.take(k) */
            pos = nxt;
        }
        if (len[0] < k || len[1] < k) continue; /* 213 */
        /* f1..f4 (214-217) and the two orientation cases of 219 */
        const uint64_t q1[2] = { first[0], first[1] }, q2[2] = { go_revcomp(first[1], k), go_revcomp(first[0], k) };
        for (int c = 0; c < 2; c++) {
            const gm_entry *p1, *p2;
            int64_t n1 = gm_get_all(gm, gn, q1[c], &p1), n2 = gm_get_all(gm, gn, q2[c], &p2);
            if (annotate_drops(p1, n1, p2, n2, k, lo, hi)) continue;
            if (n1 == 0 || n2 == 0) continue; /* no futures: `if !list.isEmpty` (238) */
            nwalked++;
            imap pe;
            imap_init(&pe, 64);
            int good = 0;
            for (int64_t i = 0; i < n1; i++)
                for (int64_t j = 0; j < n2; j++)
                    good |= walk_positions(g, p1[i].id, p1[i].dist, p2[j].id, p2[j].dist, lo, hi, &pe);
            for (int64_t i = 0; i < pe.cap; i++) if (pe.k[i] != -1) { /* pathEdgesList.reduce(_ ++ _), one increment per pair of edges */
                int64_t *v = imap_get(&paths, pe.k[i]);
                imap_put(&paths, pe.k[i], v ? *v + 1 : 1);
            }
            if (!good) bad++;
            imap_free(&pe);
        }
    }
    int64_t n = paths.n;
    int64_t *kv = (int64_t *)malloc((size_t)(n ? n : 1) * 16);
    int64_t j = 0;
    for (int64_t i = 0; i < paths.cap; i++) if (paths.k[i] != -1) { kv[2 * j] = paths.k[i]; kv[2 * j + 1] = paths.v[i]; j++; }
    qsort(kv, (size_t)n, 16, pe_cmp);
    for (int64_t i = 0; i < n && i < cap; i++) {
        e1[i] = kv[2 * i] / (g->ne + 1); e2[i] = kv[2 * i] % (g->ne + 1); cnt[i] = (int32_t)kv[2 * i + 1];
    }
    free(kv); free(gm); imap_free(&paths);
    if (bad_pairs) *bad_pairs = bad;
    if (walked) *walked = nwalked;
    return n;
}

/* MapGraph.replaceStart / replaceEnd (Graph.scala:197-209): the edge keeps its id */
static void graph_replace_start(go_graph *g, int64_t eid, int64_t new_start)
{
    edge_t *e = &g->edges[eid - 1];
    g->nodes[e->start - 1].out[e->seq[0]] = 0;
    e->start = new_start;
    g->nodes[new_start - 1].out[e->seq[0]] = eid;
}
static void graph_replace_end(go_graph *g, int64_t eid, int64_t new_end)
{
    edge_t *e = &g->edges[eid - 1];
    node_in_del(&g->nodes[e->end - 1], eid);
    e->end = new_end;
    node_in_add(&g->nodes[new_end - 1], eid);
}

typedef struct { int nin, nout; int m[4][4]; int cutoff; int col_left[4], col_right[4]; } split_ctx;
static void split_dfs_right(split_ctx *s, int j, int *l, int *r);
static void split_dfs_left(split_ctx *s, int i, int *l, int *r) /* dfsLeft 281-291; l, r = bit sets */
{
    s->col_left[i] = 1;
    *l |= 1 << i;
    for (int j = 0; j < s->nout; j++) if (!s->col_right[j] && s->m[i][j] >= s->cutoff) split_dfs_right(s, j, l, r);
}
static void split_dfs_right(split_ctx *s, int j, int *l, int *r) /* dfsRight 292-302 */
{
    s->col_right[j] = 1;
    *r |= 1 << j;
    for (int i = 0; i < s->nin; i++) if (!s->col_left[i] && s->m[i][j] >= s->cutoff) split_dfs_left(s, i, l, r);
}

/* GraphSimplifier.startup 266-317 without the final simplifyGraph: for every node with in- and out-edges, the matrix
 * m(i)(j) = pathsMap(in(i), out(j)), the connected components of the bipartite graph {m >= cutoff}; a component with
 * out-edges moves to a fresh copy of the node (addNode(node.seq), replaceEnd, replaceStart), an in-edge alone in its
 * component and every out-edge no component reached are removed.  The sweep covers the nodes present at the call
 * (the reference iterates a ConcurrentHashMap while adding to it; a copy that does get visited is one component again and
 * only moves once more, leaving an empty node for simplifyGraph -- same graph).  Returns the number of edges removed. */
int64_t go_graph_split(go_graph *g, const int64_t *e1, const int64_t *e2, const int32_t *cnt, int64_t n, int32_t cutoff,
                       int64_t *nodes_added)
{
    imap paths;
    imap_init(&paths, 1024);
    for (int64_t i = 0; i < n; i++) imap_put(&paths, pair_key(g, e1[i], e2[i]), cnt[i]);
    int64_t nn0 = g->nn, added = 0, n_rm = 0, c_rm = 64;
    int64_t *to_remove = (int64_t *)malloc((size_t)c_rm * 8);
#define TO_REMOVE(id) do { if (n_rm == c_rm) { c_rm *= 2; to_remove = (int64_t *)realloc(to_remove, (size_t)c_rm * 8); } to_remove[n_rm++] = (id); } while (0)
    for (int64_t v = 0; v < nn0; v++) {
        if (!g->nodes[v].alive) continue;
        int64_t in[4], out[4];
        split_ctx s;
        memset(&s, 0, sizeof s);
        s.cutoff = cutoff;
        if (g->nodes[v].nin > 4) { fprintf(stderr, "oracle: node with more than 4 in-edges\n"); abort(); }
        for (int j = 0; j < g->nodes[v].nin; j++) in[s.nin++] = g->nodes[v].in[j];
        for (int b = 0; b < 4; b++) if (g->nodes[v].out[b]) out[s.nout++] = g->nodes[v].out[b];
        if (s.nin == 0 || s.nout == 0) continue;
        for (int i = 0; i < s.nin; i++)
            for (int j = 0; j < s.nout; j++) {
                int64_t *c = imap_get(&paths, pair_key(g, in[i], out[j]));
                s.m[i][j] = c ? (int)*c : 0;
            }
        for (int i = 0; i < s.nin; i++) {
            if (s.col_left[i]) continue;
            int l = 0, r = 0;
            split_dfs_left(&s, i, &l, &r);
            if (r == 0) {
                TO_REMOVE(in[i]);
            } else {
                int64_t nid = graph_add_node(g, g->nodes[v].kmer);
                added++;
                for (int a = 0; a < s.nin; a++) if (l >> a & 1) graph_replace_end(g, in[a], nid);
                for (int a = 0; a < s.nout; a++) if (r >> a & 1) graph_replace_start(g, out[a], nid);
            }
        }
        for (int j = 0; j < s.nout; j++) if (!s.col_right[j]) TO_REMOVE(out[j]);
    }
#undef TO_REMOVE
    int64_t removed = 0;
    for (int64_t i = 0; i < n_rm; i++) if (g->edges[to_remove[i] - 1].alive) { graph_remove_edge(g, to_remove[i]); removed++; }
    free(to_remove);
    imap_free(&paths);
    if (nodes_added) *nodes_added = added;
    return removed;
}
