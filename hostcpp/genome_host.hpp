// genome_host.hpp -- header-only C++ host mirror of the reference's operator surface over the C ABI
// (include/genome_b200.h).  Same names and argument meaning as the Scala (paths relative to /root/reference,
// S/ = src/main/scala/ru/ifmo/genome/):
//   trait DNAMap[Int]      S/ds/ArrayDNAMap.scala:49-60      -> genome::DNAMap
//   object FreqFilter      S/data/FreqFilter.scala:25-58     -> genome::FreqFilter::extractFilteredKmers
//   object Graph / MapGraph S/data/graph/Graph.scala         -> genome::Graph::buildGraph, genome::MapGraph
//   GraphSimplifier's pair loop and node sweep S/scripts/GraphSimplifier.scala:188-317 -> MapGraph::pairSupport / splitNodes
//   MapGraph.write(file) / Graph(file)  Graph.scala:232-261,384-390 -> MapGraph::write / Graph::apply (kryo_graph.hpp)
// Errors surface as genome::Error (the reference asserts / fails its Futures).  No CPU fallback exists.
#pragma once
#include <cstdint>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/genome_b200.h"
#include "kryo_graph.hpp"

namespace genome {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc)
{
    if (rc != GB_OK) throw Error(rc, gb_last_error());
}

struct Edge {
    uint32_t start, end;        // node indices
    std::vector<uint8_t> seq;   // base codes A0 G1 C2 T3 (S/dna/Base.scala:13-16)
};

class MapGraph {
  public:
    explicit MapGraph(gb_graph *g, int k = 0) : g_(g), k_(k) {}
    MapGraph(const MapGraph &) = delete;
    MapGraph &operator=(const MapGraph &) = delete;
    MapGraph(MapGraph &&o) noexcept : g_(o.g_), k_(o.k_) { o.g_ = nullptr; }
    ~MapGraph() { gb_graph_destroy(g_); }
    int k() const { return k_; }

    std::vector<uint64_t> getNodes() const
    {
        int64_t n, e, b;
        check(gb_graph_counts(g_, &n, &e, &b));
        std::vector<uint64_t> k((size_t)n);
        check(gb_graph_export(g_, k.data(), nullptr, nullptr, nullptr, nullptr));
        return k;
    }
    std::vector<Edge> getEdges() const
    {
        int64_t n, e, b;
        check(gb_graph_counts(g_, &n, &e, &b));
        std::vector<uint32_t> s((size_t)e), t((size_t)e);
        std::vector<uint64_t> off((size_t)e + 1);
        std::vector<uint8_t> packed((size_t)(b + 3) / 4);
        check(gb_graph_export(g_, nullptr, s.data(), t.data(), off.data(), packed.data()));
        std::vector<Edge> out((size_t)e);
        for (size_t i = 0; i < (size_t)e; i++) {
            out[i].start = s[i];
            out[i].end = t[i];
            for (uint64_t j = off[i]; j < off[i + 1]; j++) out[i].seq.push_back((packed[j >> 2] >> (2 * (j & 3))) & 3);
        }
        return out;
    }
    // Graph.components (54-72): label per node, returns the number of components
    int64_t components(std::vector<uint32_t> &label) const
    {
        int64_t n, e, b, nc = 0;
        check(gb_graph_counts(g_, &n, &e, &b));
        label.assign((size_t)n, 0);
        check(gb_graph_components(g_, label.data(), &nc));
        return nc;
    }
    // Graph.getGraphMap (90-119): the (k-mer, GraphPosition) pairs fed to putNew; dist 0 = NodeGraphPosition(id)
    struct Position { uint64_t kmer; uint32_t id; uint32_t dist; };
    std::vector<Position> getGraphMap() const
    {
        int64_t n = 0;
        check(gb_graph_positions(g_, nullptr, nullptr, nullptr, 0, &n));
        std::vector<uint64_t> k((size_t)n);
        std::vector<uint32_t> id((size_t)n), d((size_t)n);
        if (n) check(gb_graph_positions(g_, k.data(), id.data(), d.data(), n, &n));
        std::vector<Position> out((size_t)n);
        for (size_t i = 0; i < (size_t)n; i++) out[i] = { k[i], id[i], d[i] };
        return out;
    }
    // Graph.getGraphMap as a device-resident DNAMap[GraphPosition] (a snapshot): size / getAll / contains in bulk
    class GraphMap {
    public:
        explicit GraphMap(gb_graph *g) { check(gb_graph_map_create(g, &m_)); }
        ~GraphMap() { gb_graph_map_destroy(m_); }
        GraphMap(const GraphMap &) = delete;
        GraphMap &operator=(const GraphMap &) = delete;
        int64_t size() const { int64_t n = 0; check(gb_graph_map_size(m_, &n)); return n; }
        // getAll (S/ds/ArrayDNAMap.scala:103-113): counts[i] positions under keys[i], the first maxPer in ids / dists
        void getAll(const std::vector<uint64_t> &keys, int maxPer, std::vector<uint32_t> &ids, std::vector<uint32_t> &dists,
                    std::vector<uint32_t> &counts) const
        {
            ids.assign(keys.size() * (size_t)maxPer, 0xFFFFFFFFu);
            dists.assign(keys.size() * (size_t)maxPer, 0xFFFFFFFFu);
            counts.assign(keys.size(), 0);
            check(gb_graph_map_get_all(m_, keys.data(), (int64_t)keys.size(), maxPer, ids.data(), dists.data(), counts.data()));
        }
        std::vector<uint32_t> contains(const std::vector<uint64_t> &keys) const // ArrayDNAMap.scala:232
        {
            std::vector<uint32_t> counts(keys.size(), 0);
            check(gb_graph_map_get_all(m_, keys.data(), (int64_t)keys.size(), 0, nullptr, nullptr, counts.data()));
            return counts;
        }
    private:
        gb_graph_map *m_ = nullptr;
    };
    // the pair loop of GraphSimplifier.startup (S/scripts/GraphSimplifier.scala:188-263): pathsMap as support[4 * e1 + b]
    // (e2 = the out-edge of e1's end node with first base b), badPairs and the number of walked orientation cases
    struct PairSupport { std::vector<uint32_t> support; int64_t badPairs = 0, walkedCases = 0; };
    PairSupport pairSupport(const uint8_t *bin, size_t nBytes, int64_t pairs, int rangeFirst = 180, int rangeLast = 250) const
    {
        int64_t n, e, b;
        check(gb_graph_counts(g_, &n, &e, &b));
        PairSupport r;
        r.support.assign((size_t)e * 4, 0);
        uint32_t none = 0;
        check(gb_graph_pair_support(g_, bin, nBytes, pairs, rangeFirst, rangeLast, e ? r.support.data() : &none, &r.badPairs, &r.walkedCases));
        return r;
    }
    // the node sweep of GraphSimplifier.startup (268-316): (edges removed, nodes added); simplifyGraph comes next (318)
    std::pair<int64_t, int64_t> splitNodes(const std::vector<uint32_t> &support, int32_t cutoff)
    {
        int64_t removed = 0, added = 0;
        check(gb_graph_split_nodes(g_, support.data(), cutoff, &removed, &added));
        return { removed, added };
    }
    void retainLargest() { check(gb_graph_retain_largest(g_)); }   // GraphBuilder.scala:52-54
    void simplifyGraph() { check(gb_graph_simplify(g_)); }          // Graph.scala:211-230
    void removeBubbles() { check(gb_graph_remove_bubbles(g_)); }    // Graph.scala:125-149
    void removeEdges(const std::vector<uint32_t> &idx) { check(gb_graph_remove_edges(g_, idx.data(), (int64_t)idx.size())); }
    // MapGraph.write(file) (Graph.scala:232-248): the Kryo `graph` file that GraphBuilder hands to GraphSimplifier
    void write(const std::string &path) const
    {
        kryo::GraphArrays a;
        a.k = k_;
        a.nodeKmer = getNodes();
        for (auto &e : getEdges()) {
            a.edgeStart.push_back(e.start);
            a.edgeEnd.push_back(e.end);
            a.edgeSeq.push_back(std::move(e.seq));
        }
        const std::vector<uint8_t> bytes = kryo::write(a);
        std::ofstream f(path, std::ios::binary);
        f.write((const char *)bytes.data(), (std::streamsize)bytes.size());
        if (!f) throw Error(GB_E_ARG, "cannot write " + path);
    }
    gb_graph *handle() const { return g_; }

  private:
    gb_graph *g_;
    int k_;
};

class DNAMap {
  public:
    DNAMap(int k, int64_t minCapacity = 0, int device = 0, uint32_t flags = 0) : k_(k) { check(gb_map_create(k, minCapacity, device, flags, &m_)); }
    DNAMap(const DNAMap &) = delete;
    DNAMap &operator=(const DNAMap &) = delete;
    ~DNAMap() { gb_map_destroy(m_); }

    int64_t size() const { int64_t n; check(gb_map_size(m_, &n)); return n; }
    // apply(key): (found, value)
    std::pair<bool, int32_t> apply(uint64_t key) const
    {
        int32_t c = 0;
        uint8_t f = 0;
        check(gb_map_lookup(m_, &key, 1, &c, &f));
        return { f != 0, c };
    }
    bool contains(uint64_t key) const { return apply(key).first; }
    void update(uint64_t key, int32_t v) { check(gb_map_update(m_, &key, &v, 1)); }
    // update(key, 1, _ + 1) for a batch (FreqFilter.scala:33)
    void updateCounts(const std::vector<uint64_t> &keys) { check(gb_map_update_counts(m_, keys.data(), (int64_t)keys.size())); }
    // deleteAll((k, v) => v < rounds) (FreqFilter.scala:55)
    void deleteBelow(int32_t rounds) { check(gb_map_delete_below(m_, rounds)); }
    // mapReduce / foreach: the host closure runs over the exported pairs
    template <class F>
    void foreach(F f) const
    {
        int64_t n = 0;
        check(gb_map_export(m_, nullptr, nullptr, 0, &n));
        std::vector<uint64_t> k((size_t)n);
        std::vector<int32_t> v((size_t)n);
        if (n) check(gb_map_export(m_, k.data(), v.data(), n, &n));
        for (size_t i = 0; i < (size_t)n; i++) f(k[i], v[i]);
    }
    // FreqFilter.add over a `.bin` stream (host buffer); returns the number of k-windows
    int64_t insertReads(const uint8_t *bin, size_t nBytes, int64_t nReads)
    {
        int64_t w = 0;
        check(gb_map_insert_reads(m_, bin, nBytes, nReads, &w));
        return w;
    }
    int k() const { return k_; }
    gb_map *handle() const { return m_; }

  private:
    gb_map *m_ = nullptr;
    int k_;
};

struct FreqFilter {
    // extractFilteredKmers(data, k, rounds) (FreqFilter.scala:25-58); `pairs` = PairedEndData.count min genome.takeFirst
    static void extractFilteredKmers(DNAMap &kmersFreq, const uint8_t *bin, size_t nBytes, int64_t pairs, int32_t rounds)
    {
        kmersFreq.insertReads(bin, nBytes, 2 * pairs);
        kmersFreq.deleteBelow(rounds);
    }
};

struct Graph {
    // Graph.buildGraph(k, kmersFreq) (Graph.scala:269-382)
    static MapGraph buildGraph(int k, const DNAMap &kmersFreq)
    {
        if (k != kmersFreq.k()) throw Error(GB_E_ARG, "k differs from the map's k");
        gb_graph *g = nullptr;
        check(gb_graph_build(kmersFreq.handle(), &g));
        return MapGraph(g, k);
    }
    // Graph(file) (Graph.scala:384-390): the Kryo `graph` file into device memory -- the graph of an empty map, filled by the
    // bulk addNode / addEdge of gb_graph_edit
    static MapGraph apply(const std::string &path, int device = 0)
    {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw Error(GB_E_ARG, "cannot read " + path);
        const std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        kryo::GraphArrays a;
        try {
            a = kryo::read(bytes);
        } catch (const std::runtime_error &e) {
            throw Error(GB_E_ARG, path + ": " + e.what());
        }
        if (a.nodeKmer.empty()) throw Error(GB_E_ARG, path + ": a graph file without nodes does not say its k");
        std::vector<uint64_t> off(a.edgeSeq.size() + 1, 0);
        std::vector<uint8_t> codes;
        for (size_t e = 0; e < a.edgeSeq.size(); e++) {
            codes.insert(codes.end(), a.edgeSeq[e].begin(), a.edgeSeq[e].end());
            off[e + 1] = codes.size();
        }
        DNAMap empty(a.k, 0, device);
        MapGraph g = buildGraph(a.k, empty);
        check(gb_graph_edit(g.handle(), 0, nullptr, nullptr, nullptr, (int64_t)a.nodeKmer.size(), a.nodeKmer.data(), (int64_t)a.edgeStart.size(),
                            a.edgeStart.data(), a.edgeEnd.data(), off.data(), codes.data(), 0, nullptr));
        return g;
    }
    // the sharded form of buildGraph (csrc/sgraph.cuh) over nShards virtual ranks on the map's one device: the multi-GPU
    // algorithm on a single GPU, same graph up to node / edge numbering
    static MapGraph buildGraphVirtualShards(int k, const DNAMap &kmersFreq, int nShards)
    {
        if (k != kmersFreq.k()) throw Error(GB_E_ARG, "k differs from the map's k");
        gb_graph *g = nullptr;
        check(gb_graph_build_virtual_shards(kmersFreq.handle(), nShards, &g));
        return MapGraph(g, k);
    }
};

} // namespace genome
