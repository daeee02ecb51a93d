// kryo_graph.hpp -- the Kryo `graph` file of the reference (MapGraph.write / Graph(file), S/data/graph/Graph.scala:232-261,
// 384-390; NodeSerializer S/data/graph/Node.scala:14-37; Edge S/data/graph/Edge.scala:11; paths relative to /root/reference):
// the file GraphBuilder hands to GraphSimplifier (S/scripts/GraphBuilder.scala:56, GraphSimplifier.scala:34).  Host I/O, no
// device work.  The serialiser is a third-party dependency that is absent from the reference tree
// (com.esotericsoftware.kryo:kryo:2.14-SNAPSHOT, project/Build.scala:39); what follows restates the wire rules of the Kryo 2.x
// line as published in its sources -- the same rules, spelled out at length, in genome_b200/formats.py, whose bytes this codec
// must reproduce (tests/test_formats_cpu.py compares them).  PARITY UNPINNED: no Kryo jar, no JVM, no graph file in the reference.
//   reference marker before every object that is not a primitive wrapper: 0 null, 1 first occurrence, n + 2 back reference
//   unregistered class: 1 (NAME + 2), name id (varint), and the first time its name (ASCII, bit 7 set on the last byte)
//   Output.writeInt / writeLong: 4 / 8 bytes big-endian; writeInt(v, true) / writeLong(v, true): 7 bits per byte, low groups
//   first, bit 7 = more, the last byte of a full-width value carries 8 bits; (v, false): zig-zag first
//   MapGraph (KryoSerializable): marker, writeInt(nodes.size), nodes, writeInt(edges.size), edges
//   Node (NodeSerializer): marker, writeLong(id), writeClassAndObject(seq), in-edge ids, (base byte, out-edge id) pairs
//   Edge (FieldSerializer, fields by name): marker, endId, id (zig-zag varlongs), seq (writeClassAndObject), startId
//   DNASeq (FieldSerializer): Long1DNASeq = len byte, long; Long2DNASeq = len, long1, long2; ArrayDNASeq = data (marker,
//   varint length + 1, bytes), length (zig-zag varint); <= 32 bases Long1, <= 64 Long2, else Array (DNASeq.scala:262-270)
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace genome {
namespace kryo {

struct GraphArrays {
    int k = 0;
    std::vector<uint64_t> nodeKmer;             // node i = k-mer nodeKmer[i], base j at bits 2j
    std::vector<uint32_t> edgeStart, edgeEnd;   // node indices
    std::vector<std::vector<uint8_t>> edgeSeq;  // base codes A0 G1 C2 T3
};

namespace detail {
inline const char *name(int cls)
{
    static const char *n[] = { "ru.ifmo.genome.dna.Long1DNASeq", "ru.ifmo.genome.dna.Long2DNASeq", "ru.ifmo.genome.dna.ArrayDNASeq" };
    return n[cls];
}
struct Writer {
    std::vector<uint8_t> out;
    int nameId[3] = { -1, -1, -1 }, names = 0;
    void fixed(uint64_t v, int bytes) { for (int i = bytes - 1; i >= 0; i--) out.push_back((uint8_t)(v >> (8 * i))); }
    void var(uint64_t v, int bits)
    {
        if (bits == 32) v &= 0xFFFFFFFFull;
        for (int i = 0; i < (bits == 32 ? 4 : 8); i++) {
            if ((v >> 7) == 0) { out.push_back((uint8_t)v); return; }
            out.push_back((uint8_t)((v & 0x7F) | 0x80));
            v >>= 7;
        }
        out.push_back((uint8_t)v);
    }
    void zig64(uint64_t v) { var((v << 1) ^ (uint64_t)((int64_t)v >> 63), 64); }
    void zig32(uint32_t v) { var((uint32_t)((v << 1) ^ (uint32_t)((int32_t)v >> 31)), 32); }
    void className(int cls)
    {
        out.push_back(1);
        if (nameId[cls] >= 0) { var((uint64_t)nameId[cls], 32); return; }
        nameId[cls] = names++;
        var((uint64_t)nameId[cls], 32);
        std::string s = name(cls);
        for (size_t i = 0; i < s.size(); i++) out.push_back((uint8_t)s[i] | (i + 1 == s.size() ? 0x80 : 0));
    }
    void seq(const uint8_t *codes, size_t n) // writeClassAndObject(out, seq: DNASeq)
    {
        if (n <= 64) {
            uint64_t w[2] = { 0, 0 };
            for (size_t i = 0; i < n; i++) w[i >> 5] |= (uint64_t)(codes[i] & 3) << (2 * (i & 31));
            className(n <= 32 ? 0 : 1);
            out.push_back(1);
            out.push_back((uint8_t)n);
            zig64(w[0]);
            if (n > 32) zig64(w[1]);
            return;
        }
        if (n >= (1ull << 31)) throw std::runtime_error("sequence longer than an Int");
        className(2);
        out.push_back(1);
        out.push_back(1);
        const size_t bytes = (n + 3) / 4;
        var(bytes + 1, 32);
        for (size_t b = 0; b < bytes; b++) {
            uint8_t v = 0;
            for (size_t j = 0; j < 4 && 4 * b + j < n; j++) v |= (uint8_t)((codes[4 * b + j] & 3) << (2 * j));
            out.push_back(v);
        }
        zig32((uint32_t)n);
    }
};
struct Reader {
    const std::vector<uint8_t> &b;
    size_t pos = 0;
    std::vector<std::string> names;
    explicit Reader(const std::vector<uint8_t> &bytes) : b(bytes) {}
    [[noreturn]] void fail(const std::string &what) const { throw std::runtime_error(what + " at byte " + std::to_string(pos)); }
    uint8_t byte() { if (pos >= b.size()) fail("truncated Kryo graph file"); return b[pos++]; }
    uint64_t fixed(int bytes) { uint64_t v = 0; for (int i = 0; i < bytes; i++) v = (v << 8) | byte(); return v; }
    uint64_t var(int bits)
    {
        uint64_t v = 0;
        int shift = 0;
        const int last = bits == 32 ? 4 : 8;
        for (int i = 0; i <= last; i++) {
            const uint8_t c = byte();
            if (i == last) { v |= (uint64_t)c << shift; break; }
            v |= (uint64_t)(c & 0x7F) << shift;
            if (!(c & 0x80)) break;
            shift += 7;
        }
        return bits == 32 ? (v & 0xFFFFFFFFull) : v;
    }
    uint64_t zig64() { const uint64_t v = var(64); return (v >> 1) ^ (0 - (v & 1)); }
    uint32_t zig32() { const uint32_t v = (uint32_t)var(32); return (v >> 1) ^ (0u - (v & 1u)); }
    void marker(const char *what) { if (var(32) != 1) fail(std::string(what) + ": a reference marker MapGraph.write does not emit"); }
    std::string className()
    {
        if (var(32) != 1) fail("registered class id: the reference registers no classes");
        const uint64_t id = var(32);
        if (id < names.size()) return names[(size_t)id];
        if (id != names.size()) fail("class name id out of order");
        std::string s;
        for (;;) {
            const uint8_t c = byte();
            s.push_back((char)(c & 0x7F));
            if (c & 0x80) break;
        }
        if (s.size() < 2) fail("UTF-8 class name: not produced for these classes");
        names.push_back(s);
        return s;
    }
    std::vector<uint8_t> seq()
    {
        const std::string cls = className();
        marker("sequence");
        std::vector<uint8_t> codes;
        if (cls == name(0) || cls == name(1)) {
            const int n = (int8_t)byte();
            uint64_t w[2] = { zig64(), 0 };
            if (cls == name(1)) w[1] = zig64();
            if (n < 0 || n > (cls == name(0) ? 32 : 64)) fail(cls + " of impossible length");
            for (int i = 0; i < n; i++) codes.push_back((uint8_t)((w[i >> 5] >> (2 * (i & 31))) & 3));
        } else if (cls == name(2)) {
            marker("ArrayDNASeq.data");
            const uint64_t lenp1 = var(32);
            if (lenp1 == 0) fail("null byte[] in an ArrayDNASeq");
            const size_t bytes = (size_t)lenp1 - 1;
            if (pos + bytes > b.size()) fail("truncated Kryo graph file");
            const size_t at = pos;
            pos += bytes;
            const uint32_t n = zig32();
            if (n > 4 * bytes) fail("ArrayDNASeq longer than its bytes");
            for (uint32_t i = 0; i < n; i++) codes.push_back((uint8_t)((b[at + i / 4] >> (2 * (i % 4))) & 3));
        } else {
            fail("unexpected class " + cls);
        }
        return codes;
    }
};
} // namespace detail

// graph.write(file): node / edge ids are index + 1 (the reference's AtomicLong generators start at 1; ids are not
// reproducible anyway); a node's in / out lists follow edge order, an out-edge is keyed by its first base
inline std::vector<uint8_t> write(const GraphArrays &g)
{
    const size_t N = g.nodeKmer.size(), E = g.edgeStart.size();
    std::vector<std::vector<uint64_t>> ins(N);
    std::vector<std::vector<std::pair<uint8_t, uint64_t>>> outs(N);
    for (size_t e = 0; e < E; e++) {
        if (g.edgeStart[e] >= N || g.edgeEnd[e] >= N || g.edgeSeq[e].empty()) throw std::runtime_error("edge does not fit the node list");
        for (auto &o : outs[g.edgeStart[e]])
            if (o.first == g.edgeSeq[e][0]) throw std::runtime_error("two out-edges of a node with the same first base: not a Map[Base, Long]");
        outs[g.edgeStart[e]].push_back({ g.edgeSeq[e][0], e + 1 });
        ins[g.edgeEnd[e]].push_back(e + 1);
    }
    detail::Writer w;
    w.out.push_back(1);
    w.fixed(N, 4);
    for (size_t i = 0; i < N; i++) {
        w.out.push_back(1);
        w.fixed(i + 1, 8);
        uint8_t codes[32];
        for (int j = 0; j < g.k; j++) codes[j] = (uint8_t)((g.nodeKmer[i] >> (2 * j)) & 3);
        w.seq(codes, (size_t)g.k);
        w.fixed(ins[i].size(), 4);
        for (uint64_t e : ins[i]) w.fixed(e, 8);
        w.fixed(outs[i].size(), 4);
        for (auto &o : outs[i]) { w.out.push_back(o.first); w.fixed(o.second, 8); }
    }
    w.fixed(E, 4);
    for (size_t e = 0; e < E; e++) {
        w.out.push_back(1);
        w.zig64((uint64_t)g.edgeEnd[e] + 1);
        w.zig64(e + 1);
        w.seq(g.edgeSeq[e].data(), g.edgeSeq[e].size());
        w.zig64((uint64_t)g.edgeStart[e] + 1);
    }
    return w.out;
}

// Graph(file): ids become positions; what Graph.read relies on is checked (unique ids, edges naming nodes of the file, one k)
inline GraphArrays read(const std::vector<uint8_t> &bytes)
{
    detail::Reader r(bytes);
    GraphArrays g;
    r.marker("MapGraph");
    std::map<uint64_t, uint32_t> index;
    const uint64_t N = r.fixed(4);
    for (uint64_t i = 0; i < N; i++) {
        r.marker("Node");
        const uint64_t id = r.fixed(8);
        const std::vector<uint8_t> codes = r.seq();
        for (uint64_t n = r.fixed(4); n > 0; n--) r.fixed(8);
        for (uint64_t n = r.fixed(4); n > 0; n--) { r.byte(); r.fixed(8); }
        if (codes.empty() || codes.size() > 31 || (i && (int)codes.size() != g.k)) r.fail("node sequences of different or impossible lengths");
        g.k = (int)codes.size();
        uint64_t x = 0;
        for (size_t j = 0; j < codes.size(); j++) x |= (uint64_t)codes[j] << (2 * j);
        if (!index.emplace(id, (uint32_t)i).second) r.fail("duplicate node id");
        g.nodeKmer.push_back(x);
    }
    const uint64_t E = r.fixed(4);
    std::map<uint64_t, bool> seen;
    for (uint64_t e = 0; e < E; e++) {
        r.marker("Edge");
        const uint64_t end = r.zig64(), id = r.zig64();
        std::vector<uint8_t> codes = r.seq();
        const uint64_t start = r.zig64();
        if (!seen.emplace(id, true).second) r.fail("duplicate edge id");
        if (!index.count(start) || !index.count(end)) r.fail("edge names a node id the file does not hold");
        g.edgeStart.push_back(index[start]);
        g.edgeEnd.push_back(index[end]);
        g.edgeSeq.push_back(std::move(codes));
    }
    if (r.pos != bytes.size()) r.fail("bytes after the graph");
    return g;
}

} // namespace kryo
} // namespace genome
