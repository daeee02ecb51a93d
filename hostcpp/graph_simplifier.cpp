// graph_simplifier.cpp -- GraphBuilder.startup followed by GraphSimplifier.startup (S/scripts/GraphSimplifier.scala:138-357,
// relative to /root/reference) against the C++ host mirror.  The reference passes the graph between the two scripts as a
// Kryo file (Graph(infile), GraphSimplifier.scala:34): with `--graph <file>` as the first two arguments this driver reads the
// file graph_builder wrote (then <k> is taken from the file and the argument is ignored); without, it runs both stages in
// one process.
// Usage: graph_simplifier [--graph <graph file>] <reads.bin> <n_pairs> <k> <cutoff> [range_first range_last [contigs_file]]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <map>

#include "genome_host.hpp"

int main(int argc, char **argv)
{
    const char *graphFile = nullptr;
    if (argc > 2 && std::string(argv[1]) == "--graph") {
        graphFile = argv[2];
        argv += 2;
        argc -= 2;
    }
    if (argc < 5) { std::fprintf(stderr, "usage: %s [--graph <graph file>] <reads.bin> <n_pairs> <k> <cutoff> [range_first range_last [contigs]]\n", argv[0]); return 2; }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        std::vector<uint8_t> bin((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        const int64_t pairs = std::atoll(argv[2]);
        int k = std::atoi(argv[3]);
        const int cutoff = std::atoi(argv[4]);                    // genome.cutoff (application.conf:71)
        const int first = argc > 6 ? std::atoi(argv[5]) : 180;    // implicit val range = 180 to 250 (GraphSimplifier.scala:153)
        const int last = argc > 6 ? std::atoi(argv[6]) : 250;
        auto builderStage = [&]() {
            genome::DNAMap kmersFreq(k);
            genome::FreqFilter::extractFilteredKmers(kmersFreq, bin.data(), bin.size(), pairs, 3);
            genome::MapGraph g = genome::Graph::buildGraph(k, kmersFreq);
            g.retainLargest();                                     // GraphBuilder.scala:52-54
            return g;
        };
        genome::MapGraph graph = graphFile ? genome::Graph::apply(graphFile) : builderStage();
        if (graphFile) k = graph.k();                              // val k = graph.getNodes.head.seq.length (GraphSimplifier.scala:36)
        std::printf("K = %d\n", k);
        auto ps = graph.pairSupport(bin.data(), bin.size(), pairs, first, last);
        std::printf("Bad pairs: %lld\n", (long long)ps.badPairs);
        std::printf("Edges before: %zu\n", graph.getEdges().size());
        graph.splitNodes(ps.support, cutoff);
        graph.simplifyGraph();
        auto edges = graph.getEdges();
        std::printf("Edges after: %zu\n", edges.size());
        int64_t total = 0;
        std::map<size_t, int64_t> lengths;
        for (auto &e : edges) { total += (int64_t)e.seq.size(); lengths[e.seq.size()]++; }
        std::printf("Total edges length: %lld\n", (long long)total);
        if (argc > 7) {
            std::FILE *out = std::fopen(argv[7], "w");
            if (!out) { std::perror(argv[7]); return 1; }
            static const char code[] = "AGCT";                     // Base.scala:13-16
            int i = 0;
            for (auto &e : edges) {                                // 338-347: the sequence line comes before its header
                for (uint8_t b : e.seq) std::fputc(code[b], out);
                std::fprintf(out, "\n>abacaba%d\n", i++);
            }
            std::fclose(out);
        }
    } catch (const genome::Error &e) {
        std::fprintf(stderr, "genome_b200 error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
