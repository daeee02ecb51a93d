// graph_builder.cpp -- GraphBuilder.startup (S/scripts/GraphBuilder.scala:18-59, relative to /root/reference) written
// against the C++ host mirror: reads a `.bin` stream, counts, filters, builds, keeps the largest component, prints the
// log lines the Scala driver prints, writes the Kryo `graph` file when asked (GraphBuilder.scala:56).
// Usage: graph_builder <reads.bin> <n_pairs> <k> [min_capacity [graph_out]]
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <map>

#include "genome_host.hpp"

int main(int argc, char **argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: %s <reads.bin> <n_pairs> <k> [min_capacity [graph_out]]\n", argv[0]); return 2; }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        std::vector<uint8_t> bin((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        const int64_t pairs = std::atoll(argv[2]);
        const int k = std::atoi(argv[3]);
        const int rounds = 3; // GraphBuilder.scala:30
        genome::DNAMap kmersFreq(k, argc > 4 ? std::atoll(argv[4]) : 0);
        genome::FreqFilter::extractFilteredKmers(kmersFreq, bin.data(), bin.size(), pairs, rounds);
        std::printf("Good reads count: %lld\n", (long long)kmersFreq.size());
        genome::MapGraph graph = genome::Graph::buildGraph(k, kmersFreq);
        std::vector<uint32_t> label;
        int64_t nc = graph.components(label);
        int64_t total = 0;
        for (auto &e : graph.getEdges()) total += (int64_t)e.seq.size();
        std::printf("Total edges length: %lld\n", (long long)total);
        std::map<uint32_t, int64_t> size;
        for (uint32_t l : label) size[l]++;
        std::map<int64_t, int64_t> hist;
        int64_t best = 0;
        for (auto &p : size) { hist[p.second]++; if (p.second > best) best = p.second; }
        std::printf("Components histogram:");
        for (auto &p : hist) std::printf(" (%lld,%lld)", (long long)p.first, (long long)p.second);
        std::printf("\nMax component size: %lld (of %lld components)\n", (long long)best, (long long)nc);
        graph.retainLargest();
        if (argc > 5) graph.write(argv[5]);                        // graph.asInstanceOf[MapGraph].write(outfile), GraphBuilder.scala:56
        graph.simplifyGraph();
        std::printf("Graph nodes: %zu\n", graph.getNodes().size());
        std::printf("Node map: %zu\n", graph.getGraphMap().size()); // Graph.scala:117: nodeMap.size
    } catch (const genome::Error &e) {
        std::fprintf(stderr, "genome_b200 error %d: %s\n", e.code, e.what());
        return 1;
    }
    return 0;
}
