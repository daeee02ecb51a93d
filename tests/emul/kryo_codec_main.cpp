// Test driver of hostcpp/kryo_graph.hpp (no GPU, no library): reads a Kryo `graph` file, decodes it into arrays, encodes the
// arrays again and writes the bytes.  tests/test_formats_cpu.py feeds it the files genome_b200/formats.py writes: the C++ codec
// must reproduce them byte for byte.   usage: kryo_codec <in> <out>     exit 1 + message on a malformed file
#include <cstdio>
#include <fstream>
#include <iterator>

#include "../../hostcpp/kryo_graph.hpp"

int main(int argc, char **argv)
{
    if (argc != 3) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    const std::vector<uint8_t> in((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    try {
        const genome::kryo::GraphArrays g = genome::kryo::read(in);
        const std::vector<uint8_t> out = genome::kryo::write(g);
        std::ofstream o(argv[2], std::ios::binary);
        o.write((const char *)out.data(), (std::streamsize)out.size());
        std::printf("k=%d nodes=%zu edges=%zu\n", g.k, g.nodeKmer.size(), g.edgeStart.size());
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
