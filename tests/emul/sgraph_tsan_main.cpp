// sgraph_tsan_main.cpp -- TEST INFRASTRUCTURE: the thread-per-rank emulation of the sharded Graph.buildGraph
// (sgraph_emul.cpp: ThreadFabric) as a stand-alone program, to be built with -fsanitize=thread.  Every access of one rank to
// another rank's window that is not ordered by a Fabric barrier shows up as a data race: this checks the BARRIER PLACEMENT of
// sg::build, which the serial emulation cannot.  Usage: sg_tsan <k> <P> <file of u64 keys>
#include <stdio.h>

#include "sgraph_emul.cpp"

int main(int argc, char **argv)
{
    if (argc != 4) return 2;
    const int k = atoi(argv[1]), P = atoi(argv[2]);
    FILE *f = fopen(argv[3], "rb");
    if (!f) return 2;
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint64_t> keys((size_t)sz / 8);
    if (fread(keys.data(), 8, keys.size(), f) != keys.size()) return 2;
    fclose(f);
    std::vector<uint64_t> off((size_t)P + 1);
    for (int r = 0; r <= P; r++) off[(size_t)r] = keys.size() * (uint64_t)r / (uint64_t)P;
    uint64_t out[8];
    const int rc = emul_sharded_build_threads(k, 0, 0, P, keys.data(), off.data(), out, nullptr, nullptr, nullptr, nullptr, nullptr);
    printf("rc %d nodes %llu edges %llu bases %llu\n", rc, (unsigned long long)out[0], (unsigned long long)out[1], (unsigned long long)out[2]);
    return rc;
}
