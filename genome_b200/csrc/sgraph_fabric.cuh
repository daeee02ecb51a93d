// sgraph_fabric.cuh -- the collective interface of the sharded Graph.buildGraph (sgraph.cuh): what the algorithm needs from
// the P ranks' interconnect.  Implemented by LocalFabric (sgraph.cuh: all ranks in one process) and by the NCCL + CUDA-IPC
// fabric of comm.cu (one process per GPU).
#pragma once
#include <vector>

#include "common.cuh"

namespace gb {
namespace sg {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int MAXR = 16; // ranks of one sharded build (one NVSwitch box has 8)

struct PeerPtrs { void *p[MAXR]; };
struct Row { u64 v[MAXR]; };

// the P ranks of one build.  `mine` lists the ranks this process drives: exactly one in the one-process-per-GPU form, all
// P in the single-process form (virtual shards / emulation).  Every call is collective over the processes; arrays indexed
// [l] run over `mine`.
struct Fabric {
    int P = 1;
    std::vector<int> mine;
    virtual ~Fabric() {}
    // host data: every rank contributes `bytes`; all[l] receives P * bytes in rank order
    virtual int allgather_host(const void *const *contrib, void *const *all, size_t bytes) = 0;
    // peer-visible memory: window[l] of bytes_of_rank[mine[l]] bytes; peers[l].p[r] = rank r's window as seen from mine[l]
    virtual int windows(const size_t *bytes_of_rank, void **window, PeerPtrs *peers) = 0;
    // u64 elements, host-known counts: send[l] + soff[l].v[p] (scnt[l].v[p] elements) lands in rank p's recv at its roff.v[mine[l]]
    virtual int alltoallv_u64(const u64 *const *send, const Row *soff, const Row *scnt, u64 *const *recv, const Row *roff,
                              const Row *rcnt) = 0;
    // all device work issued so far by every rank is complete and visible to its peers before anything issued later starts
    virtual int barrier() = 0;
    // in place: elements [off[p], off[p] + cnt[p]) of buf[l] are rank p's; afterwards every rank holds all of them
    virtual int allgatherv_u64(u64 *const *buf, const u64 *off, const u64 *cnt) = 0;
    // the process's copy of a global output array: element-wise sum over the processes (one writer per element).  The
    // ranks of one process share their copy, so the single-process form has nothing to do.
    virtual int allreduce_sum(void *buf, size_t count, int elem_bytes) = 0;
    // phase marker of sg::build (debugging aid: the NCCL fabric prints elapsed times when gb_tune("trace") is set)
    virtual void tick(const char *) {}
};

} // namespace sg

// Graph.buildGraph over the ranks of `fab`; this process's ranks take their kept k-mers from maps[l] (fab.mine[l]) and run on
// `stream`; `dual` = some rank holds keys inserted as-is (both orientations may be stored).  Collective.  Defined in sgraph.cu.
int graph_build_on_fabric(sg::Fabric &fab, Map *const *maps, cudaStream_t stream, bool dual, gb_graph **out);

} // namespace gb
