// comm.cu -- PartitionedDNAMap on one 8xB200 box: one hash shard per GPU (one rank = one process = one GPU).
// Replaces S/ds/PartitionedDNAMap.scala:15-63 (paths relative to /root/reference).
//
// Sharded insert, per batch of reads (all ranks in lock step, same number of batches), see pmap_insert:
//   communicator stream (high priority):
//     part_count                : canonical k-mers of the batch, histogram by (owner, table slice)      [partition.cu]
//     part_scatter_kernel<PEER> : every k-mer is stored straight into its owner's NVLink inbox through the peer
//                                 mapping (CUDA IPC) -- the all-to-all happens inside the bucket pass
//     counts all-to-all (NCCL)  : tiny; travels after the keys on the sender's stream, so it doubles as "data ready"
//   map stream:
//     [two-level routing only: re-bucket the received keys by fine table slice]
//     insert_keys_kernel        : update(key, 1, _ + 1) for the received keys, slice by slice (L2-blocked)
// Three inbox/buffer sets; batch b + 1 is bucketed and exchanged while batch b is upserted.
// gb_tune a2a = 1 replaces the peer stores by staged ncclSend/ncclRecv segments (what a box without peer access takes).
#include <dlfcn.h>
#include <time.h>
#include <nccl.h> // types only: the library is bound at run time (see NcclApi)

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "extract.cuh"
#include "partition.cuh"
#include "sgraph.cuh"
#include "sgraph_fabric.cuh"

namespace gb {


// buffer sets of the sharded insert (staging, inbox, events), used round-robin by the batches of one call; they live as long as the
// communicator.  A rank that has my totals of batch j knows my upsert of j - 2 is over, so 3 sets would do for the form that waits
// for the totals before it queues the next batch; the single-pass form queues the bucket pass of batch j + 1 first: one more.
constexpr int NSETS = 4;

struct BatchBufs {
    unsigned long long *send = nullptr, *recv = nullptr;
    size_t send_cap = 0, recv_cap = 0;
    unsigned long long *d_tot = nullptr, *h_tot = nullptr; // bucket totals (mine, received) and the upsert's chunk table
    PartWork work;  // bucket pass of the outgoing keys (communicator stream)
    PartWork work2; // re-bucketing of the received keys by fine table slice (map stream; two-level routing)
    cudaEvent_t exchanged = nullptr, inserted = nullptr;
    bool in_flight = false;
    int ensure(size_t ns, size_t nr);
    void release();
};

struct Comm {
    int rank = 0, n_ranks = 1, device = 0;
    ncclComm_t nccl = nullptr;
    cudaStream_t stream = nullptr; // bucketing + collectives
    BatchBufs bufs[NSETS];
    unsigned long long *d_scratch = nullptr, *h_scratch = nullptr;
    // NVLink inboxes (fused routing): inbox[i] holds n_ranks regions of region_cap keys, region s is written by rank s
    // through its peer mapping peer_inbox[i][me] obtained with CUDA IPC
    unsigned long long *inbox[NSETS] = {};
    unsigned long long *peer_inbox[NSETS][MAX_RANKS];
    size_t region_cap = 0;
    int p2p = -1; // -1 unknown, 0 NCCL send/recv staging, 1 peer stores
    // peer window of the sharded graph build (sgraph.cuh): keys, index, masks and vertex entries of this rank, mapped by all
    void *window = nullptr;
    void *peer_window[MAX_RANKS];
    size_t window_cap_of[MAX_RANKS]; // capacity of every rank's window (all ranks track all: the sizes are global knowledge)
};

// NCCL is bound with dlopen at the first gb_comm_* call, not at link time: a process that already holds an NCCL
// (a JVM with its own, or Python with torch's bundled libnccl.so.2) shares that copy through the SONAME, and a
// single-GPU user never needs NCCL at all.  GENOME_B200_NCCL names an explicit path.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.handle) return GB_OK;
    const char *names[] = { getenv("GENOME_B200_NCCL"), "libnccl.so.2", "libnccl.so" };
    void *h = nullptr;
    for (const char *n : names)
        if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) { set_error("cannot load NCCL (libnccl.so.2): %s", dlerror()); return GB_E_NCCL; }
    NcclApi a;
    a.handle = h;
#define GB_SYM(field, name)                                                       \
    *(void **)(&a.field) = dlsym(h, name);                                        \
    if (!a.field) { set_error("NCCL symbol %s missing", name); return GB_E_NCCL; }
    GB_SYM(GetUniqueId, "ncclGetUniqueId")
    GB_SYM(CommInitRank, "ncclCommInitRank")
    GB_SYM(CommDestroy, "ncclCommDestroy")
    GB_SYM(GetErrorString, "ncclGetErrorString")
    GB_SYM(GroupStart, "ncclGroupStart")
    GB_SYM(GroupEnd, "ncclGroupEnd")
    GB_SYM(Send, "ncclSend")
    GB_SYM(Recv, "ncclRecv")
    GB_SYM(AllReduce, "ncclAllReduce")
    GB_SYM(AllGather, "ncclAllGather")
    GB_SYM(Broadcast, "ncclBroadcast")
#undef GB_SYM
    g_nccl = a;
    return GB_OK;
}
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclGetErrorString g_nccl.GetErrorString
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclAllReduce g_nccl.AllReduce
#define ncclAllGather g_nccl.AllGather
#define ncclBroadcast g_nccl.Broadcast

static int nccl_fail(ncclResult_t r, const char *what, const char *file, int line)
{
    set_error("NCCL error %d (%s) at %s:%d: %s", (int)r, ncclGetErrorString(r), file, line, what);
    return GB_E_NCCL;
}
#define GB_NCCL(expr)                                                              \
    do {                                                                           \
        ncclResult_t _r = (expr);                                                  \
        if (_r != ncclSuccess) return gb::nccl_fail(_r, #expr, __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------- host helpers

// variable all-to-all of u64 elements: send segment p (send_off[p], send_cnt[p]) goes to rank p
static int all_to_all_v(Comm *c, const unsigned long long *send, const unsigned long long *send_off, const unsigned long long *send_cnt,
                        unsigned long long *recv, const unsigned long long *recv_off, const unsigned long long *recv_cnt,
                        ncclDataType_t type)
{
    GB_NCCL(ncclGroupStart());
    for (int p = 0; p < c->n_ranks; p++) {
        if (send_cnt[p]) GB_NCCL(ncclSend(send + send_off[p], send_cnt[p], type, p, c->nccl, c->stream));
        if (recv_cnt[p]) GB_NCCL(ncclRecv(recv + recv_off[p], recv_cnt[p], type, p, c->nccl, c->stream));
    }
    GB_NCCL(ncclGroupEnd());
    return GB_OK;
}

// fixed-size all-to-all: `row` u64 per pair, row p of d_send goes to rank p, row p of d_recv comes from rank p
static int all_to_all_rows(Comm *c, const unsigned long long *d_send, unsigned long long *d_recv, size_t row, cudaStream_t st = nullptr)
{
    if (!st) st = c->stream;
    GB_NCCL(ncclGroupStart());
    for (int p = 0; p < c->n_ranks; p++) {
        GB_NCCL(ncclSend(d_send + (size_t)p * row, row, ncclUint64, p, c->nccl, st));
        GB_NCCL(ncclRecv(d_recv + (size_t)p * row, row, ncclUint64, p, c->nccl, st));
    }
    GB_NCCL(ncclGroupEnd());
    return GB_OK;
}

static int comm_scratch(Comm *c)
{
    if (c->d_scratch) return GB_OK;
    GB_CUDA(cudaMalloc((void **)&c->d_scratch, 4 * MAX_BUCKETS * 8));
    GB_CUDA(cudaHostAlloc((void **)&c->h_scratch, 4 * MAX_BUCKETS * 8, cudaHostAllocDefault));
    return GB_OK;
}

// every rank contributes P counts; afterwards recv_cnt[p] = what rank p sends to me
static int exchange_counts(Comm *c, unsigned long long *d_send_cnt, unsigned long long *d_recv_cnt,
                           unsigned long long *h_send_cnt, unsigned long long *h_recv_cnt)
{
    const int P = c->n_ranks;
    GB_TRY(all_to_all_rows(c, d_send_cnt, d_recv_cnt, 1));
    GB_CUDA(cudaMemcpyAsync(h_send_cnt, d_send_cnt, P * 8, cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaMemcpyAsync(h_recv_cnt, d_recv_cnt, P * 8, cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    return GB_OK;
}

static int all_reduce_i64(Comm *c, int64_t *v, ncclRedOp_t op)
{
    GB_TRY(comm_scratch(c));
    GB_CUDA(cudaMemcpyAsync(c->d_scratch, v, 8, cudaMemcpyHostToDevice, c->stream));
    GB_NCCL(ncclAllReduce(c->d_scratch, c->d_scratch, 1, ncclInt64, op, c->nccl, c->stream));
    GB_CUDA(cudaMemcpyAsync(v, c->d_scratch, 8, cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    return GB_OK;
}

int BatchBufs::ensure(size_t ns, size_t nr)
{
    if (ns > send_cap) {
        if (send) GB_CUDA(cudaFree(send));
        send = nullptr;
        send_cap = ns + ns / 8 + 1024;
        GB_CUDA(cudaMalloc((void **)&send, send_cap * 8));
    }
    if (nr > recv_cap) {
        if (recv) GB_CUDA(cudaFree(recv));
        recv = nullptr;
        recv_cap = nr + nr / 8 + 1024;
        GB_CUDA(cudaMalloc((void **)&recv, recv_cap * 8));
    }
    if (!d_tot) {
        GB_CUDA(cudaMalloc((void **)&d_tot, (4 * MAX_BUCKETS + 8) * 8));
        GB_CUDA(cudaHostAlloc((void **)&h_tot, (4 * MAX_BUCKETS + 8) * 8, cudaHostAllocDefault));
        GB_CUDA(cudaEventCreateWithFlags(&exchanged, cudaEventDisableTiming));
        GB_CUDA(cudaEventCreateWithFlags(&inserted, cudaEventDisableTiming));
    }
    return GB_OK;
}

void BatchBufs::release()
{
    if (send) cudaFree(send);
    if (recv) cudaFree(recv);
    if (d_tot) cudaFree(d_tot);
    if (h_tot) cudaFreeHost(h_tot);
    if (exchanged) cudaEventDestroy(exchanged);
    if (inserted) cudaEventDestroy(inserted);
    work.release();
    work2.release();
    send = recv = d_tot = h_tot = nullptr;
    exchanged = inserted = nullptr;
}

static void close_inboxes(Comm *c)
{
    for (int i = 0; i < NSETS; i++) {
        for (int p = 0; p < c->n_ranks; p++)
            if (p != c->rank && c->peer_inbox[i][p]) cudaIpcCloseMemHandle(c->peer_inbox[i][p]);
        if (c->inbox[i]) cudaFree(c->inbox[i]);
        c->inbox[i] = nullptr;
        for (int p = 0; p < MAX_RANKS; p++) c->peer_inbox[i][p] = nullptr;
    }
    c->region_cap = 0;
}

// collective: make sure every rank's three inboxes hold n_ranks regions of at least `want` keys and that every rank
// has them mapped.  Returns with c->p2p = 1, or 0 when peer access is not possible (then NCCL send/recv is used).
static int ensure_inboxes(Comm *c, size_t want)
{
    const int P = c->n_ranks;
    if (c->p2p == 0) return GB_OK;
    if (c->p2p < 0) {
        int64_t ok = 1;
        if (g_tune.a2a == 1) ok = 0;
        int ndev = 0;
        cudaGetDeviceCount(&ndev);
        // ranks are the visible devices 0..P-1 of one box (one process per GPU): all of them must be peers of mine
        for (int d = 0; ok && d < P; d++) {
            if (d == c->device) continue;
            int can = 0;
            if (d >= ndev || cudaDeviceCanAccessPeer(&can, c->device, d) != cudaSuccess || !can) ok = 0;
        }
        cudaGetLastError();
        GB_TRY(all_reduce_i64(c, &ok, ncclMin));
        c->p2p = ok ? 1 : 0;
        if (!c->p2p) return GB_OK;
    }
    int64_t need = (int64_t)want;
    GB_TRY(all_reduce_i64(c, &need, ncclMax));
    if ((size_t)need <= c->region_cap) return GB_OK;
    GB_CUDA(cudaDeviceSynchronize());
    close_inboxes(c);
    const size_t cap = (size_t)need + (size_t)need / 16 + 1024;
    GB_TRY(comm_scratch(c));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    for (int i = 0; i < NSETS; i++) {
        GB_CUDA(cudaMalloc((void **)&c->inbox[i], cap * P * 8));
        cudaIpcMemHandle_t mine;
        GB_CUDA(cudaIpcGetMemHandle(&mine, c->inbox[i]));
        // all-gather the 64-byte handles through device memory
        unsigned long long *d_mine = c->d_scratch, *d_all = c->d_scratch + 8;
        GB_CUDA(cudaMemcpyAsync(d_mine, &mine, 64, cudaMemcpyHostToDevice, c->stream));
        GB_NCCL(ncclAllGather(d_mine, d_all, 8, ncclUint64, c->nccl, c->stream));
        std::vector<cudaIpcMemHandle_t> all((size_t)P);
        GB_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t)P * 64, cudaMemcpyDeviceToHost, c->stream));
        GB_CUDA(cudaStreamSynchronize(c->stream));
        for (int p = 0; p < P; p++) {
            if (p == c->rank) { c->peer_inbox[i][p] = c->inbox[i]; continue; }
            void *ptr = nullptr;
            GB_CUDA(cudaIpcOpenMemHandle(&ptr, all[(size_t)p], cudaIpcMemLazyEnablePeerAccess));
            c->peer_inbox[i][p] = (unsigned long long *)ptr;
        }
    }
    c->region_cap = cap;
    // nobody may write into a peer's inbox before that peer has finished mapping: one more collective as a barrier
    int64_t one = 1;
    GB_TRY(all_reduce_i64(c, &one, ncclSum));
    return GB_OK;
}

static void close_windows(Comm *c)
{
    for (int p = 0; p < c->n_ranks; p++)
        if (p != c->rank && c->peer_window[p]) cudaIpcCloseMemHandle(c->peer_window[p]);
    if (c->window) cudaFree(c->window);
    c->window = nullptr;
    for (int p = 0; p < MAX_RANKS; p++) { c->peer_window[p] = nullptr; c->window_cap_of[p] = 0; }
}

// The fabric of the sharded Graph.buildGraph (sgraph_fabric.cuh) for one rank per GPU: NCCL for the collectives, CUDA IPC for
// the peer windows (NVLink loads / stores from inside the kernels).  Everything runs on the communicator's stream.
struct NcclFabric : sg::Fabric {
    Comm *c;
    explicit NcclFabric(Comm *comm) : c(comm)
    {
        P = comm->n_ranks;
        mine.push_back(comm->rank);
    }
    int allgather_host(const void *const *contrib, void *const *all, size_t bytes) override
    {
        GB_TRY(comm_scratch(c));
        if (bytes > 256 || bytes * (size_t)P > 4096) { set_error("allgather_host: %zu bytes", bytes); return GB_E_ARG; }
        uint8_t *d_mine = (uint8_t *)c->d_scratch, *d_all = (uint8_t *)c->d_scratch + 256;
        GB_CUDA(cudaMemcpyAsync(d_mine, contrib[0], bytes, cudaMemcpyHostToDevice, c->stream));
        GB_NCCL(ncclAllGather(d_mine, d_all, bytes, ncclUint8, c->nccl, c->stream));
        GB_CUDA(cudaMemcpyAsync(all[0], d_all, bytes * (size_t)P, cudaMemcpyDeviceToHost, c->stream));
        GB_CUDA(cudaStreamSynchronize(c->stream));
        return GB_OK;
    }
    // every rank knows every rank's size, so all of them agree without talking on whether any window has to grow; if one
    // does, the mappings are rebuilt everywhere (rare: the windows are kept across builds and grow with 25 % slack)
    int windows(const size_t *bytes_of_rank, void **window, sg::PeerPtrs *peers) override
    {
        bool grow = false;
        for (int r = 0; r < P; r++) grow = grow || bytes_of_rank[r] > c->window_cap_of[r];
        if (grow) {
            GB_CUDA(cudaDeviceSynchronize());
            GB_TRY(barrier_sync()); // nobody is still reading a window that is about to go away
            const bool mine_grows = bytes_of_rank[c->rank] > c->window_cap_of[c->rank];
            for (int p = 0; p < P; p++)
                if (p != c->rank && c->peer_window[p]) { cudaIpcCloseMemHandle(c->peer_window[p]); c->peer_window[p] = nullptr; }
            for (int r = 0; r < P; r++)
                if (bytes_of_rank[r] > c->window_cap_of[r]) c->window_cap_of[r] = bytes_of_rank[r] + bytes_of_rank[r] / 4 + 4096;
            if (mine_grows) {
                if (c->window) GB_CUDA(cudaFree(c->window));
                c->window = nullptr;
                GB_CUDA(cudaMalloc(&c->window, c->window_cap_of[c->rank]));
            }
            GB_TRY(comm_scratch(c));
            cudaIpcMemHandle_t h;
            GB_CUDA(cudaIpcGetMemHandle(&h, c->window));
            unsigned long long *d_mine = c->d_scratch, *d_all = c->d_scratch + 8;
            GB_CUDA(cudaMemcpyAsync(d_mine, &h, 64, cudaMemcpyHostToDevice, c->stream));
            GB_NCCL(ncclAllGather(d_mine, d_all, 8, ncclUint64, c->nccl, c->stream));
            std::vector<cudaIpcMemHandle_t> all((size_t)P);
            GB_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t)P * 64, cudaMemcpyDeviceToHost, c->stream));
            GB_CUDA(cudaStreamSynchronize(c->stream));
            for (int p = 0; p < P; p++) {
                if (p == c->rank) { c->peer_window[p] = c->window; continue; }
                void *ptr = nullptr;
                GB_CUDA(cudaIpcOpenMemHandle(&ptr, all[(size_t)p], cudaIpcMemLazyEnablePeerAccess));
                c->peer_window[p] = ptr;
            }
        }
        GB_TRY(barrier_sync()); // every rank has its mappings, and has finished with the windows' previous contents
        window[0] = c->window;
        for (int p = 0; p < P; p++) peers[0].p[p] = c->peer_window[p];
        return GB_OK;
    }
    int alltoallv_u64(const sg::u64 *const *send, const sg::Row *soff, const sg::Row *scnt, sg::u64 *const *recv, const sg::Row *roff,
                      const sg::Row *rcnt) override
    {
        return all_to_all_v(c, send[0], soff[0].v, scnt[0].v, recv[0], roff[0].v, rcnt[0].v, ncclUint64);
    }
    // stream-ordered: my later kernels start after every rank has reached this point of its own stream
    int barrier() override
    {
        GB_TRY(comm_scratch(c));
        GB_NCCL(ncclAllReduce(c->d_scratch + 1024, c->d_scratch + 1024, 1, ncclUint64, ncclSum, c->nccl, c->stream));
        return GB_OK;
    }
    int barrier_sync()
    {
        GB_TRY(barrier());
        GB_CUDA(cudaStreamSynchronize(c->stream));
        return GB_OK;
    }
    int allgatherv_u64(sg::u64 *const *buf, const sg::u64 *off, const sg::u64 *cnt) override
    {
        GB_NCCL(ncclGroupStart());
        for (int p = 0; p < P; p++)
            if (cnt[p]) GB_NCCL(ncclBroadcast(buf[0] + off[p], buf[0] + off[p], cnt[p], ncclUint64, p, c->nccl, c->stream));
        GB_NCCL(ncclGroupEnd());
        return GB_OK;
    }
    int allreduce_sum(void *buf, size_t count, int elem_bytes) override
    {
        if (count) GB_NCCL(ncclAllReduce(buf, buf, count, elem_bytes == 8 ? ncclUint64 : ncclUint32, ncclSum, c->nccl, c->stream));
        return GB_OK;
    }
    double t_last = 0;
    void tick(const char *what) override
    {
        if (!g_tune.trace) return;
        cudaStreamSynchronize(c->stream);
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        const double now = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
        if (c->rank == 0 && t_last > 0) fprintf(stderr, "[sgraph] %-32s +%8.3f ms\n", what, now - t_last);
        t_last = now;
    }
};

// owner of a key: a prefix of its slot hash (orientation-blind only through FreqFilter's canonical rule: keys arrive canonical)
static inline int32_t pmap_owner_of(const Map *, unsigned long long key, int P)
{
    return (int32_t)owner_of(mix64(key), (unsigned int)P);
}

// wait for every in-flight insert, fold the new-key counter into m->size
static int drain(Map *m, BatchBufs *bufs)
{
    GB_CUDA(cudaStreamSynchronize(m->stream));
    unsigned long long c[4];
    GB_TRY(map_read_counters(m, c));
    m->size += (int64_t)c[0];
    GB_TRY(map_zero_counters(m));
    GB_CUDA(cudaStreamSynchronize(m->stream));
    for (int i = 0; i < NSETS; i++) bufs[i].in_flight = false;
    return GB_OK;
}

// The keys of an overflow list (rare: a bucket far above its share inside one CTA's tiles) get to their owners by the staged route:
// owner-major split of the list, counts and keys through NCCL, plain chunk upsert.  Collective: every rank calls it for the batch
// (with an empty list if it has none).
static int route_key_list(Map *m, Comm *c, BatchBufs &B, const unsigned long long *d_list, unsigned long long n_local, int64_t *received)
{
    const int P = c->n_ranks;
    PartLayout po;
    po.owners = P;
    po.lp_bits = 0;
    unsigned long long *d_desc = B.d_tot + 3 * MAX_BUCKETS; // { vstart[0], vstart[1], off[0] } of my list | the same for what I receive
    unsigned long long h_desc[3] = { 0, n_local, 0 };
    GB_CUDA(cudaMemcpyAsync(d_desc, h_desc, 24, cudaMemcpyHostToDevice, c->stream));
    GB_TRY(B.ensure(0, (size_t)n_local));
    KeySource ks;
    ks.keys = d_list; ks.vstart = d_desc; ks.off = d_desc + 2; ks.n_chunks = 1; ks.n_total = n_local;
    GB_TRY(part_count_keys(ks, po, B.work2, c->stream));
    GB_TRY(part_scatter_keys(ks, po, B.work2, B.recv, c->stream));
    unsigned long long scnt[MAX_RANKS], rcnt[MAX_RANKS], soff[MAX_RANKS], roff[MAX_RANKS];
    GB_TRY(comm_scratch(c));
    GB_TRY(exchange_counts(c, B.work2.bucket_total, c->d_scratch + 1536, scnt, rcnt)); // synchronises the communicator's stream
    unsigned long long st = 0, rt = 0;
    for (int p = 0; p < P; p++) { soff[p] = st; st += scnt[p]; roff[p] = rt; rt += rcnt[p]; }
    DeviceBuf d_in;
    GB_TRY(d_in.alloc((size_t)(rt + 1) * 8));
    GB_TRY(all_to_all_v(c, B.recv, soff, scnt, (unsigned long long *)d_in.p, roff, rcnt, ncclUint64));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    if (rt) {
        int64_t budget = 0;
        GB_CUDA(cudaStreamSynchronize(m->stream));
        unsigned long long cn[4];
        GB_TRY(map_read_counters(m, cn));
        m->size += (int64_t)cn[0];
        GB_TRY(map_zero_counters(m));
        GB_TRY(map_budget(m, (int64_t)rt, &budget));
        while (budget < (int64_t)rt) {
            GB_TRY(map_rebuild(m, m->cap * 2, false, 0));
            budget = (int64_t)(m->cap * 0.9) - m->size;
        }
        unsigned long long h_desc2[3] = { 0, rt, 0 };
        GB_CUDA(cudaMemcpyAsync(d_desc + 4, h_desc2, 24, cudaMemcpyHostToDevice, m->stream));
        GB_TRY(insert_key_chunks(m, (const unsigned long long *)d_in.p, d_desc + 4, d_desc + 6, 1, rt, m->stream));
        GB_CUDA(cudaStreamSynchronize(m->stream)); // d_in goes away
    }
    *received = (int64_t)rt;
    return GB_OK;
}

// Sharded FreqFilter.add, single-pass form (fixed-stride streams, one routing level, peer-mapped inboxes): NO count pass and no
// per-bucket counts on the host.  Per batch, on the communicator's stream: bucket_slabs_kernel<PEER> stores every canonical k-mer
// into a slab of its owner's inbox (slab = (source rank, table slice, CTA), sized for the CTA's share + 8 sigma) and leaves the
// slabs' fill counts beside them; one 8-byte-per-pair all-to-all follows the keys on the stream ("whoever has my total has my
// keys") and tells the owner how many keys arrived, which is all the host needs (room in the table).  On the map's stream the
// owner upserts the slabs slice-major.  The next batch's bucket pass is queued before the host waits for this batch's totals.
static int pmap_insert_slabs(Map *m, const uint8_t *d_bin, size_t n_bytes, unsigned int rec, int64_t win_max, int64_t n_reads, int64_t batch_reads,
                             int64_t batches, const PartLayout &pl, unsigned int slab, int64_t *n_windows)
{
    Comm *c = m->comm;
    const int P = c->n_ranks, k = m->k, LP = 1 << pl.lp_bits;
    BatchBufs *bufs = c->bufs;
    for (int i = 0; i < NSETS; i++) GB_TRY(bufs[i].work.ensure(c->stream));
    const unsigned int G = (unsigned int)bufs[0].work.grid;
    const unsigned long long cnt_off = (unsigned long long)LP * G * slab;
    const bool trace = g_tune.trace && c->rank == 0;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_begin = now_ms();
    if (trace) fprintf(stderr, "[pmap] single-pass: %lld reads, %lld batches of %lld, %d slices, slab %u keys, region %llu keys\n", (long long)n_reads,
                       (long long)batches, (long long)batch_reads, LP, slab, (unsigned long long)c->region_cap);
    GB_TRY(map_zero_counters(m));
    GB_CUDA(cudaStreamSynchronize(m->stream));
    GB_CUDA(cudaEventRecord(m->ev0, m->stream));
    int64_t windows = 0, pending_upper = 0;
    // d_tot / h_tot: [0, P) keys I put into owner o's slabs | [P, 2P) keys source s put into mine | [2P] overflow cursor | [2P + 1] its
    // "did not fit" flag | [2P + 2] overflow keys on all ranks
    auto stage_a = [&](int64_t b) -> int {
        BatchBufs &B = bufs[b % NSETS];
        if (b >= 2 && bufs[(b - 2) % NSETS].in_flight) GB_CUDA(cudaStreamWaitEvent(c->stream, bufs[(b - 2) % NSETS].inserted, 0));
        if (B.in_flight) GB_CUDA(cudaStreamWaitEvent(c->stream, B.inserted, 0));
        const int64_t r0 = std::min(n_reads, b * batch_reads), r1 = std::min(n_reads, r0 + batch_reads), nr = r1 - r0;
        GB_TRY(B.ensure((size_t)(nr * win_max) + 8, 0)); // the overflow list can take the whole batch
        GB_CUDA(cudaMemsetAsync(B.d_tot, 0, (size_t)(2 * P + 4) * 8, c->stream));
        PeerSlabs ps;
        memset(&ps, 0, sizeof ps);
        for (int o = 0; o < P; o++) {
            ps.keys[o] = c->peer_inbox[b % NSETS][o] + (size_t)c->rank * c->region_cap;
            ps.cnt[o] = reinterpret_cast<unsigned int *>(ps.keys[o] + cnt_off);
        }
        ps.owner_total = B.d_tot;
        ps.owners = (unsigned int)P;
        ReadBatch rb;
        rb.bin = d_bin; rb.n_bytes = n_bytes; rb.offsets = nullptr; rb.rec_bytes = rec; rb.read0 = r0; rb.n_reads = nr;
        GB_TRY(bucket_slabs_peers(rb, k, m->v210, pl, B.work, slab, ps, B.send, B.send_cap, B.d_tot + 2 * P, reinterpret_cast<unsigned int *>(B.d_tot + 2 * P + 1),
                                  c->stream));
        GB_TRY(all_to_all_rows(c, B.d_tot, B.d_tot + P, 1)); // the totals travel AFTER the keys on my stream
        GB_NCCL(ncclAllReduce(B.d_tot + 2 * P, B.d_tot + 2 * P + 2, 1, ncclUint64, ncclSum, c->nccl, c->stream));
        GB_CUDA(cudaMemcpyAsync(B.h_tot, B.d_tot, (size_t)(2 * P + 4) * 8, cudaMemcpyDeviceToHost, c->stream));
        GB_CUDA(cudaEventRecord(B.exchanged, c->stream));
        return GB_OK;
    };
    auto stage_b = [&](int64_t b) -> int {
        BatchBufs &B = bufs[b % NSETS];
        GB_CUDA(cudaEventSynchronize(B.exchanged));
        unsigned long long sent = 0, rt = 0;
        for (int p = 0; p < P; p++) { sent += B.h_tot[p]; rt += B.h_tot[P + p]; }
        const unsigned long long ovf_mine = B.h_tot[2 * P], ovf_all = B.h_tot[2 * P + 2];
        if ((unsigned int)B.h_tot[2 * P + 1]) { set_error("internal: the overflow list of a sharded batch did not hold its keys"); return GB_E_INVARIANT; }
        windows += (int64_t)(sent + ovf_mine);
        // room for the received keys (every one may be new); grow only with the pipeline drained
        int64_t cap = (int64_t)m->cap;
        if ((int64_t)(cap * 0.9) - m->size - pending_upper < (int64_t)rt || (m->size + pending_upper) * 10 > cap * 7) {
            GB_TRY(drain(m, bufs));
            pending_upper = 0;
            int64_t budget = 0;
            GB_TRY(map_budget(m, (int64_t)rt, &budget));
            while (budget < (int64_t)rt) {
                GB_TRY(map_rebuild(m, m->cap * 2, false, 0));
                budget = (int64_t)(m->cap * 0.9) - m->size;
            }
        }
        GB_CUDA(cudaStreamWaitEvent(m->stream, B.exchanged, 0));
        InboxSlabs in;
        in.sources = (unsigned int)P; in.slices = (unsigned int)LP; in.grid = G; in.region_cap = c->region_cap; in.cnt_off = cnt_off;
        GB_TRY(insert_inbox_slabs(m, c->inbox[b % NSETS], in, slab, m->stream));
        GB_CUDA(cudaEventRecord(B.inserted, m->stream));
        B.in_flight = true;
        pending_upper += (int64_t)rt;
        if (ovf_all) { // some rank's slabs overflowed in this batch: all ranks route their lists (drains the pipeline; pathological inputs only)
            int64_t got = 0;
            GB_TRY(route_key_list(m, c, B, B.send, ovf_mine, &got));
            pending_upper += got;
        }
        if (trace) fprintf(stderr, "[pmap] batch %lld: sent %llu (+ %llu overflowed), received %llu, enqueued at %.3f ms\n", (long long)b, sent, ovf_mine, rt, now_ms() - t_begin);
        return GB_OK;
    };
    if (batches) GB_TRY(stage_a(0));
    for (int64_t b = 0; b < batches; b++) {
        if (b + 1 < batches) GB_TRY(stage_a(b + 1));
        GB_TRY(stage_b(b));
    }
    GB_CUDA(cudaEventRecord(m->ev1, m->stream));
    GB_TRY(drain(m, bufs));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    if (trace) fprintf(stderr, "[pmap] drained %.3f ms\n", now_ms() - t_begin);
    float ms = 0;
    GB_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
    m->last_insert_ns = (int64_t)(ms * 1e6);
    m->windows += windows;
    if (n_windows) *n_windows = windows;
    return GB_OK;
}

// Sharded FreqFilter.add.  Per batch of reads, on the communicator's stream: bucket the canonical k-mers by
// (owner shard, table slice) [partition.cu]; exchange the per-bucket counts; all-to-all the owner segments over
// NVLink.  On the map's stream: upsert the received keys slice by slice (every source's sub-bucket of slice 0, then
// of slice 1, ...), so the resident CTAs share one L2-sized slice of the shard.  Two buffer sets overlap the
// exchange of batch b + 1 with the upsert of batch b.
static int pmap_insert(Map *m, const uint8_t *d_bin, size_t n_bytes, const unsigned long long *d_off, unsigned int rec,
                       unsigned int len0, int64_t n_reads, const int64_t *h_win_prefix, int64_t *n_windows)
{
    Comm *c = m->comm;
    const int P = c->n_ranks;
    const int k = m->k;
    const bool fixed = d_off == nullptr;
    const int64_t win_max = fixed ? std::max<int64_t>(0, (int64_t)len0 - k + 1) : (255 - k + 1);
    // batches: enough of them to overlap exchange with upsert, bounded staging memory (<= 2^26 k-mers = 512 MiB each)
    int64_t batch_reads = std::max<int64_t>(TILE_READS, (((int64_t)1 << 26) / std::max<int64_t>(win_max, 1)) / TILE_READS * TILE_READS);
    const int want_batches = g_tune.batches > 0 ? (int)g_tune.batches : 2;
    int64_t part = ((n_reads + want_batches - 1) / want_batches + TILE_READS - 1) / TILE_READS * TILE_READS;
    if (part >= TILE_READS * 64) batch_reads = std::min(batch_reads, part);
    int64_t batches = n_reads ? (n_reads + batch_reads - 1) / batch_reads : 0;
    int64_t lp = slice_bits_for(m->cap, P);
    // fewer buckets = longer runs per bucket in the staged bucket pass = fuller NVLink write packets; measured at
    // P = 2: 16 slices/shard 4.1 ms, 64 slices/shard 6.0 ms per 96.6 M k-mers.  Below 8 slices the upsert leaves L2.
    while (lp > 3 && (P << lp) > 32) lp--;
    while (lp > 0 && (P << lp) > 128) lp--; // stay in the staged (sector-coalesced) regime
    if (g_tune.slice_bits >= 0) lp = std::min<int64_t>(lp, g_tune.slice_bits);
    // Routing levels.  ONE: the wire buckets are (owner, table slice) and the receiver upserts them as they arrive --
    // fewest passes, but #buckets = P x slices must stay small for NVLink (long store runs), so slices grow with the
    // shard.  TWO: the wire buckets are owners only (longest runs) and the receiver re-buckets each batch by fine slice
    // (one more local pass over 8 B keys) -- keeps the upsert L2-resident for multi-GB shards.
    int64_t two_level = (int64_t)(SLOT_BYTES * m->cap > (4ull << 30));
    if (g_tune.route) two_level = g_tune.route == 2;
    GB_TRY(all_reduce_i64(c, &two_level, ncclMax));
    if (two_level) lp = 0;
    GB_TRY(all_reduce_i64(c, &batches, ncclMax));
    GB_TRY(all_reduce_i64(c, &lp, ncclMin)); // every rank must cut the same buckets
    PartLayout fine; // receiver-side slices of two-level routing
    fine.owners = 1;
    fine.lp_bits = slice_bits_for(m->cap, 1);
    PartLayout pl;
    pl.owners = P;
    pl.lp_bits = (int)lp;
    const int LP = 1 << pl.lp_bits, NB = pl.nb();

    const bool trace = g_tune.trace && c->rank == 0;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_begin = now_ms();
    if (trace) fprintf(stderr, "[pmap] %lld reads, %lld batches of %lld, LP=%d\n", (long long)n_reads, (long long)batches, (long long)batch_reads, LP);
    std::vector<cudaEvent_t> tev; // trace only: 7 events per batch
    auto mark = [&](cudaStream_t strm) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, strm);
        tev.push_back(e);
    };
    BatchBufs *bufs = c->bufs;
    // the single-pass form (pmap_insert_slabs) when every rank can take it: fixed-stride stream, one routing level, peer stores,
    // batches large enough for slabs; its inbox regions hold slabs + their counts
    const int64_t batch_windows = std::min<int64_t>(batch_reads, std::max<int64_t>(n_reads, 1)) * std::max<int64_t>(win_max, 1);
    int64_t slab_keys_cta = 0, can_slabs = fixed && !two_level && g_tune.a2a == 0 && g_tune.single_pass && (int64_t)len0 >= k;
    // its wire buckets may be as many as the staged pass handles (128): the runs it stores are 2048 / buckets keys, and with 16
    // slices per shard the owner's upsert stays L2-blocked up to 8 GPUs (the 32-bucket limit of the counted pass below leaves 4
    // slices at P = 8 -- 480 MB each on C2 -- which is why that one loses the L2 blocking as P grows)
    PartLayout pls;
    pls.owners = P;
    {
        int64_t lps = std::min<int64_t>(4, slice_bits_for(m->cap, P)); // measured at P = 2 on C2: 16 slices 2.41 ms, 32 slices 2.50 (shorter store runs)
        if (g_tune.slice_bits >= 0) lps = std::min<int64_t>(lps, g_tune.slice_bits);
        GB_TRY(all_reduce_i64(c, &lps, ncclMin));
        pls.lp_bits = (int)lps;
    }
    const int LPS = 1 << pls.lp_bits, NBS = pls.nb();
    {
        GB_TRY(bufs[0].work.ensure(c->stream));
        slab_keys_cta = (int64_t)slab_cta_keys(std::min<int64_t>(batch_reads, n_reads), (unsigned long long)batch_windows, bufs[0].work.grid);
        int64_t big_enough = batch_windows >= std::max<int64_t>(1, g_tune.single_pass_min) ? 1 : 0;
        GB_TRY(all_reduce_i64(c, &big_enough, ncclMax)); // a rank with few (or no) reads follows the others
        GB_TRY(all_reduce_i64(c, &can_slabs, ncclMin));
        GB_TRY(all_reduce_i64(c, &slab_keys_cta, ncclMax));
        can_slabs = can_slabs && big_enough;
    }
    unsigned int slab = can_slabs && NBS <= 128 ? slab_keys_for((unsigned long long)slab_keys_cta, (unsigned int)NBS, bufs[0].work.grid) : 0;
    if (slab && (unsigned long long)LPS * bufs[0].work.grid * slab + (unsigned long long)LPS * bufs[0].work.grid / 2 + 8 >= (1ull << P2P_REL_BITS)) slab = 0;
    const size_t region_need = slab ? (size_t)LPS * bufs[0].work.grid * slab + (size_t)LPS * bufs[0].work.grid / 2 + 8 : (size_t)batch_windows;
    // fused routing: every rank's bucket pass stores straight into the owners' inboxes over NVLink
    GB_TRY(ensure_inboxes(c, region_need));
    if (slab && c->p2p == 1) return pmap_insert_slabs(m, d_bin, n_bytes, rec, win_max, n_reads, batch_reads, batches, pls, slab, n_windows);
    const bool p2p = c->p2p == 1;
    if (trace) fprintf(stderr, "[pmap] routing: %s, %s (%d wire buckets, %d slices)\n", p2p ? "peer stores into NVLink inboxes" : "NCCL send/recv",
                       two_level ? "two-level" : "one-level", NB, two_level ? 1 << fine.lp_bits : LP);
    GB_TRY(map_zero_counters(m));
    GB_CUDA(cudaStreamSynchronize(m->stream));
    GB_CUDA(cudaEventRecord(m->ev0, m->stream));

    int64_t windows = 0, pending_upper = 0; // pending_upper: keys handed to upserts not yet folded into m->size
    double t_issue = 0;
    // ---- stage A of batch b, on the communicator's stream: bucket the canonical k-mers of the batch's reads by (owner, slice) and
    // get them to their owners, followed by the per-bucket counts ("whoever has my counts has my keys").
    // Buffer sets: a rank that has received my counts of batch j knows my upsert of batch j - 2 is over (the wait below precedes
    // them on my stream).  With peer stores it has them before it writes batch j + 1, which may therefore reuse set (j - 2) % 3;
    // the copy-engine pipeline issues the transfer of batch j + 1 before it has the counts of batch j (only those of j - 1): one
    // more set.  Both use NSETS = 4.
    auto stage_a = [&](int64_t b) -> int {
        BatchBufs &B = bufs[b % NSETS];
        if (b >= 2 && bufs[(b - 2) % NSETS].in_flight) GB_CUDA(cudaStreamWaitEvent(c->stream, bufs[(b - 2) % NSETS].inserted, 0));
        const int64_t r0 = std::min(n_reads, b * batch_reads), r1 = std::min(n_reads, r0 + batch_reads), nr = r1 - r0;
        const int64_t w_upper = fixed ? nr * win_max : (h_win_prefix ? h_win_prefix[r1] - h_win_prefix[r0] : nr * win_max);
        if (B.in_flight) GB_CUDA(cudaStreamWaitEvent(c->stream, B.inserted, 0)); // its buffers are still being read
        GB_TRY(B.ensure(p2p ? 0 : (size_t)w_upper, 0));
        ReadBatch rb;
        rb.bin = d_bin; rb.n_bytes = n_bytes; rb.offsets = d_off; rb.rec_bytes = rec; rb.read0 = r0; rb.n_reads = nr;
        // d_tot: [0, NB) my bucket totals (row o = what I send to owner o), [NB, 2NB) row s = what source s sends me
        mark(c->stream);
        GB_TRY(part_count(rb, k, m->v210, pl, B.work, c->stream));
        mark(c->stream);
        GB_CUDA(cudaMemcpyAsync(B.d_tot, B.work.bucket_total, NB * 8, cudaMemcpyDeviceToDevice, c->stream));
        if (!p2p) {
            GB_TRY(all_to_all_rows(c, B.d_tot, B.d_tot + NB, (size_t)LP));
            GB_CUDA(cudaMemcpyAsync(B.h_tot, B.d_tot, 2 * NB * 8, cudaMemcpyDeviceToHost, c->stream));
        }
        mark(c->stream);
        if (p2p) {
            PeerOut po;
            memset(&po, 0, sizeof po);
            for (int p = 0; p < P; p++) po.base[p] = c->peer_inbox[b % NSETS][p] + (size_t)c->rank * c->region_cap;
            GB_TRY(part_scatter_peers(rb, k, m->v210, pl, B.work, po, c->stream));
            // the counts travel AFTER the keys on my stream: whoever has my counts has my keys
            GB_TRY(all_to_all_rows(c, B.d_tot, B.d_tot + NB, (size_t)LP));
            GB_CUDA(cudaMemcpyAsync(B.h_tot, B.d_tot, 2 * NB * 8, cudaMemcpyDeviceToHost, c->stream));
        } else {
            GB_TRY(part_scatter(rb, k, m->v210, pl, B.work, B.send, c->stream)); // runs while the host waits for the counts
        }
        mark(c->stream);
        t_issue = now_ms();
        return GB_OK;
    };
    // ---- stage B of batch b: with everybody's counts on the host, the chunk table of the received keys (slice-major) and their
    // upsert on the map's stream
    auto stage_b = [&](int64_t b) -> int {
        BatchBufs &B = bufs[b % NSETS];
        GB_CUDA(cudaStreamSynchronize(c->stream));
        const double t_counts = now_ms();
        unsigned long long scnt[MAX_RANKS], soff[MAX_RANKS], rcnt[MAX_RANKS], roff[MAX_RANKS];
        unsigned long long st = 0, rt = 0;
        for (int p = 0; p < P; p++) {
            unsigned long long s1 = 0, s2 = 0;
            for (int l = 0; l < LP; l++) { s1 += B.h_tot[p * LP + l]; s2 += B.h_tot[NB + p * LP + l]; }
            scnt[p] = s1; soff[p] = st; st += s1;
            rcnt[p] = s2; roff[p] = p2p ? (unsigned long long)p * c->region_cap : rt; rt += s2;
        }
        windows += (int64_t)st;
        const unsigned long long *recv_keys = nullptr;
        if (p2p) {
            recv_keys = c->inbox[b % NSETS];
        } else {
            GB_TRY(B.ensure(0, (size_t)rt));
            GB_TRY(all_to_all_v(c, B.send, soff, scnt, B.recv, roff, rcnt, ncclUint64));
            recv_keys = B.recv;
        }
        mark(c->stream);
        // chunk table of the upsert, slice-major: chunk (l, s) = source s's keys of slice l
        unsigned long long *vstart = B.h_tot + 2 * NB, *coff = vstart + NB + 1;
        {
            // offset of (s, l) inside source s's segment = prefix over l; walk slices outermost to fill vstart in order
            std::vector<unsigned long long> seg_off((size_t)NB);
            for (int s = 0; s < P; s++) {
                unsigned long long acc = 0;
                for (int l = 0; l < LP; l++) { seg_off[(size_t)s * LP + l] = roff[s] + acc; acc += B.h_tot[NB + s * LP + l]; }
            }
            unsigned long long v = 0;
            int ci = 0;
            for (int l = 0; l < LP; l++)
                for (int s = 0; s < P; s++, ci++) {
                    vstart[ci] = v;
                    coff[ci] = seg_off[(size_t)s * LP + l];
                    v += B.h_tot[NB + s * LP + l];
                }
            vstart[NB] = v;
        }
        GB_CUDA(cudaMemcpyAsync(B.d_tot + 2 * NB, vstart, (2 * NB + 1) * 8, cudaMemcpyHostToDevice, c->stream));
        GB_CUDA(cudaEventRecord(B.exchanged, c->stream));

        // room for the received keys (every one may be new); grow only with the pipeline drained
        int64_t cap = (int64_t)m->cap;
        if ((int64_t)(cap * 0.9) - m->size - pending_upper < (int64_t)rt || (m->size + pending_upper) * 10 > cap * 7) {
            GB_TRY(drain(m, bufs));
            pending_upper = 0;
            int64_t budget = 0;
            GB_TRY(map_budget(m, (int64_t)rt, &budget));
            while (budget < (int64_t)rt) { // map_budget guarantees only a minimum batch: force the size we need
                GB_TRY(map_rebuild(m, m->cap * 2, false, 0));
                budget = (int64_t)(m->cap * 0.9) - m->size;
            }
        }
        GB_CUDA(cudaStreamWaitEvent(m->stream, B.exchanged, 0));
        mark(m->stream);
        if (two_level && rt) {
            // re-bucket the received segments by fine slice into B.recv (unused by peer routing), then one ordered upsert
            if (!p2p) { set_error("two-level routing needs peer routing (inbox + local staging)"); return GB_E_STATE; }
            GB_TRY(B.ensure(0, (size_t)rt));
            KeySource ks;
            ks.keys = recv_keys; ks.vstart = B.d_tot + 2 * NB; ks.off = B.d_tot + 3 * NB + 1; ks.n_chunks = NB; ks.n_total = rt;
            GB_TRY(part_count_keys(ks, fine, B.work2, m->stream));
            GB_TRY(part_scatter_keys(ks, fine, B.work2, B.recv, m->stream));
            unsigned long long *desc = B.h_tot + 4 * NB + 1; // vstart[0], vstart[1], off[0] of the single chunk
            desc[0] = 0; desc[1] = rt; desc[2] = 0;
            GB_CUDA(cudaMemcpyAsync(B.d_tot + 4 * NB + 1, desc, 3 * 8, cudaMemcpyHostToDevice, m->stream));
            GB_TRY(insert_key_chunks(m, B.recv, B.d_tot + 4 * NB + 1, B.d_tot + 4 * NB + 3, 1, rt, m->stream));
        } else {
            GB_TRY(insert_key_chunks(m, recv_keys, B.d_tot + 2 * NB, B.d_tot + 3 * NB + 1, NB, rt, m->stream));
        }
        mark(m->stream);
        GB_CUDA(cudaEventRecord(B.inserted, m->stream));
        B.in_flight = true;
        pending_upper += (int64_t)rt;
        if (trace) fprintf(stderr, "[pmap] batch %lld: issued %.3f  bucketed+counts %.3f  enqueued %.3f ms (send %llu recv %llu)\n", (long long)b,
                           t_issue - t_begin, t_counts - t_begin, now_ms() - t_begin, st, rt);
        return GB_OK;
    };
    for (int64_t b = 0; b < batches; b++) {
        GB_TRY(stage_a(b));
        GB_TRY(stage_b(b));
    }
    GB_CUDA(cudaEventRecord(m->ev1, m->stream));
    GB_TRY(drain(m, bufs));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    if (trace) {
        fprintf(stderr, "[pmap] drained %.3f ms\n", now_ms() - t_begin);
        for (size_t b = 0; b + 7 <= tev.size(); b += 7) {
            float t[7];
            for (int i = 0; i < 7; i++) cudaEventElapsedTime(&t[i], tev[0], tev[b + i]);
            fprintf(stderr, "[pmap] gpu batch %zu: count %.3f-%.3f  counts-exchange -%.3f  scatter -%.3f  all-to-all -%.3f | upsert %.3f-%.3f\n",
                    b / 7, t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
        }
        for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    float ms = 0;
    GB_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
    m->last_insert_ns = (int64_t)(ms * 1e6);
    m->windows += windows;
    if (n_windows) *n_windows = windows;
    return GB_OK;
}

static int check_pmap(gb_map *h, Map **m)
{
    GB_TRY(check_map(h, m));
    if (!(*m)->comm) { set_error("map is not bound to a communicator (use gb_pmap_create)"); return GB_E_STATE; }
    return GB_OK;
}

} // namespace gb

using namespace gb;

extern "C" {

int gb_comm_unique_id(uint8_t id[GB_UNIQUE_ID_BYTES])
{
    if (!id) { set_error("null argument"); return GB_E_ARG; }
    static_assert(sizeof(ncclUniqueId) <= GB_UNIQUE_ID_BYTES, "unique id does not fit");
    GB_TRY(nccl_load());
    ncclUniqueId u;
    GB_NCCL(ncclGetUniqueId(&u));
    memset(id, 0, GB_UNIQUE_ID_BYTES);
    memcpy(id, &u, sizeof u);
    return GB_OK;
}

int gb_comm_create(const uint8_t id[GB_UNIQUE_ID_BYTES], int rank, int n_ranks, int device, gb_comm **out)
{
    if (!id || !out) { set_error("null argument"); return GB_E_ARG; }
    *out = nullptr;
    if (n_ranks < 1 || n_ranks > MAX_RANKS || rank < 0 || rank >= n_ranks) { set_error("bad rank %d of %d", rank, n_ranks); return GB_E_ARG; }
    GB_TRY(nccl_load());
    int ndev = 0;
    GB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("device %d not present (%d devices)", device, ndev); return GB_E_CUDA; }
    GB_CUDA(cudaSetDevice(device));
    Comm *c = new Comm();
    c->rank = rank; c->n_ranks = n_ranks; c->device = device;
    memset(c->peer_inbox, 0, sizeof c->peer_inbox);
    memset(c->peer_window, 0, sizeof c->peer_window);
    memset(c->window_cap_of, 0, sizeof c->window_cap_of);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclResult_t r = ncclCommInitRank(&c->nccl, n_ranks, u, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank", __FILE__, __LINE__); }
    // the bucketing + exchange stream is the critical path of the sharded insert: its CTAs go first when SM slots free
    // up under the (much larger) upsert grid that runs on the map's stream
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
        ncclCommDestroy(c->nccl);
        delete c;
        set_error("stream creation failed");
        return GB_E_CUDA;
    }
    *out = reinterpret_cast<gb_comm *>(c);
    return GB_OK;
}

int gb_comm_destroy(gb_comm *h)
{
    if (!h) return GB_OK;
    Comm *c = reinterpret_cast<Comm *>(h);
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    close_inboxes(c);
    close_windows(c);
    for (int i = 0; i < NSETS; i++) c->bufs[i].release();
    if (c->d_scratch) cudaFree(c->d_scratch);
    if (c->h_scratch) cudaFreeHost(c->h_scratch);
    if (c->nccl) ncclCommDestroy(c->nccl);
    delete c;
    return GB_OK;
}

// element-wise sum of a HOST array over the ranks, in place (host composition of per-rank partial results: the paired-end
// support counts of GraphSimplifier.scala:239-245 when every rank walks its own slice of the pairs on its copy of the graph)
static int comm_allreduce_host(gb_comm *h, void *host, int64_t n, size_t elem, ncclDataType_t type)
{
    if (!h) { set_error("null communicator"); return GB_E_ARG; }
    if (n < 0 || (n > 0 && !host)) { set_error("bad arguments"); return GB_E_ARG; }
    Comm *c = reinterpret_cast<Comm *>(h);
    GB_CUDA(cudaSetDevice(c->device));
    // collective even when n == 0 on this rank would be a mismatch: n must be the same everywhere (documented)
    if (n == 0) return GB_OK;
    void *d = nullptr;
    GB_CUDA(cudaMalloc(&d, (size_t)n * elem));
    int rc = GB_OK;
    if (cudaMemcpyAsync(d, host, (size_t)n * elem, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rc = GB_E_CUDA;
    if (rc == GB_OK && ncclAllReduce(d, d, (size_t)n, type, ncclSum, c->nccl, c->stream) != ncclSuccess) rc = GB_E_NCCL;
    if (rc == GB_OK && cudaMemcpyAsync(host, d, (size_t)n * elem, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = GB_E_CUDA;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && rc == GB_OK) rc = GB_E_CUDA;
    cudaFree(d);
    if (rc != GB_OK) set_error("all-reduce of %lld host elements failed", (long long)n);
    return rc;
}
int gb_comm_allreduce_sum_u32(gb_comm *h, uint32_t *host, int64_t n) { return comm_allreduce_host(h, host, n, 4, ncclUint32); }
int gb_comm_allreduce_sum_i64(gb_comm *h, int64_t *host, int64_t n) { return comm_allreduce_host(h, host, n, 8, ncclInt64); }

int gb_pmap_create(gb_comm *ch, int k, int64_t min_capacity_per_shard, uint32_t flags, gb_map **out)
{
    if (!ch) { set_error("null communicator"); return GB_E_ARG; }
    Comm *c = reinterpret_cast<Comm *>(ch);
    GB_TRY(gb_map_create(k, min_capacity_per_shard, c->device, flags, out));
    reinterpret_cast<Map *>(*out)->comm = c;
    return GB_OK;
}

int gb_pmap_insert_reads_device(gb_map *h, const uint8_t *d_bin, size_t n_bytes, const uint64_t *d_offsets, int64_t n_reads,
                                int64_t *n_windows)
{
    Map *m;
    GB_TRY(check_pmap(h, &m));
    if (n_windows) *n_windows = 0;
    if (n_reads < 0 || (!d_bin && n_reads > 0)) { set_error("bad arguments"); return GB_E_ARG; }
    // NOTE: collective -- a rank with no reads still takes part in every exchange
    if (d_offsets || n_reads == 0)
        return pmap_insert(m, d_bin, n_bytes, n_reads ? (const unsigned long long *)d_offsets : nullptr, 0, 0, n_reads, nullptr, n_windows);
    uint8_t len0 = 0;
    GB_CUDA(cudaMemcpyAsync(&len0, d_bin, 1, cudaMemcpyDeviceToHost, m->stream));
    GB_CUDA(cudaStreamSynchronize(m->stream));
    unsigned int rec = 1 + (len0 + 3) / 4;
    if ((unsigned long long)n_reads * rec > n_bytes) { set_error("truncated .bin stream"); return GB_E_ARG; }
    unsigned long long bad = 0;
    GB_TRY(map_verify_fixed(m, d_bin, rec, len0, n_reads, &bad));
    if (bad) { set_error("records are not fixed-length: pass d_offsets"); return GB_E_ARG; }
    m->fixed_stride = 1;
    return pmap_insert(m, d_bin, n_bytes, nullptr, rec, len0, n_reads, nullptr, n_windows);
}

int gb_pmap_insert_reads(gb_map *h, const uint8_t *bin, size_t n_bytes, int64_t n_reads, int64_t *n_windows)
{
    Map *m;
    GB_TRY(check_pmap(h, &m));
    ArenaScope scope(&m->arena);
    if (n_windows) *n_windows = 0;
    if (n_reads < 0 || (!bin && n_reads > 0)) { set_error("bad arguments"); return GB_E_ARG; }
    if (n_reads == 0) return pmap_insert(m, nullptr, 0, nullptr, 0, 0, 0, nullptr, n_windows);
    if (n_bytes == 0) { set_error("truncated .bin stream at read 0"); return GB_E_ARG; }
    // fixed-stride fast path: copy n_reads records of the first record's size and let the device check every
    // length byte (sound: equal length bytes at i * rec imply the record chain is i * rec); else scan on the host.
    // The choice is local to the rank: the collective schedule (batch count) is agreed inside pmap_insert.
    const unsigned int len0 = bin[0], rec = 1 + (len0 + 3) / 4;
    DeviceBuf d_bin, d_off;
    if ((unsigned long long)n_reads * rec <= n_bytes) {
        const size_t used = (size_t)n_reads * rec;
        GB_TRY(d_bin.alloc(used + 16, m->stream));
        GB_CUDA(cudaMemcpyAsync(d_bin.p, bin, used, cudaMemcpyHostToDevice, m->stream));
        unsigned long long bad = 0;
        GB_TRY(map_verify_fixed(m, (const uint8_t *)d_bin.p, rec, len0, n_reads, &bad));
        if (!bad) {
            m->fixed_stride = 1;
            return pmap_insert(m, (const uint8_t *)d_bin.p, used, nullptr, rec, len0, n_reads, nullptr, n_windows);
        }
    }
    m->fixed_stride = 0;
    std::vector<unsigned long long> off;
    std::vector<int64_t> winp;
    GB_TRY(scan_records(bin, n_bytes, n_reads, m->k, off, winp));
    const size_t used = (size_t)off[(size_t)n_reads];
    GB_TRY(d_bin.alloc(used + 16, m->stream));
    GB_CUDA(cudaMemcpyAsync(d_bin.p, bin, used, cudaMemcpyHostToDevice, m->stream));
    GB_TRY(d_off.alloc(off.size() * 8, m->stream));
    GB_CUDA(cudaMemcpyAsync(d_off.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, m->stream));
    GB_CUDA(cudaStreamSynchronize(m->stream));
    return pmap_insert(m, (const uint8_t *)d_bin.p, used, (const unsigned long long *)d_off.p, 0, 0, n_reads, winp.data(), n_windows);
}

int gb_pmap_size(gb_map *h, int64_t *size)
{
    Map *m;
    GB_TRY(check_pmap(h, &m));
    if (!size) { set_error("null argument"); return GB_E_ARG; }
    int64_t s = m->size;
    GB_TRY(all_reduce_i64(m->comm, &s, ncclSum));
    *size = s;
    return GB_OK;
}

int gb_pmap_delete_below(gb_map *h, int32_t min_count)
{
    Map *m;
    GB_TRY(check_pmap(h, &m));
    return gb_map_delete_below(h, min_count);
}

int gb_pmap_owner(gb_map *h, const uint64_t *keys, int64_t n, int32_t *owner)
{
    Map *m;
    GB_TRY(check_pmap(h, &m));
    if (n < 0 || (n > 0 && (!keys || !owner))) { set_error("bad arguments"); return GB_E_ARG; }
    return gb_owner_of(keys, n, m->comm->n_ranks, owner);
}

// partition(key) (PartitionedDNAMap.scala:60-63); pure host arithmetic, usable without a GPU
int gb_owner_of(const uint64_t *keys, int64_t n, int n_parts, int32_t *owner)
{
    if (n < 0 || n_parts < 1 || (n > 0 && (!keys || !owner))) { set_error("bad arguments"); return GB_E_ARG; }
    for (int64_t i = 0; i < n; i++) owner[i] = (int32_t)owner_of(mix64(keys[i]), (unsigned int)n_parts);
    return GB_OK;
}

// the ownership rule of the sharded graph build (sgraph.cuh): minimizer owner; host arithmetic
int gb_owner_of_minimizer(const uint64_t *keys, int64_t n, int k, int n_parts, int32_t *owner)
{
    if (n < 0 || n_parts < 1 || (n > 0 && (!keys || !owner))) { set_error("bad arguments"); return GB_E_ARG; }
    if (k < 1 || k > 31) { set_error("k = %d outside 1..31", k); return GB_E_K_RANGE; }
    const int m = sg::minimizer_len(k);
    for (int64_t i = 0; i < n; i++) owner[i] = (int32_t)sg::owner_of_kmer(keys[i], k, m, n_parts);
    return GB_OK;
}

int gb_pmap_lookup(gb_map *h, const uint64_t *keys, int64_t n, int32_t *counts, uint8_t *found)
{
    Map *m;
    GB_TRY(check_pmap(h, &m));
    ArenaScope scope(&m->arena);
    if (n < 0 || (n > 0 && !keys)) { set_error("bad arguments"); return GB_E_ARG; }
    {   // collective: every rank must take the same exit
        int64_t bad = check_keys(m, keys, n) != GB_OK;
        GB_TRY(all_reduce_i64(m->comm, &bad, ncclMax));
        if (bad) { set_error("a query key is longer than k = %d (on some rank)", m->k); return GB_E_K_RANGE; }
    }
    Comm *c = m->comm;
    const int P = c->n_ranks;
    // bucket the queries by owner on the host (they come from the host), remember where each one came from
    std::vector<unsigned long long> h_send((size_t)n), scnt(P, 0), soff(P, 0), rcnt(P, 0), roff(P, 0), fill(P, 0);
    std::vector<int64_t> origin((size_t)n);
    std::vector<int32_t> own((size_t)n);
    for (int64_t i = 0; i < n; i++) { own[(size_t)i] = pmap_owner_of(m, keys[i], P); scnt[own[(size_t)i]]++; }
    unsigned long long st = 0;
    for (int p = 0; p < P; p++) { soff[p] = st; st += scnt[p]; }
    for (int64_t i = 0; i < n; i++) {
        size_t pos = (size_t)(soff[own[(size_t)i]] + fill[own[(size_t)i]]++);
        h_send[pos] = keys[i];
        origin[pos] = i;
    }
    DeviceBuf d_cnt, d_send, d_recv, d_ans, d_back;
    GB_TRY(d_cnt.alloc(2 * MAX_RANKS * 8));
    GB_CUDA(cudaMemcpyAsync(d_cnt.p, scnt.data(), P * 8, cudaMemcpyHostToDevice, c->stream));
    std::vector<unsigned long long> tmp(P);
    GB_TRY(exchange_counts(c, (unsigned long long *)d_cnt.p, (unsigned long long *)d_cnt.p + MAX_RANKS, tmp.data(), rcnt.data()));
    unsigned long long rt = 0;
    for (int p = 0; p < P; p++) { roff[p] = rt; rt += rcnt[p]; }
    GB_TRY(d_send.alloc((size_t)n * 8));
    GB_TRY(d_recv.alloc((size_t)rt * 8));
    GB_TRY(d_ans.alloc((size_t)rt * 8));
    GB_TRY(d_back.alloc((size_t)n * 8));
    if (n) GB_CUDA(cudaMemcpyAsync(d_send.p, h_send.data(), (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    GB_TRY(all_to_all_v(c, (unsigned long long *)d_send.p, soff.data(), scnt.data(), (unsigned long long *)d_recv.p, roff.data(), rcnt.data(), ncclUint64));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    // local probe: answers packed as found << 32 | count
    std::vector<unsigned long long> h_q((size_t)rt), h_a((size_t)rt);
    if (rt) {
        GB_CUDA(cudaMemcpy(h_q.data(), d_recv.p, (size_t)rt * 8, cudaMemcpyDeviceToHost));
        std::vector<int32_t> cnts((size_t)rt);
        std::vector<uint8_t> fnd((size_t)rt);
        GB_TRY(gb_map_lookup(h, (const uint64_t *)h_q.data(), (int64_t)rt, cnts.data(), fnd.data()));
        for (size_t i = 0; i < (size_t)rt; i++) h_a[i] = ((unsigned long long)fnd[i] << 32) | (uint32_t)cnts[i];
        GB_CUDA(cudaMemcpyAsync(d_ans.p, h_a.data(), (size_t)rt * 8, cudaMemcpyHostToDevice, c->stream));
    }
    GB_TRY(all_to_all_v(c, (unsigned long long *)d_ans.p, roff.data(), rcnt.data(), (unsigned long long *)d_back.p, soff.data(), scnt.data(), ncclUint64));
    std::vector<unsigned long long> h_back((size_t)n);
    if (n) GB_CUDA(cudaMemcpyAsync(h_back.data(), d_back.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t pos = 0; pos < n; pos++) {
        int64_t i = origin[(size_t)pos];
        if (counts) counts[i] = (int32_t)(uint32_t)h_back[(size_t)pos];
        if (found) found[i] = (uint8_t)(h_back[(size_t)pos] >> 32);
    }
    return GB_OK;
}

// Graph.buildGraph over all shards: every shard's (key, count) pairs are all-gathered into a replica of the whole
// filtered table on every rank, and the single-GPU build runs on the replica (identical result on every rank).
int gb_pmap_graph_build(gb_map *h, gb_graph **out)
{
    const bool trace = g_tune.trace != 0;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_begin = now_ms();
    auto tick = [&](const char *what) { if (trace) fprintf(stderr, "[pgraph] %-26s %9.3f ms\n", what, now_ms() - t_begin); };
    Map *m;
    GB_TRY(check_pmap(h, &m));
    ArenaScope scope(&m->arena);
    if (!out) { set_error("null out pointer"); return GB_E_ARG; }
    *out = nullptr;
    Comm *c = m->comm;
    const int P = c->n_ranks;
    int64_t total = m->size, dual = m->noncanonical;
    GB_TRY(all_reduce_i64(c, &total, ncclSum));
    GB_TRY(all_reduce_i64(c, &dual, ncclMax));
    // gb_tune pgraph_sharded (the default): no replica -- minimizer re-routing, rank-local list ranking, segment list (sgraph.cuh).
    // Needs peer access between all ranks; the tuning key is process-wide and set alike on every rank, so the choice is collective.
    if (g_tune.pgraph_sharded && P <= sg::MAXR) {
        GB_TRY(ensure_inboxes(c, 0)); // settles c->p2p (collective)
        if (c->p2p == 1) {
            GB_CUDA(cudaStreamSynchronize(m->stream));
            // scratch for this rank's share after the re-routing (about its own size; twice that leaves room for imbalance)
            GB_TRY(m->arena.reserve((size_t)m->size * 2 * 140 + (size_t)total * 4 + ((size_t)64 << 20)));
            NcclFabric fab(c);
            Map *maps[1] = { m };
            const int rc = graph_build_on_fabric(fab, maps, c->stream, dual != 0, out);
            tick("sharded graph built");
            return rc;
        }
    }
    // per-rank sizes
    DeviceBuf d_sizes;
    GB_TRY(d_sizes.alloc(MAX_RANKS * 8 * 2));
    unsigned long long mine = (unsigned long long)m->size;
    GB_CUDA(cudaMemcpyAsync(d_sizes.p, &mine, 8, cudaMemcpyHostToDevice, c->stream));
    GB_NCCL(ncclAllGather(d_sizes.p, (unsigned long long *)d_sizes.p + MAX_RANKS, 1, ncclUint64, c->nccl, c->stream));
    std::vector<unsigned long long> sizes(P), offs(P);
    GB_CUDA(cudaMemcpyAsync(sizes.data(), (unsigned long long *)d_sizes.p + MAX_RANKS, P * 8, cudaMemcpyDeviceToHost, c->stream));
    GB_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long t = 0;
    for (int p = 0; p < P; p++) { offs[p] = t; t += sizes[p]; }

    tick("sizes exchanged");
    // every shard exports straight into its segment of the gathered arrays
    DeviceBuf all_keys, all_vals;
    GB_TRY(all_keys.alloc((size_t)t * 8));
    GB_TRY(all_vals.alloc((size_t)t * 4));
    GB_TRY(map_export_device(m, (unsigned long long *)all_keys.p + offs[c->rank], (int *)all_vals.p + offs[c->rank]));
    GB_NCCL(ncclGroupStart());
    for (int p = 0; p < P; p++) {
        if (!sizes[p]) continue;
        GB_NCCL(ncclBroadcast((unsigned long long *)all_keys.p + offs[p], (unsigned long long *)all_keys.p + offs[p], sizes[p], ncclUint64, p, c->nccl, c->stream));
        GB_NCCL(ncclBroadcast((int *)all_vals.p + offs[p], (int *)all_vals.p + offs[p], sizes[p], ncclInt32, p, c->nccl, c->stream));
    }
    GB_NCCL(ncclGroupEnd());
    GB_CUDA(cudaStreamSynchronize(c->stream));

    tick("keys gathered");
    // the replica map is kept with the shard and reused by the next build (its table is GBs: no malloc per call)
    gb_map *rh = reinterpret_cast<gb_map *>(m->replica);
    if (!rh) {
        GB_TRY(gb_map_create(m->k, (int64_t)t, m->device, m->v210 ? GB_FLAG_HASH_SCALA_210 : 0, &rh));
        m->replica = reinterpret_cast<Map *>(rh);
    } else {
        GB_TRY(gb_map_clear(rh, (int64_t)t));
    }
    Map *r = reinterpret_cast<Map *>(rh);
    r->noncanonical = dual != 0;
    // the gathered array is identical on every rank: its index is the vertex id everywhere, so the ranks can split the
    // membership probes (the heaviest kernel of the build) and exchange the results
    const bool as_vertices = t < (1ull << 30);
    int rc = map_zero_counters(r);
    if (rc == GB_OK) rc = map_launch_update_set(r, (const unsigned long long *)all_keys.p, (const int *)all_vals.p, (int64_t)t, r->stream, as_vertices);
    unsigned long long cn[4];
    if (rc == GB_OK) rc = map_read_counters(r, cn);
    if (rc != GB_OK) return rc;
    r->size = (int64_t)cn[0];
    tick("replica filled");
    if (!as_vertices) return gb_graph_build(rh, out);
    r->kept_keys = (const unsigned long long *)all_keys.p;
    r->kept_n = (int64_t)t;
    r->kept_valid = true;
    struct Ctx { Comm *c; const unsigned long long *offs, *sizes; } ctx{ c, offs.data(), sizes.data() };
    ShardPlan sp;
    sp.lo = offs[c->rank];
    sp.hi = offs[c->rank] + sizes[c->rank];
    sp.ctx = &ctx;
    sp.gather = [](void *vctx, void *base, size_t elem) -> int {
        Ctx *x = (Ctx *)vctx;
        Comm *cc = x->c;
        GB_NCCL(ncclGroupStart());
        for (int p = 0; p < cc->n_ranks; p++) {
            if (!x->sizes[p]) continue;
            char *seg = (char *)base + x->offs[p] * elem;
            GB_NCCL(ncclBroadcast(seg, seg, x->sizes[p] * elem, ncclUint8, p, cc->nccl, cc->stream));
        }
        GB_NCCL(ncclGroupEnd());
        GB_CUDA(cudaStreamSynchronize(cc->stream));
        return GB_OK;
    };
    rc = graph_build_sharded(rh, out, &sp);
    tick("graph built");
    r->kept_valid = false; // all_keys goes back to the arena with this call
    return rc;
}

} // extern "C"
