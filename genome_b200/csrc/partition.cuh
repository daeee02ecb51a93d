// partition.cuh -- host API of partition.cu: canonical k-mers of a read batch grouped into buckets
// (owner shard, table slice), and the L2-blocked bulk upsert that consumes such buckets.
#pragma once
#include "common.cuh"

namespace gb {

constexpr int MAX_BUCKETS = 512;
constexpr int MAX_RANKS = 64;
constexpr int P2P_REL_BITS = 26; // a sharded batch holds at most 2^26 k-mers

// destination of pass 2 when the buckets of owner o go straight into o's inbox over NVLink (peer-mapped memory):
// key number r of owner o's segment is stored at base[o][r]
struct PeerOut {
    unsigned long long *base[MAX_RANKS];
};
constexpr int SLICE_LOG2_BYTES = 26; // a table slice of 64 MiB stays resident in the 126 MB L2 while it is filled;
                                     // measured on C2 (2.1 GB table): 8..32 slices within 1%, 64 slices 10% slower (bucket pass)

// bucket = owner * (1 << lp_bits) + slice, owner = owner_of(h, owners), slice = top lp_bits of h (= of the slot index)
struct PartLayout {
    int owners = 1;
    int lp_bits = 0;
    int nb() const { return owners << lp_bits; }
};

struct ReadBatch {
    const uint8_t *bin = nullptr;
    unsigned long long n_bytes = 0;
    const unsigned long long *offsets = nullptr; // device, n_reads + 1 entries; nullptr = fixed stride
    unsigned int rec_bytes = 0;                  // fixed stride only
    long long read0 = 0, n_reads = 0;
};

// keys already extracted (a received inbox): `n_chunks` ranges, chunk c = virtual positions [vstart[c], vstart[c+1])
// starting at keys[off[c]]; all arrays on the device
struct KeySource {
    const unsigned long long *keys = nullptr, *vstart = nullptr, *off = nullptr;
    int n_chunks = 0;
    unsigned long long n_total = 0;
};

// per-CTA bucket histograms -> bucket bases; device arrays live in PartWork
struct PartWork {
    int grid = 0;
    unsigned int *cta_hist = nullptr;        // [grid][nb] counts, then exclusive offsets inside the bucket
    unsigned long long *bucket_base = nullptr; // [MAX_BUCKETS + 1] exclusive prefix of bucket totals; [nb] = total keys
    unsigned long long *bucket_total = nullptr; // [MAX_BUCKETS]
    cudaStream_t owner_stream = nullptr;
    int ensure(cudaStream_t st);
    void release();
};

inline int slice_bits_for(unsigned long long table_slots, int owners)
{
    int lp = 0; // smallest lp with table bytes / 2^lp <= slice size (16-byte slots)
    while (((table_slots * 16) >> lp) > (1ull << SLICE_LOG2_BYTES)) lp++;
    while (lp > 0 && (owners << lp) > 128) lp--; // the staged bucket pass handles at most 128 buckets
    return lp < 0 ? 0 : lp;
}

// pass 1: count; fills w.cta_hist / w.bucket_total / w.bucket_base (all on `st`, no host synchronisation)
int part_count(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, cudaStream_t st);
// pass 2: write every canonical k-mer of the batch into its bucket's range of `out` (bucket-major order)
int part_scatter(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, unsigned long long *out, cudaStream_t st);
// same, but owner o's buckets are written to peers.base[o] (its inbox region for this rank), not to one local array
int part_scatter_peers(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, const PeerOut &peers, cudaStream_t st);

// the same two passes over keys that are already extracted: re-bucket a received batch by (fine) table slice
int part_count_keys(const KeySource &ks, const PartLayout &pl, PartWork &w, cudaStream_t st);
int part_scatter_keys(const KeySource &ks, const PartLayout &pl, PartWork &w, unsigned long long *out, cudaStream_t st);

// update(key, 1, _ + 1) for the keys of `n_chunks` ranges visited in order: chunk c holds the virtual positions
// [vstart[c], vstart[c+1]) and starts at keys[off[c]].  d_vstart has n_chunks + 1 entries.  n_total = vstart[n_chunks].
// total_is_upper_bound: n_total only sizes the launch, the exact number of keys is d_vstart[n_chunks] on the device.
int insert_key_chunks(Map *m, const unsigned long long *d_keys, const unsigned long long *d_vstart, const unsigned long long *d_off,
                      int n_chunks, unsigned long long n_total, cudaStream_t st, bool total_is_upper_bound = false);

// the single-pass bucket pass (no count pass; gb_tune single_pass): see bucket_slabs_kernel
unsigned int slab_keys_for(unsigned long long cta_keys, unsigned int nb, int grid);
unsigned long long slab_cta_keys(long long n_reads, unsigned long long windows, int grid);
// the same in LIST mode for the chunked host insert: begin / one launch per read range / end (see partition.cu)
int slab_list_begin(const PartLayout &pl, PartWork &w, cudaStream_t st);
int slab_list_range(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, unsigned long long *out, unsigned int slab,
                    unsigned long long ovf_cap, cudaStream_t st);
int slab_list_end(const PartLayout &pl, PartWork &w, unsigned int slab, unsigned long long ovf_cap, unsigned long long *d_desc, Map *m, cudaStream_t st);
int part_scatter_slabs(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, unsigned long long *out, unsigned int slab,
                       Map *m, cudaStream_t st);
// the upsert over those slabs (slab s holds min(d_count[s], slab) keys at d_keys + s * slab), in slab order = slice order
int insert_slabs(Map *m, const unsigned long long *d_keys, const unsigned int *d_count, unsigned int slab, unsigned int n_slabs, cudaStream_t st);

// The same pass on a sharded map (comm.cu): bucket = (owner shard, table slice), and the slabs of owner o live in o's NVLink inbox --
// the all-to-all happens inside the kernel, store by store, in the staged runs.  keys[o] / cnt[o] = this rank's region of owner o's
// inbox through its peer mapping: slice-major slabs [slice][CTA], then their fill counts.  No count pass, no counts on the host: the
// owner upserts straight from the slabs (insert_slabs_kernel, InboxSlabs).  owner_total[o] += keys put into owner o's slabs.
struct PeerSlabs {
    unsigned long long *keys[MAX_RANKS];
    unsigned int *cnt[MAX_RANKS];
    unsigned long long *owner_total = nullptr;
    unsigned int owners = 0;
};
// what the owner sees of them: its inbox holds `sources` regions of region_cap keys; a region = slice-major slabs [slices][grid],
// then (at key offset cnt_off) their fill counts as u32
struct InboxSlabs {
    unsigned int sources = 0, slices = 0, grid = 0;
    unsigned long long region_cap = 0, cnt_off = 0;
};
int bucket_slabs_peers(const ReadBatch &rb, int k, bool v210, const PartLayout &pl, PartWork &w, unsigned int slab, const PeerSlabs &ps,
                       unsigned long long *ovf, unsigned long long ovf_cap, unsigned long long *d_cursor, unsigned int *d_failed, cudaStream_t st);
int insert_inbox_slabs(Map *m, const unsigned long long *d_inbox, const InboxSlabs &in, unsigned int slab, cudaStream_t st);

// desc = { 0, *d_total, 0 } (the chunk table of ONE contiguous range) and counters[3] += *d_total, all on the stream
int make_single_chunk(const unsigned long long *d_total, unsigned long long *d_desc, unsigned long long *d_counters, cudaStream_t st);

} // namespace gb
