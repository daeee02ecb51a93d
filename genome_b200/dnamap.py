"""Host-side mirror of the reference's DNAMap surface over the C ABI (the role the Scala shim of INTEGRATION.md
plays on the JVM).  Names and argument meaning follow the reference (paths relative to /root/reference,
S/ = src/main/scala/ru/ifmo/genome/):

  trait DNAMap[T]          S/ds/ArrayDNAMap.scala:49-60      -> ArrayDNAMap / PartitionedDNAMap below (T = Int)
  class ArrayDNAMap        S/ds/ArrayDNAMap.scala:62-243     -> one shard on one GPU
  class PartitionedDNAMap  S/ds/PartitionedDNAMap.scala:15-63 -> one shard per rank, NCCL all-to-all routing
  object FreqFilter        S/data/FreqFilter.scala:25-58
  class PairedEndData      S/data/PairedEndData.scala:12-41   (the `.bin` stream + the header object file)

Keys are Long1DNASeq values: one uint64 per k-mer, base i at bits 2i (S/dna/DNASeq.scala:80-85).  Bulk calls take
numpy arrays.  Scala closures cannot cross a C ABI; the three the hot path uses are dedicated entry points
(`update_counts` = update(key, 1, _ + 1), `delete_below` = deleteAll((k, v) => v < rounds), the terminal
classifier inside Graph.buildGraph); `mapReduce` / `foreach` / `deleteAll` with an arbitrary Python callable run
over the exported arrays on the host.
"""
import ctypes as C

import numpy as np

from . import capi


class PairedEndData:
    """A read set in the reference `.bin` layout: per read 1 length byte + (len+3)/4 packed bytes; reads come in
    pairs (S/data/PairedEndData.scala:20-36).  `count` is the number of PAIRS like the reference's field."""

    def __init__(self, bin_bytes, count, insert=0):
        self.bin = np.ascontiguousarray(bin_bytes, dtype=np.uint8)
        self.count = int(count)
        self.insert = int(insert)

    @property
    def n_reads(self):
        return 2 * self.count

    @staticmethod
    def apply(path):
        """object PairedEndData.apply(f) (PairedEndData.scala:38-41): read the Java-serialised header at `path`, then the
        `.bin` stream it names (a relative name is resolved like java.io.File: against the working directory)."""
        from . import formats
        with open(path, "rb") as f:
            count, insert, bin_path = formats.read_paired_end_header(f.read())
        return PairedEndData(np.fromfile(bin_path, dtype=np.uint8), count, insert)

    def write(self, path, bin_path):
        """PairedEndData.write(f) (14-18) plus the stream itself: `bin_path` receives the records, `path` the header."""
        from . import formats
        self.bin.tofile(bin_path)
        with open(path, "wb") as f:
            f.write(formats.write_paired_end_header(self.count, self.insert, bin_path))

    def record_offsets(self):
        """Byte offset of every record (n_reads + 1 entries); raises on a truncated stream."""
        off = np.empty(self.n_reads + 1, np.uint64)
        pos = 0
        b = self.bin
        for r in range(self.n_reads):
            if pos >= b.size:
                raise ValueError("truncated .bin stream at read %d" % r)
            off[r] = pos
            pos += 1 + (int(b[pos]) + 3) // 4
        if pos > b.size:
            raise ValueError("truncated .bin stream")
        off[self.n_reads] = pos
        return off

    def take(self, n_pairs):
        """The first n_pairs pairs (`data.getPairs.take(takeFirst)`, GraphSimplifier.scala:211)."""
        n_pairs = min(int(n_pairs), self.count)
        if n_pairs == self.count:
            return self
        if self.bin.size and self.n_reads and _fixed_record(self.bin, self.n_reads):
            rec = 1 + (int(self.bin[0]) + 3) // 4
            return PairedEndData(self.bin[:2 * n_pairs * rec], n_pairs, self.insert)
        off = self.record_offsets()
        return PairedEndData(self.bin[:int(off[2 * n_pairs])], n_pairs, self.insert)

    def shard(self, rank, world):
        """This rank's slice of the pair stream, in file order (pairs are never split)."""
        lo, hi = shard_range(self.count, rank, world)
        if self.bin.size and self.n_reads and _fixed_record(self.bin, self.n_reads):
            rec = 1 + (int(self.bin[0]) + 3) // 4
            return PairedEndData(self.bin[2 * lo * rec:2 * hi * rec], hi - lo, self.insert)
        off = self.record_offsets()
        return PairedEndData(self.bin[int(off[2 * lo]):int(off[2 * hi])], hi - lo, self.insert)


def _fixed_record(b, n_reads):
    rec = 1 + (int(b[0]) + 3) // 4
    if n_reads * rec != b.size:
        return False
    return bool(np.all(b[::rec] == b[0]))


def shard_range(count, rank, world):
    """[lo, hi) of `count` items for `rank` of `world`: contiguous, balanced to within one item."""
    base, rem = divmod(count, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _MapBase:
    """trait DNAMap[Int] (S/ds/ArrayDNAMap.scala:49-60)."""

    _insert = "gb_map_insert_reads"
    _insert_device = "gb_map_insert_reads_device"
    _size = "gb_map_size"
    _lookup = "gb_map_lookup"
    _delete_below = "gb_map_delete_below"

    def __init__(self):
        self.h = None
        self.k = 0

    def close(self):
        if getattr(self, "h", None):
            capi.lib().gb_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- size / apply / contains
    @property
    def size(self):
        n = C.c_int64()
        capi.check(getattr(capi.lib(), self._size)(self.h, C.byref(n)))
        return n.value

    def lookup(self, keys):
        """Bulk apply/contains: (counts int32[n], found bool[n])."""
        keys = capi.as_u64(keys)
        counts = np.zeros(keys.size, np.int32)
        found = np.zeros(keys.size, np.uint8)
        capi.check(getattr(capi.lib(), self._lookup)(self.h, capi.ptr(keys), keys.size, capi.ptr(counts), capi.ptr(found)))
        return counts, found.astype(bool)

    def apply(self, key):
        """apply(key): Option[Int] -> the count or None."""
        c, f = self.lookup(np.array([key], np.uint64))
        return int(c[0]) if f[0] else None

    def contains(self, key):
        return bool(self.lookup(np.array([key], np.uint64))[1][0])

    # ---- updates
    def update_counts(self, keys):
        """update(key, 1, _ + 1) for every key, taken as it is (FreqFilter.scala:33)."""
        keys = capi.as_u64(keys)
        capi.check(capi.lib().gb_map_update_counts(self.h, capi.ptr(keys), keys.size))

    def update(self, keys, vals):
        """update(key, v) for every (key, v)."""
        keys = capi.as_u64(keys)
        vals = np.ascontiguousarray(vals, dtype=np.int32)
        if vals.size != keys.size:
            raise ValueError("keys and vals differ in length")
        capi.check(capi.lib().gb_map_update(self.h, capi.ptr(keys), capi.ptr(vals), keys.size))

    def insert_reads(self, data, n_reads=None):
        """FreqFilter.add over a `.bin` stream held in HOST memory; returns the number of k-windows inserted."""
        b = data.bin if isinstance(data, PairedEndData) else np.ascontiguousarray(data, dtype=np.uint8)
        if n_reads is None:
            n_reads = data.n_reads
        w = C.c_int64()
        capi.check(getattr(capi.lib(), self._insert)(self.h, capi.ptr(b), b.size, int(n_reads), C.byref(w)))
        return w.value

    def insert_reads_device(self, d_bin_ptr, n_bytes, n_reads, d_offsets_ptr=None):
        """Same with the stream already in device memory (raw device addresses)."""
        w = C.c_int64()
        capi.check(getattr(capi.lib(), self._insert_device)(self.h, capi.ptr(int(d_bin_ptr)), int(n_bytes),
                                                            capi.ptr(d_offsets_ptr), int(n_reads), C.byref(w)))
        return w.value

    def insert_records_device(self, d_bin_ptr, n_bytes, rec_bytes, n_records, max_len):
        """Records at a fixed stride with their own length bytes (padded ragged reads; super-k-mer records), in device memory."""
        w = C.c_int64()
        capi.check(capi.lib().gb_map_insert_records_device(self.h, capi.ptr(int(d_bin_ptr)), int(n_bytes), int(rec_bytes), int(n_records),
                                                           int(max_len), C.byref(w)))
        return w.value

    # ---- deleteAll / mapReduce / foreach
    def delete_below(self, rounds):
        """deleteAll((k, v) => v < rounds) (FreqFilter.scala:55)."""
        capi.check(getattr(capi.lib(), self._delete_below)(self.h, int(rounds)))

    def export(self):
        """(keys uint64[n], counts int32[n]) of this shard, order unspecified."""
        n = C.c_int64()
        capi.check(capi.lib().gb_map_export(self.h, None, None, 0, C.byref(n)))
        keys = np.empty(n.value, np.uint64)
        vals = np.empty(n.value, np.int32)
        if n.value:
            capi.check(capi.lib().gb_map_export(self.h, capi.ptr(keys), capi.ptr(vals), n.value, C.byref(n)))
        return keys, vals

    def export_sorted(self):
        keys, vals = self.export()
        o = np.argsort(keys, kind="stable")
        return keys[o], vals[o]

    def mapReduce(self, map_fn, reduce_fn):
        """mapReduce(map, reduce): the host closure runs over the exported (key, value) pairs."""
        keys, vals = self.export()
        out = []
        for kk, vv in zip(keys.tolist(), vals.tolist()):
            r = map_fn((kk, vv))
            if r is not None:
                out.append(r)
        return reduce_fn(out)

    def foreach(self, f):
        self.mapReduce(lambda p: (f(p), None)[1], lambda _: None)

    def neighbour_masks(self, keys):
        """incoming/outcoming (Graph.scala:272-282) of arbitrary k-mers: (out_mask, in_mask) uint8 arrays, bit b = base b."""
        keys = capi.as_u64(keys)
        m = np.zeros(keys.size, np.uint8)
        capi.check(capi.lib().gb_map_neighbour_masks(self.h, capi.ptr(keys), keys.size, capi.ptr(m)))
        return m & 15, m >> 4

    def clear(self, min_capacity=0):
        """A fresh map on the same handle (new ArrayDNAMap[Int](k)), sized for min_capacity distinct keys."""
        capi.check(capi.lib().gb_map_clear(self.h, int(min_capacity)))

    def sync(self):
        capi.check(capi.lib().gb_sync(self.h))

    def timer_start(self):
        capi.check(capi.lib().gb_timer_start(self.h))

    def timer_stop(self):
        """Device time in ns since timer_start on the handle's stream (CUDA events)."""
        ns = C.c_int64()
        capi.check(capi.lib().gb_timer_stop(self.h, C.byref(ns)))
        return ns.value

    def stats(self):
        s = (C.c_int64 * 8)()
        capi.check(capi.lib().gb_map_stats(self.h, s))
        return dict(capacity=s[0], table_bytes=s[1], grows=s[2], windows=s[3], last_insert_ns=s[4], fixed_stride=s[5],
                    bucket_ns=s[6], upsert_ns=s[7])


    def phase_ns(self):
        s = (C.c_int64 * 8)()
        capi.check(capi.lib().gb_map_phase_ns(self.h, s))
        return dict(bucket_ns=s[0], upsert_ns=s[1], filter_sweep_ns=s[2], filter_reinsert_ns=s[3], slots_swept=s[4],
                    graph_masks_ns=s[5], graph_rank_ns=s[6], graph_jump_launches=s[7])


class ArrayDNAMap(_MapBase):
    """new ArrayDNAMap[Int](k) (S/ds/ArrayDNAMap.scala:62-72): one open-addressing table in one GPU's HBM."""

    def __init__(self, k, min_capacity=0, device=0, flags=0):
        super().__init__()
        h = C.c_void_p()
        capi.check(capi.lib().gb_map_create(int(k), int(min_capacity), int(device), int(flags), C.byref(h)))
        self.h = h
        self.k = int(k)
        self.device = device


class Communicator:
    """The set of shards of a PartitionedDNAMap: one rank per GPU.  `broadcast` ships the 128-byte NCCL id from
    rank 0 to the others (any host transport: torch.distributed, a file, a socket)."""

    def __init__(self, rank, world, device, broadcast):
        _prefer_torch_nccl()
        ident = np.zeros(capi.GB_UNIQUE_ID_BYTES, np.uint8)
        if rank == 0:
            capi.check(capi.lib().gb_comm_unique_id(capi.ptr(ident)))
        ident = np.ascontiguousarray(broadcast(ident), dtype=np.uint8)
        h = C.c_void_p()
        capi.check(capi.lib().gb_comm_create(capi.ptr(ident), int(rank), int(world), int(device), C.byref(h)))
        self.h = h
        self.rank, self.world, self.device = rank, world, device

    def allreduce_sum(self, a):
        """Element-wise sum over the ranks of a uint32 or int64 host array, in place (same shape on every rank)."""
        if a.dtype == np.uint32:
            capi.check(capi.lib().gb_comm_allreduce_sum_u32(self.h, capi.ptr(a), a.size))
        elif a.dtype == np.int64:
            capi.check(capi.lib().gb_comm_allreduce_sum_i64(self.h, capi.ptr(a), a.size))
        else:
            raise TypeError("uint32 or int64")
        return a

    def close(self):
        if getattr(self, "h", None):
            capi.lib().gb_comm_destroy(self.h)
            self.h = None


def _prefer_torch_nccl():
    """The library binds NCCL with dlopen("libnccl.so.2").  If torch may be imported later in this process, make sure
    the copy that gets bound is torch's bundled one (a newer system-wide copy loaded first would break torch)."""
    import os
    import sys
    if "torch" in sys.modules or os.environ.get("GENOME_B200_NCCL"):
        return
    for p in sys.path:
        cand = os.path.join(p, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["GENOME_B200_NCCL"] = cand
            return


def torch_broadcast(ident):
    """Broadcast helper over an initialised torch.distributed process group (plumbing only)."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(ident.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, 0)
    return t.cpu().numpy()


class PartitionedDNAMap(_MapBase):
    """new PartitionedDNAMap[Int](k) (S/ds/PartitionedDNAMap.scala:15-28): this rank's shard.  Every method that
    the reference routes or broadcasts is collective here: all ranks call it in the same order."""

    _insert = "gb_pmap_insert_reads"
    _insert_device = "gb_pmap_insert_reads_device"
    _size = "gb_pmap_size"
    _lookup = "gb_pmap_lookup"
    _delete_below = "gb_pmap_delete_below"

    def __init__(self, k, comm, min_capacity_per_shard=0, flags=0):
        super().__init__()
        h = C.c_void_p()
        capi.check(capi.lib().gb_pmap_create(comm.h, int(k), int(min_capacity_per_shard), int(flags), C.byref(h)))
        self.h = h
        self.k = int(k)
        self.comm = comm

    @property
    def local_size(self):
        n = C.c_int64()
        capi.check(capi.lib().gb_map_size(self.h, C.byref(n)))
        return n.value

    def owner(self, keys):
        keys = capi.as_u64(keys)
        o = np.zeros(keys.size, np.int32)
        capi.check(capi.lib().gb_pmap_owner(self.h, capi.ptr(keys), keys.size, capi.ptr(o)))
        return o


def owner_of(keys, n_parts):
    """partition(key) (PartitionedDNAMap.scala:60-63) as the library computes it; host arithmetic, no GPU."""
    keys = capi.as_u64(keys)
    o = np.zeros(keys.size, np.int32)
    capi.check(capi.lib().gb_owner_of(capi.ptr(keys), keys.size, int(n_parts), capi.ptr(o)))
    return o


class FreqFilter:
    """object FreqFilter (S/data/FreqFilter.scala)."""

    @staticmethod
    def extractFilteredKmers(data, k, rounds, comm=None, min_capacity=0, device=0, take_first=None):
        """extractFilteredKmers(data, k, rounds): count every canonical k-window of the first `take_first` pairs
        (genome.takeFirst), then deleteAll(v < rounds).  `data` is always the WHOLE data set -- the same convention as
        GraphBuilder.startup, GraphSimplifier.startup and MapGraph.pairSupport: with `comm` the first `take_first` pairs of the
        file are selected and THEN every rank takes its slice of them (PairedEndData.take(...).shard(rank, world))."""
        if take_first is not None:
            data = data.take(int(take_first))
        if comm is not None:
            data = data.shard(comm.rank, comm.world)
        n_reads = data.n_reads
        kmers = PartitionedDNAMap(k, comm, min_capacity) if comm is not None else ArrayDNAMap(k, min_capacity, device)
        kmers.insert_reads(data, n_reads)
        kmers.delete_below(rounds)
        return kmers
