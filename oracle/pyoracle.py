"""ctypes view of oracle/liboracle.so -- the CPU restatement of the reference's k-mer -> graph path.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.h): importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never from genome_b200/.  PARITY UNPINNED.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    u64, i64, i32, vp = C.c_uint64, C.c_int64, C.c_int32, C.c_void_p
    sig = {
        "go_hash": (i32, [u64, C.c_int]),
        "go_revcomp": (u64, [u64, C.c_int]),
        "go_canonical": (u64, [u64, C.c_int, C.c_int]),
        "go_improve": (i32, [i32]),
        "go_prepend": (u64, [u64, C.c_int, C.c_int]),
        "go_append": (u64, [u64, C.c_int, C.c_int]),
        "go_map_new": (vp, [C.c_int, C.c_int, C.c_int]),
        "go_map_free": (None, [vp]),
        "go_map_k": (C.c_int, [vp]),
        "go_map_update1": (None, [vp, u64]),
        "go_map_update": (None, [vp, u64, i32]),
        "go_map_apply": (C.c_int, [vp, u64, C.POINTER(i32)]),
        "go_map_size": (i64, [vp]),
        "go_map_delete_below": (None, [vp, i32]),
        "go_map_export": (i64, [vp, vp, vp, i64]),
        "go_map_bins": (i64, [vp]),
        "go_count_windows": (i64, [vp, C.c_size_t, i64, C.c_int]),
        "go_insert_reads": (i64, [vp, vp, C.c_size_t, i64]),
        "go_insert_reads_mt": (i64, [vp, vp, C.c_size_t, i64, C.c_int]),
        "go_extract_canonical": (i64, [vp, C.c_size_t, i64, C.c_int, C.c_int, vp, i64]),
        "go_build_graph": (vp, [vp]),
        "go_graph_free": (None, [vp]),
        "go_graph_counts": (None, [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
        "go_graph_export": (None, [vp, vp, vp, vp, vp, vp, vp]),
        "go_graph_edge_ids": (None, [vp, vp]),
        "go_graph_components": (i64, [vp, vp]),
        "go_graph_retain_largest": (None, [vp]),
        "go_graph_simplify": (None, [vp]),
        "go_graph_remove_bubbles": (None, [vp]),
        "go_graph_remove_edges": (i64, [vp, vp, i64]),
        "go_graph_clip_tips": (i64, [vp, i64]),
        "go_graph_check": (C.c_int, [vp]),
        "go_graph_map": (i64, [vp, vp, vp, vp, i64]),
        "go_walk": (C.c_int, [vp, i64, i32, i64, i32, C.c_int, C.c_int, vp, i64, C.POINTER(i64)]),
        "go_pair_support": (i64, [vp, vp, C.c_size_t, i64, C.c_int, C.c_int, vp, vp, vp, i64, C.POINTER(i64), C.POINTER(i64)]),
        "go_graph_split": (i64, [vp, vp, vp, vp, i64, i32, C.POINTER(i64)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def hash_(v, variant=291):
    return lib().go_hash(int(v), variant)


def revcomp(x, k):
    return lib().go_revcomp(int(x), k)


def canonical(x, k, variant=291):
    return lib().go_canonical(int(x), k, variant)


class OracleMap:
    """PartitionedDNAMap[Int] over `partitions` ArrayDNAMaps (S/ds/PartitionedDNAMap.scala, ArrayDNAMap.scala)."""

    def __init__(self, k, partitions=1, variant=291):
        self.k = k
        self.h = lib().go_map_new(k, partitions, variant)
        if not self.h:
            raise ValueError("k must be in 1..31")

    def __del__(self):
        if getattr(self, "h", None):
            lib().go_map_free(self.h)
            self.h = None

    def update1(self, key):
        lib().go_map_update1(self.h, int(key))

    def update(self, key, v):
        lib().go_map_update(self.h, int(key), int(v))

    def apply(self, key):
        v = C.c_int32(0)
        return v.value if lib().go_map_apply(self.h, int(key), C.byref(v)) else None

    def contains(self, key):
        return bool(lib().go_map_apply(self.h, int(key), None))

    def size(self):
        return lib().go_map_size(self.h)

    def bins(self):
        return lib().go_map_bins(self.h)

    def delete_below(self, rounds):
        lib().go_map_delete_below(self.h, rounds)

    def insert_reads(self, bin_bytes, n_reads, threads=0):
        buf = np.ascontiguousarray(bin_bytes, dtype=np.uint8)
        if threads:
            return lib().go_insert_reads_mt(self.h, _ptr(buf), buf.size, n_reads, threads)
        return lib().go_insert_reads(self.h, _ptr(buf), buf.size, n_reads)

    def export(self):
        n = self.size()
        keys = np.empty(n, np.uint64)
        vals = np.empty(n, np.int32)
        lib().go_map_export(self.h, _ptr(keys), _ptr(vals), n)
        return keys, vals

    def export_sorted(self):
        keys, vals = self.export()
        o = np.argsort(keys, kind="stable")
        return keys[o], vals[o]


def count_windows(bin_bytes, n_reads, k):
    buf = np.ascontiguousarray(bin_bytes, dtype=np.uint8)
    return lib().go_count_windows(_ptr(buf), buf.size, n_reads, k)


def extract_canonical(bin_bytes, n_reads, k, variant=291):
    buf = np.ascontiguousarray(bin_bytes, dtype=np.uint8)
    n = lib().go_extract_canonical(_ptr(buf), buf.size, n_reads, k, variant, None, 0)
    out = np.empty(n, np.uint64)
    lib().go_extract_canonical(_ptr(buf), buf.size, n_reads, k, variant, _ptr(out), n)
    return out


class OracleGraph:
    """MapGraph built by Graph.buildGraph (S/data/graph/Graph.scala:269-382)."""

    def __init__(self, omap):
        self.k = omap.k
        self.h = lib().go_build_graph(omap.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().go_graph_free(self.h)
            self.h = None

    def counts(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        lib().go_graph_counts(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def export(self):
        nn, ne, nb = self.counts()
        node_kmer = np.empty(nn, np.uint64)
        node_id = np.empty(nn, np.int64)
        es = np.empty(ne, np.int64)
        ee = np.empty(ne, np.int64)
        off = np.empty(ne + 1, np.int64)
        bases = np.empty(nb, np.uint8)
        lib().go_graph_export(self.h, _ptr(node_kmer), _ptr(node_id), _ptr(es), _ptr(ee), _ptr(off), _ptr(bases))
        return node_kmer, node_id, es, ee, off, bases

    def edge_ids(self):
        ids = np.empty(self.counts()[1], np.int64)
        lib().go_graph_edge_ids(self.h, _ptr(ids))
        return ids

    def components(self):
        nn = self.counts()[0]
        label = np.empty(nn, np.int64)
        nc = lib().go_graph_components(self.h, _ptr(label))
        return nc, label

    def retain_largest(self):
        lib().go_graph_retain_largest(self.h)

    def simplify(self):
        lib().go_graph_simplify(self.h)

    def remove_bubbles(self):
        lib().go_graph_remove_bubbles(self.h)

    def clip_tips(self, max_len):
        return lib().go_graph_clip_tips(self.h, max_len)

    def check(self):
        return lib().go_graph_check(self.h)

    def graph_map(self):
        """Graph.getGraphMap (Graph.scala:90-119): (kmer, id, dist) arrays, dist 0 = node position."""
        n = lib().go_graph_map(self.h, None, None, None, 0)
        kmer = np.empty(n, np.uint64)
        ident = np.empty(n, np.int64)
        dist = np.empty(n, np.int32)
        lib().go_graph_map(self.h, _ptr(kmer), _ptr(ident), _ptr(dist), n)
        return kmer, ident, dist

    def walk(self, pos1, pos2, lo, hi):
        """WalkingActor.receive (GraphSimplifier.scala:77-126) for positions (id, dist): (good, sorted [(e1, e2)])."""
        n = C.c_int64()
        good = lib().go_walk(self.h, pos1[0], pos1[1], pos2[0], pos2[1], lo, hi, None, 0, C.byref(n))
        pairs = np.empty(2 * n.value, np.int64)
        lib().go_walk(self.h, pos1[0], pos1[1], pos2[0], pos2[1], lo, hi, _ptr(pairs), n.value, C.byref(n))
        return bool(good), sorted((int(pairs[2 * i]), int(pairs[2 * i + 1])) for i in range(n.value))

    def pair_support(self, bin_bytes, n_pairs, lo, hi):
        """The pair loop of GraphSimplifier.startup (188-263): (e1 ids, e2 ids, counts, badPairs, walked cases)."""
        buf = np.ascontiguousarray(bin_bytes, dtype=np.uint8)
        bad, walked = C.c_int64(), C.c_int64()
        n = lib().go_pair_support(self.h, _ptr(buf), buf.size, n_pairs, lo, hi, None, None, None, 0, C.byref(bad), C.byref(walked))
        if n < 0:
            raise ValueError("truncated .bin stream")
        e1 = np.empty(n, np.int64)
        e2 = np.empty(n, np.int64)
        cnt = np.empty(n, np.int32)
        lib().go_pair_support(self.h, _ptr(buf), buf.size, n_pairs, lo, hi, _ptr(e1), _ptr(e2), _ptr(cnt), n, C.byref(bad), C.byref(walked))
        return e1, e2, cnt, bad.value, walked.value

    def split(self, e1, e2, cnt, cutoff):
        """GraphSimplifier.startup 266-317 without simplifyGraph: (edges removed, nodes added)."""
        e1 = np.ascontiguousarray(e1, np.int64)
        e2 = np.ascontiguousarray(e2, np.int64)
        cnt = np.ascontiguousarray(cnt, np.int32)
        added = C.c_int64()
        removed = lib().go_graph_split(self.h, _ptr(e1), _ptr(e2), _ptr(cnt), e1.size, cutoff, C.byref(added))
        return removed, added.value
