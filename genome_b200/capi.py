"""ctypes binding of libgenome_b200.so (include/genome_b200.h).  This is the harness a JNA/JNI shim stands in for
on the JVM: plain pointers and sizes, no torch types.

There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgenome_b200.so")

GB_OK = 0
ERRORS = {-1: "GB_E_ARG", -2: "GB_E_K_RANGE", -3: "GB_E_OOM", -4: "GB_E_CUDA", -5: "GB_E_NCCL", -6: "GB_E_CAPACITY",
          -7: "GB_E_INVARIANT", -8: "GB_E_STATE"}
GB_FLAG_HASH_SCALA_210 = 1
GB_UNIQUE_ID_BYTES = 128


class GenomeError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "GB_E_?"), code, text))
        self.code = code
        self.name = ERRORS.get(code, "GB_E_?")


_u64, _i64, _i32, _u32, _vp, _sz = C.c_uint64, C.c_int64, C.c_int32, C.c_uint32, C.c_void_p, C.c_size_t
_pp = C.POINTER(C.c_void_p)
_pi64 = C.POINTER(C.c_int64)

# every symbol include/genome_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "gb_last_error": (C.c_char_p, []),
    "gb_version": (C.c_int, []),
    "gb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "gb_host_alloc": (C.c_int, [_sz, _pp]),
    "gb_host_free": (C.c_int, [_vp]),
    "gb_map_create": (C.c_int, [C.c_int, _i64, C.c_int, _u32, _pp]),
    "gb_map_destroy": (C.c_int, [_vp]),
    "gb_map_insert_reads": (C.c_int, [_vp, _vp, _sz, _i64, _pi64]),
    "gb_map_insert_reads_device": (C.c_int, [_vp, _vp, _sz, _vp, _i64, _pi64]),
    "gb_map_insert_records_device": (C.c_int, [_vp, _vp, _sz, C.c_uint32, _i64, C.c_uint32, _pi64]),
    "gb_map_update_counts": (C.c_int, [_vp, _vp, _i64]),
    "gb_map_update": (C.c_int, [_vp, _vp, _vp, _i64]),
    "gb_map_size": (C.c_int, [_vp, _pi64]),
    "gb_map_lookup": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "gb_map_delete_below": (C.c_int, [_vp, _i32]),
    "gb_map_export": (C.c_int, [_vp, _vp, _vp, _i64, _pi64]),
    "gb_map_neighbour_masks": (C.c_int, [_vp, _vp, _i64, _vp]),
    "gb_map_clear": (C.c_int, [_vp, _i64]),
    "gb_sync": (C.c_int, [_vp]),
    "gb_timer_start": (C.c_int, [_vp]),
    "gb_timer_stop": (C.c_int, [_vp, _pi64]),
    "gb_launch_count": (C.c_longlong, []),
    "gb_bench_random_atomics": (C.c_int, [C.c_int, _sz, _i64, C.c_int, _pi64]),
    "gb_bench_smem_upsert": (C.c_int, [C.c_int, C.c_int, _i64, _i64, C.c_int, C.c_int, _pi64, _pi64]),
    "gb_bench_l2_requests": (C.c_int, [C.c_int, _sz, _i64, C.c_int, C.c_int, _pi64]),
    "gb_tune": (C.c_int, [C.c_char_p, _i64, _pi64]),
    "gb_tune_get": (C.c_int, [C.c_char_p, _pi64]),
    "gb_map_stats": (C.c_int, [_vp, _pi64]),
    "gb_map_phase_ns": (C.c_int, [_vp, _pi64]),
    "gb_graph_build": (C.c_int, [_vp, _pp]),
    "gb_graph_build_virtual_shards": (C.c_int, [_vp, C.c_int, _pp]),
    "gb_graph_destroy": (C.c_int, [_vp]),
    "gb_graph_counts": (C.c_int, [_vp, _pi64, _pi64, _pi64]),
    "gb_graph_export": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "gb_graph_components": (C.c_int, [_vp, _vp, _pi64]),
    "gb_graph_retain_largest": (C.c_int, [_vp]),
    "gb_graph_retain": (C.c_int, [_vp, _vp]),
    "gb_graph_edit": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "gb_graph_simplify": (C.c_int, [_vp]),
    "gb_graph_remove_bubbles": (C.c_int, [_vp]),
    "gb_graph_remove_edges": (C.c_int, [_vp, _vp, _i64]),
    "gb_graph_clip_tips": (C.c_int, [_vp, _i64, _pi64]),
    "gb_graph_check": (C.c_int, [_vp]),
    "gb_graph_stats": (C.c_int, [_vp, _pi64]),
    "gb_graph_positions": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _pi64]),
    "gb_graph_map_create": (C.c_int, [_vp, _pp]),
    "gb_graph_map_destroy": (C.c_int, [_vp]),
    "gb_graph_map_size": (C.c_int, [_vp, _pi64]),
    "gb_graph_map_get_all": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp, _vp, _vp]),
    "gb_graph_pair_support": (C.c_int, [_vp, _vp, _sz, _i64, C.c_int, C.c_int, _vp, _pi64, _pi64]),
    "gb_graph_split_nodes": (C.c_int, [_vp, _vp, _i32, _pi64, _pi64]),
    "gb_comm_allreduce_sum_u32": (C.c_int, [_vp, _vp, _i64]),
    "gb_comm_allreduce_sum_i64": (C.c_int, [_vp, _vp, _i64]),
    "gb_comm_unique_id": (C.c_int, [_vp]),
    "gb_comm_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _pp]),
    "gb_comm_destroy": (C.c_int, [_vp]),
    "gb_pmap_create": (C.c_int, [_vp, C.c_int, _i64, _u32, _pp]),
    "gb_pmap_insert_reads": (C.c_int, [_vp, _vp, _sz, _i64, _pi64]),
    "gb_pmap_insert_reads_device": (C.c_int, [_vp, _vp, _sz, _vp, _i64, _pi64]),
    "gb_pmap_size": (C.c_int, [_vp, _pi64]),
    "gb_pmap_delete_below": (C.c_int, [_vp, _i32]),
    "gb_pmap_lookup": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "gb_pmap_graph_build": (C.c_int, [_vp, _pp]),
    "gb_pmap_owner": (C.c_int, [_vp, _vp, _i64, _vp]),
    "gb_owner_of": (C.c_int, [_vp, _i64, C.c_int, _vp]),
    "gb_owner_of_minimizer": (C.c_int, [_vp, _i64, C.c_int, C.c_int, _vp]),
}

_LIB = None


def lib():
    """Load libgenome_b200.so; raise if it has not been built (python -m genome_b200.build)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libgenome_b200.so is missing: build it with `python -m genome_b200.build` "
                               "(there is no CPU fallback)")
        # kernels are loaded when the library is, not at their first launch (lazy loading stalls first calls by 10..400 ms);
        # only effective if CUDA has not been initialised in this process yet
        os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _LIB = L
    return _LIB


TUNE_ENV = "GENOME_B200_TUNE"   # "key=value,key=value": read by THIS harness (bench.py, the multi-process test workers), not by the library


def tune_from_env():
    """Apply GENOME_B200_TUNE (harness-side convenience for processes started by torchrun / gpurun scripts)."""
    spec = os.environ.get(TUNE_ENV, "")
    out = {}
    for item in filter(None, (x.strip() for x in spec.split(","))):
        key, _, val = item.partition("=")
        set_tune(key.strip(), int(val or "1"))
        out[key.strip()] = int(val or "1")
    return out


def set_tune(name, value):
    """gb_tune: returns the previous value."""
    prev = C.c_int64(0)
    check(lib().gb_tune(name.encode(), int(value), C.byref(prev)))
    return prev.value


def get_tune(name):
    v = C.c_int64(0)
    check(lib().gb_tune_get(name.encode(), C.byref(v)))
    return v.value


class tuned:
    """with capi.tuned(insert_path=2, single_pass_min=1): ...  -- sets the keys, restores them on exit."""

    def __init__(self, **kw):
        self.kw = kw
        self.prev = {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.prev[k] = set_tune(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.prev.items():
            set_tune(k, v)
        return False


def check(code):
    if code != GB_OK:
        raise GenomeError(code, lib().gb_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a numpy array (None -> NULL) or of a raw integer address."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def as_u64(keys):
    return np.ascontiguousarray(keys, dtype=np.uint64)
