"""Worker of tests/test_sgraph_emul_cpu.py::test_one_process_per_rank_gloo (launched by torch.distributed.run, backend gloo).

The sharded Graph.buildGraph (genome_b200/csrc/sgraph.cuh, compiled with g++ in tests/emul/sgraph_emul.cpp) with ONE PROCESS
PER RANK: the Fabric's collectives are callbacks implemented here with torch.distributed over gloo, the peer windows are
files in /dev/shm mapped by every process -- peers live in other address spaces and are seen at addresses of this process's
own, like CUDA-IPC mappings; what NCCL does on the GPU box, gloo does here.  Checks every rank's result against the oracle."""
import ctypes as C
import mmap
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle  # noqa: E402
from tests import helpers as H  # noqa: E402

MAXR = 16
_vp = C.c_void_p


class Callbacks(C.Structure):
    _fields_ = [("allgather", C.CFUNCTYPE(C.c_int, _vp, _vp, C.c_uint64)),
                ("window", C.CFUNCTYPE(C.c_int, C.c_uint64, C.POINTER(_vp), C.POINTER(_vp))),
                ("alltoallv", C.CFUNCTYPE(C.c_int, _vp, _vp, _vp, _vp, _vp, _vp)),
                ("barrier", C.CFUNCTYPE(C.c_int)),
                ("allgatherv", C.CFUNCTYPE(C.c_int, _vp, _vp, _vp)),
                ("allreduce_sum", C.CFUNCTYPE(C.c_int, _vp, C.c_uint64, C.c_int))]


def view(ptr, n, ctype, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(int(n),))


class GlooFabric:
    def __init__(self, rank, world, tag):
        self.rank, self.world, self.tag = rank, world, tag
        self.maps, self.files, self.seq = [], [], 0
        self.cb = Callbacks(Callbacks._fields_[0][1](self.allgather), Callbacks._fields_[1][1](self.window),
                            Callbacks._fields_[2][1](self.alltoallv), Callbacks._fields_[3][1](self.barrier),
                            Callbacks._fields_[4][1](self.allgatherv), Callbacks._fields_[5][1](self.allreduce_sum))

    def guarded(fn):
        def wrapper(self, *a):
            try:
                return fn(self, *a)
            except Exception as e:   # an exception must not unwind through the C caller
                sys.stderr.write("fabric callback %s failed: %r\n" % (fn.__name__, e))
                return -7
        return wrapper

    @guarded
    def allgather(self, mine, all_, nbytes):
        t = torch.from_numpy(view(mine, nbytes, C.c_uint8, np.uint8).copy())
        outs = [torch.empty(int(nbytes), dtype=torch.uint8) for _ in range(self.world)]
        dist.all_gather(outs, t)
        view(all_, nbytes * self.world, C.c_uint8, np.uint8)[:] = torch.cat(outs).numpy()
        return 0

    def _map(self, path, size, create):
        fd = os.open(path, os.O_RDWR | (os.O_CREAT if create else 0), 0o600)
        if create:
            os.ftruncate(fd, size)
        mm = mmap.mmap(fd, size)
        os.close(fd)
        self.maps.append(mm)
        return C.addressof(C.c_char.from_buffer(mm))

    @guarded
    def window(self, my_bytes, mine_pp, peers_pp):
        self.seq += 1
        name = lambda r: "/dev/shm/gbsg_%s_%d_%d" % (self.tag, self.seq, r)
        size = max(int(my_bytes), 16)
        mine = self._map(name(self.rank), size, True)
        self.files.append(name(self.rank))
        dist.barrier()   # every window exists
        for r in range(self.world):
            peers_pp[r] = mine if r == self.rank else self._map(name(r), os.path.getsize(name(r)), False)
        mine_pp[0] = mine
        dist.barrier()   # every rank has its mappings
        return 0

    @guarded
    def alltoallv(self, send, soff, scnt, recv, roff, rcnt):
        soff, scnt = view(soff, MAXR, C.c_uint64, np.uint64), view(scnt, MAXR, C.c_uint64, np.uint64)
        roff, rcnt = view(roff, MAXR, C.c_uint64, np.uint64), view(rcnt, MAXR, C.c_uint64, np.uint64)
        reqs, keep = [], []
        for p in range(self.world):
            ns, nr = int(scnt[p]), int(rcnt[p])
            if p == self.rank:
                assert ns == nr
                if ns:
                    C.memmove(recv + 8 * int(roff[p]), send + 8 * int(soff[p]), 8 * ns)
                continue
            if nr:
                t = torch.from_numpy(view(recv + 8 * int(roff[p]), nr, C.c_int64, np.int64))
                keep.append(t)
                reqs.append(dist.irecv(t, src=p))
            if ns:
                t = torch.from_numpy(view(send + 8 * int(soff[p]), ns, C.c_int64, np.int64))
                keep.append(t)
                reqs.append(dist.isend(t, dst=p))
        for r in reqs:
            r.wait()
        return 0

    @guarded
    def barrier(self):
        dist.barrier()
        return 0

    @guarded
    def allgatherv(self, buf, off, cnt):
        off, cnt = view(off, self.world, C.c_uint64, np.uint64), view(cnt, self.world, C.c_uint64, np.uint64)
        for p in range(self.world):
            if int(cnt[p]):
                dist.broadcast(torch.from_numpy(view(buf + 8 * int(off[p]), int(cnt[p]), C.c_int64, np.int64)), src=p)
        return 0

    @guarded
    def allreduce_sum(self, buf, count, elem):
        if count:
            # two's-complement sums wrap like the unsigned ones
            a = view(buf, count, C.c_int64, np.int64) if elem == 8 else view(buf, count, C.c_int32, np.int32)
            dist.all_reduce(torch.from_numpy(a))
        return 0

    def close(self):
        dist.barrier()
        self.maps.clear()
        for f in self.files:
            try:
                os.unlink(f)
            except OSError:
                pass


def run_build(lib, fab, k, dual, mine, rank, world):
    mine = np.ascontiguousarray(mine, np.uint64)
    out = np.zeros(8, np.uint64)
    p = lambda a: a.ctypes.data_as(_vp)
    rc = lib.emul_sharded_build_rank(k, int(dual), 0, world, rank, p(mine), C.c_uint64(mine.size), C.byref(fab.cb), p(out), None, None, None, None, None)
    assert rc == 0, rc
    N, E, B = int(out[0]), int(out[1]), int(out[2])
    node_kmer, es, ee = np.zeros(max(N, 1), np.uint64), np.zeros(max(E, 1), np.uint32), np.zeros(max(E, 1), np.uint32)
    eo, words = np.zeros(E + 1, np.uint64), np.zeros((B + 15) // 16 + 1, np.uint32)
    rc = lib.emul_sharded_build_rank(k, int(dual), 0, world, rank, p(mine), C.c_uint64(mine.size), C.byref(fab.cb), p(out), p(node_kmer), p(es), p(ee),
                                     p(eo), p(words))
    assert rc == 0, rc
    bases = np.zeros(words.size * 16, np.uint8)
    for j in range(16):
        bases[j::16] = (words >> np.uint32(2 * j)) & 3
    nk = [int(x) for x in node_kmer[:N]]
    edges = sorted((nk[int(es[i])], nk[int(ee[i])], bases[int(eo[i]):int(eo[i + 1])].tobytes()) for i in range(E))
    return (sorted(nk), edges), (N, E, B), dict(segments=int(out[4]), cycle_vertices=int(out[5]))


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = C.CDLL(os.path.join(ROOT, "tests", "_build", "libsgraph_emul.so"))
    fab = GlooFabric(rank, world, os.environ.get("MASTER_PORT", "0"))
    for (k, glen, rl, cov, err, rounds) in [(31, 20000, 100, 30, 0.01, 3), (15, 5000, 60, 30, 0.01, 2), (8, 1500, 40, 10, 0.0, 1), (4, 120, 20, 6, 0.0, 1)]:
        b, n, _ = H.small_reads(glen, rl, cov, err, seed=2000 + k)   # same bytes on every rank
        om, _ = H.oracle_counts(b, n, k)
        om.delete_below(rounds)
        og = pyoracle.OracleGraph(om)
        keys, _ = om.export()
        keys = np.ascontiguousarray(keys, np.uint64)
        want = H.canon_oracle_graph(og)
        # even slices, then everything on the last rank (the others still take part in every collective)
        for cuts in ([keys.size * r // world for r in range(world + 1)], [0] * world + [keys.size]):
            got, counts, st = run_build(lib, fab, k, False, keys[cuts[rank]:cuts[rank + 1]], rank, world)
            assert counts == og.counts(), (rank, k, counts, og.counts())
            assert got == want, (rank, k)
            if world > 1 and k >= 15:
                assert st["segments"] > 0
    fab.close()
    if rank == 0:
        print("SGRAPH GLOO OK world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
