"""Host-side mirror of trait Graph / MapGraph / object Graph (S/data/graph/Graph.scala, paths relative to
/root/reference) over the C ABI.  Nodes and edges are addressed by their index in the current export: the
reference's ids are not reproducible even reference-vs-reference."""
import ctypes as C

import numpy as np

from . import capi
from .dnamap import PartitionedDNAMap


class MapGraph:
    """class MapGraph (Graph.scala:152-262) held in device memory."""

    def __init__(self, handle, k):
        self.h = handle
        self.k = k

    def close(self):
        if getattr(self, "h", None):
            capi.lib().gb_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counts(self):
        """(getNodes.size, getEdges.size, getEdges.map(_.seq.length).sum)"""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        capi.check(capi.lib().gb_graph_counts(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def export(self):
        """node_kmer u64[N], edge_start u32[E], edge_end u32[E], edge_off u64[E+1] (bases), bases u8[n_bases] (codes)."""
        nn, ne, nb = self.counts()
        node_kmer = np.empty(nn, np.uint64)
        es = np.empty(ne, np.uint32)
        ee = np.empty(ne, np.uint32)
        off = np.empty(ne + 1, np.uint64)
        packed = np.zeros((nb + 3) // 4, np.uint8)
        capi.check(capi.lib().gb_graph_export(self.h, capi.ptr(node_kmer), capi.ptr(es), capi.ptr(ee), capi.ptr(off), capi.ptr(packed)))
        bases = np.empty(packed.size * 4, np.uint8)
        for j in range(4):
            bases[j::4] = (packed >> (2 * j)) & 3
        return node_kmer, es, ee, off, bases[:nb]

    # ---- trait Graph
    def getNodes(self):
        return self.export()[0]

    def getEdges(self):
        """list of (start index, end index, seq codes)"""
        _, es, ee, off, bases = self.export()
        return [(int(es[i]), int(ee[i]), bases[int(off[i]):int(off[i + 1])]) for i in range(es.size)]

    def components(self):
        """Graph.components (54-72): (number of components, label per node)."""
        nn = self.counts()[0]
        label = np.zeros(nn, np.uint32)
        nc = C.c_int64()
        capi.check(capi.lib().gb_graph_components(self.h, capi.ptr(label), C.byref(nc)))
        return nc.value, label

    def retain_largest(self):
        """graph.retain(components.maxBy(_.size)) (GraphBuilder.scala:52-54)."""
        capi.check(capi.lib().gb_graph_retain_largest(self.h))

    def retain(self, node_keep):
        """MapGraph.retain(nodesSet) (161-165): `node_keep` = boolean per node of the current export."""
        keep = np.ascontiguousarray(node_keep, dtype=np.uint8)
        if keep.size != self.counts()[0]:
            raise ValueError("one flag per node of the current graph")
        capi.check(capi.lib().gb_graph_retain(self.h, capi.ptr(keep)))

    def edit(self, replace=(), add_nodes=(), add_edges=(), remove_nodes=()):
        """The fine-grained mutators of trait Graph in bulk (gb_graph_edit), applied in this order:
        replace = [(edge, new_start or None, new_end or None)]  -- replaceStart / replaceEnd (197-209)
        add_nodes = [kmer u64], add_edges = [(start, end, base codes)]  -- addNode / addEdge (172-184); returns their indices
        remove_nodes = [node]  -- removeNode (185-187)."""
        NONE = 0xFFFFFFFF
        nn, ne, _ = self.counts()
        ridx = np.array([r[0] for r in replace], np.uint32)
        rs = np.array([NONE if r[1] is None else r[1] for r in replace], np.uint32)
        re_ = np.array([NONE if r[2] is None else r[2] for r in replace], np.uint32)
        nk = np.array(list(add_nodes), np.uint64)
        es = np.array([e[0] for e in add_edges], np.uint32)
        ee = np.array([e[1] for e in add_edges], np.uint32)
        off = np.zeros(len(add_edges) + 1, np.uint64)
        off[1:] = np.cumsum([len(e[2]) for e in add_edges]) if add_edges else []
        codes = np.concatenate([np.asarray(e[2], np.uint8) for e in add_edges]) if add_edges else np.zeros(0, np.uint8)
        rm = np.array(list(remove_nodes), np.uint32)
        capi.check(capi.lib().gb_graph_edit(self.h, ridx.size, capi.ptr(ridx), capi.ptr(rs), capi.ptr(re_), nk.size, capi.ptr(nk), es.size,
                                            capi.ptr(es), capi.ptr(ee), capi.ptr(off), capi.ptr(codes), rm.size, capi.ptr(rm)))
        return list(range(nn, nn + nk.size)), list(range(ne, ne + es.size))

    def write(self, path):
        """MapGraph.write(file) (Graph.scala:232-248): the Kryo `graph` file (formats.write_kryo_graph; ids = index + 1)."""
        from . import formats
        node_kmer, es, ee, off, bases = self.export()
        seqs = [bases[int(off[i]):int(off[i + 1])] for i in range(es.size)]
        with open(path, "wb") as f:
            f.write(formats.write_kryo_graph(self.k, node_kmer, es, ee, seqs))

    def simplifyGraph(self):
        capi.check(capi.lib().gb_graph_simplify(self.h))

    def removeBubbles(self):
        capi.check(capi.lib().gb_graph_remove_bubbles(self.h))

    def removeEdges(self, edge_idx):
        idx = np.ascontiguousarray(edge_idx, dtype=np.uint32)
        capi.check(capi.lib().gb_graph_remove_edges(self.h, capi.ptr(idx), idx.size))

    def clipTips(self, max_len):
        """EXTENSION (no reference counterpart)."""
        n = C.c_int64()
        capi.check(capi.lib().gb_graph_clip_tips(self.h, int(max_len), C.byref(n)))
        return n.value

    def check(self):
        capi.check(capi.lib().gb_graph_check(self.h))

    def getGraphMap(self):
        """Graph.getGraphMap (90-119) as arrays: (kmer u64[n], id u32[n], dist u32[n]); dist 0 = NodeGraphPosition(id),
        dist >= 1 = EdgeGraphPosition(id, dist)."""
        n = C.c_int64()
        capi.check(capi.lib().gb_graph_positions(self.h, None, None, None, 0, C.byref(n)))
        kmer = np.empty(n.value, np.uint64)
        ident = np.empty(n.value, np.uint32)
        dist = np.empty(n.value, np.uint32)
        if n.value:
            capi.check(capi.lib().gb_graph_positions(self.h, capi.ptr(kmer), capi.ptr(ident), capi.ptr(dist), n.value, C.byref(n)))
        return kmer, ident, dist

    def graphMap(self):
        """Graph.getGraphMap (90-119) as a device-resident DNAMap[GraphPosition] (a snapshot of the current graph)."""
        h = C.c_void_p()
        capi.check(capi.lib().gb_graph_map_create(self.h, C.byref(h)))
        return GraphPositionMap(h, self.k)

    def pairSupport(self, data, takeFirst=None, range_=(180, 250), comm=None):
        """The pair loop of GraphSimplifier.startup (S/scripts/GraphSimplifier.scala:188-263): graphMap.getAll of the first
        k-mers of both reads in both orientations, annotate, and the WalkingActor walks (33-127), for the first `takeFirst`
        pairs of `data` (a PairedEndData).  Returns (support uint32[n_edges, 4], badPairs, walked cases) with
        support[e1, b] = pathsMap((e1, e2)), e2 = the out-edge of e1's end node starting with base b; `range_` is the
        reference's `180 to 250` (153).  `comm` (a Communicator): the pair loop is split over its ranks."""
        n_pairs = data.count if takeFirst is None else min(int(takeFirst), data.count)
        ne = self.counts()[1]
        support = np.zeros((ne, 4), np.uint32)
        bad, walked = C.c_int64(), C.c_int64()
        if comm is not None and comm.world > 1:
            # every rank holds the same graph (Graph.buildGraph over a PartitionedDNAMap) and ALL pairs: it walks its own
            # slice of the first n_pairs pairs, then the per-rank counts are summed (pairs are independent, 213-248)
            mine = data.take(n_pairs).shard(comm.rank, comm.world)
            capi.check(capi.lib().gb_graph_pair_support(self.h, capi.ptr(mine.bin), mine.bin.size, mine.count, int(range_[0]),
                                                        int(range_[1]), capi.ptr(support), C.byref(bad), C.byref(walked)))
            comm.allreduce_sum(support)
            tot = comm.allreduce_sum(np.array([bad.value, walked.value], np.int64))
            return support, int(tot[0]), int(tot[1])
        capi.check(capi.lib().gb_graph_pair_support(self.h, capi.ptr(data.bin), data.bin.size, n_pairs, int(range_[0]), int(range_[1]),
                                                    capi.ptr(support), C.byref(bad), C.byref(walked)))
        return support, bad.value, walked.value

    def splitNodes(self, support, cutoff):
        """The node sweep of GraphSimplifier.startup (268-316): in x out matrices thresholded at `cutoff`, one node copy per
        supported component, unsupported edges removed.  Returns (edges removed, nodes added); call simplifyGraph next (318)."""
        support = np.ascontiguousarray(support, dtype=np.uint32)
        if support.size != 4 * self.counts()[1]:
            raise ValueError("support must have 4 entries per edge of the current graph")
        removed, added = C.c_int64(), C.c_int64()
        capi.check(capi.lib().gb_graph_split_nodes(self.h, capi.ptr(support), int(cutoff), C.byref(removed), C.byref(added)))
        return removed.value, added.value

    def stats(self):
        s = (C.c_int64 * 8)()
        capi.check(capi.lib().gb_graph_stats(self.h, s))
        # s[4] is shared: a graph fresh from the sharded build reports its segment count there until pairSupport overwrites it
        return dict(kept_kmers=s[0], jump_launches=s[1], cycle_vertices=s[2], build_ns=s[3], segments=s[4],
                    pair_support_map_ns=s[4], pair_support_filter_ns=s[5], pair_support_walk_ns=s[6], pair_support_cases=s[7])


class GraphPositionMap:
    """The DNAMap[GraphPosition] of Graph.getGraphMap (Graph.scala:92-117): `size`, `getAll`, `contains` of trait DNAMap
    (S/ds/ArrayDNAMap.scala:49-60) in bulk.  Keys are k-mers as u64, oriented as given (the map holds both strands)."""

    def __init__(self, handle, k):
        self.h = handle
        self.k = k

    def close(self):
        if getattr(self, "h", None):
            capi.lib().gb_graph_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self):
        n = C.c_int64()
        capi.check(capi.lib().gb_graph_map_size(self.h, C.byref(n)))
        return n.value

    def getAll(self, keys, max_per_key=4):
        """(counts u32[n], ids u32[n, max_per_key], dists u32[n, max_per_key]); dist 0 = NodeGraphPosition(id), dist >= 1 =
        EdgeGraphPosition(id, dist); unused entries are 0xFFFFFFFF.  counts[i] may exceed max_per_key (multimap, Q13)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        n = keys.size
        counts = np.zeros(n, np.uint32)
        ids = np.full((n, max_per_key), 0xFFFFFFFF, np.uint32)
        dists = np.full((n, max_per_key), 0xFFFFFFFF, np.uint32)
        capi.check(capi.lib().gb_graph_map_get_all(self.h, capi.ptr(keys), n, int(max_per_key), capi.ptr(ids), capi.ptr(dists),
                                                   capi.ptr(counts)))
        return counts, ids, dists

    def contains(self, keys):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.zeros(keys.size, np.uint32)
        capi.check(capi.lib().gb_graph_map_get_all(self.h, capi.ptr(keys), keys.size, 0, None, None, capi.ptr(counts)))
        return counts > 0


class Graph:
    """object Graph (Graph.scala:264-392)."""

    @staticmethod
    def buildGraph(k, kmersFreq):
        """Graph.buildGraph(k, kmersFreq) (269-382) on the map's current contents."""
        if k != kmersFreq.k:
            raise ValueError("k differs from the map's k")
        h = C.c_void_p()
        fn = capi.lib().gb_pmap_graph_build if isinstance(kmersFreq, PartitionedDNAMap) else capi.lib().gb_graph_build
        capi.check(fn(kmersFreq.h, C.byref(h)))
        return MapGraph(h, k)

    @staticmethod
    def apply(path, device=0):
        """Graph(file) (Graph.scala:384-390): reads a Kryo `graph` file into device memory -- an empty graph (buildGraph of an
        empty map) filled by the bulk addNode / addEdge of gb_graph_edit."""
        from . import formats
        from .dnamap import ArrayDNAMap
        with open(path, "rb") as f:
            nodes, edges = formats.read_kryo_graph(f.read())
        k, node_kmer, es, ee, seqs = formats.kryo_graph_arrays(nodes, edges)
        if k == 0:
            raise ValueError("a graph file without nodes does not say its k")
        empty = ArrayDNAMap(k, device=device)
        try:
            g = Graph.buildGraph(k, empty)
        finally:
            empty.close()
        g.edit(add_nodes=node_kmer.tolist(), add_edges=[(int(es[i]), int(ee[i]), seqs[i]) for i in range(es.size)])
        return g

    @staticmethod
    def buildGraphVirtualShards(k, kmersFreq, n_shards):
        """The sharded form of Graph.buildGraph (csrc/sgraph.cuh: minimizer re-routing, rank-local list ranking, segment list)
        over `n_shards` virtual ranks on the map's one device -- the multi-GPU algorithm, checkable on a single GPU."""
        if k != kmersFreq.k:
            raise ValueError("k differs from the map's k")
        if isinstance(kmersFreq, PartitionedDNAMap):
            raise ValueError("virtual shards run on a single-GPU map; a PartitionedDNAMap shards over its own GPUs (Graph.buildGraph)")
        h = C.c_void_p()
        capi.check(capi.lib().gb_graph_build_virtual_shards(kmersFreq.h, int(n_shards), C.byref(h)))
        return MapGraph(h, k)
