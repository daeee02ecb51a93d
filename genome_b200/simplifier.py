"""Host-side mirror of the second script of the reference, GraphSimplifier.startup (S/scripts/GraphSimplifier.scala:138-357,
relative to /root/reference), over the C ABI -- SURVEY 8(f) row 4.  The Akka plumbing (WalkingActor deployment, futures, the
500-permit semaphore of 250-261) has no counterpart: the pair loop is one bulk call."""
import numpy as np

from . import formats


class GraphSimplifier:
    """`range` is hard-coded to 180 to 250 in the reference (153); `cutoff` and `takeFirst` come from its config
    (genome.cutoff = 200, genome.takeFirst, application.conf:70-71)."""

    def __init__(self, range_=(180, 250), cutoff=200, takeFirst=None):
        self.range = (int(range_[0]), int(range_[1]))
        self.cutoff = int(cutoff)
        self.takeFirst = takeFirst
        self.log = {}

    def startup(self, graph, data, outfile=None, contigs=None, comm=None):
        """One iteration (`for (it <- 0 until 1)`, 156) on a MapGraph and a PairedEndData; the graph is modified in place.
        `outfile` receives the per-node matrices (line format of 312), `contigs` the contig file (338-347)."""
        graph.check()                                                   # the asserts of 159-170
        support, bad, walked = graph.pairSupport(data, self.takeFirst, self.range, comm)  # comm: pairs split over the GPUs
        self.log["bad_pairs"] = bad                                     # "Bad pairs: " (265)
        self.log["walked_cases"] = walked
        if outfile is not None:
            self._write_matrices(graph, support, outfile)
        self.log["edges_before"] = graph.counts()[1]                    # "Edges before: " (315)
        removed, added = graph.splitNodes(support, self.cutoff)
        self.log["edges_removed"], self.log["nodes_added"] = removed, added
        graph.simplifyGraph()                                           # 318 (removeBubbles stays commented out, 317)
        nn, ne, nb = graph.counts()
        self.log["edges_after"] = ne                                    # "Edges after: " (319)
        self.log["total_edges_length"] = nb                             # 323
        n_comp, label = graph.components()
        label = label.astype(np.int64)
        _, es, ee, off, _ = graph.export()
        lens = np.diff(off.astype(np.int64))
        # hist2 (325-328): components grouped by (node count, summed out-edge length)
        comp_nodes = np.bincount(label, minlength=n_comp)
        comp_len = np.bincount(label[es], weights=lens, minlength=n_comp).astype(np.int64) if ne else np.zeros(n_comp, np.int64)
        hist = {}
        for a, b in zip(comp_nodes.tolist(), comp_len.tolist()):
            hist[(a, b)] = hist.get((a, b), 0) + 1
        self.log["components_histogram"] = sorted(hist.items())
        self.log["max_component_size"] = int(comp_nodes.max()) if n_comp else 0   # 330-331
        vals, counts = np.unique(lens, return_counts=True)
        self.log["edge_lengths"] = dict(zip(vals.tolist(), counts.tolist()))     # 335-336
        if contigs is not None:
            formats.write_contigs([seq for (_, _, seq) in graph.getEdges()], contigs)
        return graph

    @staticmethod
    def _write_matrices(graph, support, path):
        """`node.id -> matrix in-lengths out-lengths` for every node with in- and out-edges (268-269, 312).  Node and edge
        order inside a line follow this library's indices (the reference's follow Scala Set / Map iteration order)."""
        node_kmer, es, ee, off, bases = graph.export()
        lens = np.diff(off.astype(np.int64))
        ins = [[] for _ in range(node_kmer.size)]
        outs = [[] for _ in range(node_kmer.size)]
        for e in range(es.size):
            outs[es[e]].append(e)
            ins[ee[e]].append(e)
        with open(path, "w") as f:
            for v in range(node_kmer.size):
                if not ins[v] or not outs[v]:
                    continue
                m = [[int(support[i, bases[int(off[j])]]) for j in outs[v]] for i in ins[v]]
                rows = ", ".join("Array(" + ", ".join(str(x) for x in r) + ")" for r in m)
                f.write("%d -> Array(%s) List(%s) List(%s)\n" % (v, rows, ", ".join(str(int(lens[i])) for i in ins[v]),
                                                                ", ".join(str(int(lens[j])) for j in outs[v])))
