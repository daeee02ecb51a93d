"""Host-side mirror of the first script of the reference, GraphBuilder.startup (S/scripts/GraphBuilder.scala:18-59, relative to
/root/reference), over the C ABI: extractFilteredKmers -> buildGraph -> components and their two
histograms -> retain(largest component).  The Kryo `write(outfile)` (54) has no counterpart: the graph stays
on the device for GraphSimplifier, or is exported as arrays."""
import numpy as np

from .dnamap import FreqFilter
from .graph import Graph


def component_histograms(n_components, label, edge_start, edge_off):
    """GraphBuilder.scala:41-47 on exported arrays.  `hist`: components grouped by node count -> [(size, how many)];
    `hist2`: components grouped by the summed length of their nodes' out-edges -> [(length, how many)]; both sorted by key."""
    label = np.asarray(label, np.int64)
    lens = np.diff(np.asarray(edge_off, np.int64))
    comp_nodes = np.bincount(label, minlength=n_components)
    if lens.size:
        comp_len = np.bincount(label[np.asarray(edge_start, np.int64)], weights=lens, minlength=n_components).astype(np.int64)
    else:
        comp_len = np.zeros(n_components, np.int64)
    v1, c1 = np.unique(comp_nodes, return_counts=True)
    v2, c2 = np.unique(comp_len, return_counts=True)
    return list(zip(v1.tolist(), c1.tolist())), list(zip(v2.tolist(), c2.tolist())), comp_nodes


class GraphBuilder:
    """`rounds = 3` is hard-coded in the reference (30); `genome.k` is read from a config key no config file defines (29,
    SURVEY Q15), so k is an argument here."""

    def __init__(self, k, rounds=3):
        self.k = int(k)
        self.rounds = int(rounds)
        self.log = {}

    def startup(self, data, comm=None, min_capacity=0):
        """Returns the MapGraph after `retain(maxComponent)`; self.log holds what the reference logs.  `comm`: a Communicator,
        then the map is a PartitionedDNAMap over its GPUs; `data` is the WHOLE data set on every rank (each takes its own slice, the
        same convention as GraphSimplifier.startup)."""
        kmers = FreqFilter.extractFilteredKmers(data, self.k, self.rounds, comm=comm, min_capacity=min_capacity)   # 32
        self.log["good_reads_count"] = kmers.size                      # "Good reads count: " (34) -- the kept k-mers, sic
        graph = Graph.buildGraph(self.k, kmers)                         # 36
        nn, ne, nb = graph.counts()
        self.log["total_edges_length"] = nb                             # 39
        n_comp, label = graph.components()                              # 37
        _, es, _, off, _ = graph.export()
        hist, hist2, comp_nodes = component_histograms(n_comp, label, es, off)
        self.log["components_histogram"] = hist                         # 41-42
        self.log["components_histogram_2"] = hist2                      # 44-47
        self.log["max_component_size"] = int(comp_nodes.max()) if n_comp else 0   # 52-53
        if n_comp:
            graph.retain_largest()                                      # 54; `maxBy` of an empty collection throws in the reference
        self.kmers = kmers
        return graph
