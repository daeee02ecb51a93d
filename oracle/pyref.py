"""A SECOND, independent reading of the reference's algorithm for this path, in plain Python: dictionaries, sets and lists
standing where the Scala has its collections, one statement per statement.  Test infrastructure like the rest of oracle/ (only
tests/ imports it; small inputs only).  It exists because the C oracle (oracle.c) is itself a restatement that nothing in
the reference can pin (no tests, no goldens, no JVM here: DESIGN.md section 6): two restatements written apart from each other
-- this one from the Scala alone, with Python's big integers masked to 64 bits where the Scala has a Long -- that agree on
every seeded input are the strongest check of the oracle available in this image (tests/test_pyref_cpu.py).

Citations are relative to /root/reference, S/ = src/main/scala/ru/ifmo/genome/.  What is NOT reproduced: iteration orders of
JDK / Scala hash collections (ids, the order of the simplifyGraph sweep) -- results are compared on canonical forms, and the
sweep is run in several orders by the tests to show that the canonical result does not depend on it.
"""
M64 = (1 << 64) - 1
BITS, N = 2, 32          # S/dna/DNASeq.scala:232-233 (bits, n = 64 / bits)


# ---------------------------------------------------------------------------------------------- S/dna/Base.scala:13-19
A, G, C, T = 0, 1, 2, 3
COMPLEMENT = {A: T, T: A, G: C, C: G}


# ---------------------------------------------------------------------------------------------- Long1DNASeq (DNASeq.scala:73-169)
class Seq1:
    """class Long1DNASeq(val long: Long, val len: Byte): base i at bits 2i."""
    __slots__ = ("long", "len")

    def __init__(self, long, length):
        self.long = long & M64
        self.len = length

    def key(self):
        return (self.long, self.len)

    def __eq__(self, o):                                   # 110-113
        return self.len == o.len and self.long == o.long

    def __hash__(self):
        return hash(self.key())

    def apply(self, i):                                    # 80-85
        assert 0 <= i < self.len
        return (self.long >> (i * BITS)) & 3

    def subseq(self, l, r):                                # 117-121
        assert 0 <= l <= r <= self.len
        mask = M64 if r - l == N else (1 << (BITS * (r - l))) - 1
        return Seq1((self.long >> (BITS * l)) & mask, r - l)

    def take(self, n):                                     # 123
        return self.subseq(0, min(max(n, 0), self.len))

    def drop(self, n):                                     # 125
        return self.subseq(min(max(n, 0), self.len), self.len)

    def sliding(self, size):                               # 127-133 (step 1)
        if self.len < size:
            return [self]
        return [self.subseq(i, i + size) for i in range(0, self.len - size + 1)]

    def prepend(self, base):                               # +: 135-143
        assert self.len < N, "the generic builder path is not needed for k <= 31"
        return Seq1((self.long << BITS) | base, self.len + 1)

    def append(self, base):                                # :+ 145-153
        assert self.len < N, "for length == n the reference falls into super.+: (a prepend): k = 32 is broken there"
        return Seq1(self.long | (base << (self.len * BITS)), self.len + 1)

    def reverse(self):                                     # 155-163; Scala precedence: shifts bind tighter than &, & tighter than |
        i = self.long
        i = (((i & 0x3333333333333333) << 2) | ((i >> 2) & 0x3333333333333333)) & M64
        i = (((i & 0x0f0f0f0f0f0f0f0f) << 4) | ((i >> 4) & 0x0f0f0f0f0f0f0f0f)) & M64
        i = (((i & 0x00ff00ff00ff00ff) << 8) | ((i >> 8) & 0x00ff00ff00ff00ff)) & M64
        i = ((i << 48) | ((i & 0xffff0000) << 16) | ((i >> 16) & 0xffff0000) | (i >> 48)) & M64
        return Seq1(i >> (BITS * (N - self.len)), self.len)

    def complement(self):                                  # 165-168
        # `if (length == 64) -1L` never holds (a Long1DNASeq has <= 32 bases) and the JVM takes shift counts mod 64, so a
        # 32-base sequence gets (1L << 0) - 1 = 0 and is NOT complemented there; k <= 31 is what the path uses
        mask = M64 if self.len == 64 else (1 << ((BITS * self.len) & 63)) - 1
        return Seq1(self.long ^ mask, self.len)

    def rev_complement(self):                              # DNASeq.scala:28: complement.reverse
        return self.complement().reverse()

    def hash_code(self, variant=291):                      # 103: long.##
        return long_hash(self.long, variant)

    def bases(self):
        return [self.apply(i) for i in range(self.len)]


def to_int32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >> 31 else v


def long_hash(v, variant=291):
    """`long.##` = ScalaRunTime.hash(Long).  scala-library 2.9.1 (the version the reference builds with, project/Build.scala):
    `val iv = lv.toInt; if (iv == lv) iv else lv.##` with java.lang.Long.hashCode = (int)(v ^ (v >>> 32)) behind the boxed ##
    -- for a non-negative v below 2^62 both branches are (int)(v ^ v >>> 32) unless v fits an Int, where the hash is v itself
    (and v ^ 0 is v again).  variant 210: scala >= 2.10, `val lo = lv.toInt; val hi = (lv >>> 32).toInt; lo ^ (hi + (lo >>> 31))`."""
    v &= M64
    lo, hi = v & 0xFFFFFFFF, v >> 32
    if variant == 210:
        return to_int32(lo ^ ((hi + (lo >> 31)) & 0xFFFFFFFF))
    signed = v - (1 << 64) if v >> 63 else v
    iv = to_int32(v)
    if iv == signed:
        return iv
    return to_int32(lo ^ hi)


def seq_from_codes(codes):
    """DNASeq.newBuilder (235-272) for <= 32 bases: base i at bits 2i of l1."""
    assert len(codes) <= N
    v = 0
    for i, c in enumerate(codes):
        v |= int(c) << (i * BITS)
    return Seq1(v, len(codes))


# ---------------------------------------------------------------------------------------------- PairedEndData.getPairs (20-36)
def read_pairs(bin_bytes, count):
    """`count` pairs from the .bin stream: per read 1 length byte, (len + 3) / 4 packed bytes (4 bases per byte, low bits first,
    DNASeq.apply(ar, length) 285-303).  Reads longer than 32 bases are Long2DNASeq / ArrayDNASeq in the reference; their generic
    sliding(k) yields the same windows, so a read is a plain list of base codes here."""
    b = bytes(bin_bytes)
    pos = 0

    def read():
        nonlocal pos
        ln = b[pos]
        bl = (ln + 3) // 4
        packed = b[pos + 1:pos + 1 + bl]
        pos += 1 + bl
        return [(packed[i // 4] >> (2 * (i % 4))) & 3 for i in range(ln)]

    return [(read(), read()) for _ in range(count)]


def windows(read, k):
    """seq.sliding(k) for seq.length >= k (IterableLike.sliding / Long1DNASeq.sliding): every window, as Long1DNASeq."""
    return [seq_from_codes(read[i:i + k]) for i in range(len(read) - k + 1)]


# ---------------------------------------------------------------------------------------------- FreqFilter (S/data/FreqFilter.scala:25-58)
def extract_filtered_kmers(bin_bytes, count, k, rounds, take_first=None, variant=291, filter_=True):
    """Returns the DNAMap[Int] as a dict {Seq1.key(): count}."""
    freq = {}

    def add(seq):                                          # 28-36
        if len(seq) >= k:
            for x in windows(seq, k):
                rcx = x.rev_complement()
                y = x if x.hash_code(variant) < rcx.hash_code(variant) else rcx
                freq[y.key()] = freq[y.key()] + 1 if y.key() in freq else 1     # kmersFreq.update(y, 1, _ + 1)

    pairs = read_pairs(bin_bytes, count)
    if take_first is not None:
        pairs = pairs[:take_first]                         # data.getPairs.take(max), 44
    for p1, p2 in pairs:
        add(p1)
        add(p2)
    if filter_:
        for key in [key for key, v in freq.items() if v < rounds]:             # deleteAll((k, v) => v < rounds), 55
            del freq[key]
    return freq


# ---------------------------------------------------------------------------------------------- MapGraph (Graph.scala:152-230)
class Node:
    def __init__(self, id_, seq):
        self.id, self.seq, self.in_edge_ids, self.out_edge_ids = id_, seq, [], {}   # Set[Long], Map[Base, Long]


class Edge:
    def __init__(self, id_, start_id, end_id, seq):
        self.id, self.start_id, self.end_id, self.seq = id_, start_id, end_id, seq  # seq: list of base codes


class MapGraph:
    def __init__(self):
        self.nodes, self.edges = {}, {}
        self.node_id_gen = self.edge_id_gen = 0

    def add_node(self, seq):                               # 172-176
        self.node_id_gen += 1
        n = Node(self.node_id_gen, seq)
        self.nodes[n.id] = n
        return n

    def add_edge(self, start, end, seq):                   # 178-184
        self.edge_id_gen += 1
        e = Edge(self.edge_id_gen, start.id, end.id, list(seq))
        start.out_edge_ids[seq[0]] = e.id                  # outEdgeIds += seq.head -> id (a Map: a second edge with that base replaces)
        if e.id not in end.in_edge_ids:
            end.in_edge_ids.append(e.id)
        self.edges[e.id] = e
        return e

    def remove_node(self, node):                           # 187-189
        self.nodes.pop(node.id, None)

    def remove_edge(self, edge):                           # 191-195
        s = self.nodes.get(edge.start_id)
        if s is not None:
            s.out_edge_ids.pop(edge.seq[0], None)          # outEdgeIds -= edge.seq(0): by BASE, whichever edge holds it
        t = self.nodes.get(edge.end_id)
        if t is not None and edge.id in t.in_edge_ids:
            t.in_edge_ids.remove(edge.id)
        self.edges.pop(edge.id, None)

    def in_edges(self, node):                              # Node.scala:50
        return [self.edges[i] for i in node.in_edge_ids]

    def out_edges(self, node):                             # Node.scala:52 (.values)
        return [self.edges[i] for i in node.out_edge_ids.values()]

    def retain(self, keep_ids):                            # 161-165
        keep_ids = set(keep_ids)
        self.nodes = {i: n for i, n in self.nodes.items() if i in keep_ids}
        self.edges = {i: e for i, e in self.edges.items() if e.start_id in keep_ids and e.end_id in keep_ids}

    def components(self):                                  # Graph.scala:54-72
        col, out = set(), []
        for start in list(self.nodes.values()):
            if start.id in col:
                continue
            stack, visited = [start], set()
            while stack:
                node = stack.pop()
                visited.add(node.id)
                nb = [self.nodes[e.start_id] for e in self.in_edges(node)] + [self.nodes[e.end_id] for e in self.out_edges(node)]
                nb = [x for x in nb if x.id not in col]
                col.update(x.id for x in nb)
                stack.extend(nb)
            out.append(visited)
        return out

    def simplify_graph(self, order=None):                  # 211-230; `order`: the node sweep order (the JVM's is a hash map's)
        ids = list(self.nodes) if order is None else list(order)
        for nid in ids:
            node = self.nodes.get(nid)
            if node is None:
                continue                                   # (the reference sweeps a snapshot; a removed node is never revisited either)
            in_, out = self.in_edges(node), self.out_edges(node)
            if len(in_) == 0 and len(out) == 0:
                self.remove_node(node)
            elif len(in_) == 1 and len(out) == 1:
                e1, e2 = in_[0], out[0]
                if e1.id == e2.id:
                    self.remove_edge(e1)
                else:
                    self.remove_edge(e1)
                    self.remove_edge(e2)
                    self.add_edge(self.nodes[e1.start_id], self.nodes[e2.end_id], e1.seq + e2.seq)
                self.remove_node(node)

    def remove_bubbles(self):                              # Graph.scala:121-149 (similar: 117-119)
        def similar(a, b):
            return abs(len(a) - len(b)) * 5 < max(len(a), len(b))

        for node in list(self.nodes.values()):
            out = self.out_edges(node)                     # outEdges.values.toArray: a Map[Base, Long] of <= 4 entries keeps
            to_remove = []                                 # insertion order, like the dict here
            for i in range(len(out)):
                if out[i].id in [e.id for e in to_remove]:
                    continue
                for j in range(i + 1, len(out)):
                    if out[i].end_id == out[j].end_id and similar(out[i].seq, out[j].seq):
                        if out[j].id not in [e.id for e in to_remove]:
                            to_remove.append(out[j])
            for e in to_remove:
                self.remove_edge(e)

    def canonical(self):
        """(sorted node k-mers, sorted (start k-mer, end k-mer, seq bytes)) -- the form tests/helpers.py compares."""
        nodes = sorted(n.seq.long for n in self.nodes.values())
        edges = sorted((self.nodes[e.start_id].seq.long, self.nodes[e.end_id].seq.long, bytes(e.seq)) for e in self.edges.values())
        return nodes, edges


# ---------------------------------------------------------------------------------------------- Graph.buildGraph (Graph.scala:269-382)
def build_graph(k, kmers_freq):
    """kmers_freq: dict {(long, len): count} as returned by extract_filtered_kmers."""
    def contains(x):                                       # 270
        return x.key() in kmers_freq or x.rev_complement().key() in kmers_freq

    def incoming(x):                                       # 272-276
        return [base for base in (A, G, C, T) if contains(x.take(k - 1).prepend(base))]

    def outcoming(x):                                      # 278-282
        return [base for base in (A, G, C, T) if contains(x.drop(1).append(base))]

    term = set()                                           # 319-331
    for (long_, len_) in kmers_freq:
        read = Seq1(long_, len_)
        in_, out = len(incoming(read)), len(outcoming(read))
        if (in_ != 1 or out != 1) and (in_ != 0 or out != 0):
            term.add(read)
    term_kmers = set(term) | {x.rev_complement() for x in term}

    graph = MapGraph()
    node_map = {read: graph.add_node(read) for read in sorted(term_kmers, key=Seq1.key)}   # 343-347

    def build_edges(read):                                 # 349-365
        node = node_map[read]
        for base in outcoming(read):
            builder = [base]
            seq = read.drop(1).append(base)
            while seq not in node_map:
                out = outcoming(seq)
                assert len(out) == 1, (seq.key(), out)
                builder.append(out[0])
                seq = seq.drop(1).append(out[0])
            graph.add_edge(node, node_map[seq], builder)

    for read in sorted(term_kmers, key=Seq1.key):          # 367-374
        build_edges(read)
    return graph                                           # perfect cycles are ignored (375)


# ---------------------------------------------------------------------------------------------- Graph.getGraphMap (Graph.scala:90-119)
def graph_positions(graph):
    """[(k-mer long, ('node', id) | ('edge', id, dist))] in the reference's order of putNew calls (nodes, then edges)."""
    out = [(n.seq.long, ("node", n.id)) for n in graph.nodes.values()]
    for e in graph.edges.values():
        seq = graph.nodes[e.start_id].seq.drop(1).append(e.seq[0])
        dist = 1
        for base in e.seq[1:]:
            out.append((seq.long, ("edge", e.id, dist)))
            seq = seq.drop(1).append(base)
            dist += 1
    return out


# ---------------------------------------------------------------------------------------------- MapGraph.replaceStart / replaceEnd (197-209)
def replace_start(graph, edge_id, new_start):
    edge = graph.edges[edge_id]
    graph.edges[edge.id] = Edge(edge.id, new_start.id, edge.end_id, edge.seq)
    graph.nodes[edge.start_id].out_edge_ids.pop(edge.seq[0], None)
    new_start.out_edge_ids[edge.seq[0]] = edge.id


def replace_end(graph, edge_id, new_end):
    edge = graph.edges[edge_id]
    graph.edges[edge.id] = Edge(edge.id, edge.start_id, new_end.id, edge.seq)
    old = graph.nodes[edge.end_id]
    if edge.id in old.in_edge_ids:
        old.in_edge_ids.remove(edge.id)
    if edge.id not in new_end.in_edge_ids:
        new_end.in_edge_ids.append(edge.id)


# ---------------------------------------------------------------------------------------------- WalkingActor (S/scripts/GraphSimplifier.scala:33-127)
class WalkingActor:
    """Positions are ('node', id) or ('edge', id, dist) like NodeGraphPosition / EdgeGraphPosition."""

    def __init__(self, graph, range_first, range_last):
        self.graph, self.first, self.last = graph, range_first, range_last
        self.cache = {}

    def in_range(self, v):
        return self.first <= v <= self.last

    def reachable(self, node_id):                          # 42-72: Dijkstra backwards over in-edges, distances <= range.last
        if node_id in self.cache:
            return self.cache[node_id]
        import heapq
        g = self.graph
        queue, seen = [(0, node_id)], {}
        while queue:
            dist, u = heapq.heappop(queue)                 # the ordering y._1 - x._1 makes the smallest distance the head
            if u not in seen:
                seen[u] = dist
                for edge in g.in_edges(g.nodes[u]):
                    d2 = dist + len(edge.seq)
                    if d2 <= self.last:
                        heapq.heappush(queue, (d2, edge.start_id))
        if len(self.cache) < 50000:
            self.cache[node_id] = seen
        return seen

    def receive(self, pos1, pos2):                         # 77-126: (good, pathEdges)
        g = self.graph
        path_edges = set()
        if pos2[0] == "node":
            node2, dist2 = pos2[1], 0
        else:
            node2, dist2 = g.edges[pos2[1]].start_id, pos2[2]
        start_edge = g.edges[pos1[1]] if pos1[0] == "edge" else None
        end_edge = g.edges[pos2[1]] if pos2[0] == "edge" else None
        reach = self.reachable(node2)
        memo = {}

        def dfs(node1, dist1, prev_edge):
            key = (prev_edge.id if prev_edge is not None else None, dist1)
            if key in memo:
                return memo[key]
            if dist1 + dist2 + reach.get(node1, self.last + 1) > self.last:
                return False
            if node1 == node2 and self.in_range(dist1 + dist2):
                if prev_edge is not None and end_edge is not None:
                    path_edges.add((prev_edge.id, end_edge.id))
                cur = True
            else:
                cur = False
            for e in g.out_edges(g.nodes[node1]):
                res = dfs(e.end_id, dist1 + len(e.seq), e)
                if res and prev_edge is not None:
                    path_edges.add((prev_edge.id, e.id))
                cur |= res
            memo[key] = cur
            return cur

        if pos1[0] == "node":
            node0, dist0 = pos1[1], 0
        else:
            node0, dist0 = g.edges[pos1[1]].end_id, len(g.edges[pos1[1]].seq) - pos1[2]
        import sys
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(20000)
        try:
            good = dfs(node0, dist0, start_edge)
        finally:
            sys.setrecursionlimit(old)
        return good, path_edges


# ---------------------------------------------------------------------------------------------- GraphSimplifier.startup (138-357)
def pair_support(graph, bin_bytes, count, k, range_first, range_last, take_first=None):
    """The pair loop (188-263): returns (pathsMap {(e1 id, e2 id): count}, badPairs, walked orientation cases)."""
    gmap = {}
    for kmer, pos in graph_positions(graph):               # graphMap = graph.getGraphMap: putNew keeps every position of a k-mer
        gmap.setdefault(kmer, []).append(pos)
    actor = WalkingActor(graph, range_first, range_last)

    def annotate(positions1, positions2):                  # 192-206
        for p in positions1:
            if p[0] == "edge":
                for q in positions2:
                    if q[0] == "edge" and p[1] == q[1] and range_first <= (q[2] - p[2]) + k <= range_last:
                        return None
        return positions1, positions2

    paths, bad, walked = {}, 0, 0
    pairs = read_pairs(bin_bytes, count)
    if take_first is not None:
        pairs = pairs[:take_first]
    for p1, p2 in pairs:
        if len(p1) < k or len(p2) < k:
            continue
        a, b = seq_from_codes(p1[:k]), seq_from_codes(p2[:k])
        f1, f2 = gmap.get(a.long, []), gmap.get(b.rev_complement().long, [])
        f3, f4 = gmap.get(b.long, []), gmap.get(a.rev_complement().long, [])
        for l1, l2 in ((f1, f2), (f3, f4)):
            if annotate(l1, l2) is None:
                continue
            todo = [(x, y) for x in l1 for y in l2]
            if not todo:
                continue                                   # `for (list <- Future.sequence(futures) if !list.isEmpty)`
            walked += 1
            results = [actor.receive(x, y) for x, y in todo]
            good = any(r[0] for r in results)
            path_edges = set().union(*[r[1] for r in results])
            for es in path_edges:
                paths[es] = paths.get(es, 0) + 1
            if not good:
                bad += 1
    return paths, bad, walked


def split_nodes(graph, paths, cutoff):
    """The node sweep (266-317) without the simplifyGraph that follows: returns (edges removed, nodes added)."""
    to_remove, added = set(), 0
    for node in list(graph.nodes.values()):
        in_, out = list(node.in_edge_ids), list(node.out_edge_ids.values())
        if not in_ or not out:
            continue
        matrix = [[paths.get((i, j), 0) for j in out] for i in in_]
        col_left, col_right = [False] * len(in_), [False] * len(out)

        def dfs_left(i):
            assert not col_left[i]
            col_left[i] = True
            l, r = {i}, set()
            for j in range(len(out)):
                if not col_right[j] and matrix[i][j] >= cutoff:
                    l1, r1 = dfs_right(j)
                    l |= l1
                    r |= r1
            return l, r

        def dfs_right(j):
            assert not col_right[j]
            col_right[j] = True
            l, r = set(), {j}
            for i in range(len(in_)):
                if not col_left[i] and matrix[i][j] >= cutoff:
                    l1, r1 = dfs_left(i)
                    l |= l1
                    r |= r1
            return l, r

        for i in range(len(in_)):
            if col_left[i]:
                continue
            l, r = dfs_left(i)
            if not r:
                to_remove.add(in_[i])
            else:
                new_node = graph.add_node(node.seq)
                added += 1
                for x in l:
                    replace_end(graph, in_[x], new_node)
                for x in r:
                    replace_start(graph, out[x], new_node)
        to_remove |= {out[j] for j in range(len(out)) if not col_right[j]}
    for eid in to_remove:                                  # 320: toRemove.foreach(id => graph.removeEdge(graph.getEdge(id)))
        graph.remove_edge(graph.edges[eid])
    return len(to_remove), added
