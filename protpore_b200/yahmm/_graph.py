"""Insertion-ordered directed graph used by `Model` before `bake`.

The reference stores the pre-bake topology in a networkx-1.11 `DiGraph` (yahmm.pyx:2057) whose node
and adjacency orders are whatever Python-2 dict hashing of `State` objects gives, i.e. they depend
on memory addresses.  Those orders are observable: they fix the baked state order (2441-2469) and
the in-edge CSR order that breaks Viterbi ties (2575-2592).  This class *defines* both as insertion
order (DESIGN.md "documented exact ties").  Only the graph operations `Model` needs are provided,
under the names the reference calls them by so user code that pokes `model.graph` keeps working.
"""


class OrderedDiGraph(object):
    __slots__ = ("node", "edge", "pred")

    def __init__(self):
        self.node = {}   # node -> attr dict (insertion ordered)
        self.edge = {}   # edge[a][b] -> attr dict ('weight', 'pseudocount', 'group')
        self.pred = {}   # pred[b][a] -> the same attr dict object

    # ---- construction -------------------------------------------------
    def add_node(self, n):
        if n not in self.node:
            self.node[n] = {}
            self.edge[n] = {}
            self.pred[n] = {}

    def add_edge(self, a, b, **attr):
        self.add_node(a)
        self.add_node(b)
        cur = self.edge[a].get(b)
        if cur is None:
            cur = {}
            self.edge[a][b] = cur
            self.pred[b][a] = cur
        cur.update(attr)

    def remove_edge(self, a, b):
        del self.edge[a][b]
        del self.pred[b][a]

    def remove_node(self, n):
        for b in list(self.edge[n]):
            del self.pred[b][n]
        for a in list(self.pred[n]):
            del self.edge[a][n]
        del self.edge[n], self.pred[n], self.node[n]

    # ---- queries (list snapshots, as networkx 1.11 returned) -----------
    def nodes(self):
        return list(self.node)

    def edges(self, data=False):
        out = []
        for a, nbrs in self.edge.items():
            for b, d in nbrs.items():
                out.append((a, b, d) if data else (a, b))
        return out

    def edges_iter(self, data=False):
        return iter(self.edges(data=data))

    def number_of_edges(self):
        return sum(len(nbrs) for nbrs in self.edge.values())

    def __contains__(self, n):
        return n in self.node

    def __len__(self):
        return len(self.node)

    def __getitem__(self, n):
        return self.edge[n]

    def subgraph(self, nbunch):
        keep = set(nbunch)
        h = OrderedDiGraph()
        for n in self.node:
            if n in keep:
                h.add_node(n)
        for a, nbrs in self.edge.items():
            if a not in keep:
                continue
            for b, d in nbrs.items():
                if b in keep:
                    h.edge[a][b] = d
                    h.pred[b][a] = d
        return h


def union(g, h):
    """Disjoint union keeping g's nodes first (Model.add_model / concatenate_model, yahmm.pyx:2225, 2239)."""
    r = OrderedDiGraph()
    for x in (g, h):
        for n in x.node:
            r.add_node(n)
        for a, b, d in x.edges(data=True):
            r.add_edge(a, b, **d)
    return r


def topological_order(g):
    """Topological order of a DAG exactly as networkx 1.11 `topological_sort` produces it.

    Third-party algorithm restated (networkx 1.11, algorithms/dag.py): iterative depth-first search
    seeded in node order, successors pushed in adjacency order, reverse post-order returned.  The
    reference calls it on the silent-state subgraph (yahmm.pyx:2457-2461) and the result fixes the
    baked index of every silent state, so the traversal order matters, not just validity.
    """
    seen, explored, post = set(), set(), []
    for root in g.node:
        if root in explored:
            continue
        stack = [root]
        while stack:
            w = stack[-1]
            if w in explored:
                stack.pop()
                continue
            seen.add(w)
            fresh = []
            for nxt in g.edge[w]:
                if nxt not in explored:
                    if nxt in seen:
                        raise RuntimeError("Graph contains a cycle.")
                    fresh.append(nxt)
            if fresh:
                stack.extend(fresh)
            else:
                explored.add(w)
                post.append(w)
                stack.pop()
    post.reverse()
    return post
