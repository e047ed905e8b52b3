"""networkx-1.11 API surface that the reference's yahmm.pyx touches, with *insertion-ordered* dicts.

TEST INFRASTRUCTURE ONLY.  The reference README pins networkx 1.11 (README.md:5); networkx 3.x dropped
`G.edge`, `edges_iter`, list-returning `nodes()`/`edges()` and list-returning `topological_sort`,
all of which `Model.bake` relies on (yahmm.pyx:2319-2469, 2549, 2575).  Under Python 2 the node
and adjacency dict orders are hash(id)-dependent, i.e. run-dependent; this shim *defines* them as
insertion order, which is the order our product bake uses too (DESIGN.md "tie order").

Call sites served: DiGraph() 2057; add_node 2066/2125; add_edge(**attr) 2156/2418; nodes() 2319/
2329/2364/2392/2441; edges()/edges(data=True) list snapshots 2333/2389/2405/2436/2442; G.edge[s]
2368/2379; remove_node 2345/2353/2428; remove_edge 2413; subgraph 2457; edges_iter 2549/2575;
topological_sort 2461; union 2225/2239; draw 2258.
"""


class DiGraph(object):
    def __init__(self):
        self.node = {}
        self.edge = {}   # succ: edge[a][b] -> attr dict
        self.pred = {}   # pred[b][a] -> same attr dict

    # -- mutation -------------------------------------------------------
    def add_node(self, n, **attr):
        if n not in self.node:
            self.node[n] = {}
            self.edge[n] = {}
            self.pred[n] = {}
        self.node[n].update(attr)

    def add_edge(self, a, b, **attr):
        self.add_node(a)
        self.add_node(b)
        d = self.edge[a].get(b, {})
        d.update(attr)
        self.edge[a][b] = d
        self.pred[b][a] = d

    def remove_edge(self, a, b):
        del self.edge[a][b]
        del self.pred[b][a]

    def remove_node(self, n):
        for b in list(self.edge[n]):
            del self.pred[b][n]
        for a in list(self.pred[n]):
            del self.edge[a][n]
        del self.edge[n]
        del self.pred[n]
        del self.node[n]

    # -- views (1.11 returned lists; bake mutates while iterating them) --
    def nodes(self):
        return list(self.node)

    def nodes_iter(self):
        return iter(list(self.node))

    def edges_iter(self, data=False):
        for a, nbrs in list(self.edge.items()):
            for b, d in list(nbrs.items()):
                yield (a, b, d) if data else (a, b)

    def edges(self, data=False):
        return list(self.edges_iter(data=data))

    def __getitem__(self, n):
        return self.edge[n]

    def __contains__(self, n):
        return n in self.node

    def __len__(self):
        return len(self.node)

    def __iter__(self):
        return iter(list(self.node))

    def subgraph(self, nbunch):
        keep = set(nbunch)
        h = DiGraph()
        for n in self.node:
            if n in keep:
                h.add_node(n)
        for a, b, d in self.edges_iter(data=True):
            if a in keep and b in keep:
                h.add_edge(a, b, **d)
        return h


def union(g, h):
    r = DiGraph()
    for x in (g, h):
        for n in x.nodes():
            r.add_node(n)
        for a, b, d in x.edges_iter(data=True):
            r.add_edge(a, b, **d)
    return r


def topological_sort(g, nbunch=None, reverse=False):
    """networkx 1.11 algorithm: iterative DFS, reverse post-order, seeded in nodes() order."""
    seen, order, explored = set(), [], set()
    for v in (g.nodes_iter() if nbunch is None else nbunch):
        if v in explored:
            continue
        fringe = [v]
        while fringe:
            w = fringe[-1]
            if w in explored:
                fringe.pop()
                continue
            seen.add(w)
            new_nodes = []
            for n in g[w]:
                if n not in explored:
                    if n in seen:
                        raise RuntimeError("Graph contains a cycle.")
                    new_nodes.append(n)
            if new_nodes:
                fringe.extend(new_nodes)
            else:
                explored.add(w)
                order.append(w)
                fringe.pop()
    return order if reverse else list(reversed(order))


def draw(*args, **kwargs):
    pass
