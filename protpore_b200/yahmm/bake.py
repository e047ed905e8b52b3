"""`Model.bake`: lower the state graph to CSR arrays (restates yahmm.pyx:2280-2655).

Output (`Baked`): state order (emitting in graph order, then silent in topological order), in-edge
and out-edge CSR (`*_ptr[m+1]`, `*_idx[E]`, `*_logp[E]`, `*_pseudo[E]`), tied-state CSR, tied edge
groups, `finite`, `start_index`, `end_index`, `silent_start`.  All float work uses the same libm
calls in the same order as the reference so the arrays are bit-identical to the compiled reference
on the same insertion-ordered graph (tests/test_bake.py).

Deliberately reproduced quirks (each verified against the compiled reference):
  * the orphan-removal counters are allocated once and never cleared between sweeps (2319-2360);
  * out-edge normalisation divides by `round(sum(e**w), 8)` and only when that differs from 1 (2364-2380);
  * a silent state is merged away when ANY of its out-edges has weight exactly 0.0 (2385-2431);
  * in-edge CSR slots are filled source-major, so the in-edge order of a state is the pre-bake node
    order of its sources, not their baked index (2575-2592) -- this is what breaks Viterbi ties.
"""
from __future__ import print_function

import math

import numpy

from ._graph import topological_order

NEGINF = float("-inf")


def _plog(value):
    """Module-level `log` of the reference for a scalar (yahmm.pyx:73-84)."""
    return math.log(value) if value > 0 else NEGINF


class Baked(object):
    """Plain container for the baked arrays (names follow the reference's private memoryviews, 2029-2040)."""

    __slots__ = ("states", "silent_start", "start_index", "end_index", "finite",
                 "in_ptr", "in_idx", "in_logp", "in_pseudo",
                 "out_ptr", "out_idx", "out_logp", "out_pseudo",
                 "state_weights", "tied_ptr", "tied_idx",
                 "tied_edge_group_size", "tied_edges_starts", "tied_edges_ends")


def _remove_orphans(model, verbose):
    g = model.graph
    n0 = len(g)
    in_count = numpy.zeros(n0, dtype=numpy.int32)
    out_count = numpy.zeros(n0, dtype=numpy.int32)
    while True:
        removed = 0
        pre = g.nodes()
        index = {s: i for i, s in enumerate(pre)}
        for a, b in g.edges():
            out_count[index[a]] += 1
            in_count[index[b]] += 1
        for i, s in enumerate(pre):
            if s is model.start or s is model.end:
                continue
            if in_count[i] == 0:
                removed += 1
                g.remove_node(s)
                if verbose:
                    print("Orphan state {} removed due to no edges \
							leading to it".format(s.name))
            elif out_count[i] == 0:
                removed += 1
                g.remove_node(s)
                if verbose:
                    print("Orphan state {} removed due to no edges \
							leaving it".format(s.name))
        if removed == 0:
            break


def _normalise_out_edges(model, verbose):
    g = model.graph
    for s in g.nodes():
        total = round(sum(math.e ** d['weight'] for d in g.edge[s].values()), 8)
        if total != 1. and s is not model.end:
            if verbose:
                print("{} : {} summed to {}, normalized to 1.0".format(model.name, s.name, total))
            shift = _plog(total)
            for d in g.edge[s].values():
                d['weight'] = d['weight'] - shift


def _merge_silent(model, merge, verbose):
    g = model.graph
    while True:
        merged = 0
        for a, b, e in g.edges(data=True):
            if a not in g or b not in g:
                continue
            if a is model.start or b is model.end:
                continue
            if not (e['weight'] == 0.0 and a.is_silent()):
                continue
            if not (merge == 'all' or (merge == 'partial' and b.is_silent())):
                continue
            # redirect every edge x -> a to x -> b, visiting x in node order (what a scan of
            # graph.edges() in the reference amounts to)
            for x in [x for x in g.node if a in g.edge[x]]:
                d = g.edge[x][a]
                merged += 1
                g.remove_edge(x, a)
                pseudo = max(e['pseudocount'], d['pseudocount'])
                group = e['group'] if e['group'] == d['group'] else None
                g.add_edge(x, b, weight=d['weight'], pseudocount=pseudo, group=group)
                if verbose:
                    print("{} : {} - {} merged".format(model.name, a, b))
            g.remove_node(a)
        if merged == 0:
            break


def _report_silent_two_cycles(model):
    g = model.graph
    for a, b in g.edges():
        if a.is_silent() and b.is_silent() and a in g.edge[b]:
            print("Loop: {} - {}".format(a.name, b.name))


def bake_graph(model, verbose=False, merge="all"):
    """Mutates `model.graph` exactly as the reference does and returns the `Baked` arrays."""
    g = model.graph
    merge = merge.lower() if merge else None

    if merge == 'all':
        _remove_orphans(model, verbose)
    _normalise_out_edges(model, verbose)
    if merge in ('all', 'partial'):
        _merge_silent(model, merge, verbose)
    _report_silent_two_cycles(model)

    nodes = g.nodes()
    silent = [s for s in nodes if s.is_silent()]
    emitting = [s for s in nodes if not s.is_silent()]
    silent_sorted = topological_order(g.subgraph(silent))

    out = Baked()
    out.silent_start = len(emitting)
    out.states = emitting + silent_sorted
    m = len(out.states)
    index = {s: i for i, s in enumerate(out.states)}

    # tied states: for state i, every j != i (ascending) sharing the same distribution object (2479-2512)
    by_dist = {}
    for i in range(out.silent_start):
        by_dist.setdefault(id(out.states[i].distribution), []).append(i)
    tied_ptr = numpy.zeros(out.silent_start + 1, dtype=numpy.int32)
    tied_idx = []
    for i in range(out.silent_start):
        mates = [j for j in by_dist[id(out.states[i].distribution)] if j != i]
        tied_idx.extend(mates)
        tied_ptr[i + 1] = tied_ptr[i] + len(mates)
    out.tied_ptr = tied_ptr
    out.tied_idx = numpy.asarray(tied_idx, dtype=numpy.int32)

    out.state_weights = numpy.array([math.log(out.states[i].weight) for i in range(out.silent_start)],
                                    dtype=numpy.float64)

    edges = g.edges(data=True)          # source-major, adjacency order
    n_edges = len(edges)
    in_count = numpy.zeros(m + 1, dtype=numpy.int32)
    out_count = numpy.zeros(m + 1, dtype=numpy.int32)
    for a, b, _ in edges:
        in_count[index[b] + 1] += 1
        out_count[index[a] + 1] += 1
    if model.end not in index:
        raise SyntaxError("Model.end has been deleted, leaving the \
				model with no end. Please ensure it has an end.")
    out.finite = 0 if in_count[index[model.end] + 1] == 0 else 1
    out.in_ptr = numpy.cumsum(in_count, dtype=numpy.int32)
    out.out_ptr = numpy.cumsum(out_count, dtype=numpy.int32)

    out.in_idx = numpy.full(n_edges, -1, dtype=numpy.int32)
    out.out_idx = numpy.full(n_edges, -1, dtype=numpy.int32)
    out.in_logp = numpy.zeros(n_edges)
    out.out_logp = numpy.zeros(n_edges)
    out.in_pseudo = numpy.zeros(n_edges)
    out.out_pseudo = numpy.zeros(n_edges)
    in_fill = out.in_ptr[:-1].copy()
    out_fill = out.out_ptr[:-1].copy()
    groups = {}
    for a, b, d in edges:
        ia, ib = index[a], index[b]
        k = in_fill[ib]
        in_fill[ib] += 1
        out.in_idx[k], out.in_logp[k], out.in_pseudo[k] = ia, d['weight'], d['pseudocount']
        k = out_fill[ia]
        out_fill[ia] += 1
        out.out_idx[k], out.out_logp[k], out.out_pseudo[k] = ib, d['weight'], d['pseudocount']
        if d['group'] is not None:
            groups.setdefault(d['group'], []).append((ia, ib))

    sizes = numpy.zeros(len(groups) + 1, dtype=numpy.int32)
    starts, ends = [], []
    for i, members in enumerate(groups.values()):
        sizes[i + 1] = sizes[i] + len(members)
        for s, e in members:
            starts.append(s)
            ends.append(e)
    out.tied_edge_group_size = sizes
    out.tied_edges_starts = numpy.asarray(starts, dtype=numpy.int32)
    out.tied_edges_ends = numpy.asarray(ends, dtype=numpy.int32)

    if model.start not in index:
        raise SyntaxError("Model.start has been deleted, leaving the \
				model with no start. Please ensure it has a start.")
    out.start_index = index[model.start]
    out.end_index = index[model.end]
    return out
