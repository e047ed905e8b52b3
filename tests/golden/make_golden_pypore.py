#!/usr/bin/env python
"""Golden vectors for the models PyPore/hmm.py builds and for the model text format (8f-3).

Run in the build container only (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden_pypore.py

The reference's builders (ModularProfileModel over Phi29GlobalAlignmentModule / NanoporeGlobalAlignmentModule /
GlobalAlignmentModule incl. a fork, Hel308ProfileHMM, Phi29ProfileHMM; /root/reference/PyPore/hmm.py:38-651) are
executed TWICE from the reference's own source file: once on the compiled, unmodified reference yahmm (oracle/_ref) to
get the results, once on the drop-in yahmm of this repository to get a REPLAY LOG of the graph they construct -- nodes
and edges in insertion order, which is what defines state order and tie order.  The GPU box has no /root/reference: the
parity test rebuilds each model from the log (tests/util.py: replay_model), bakes it, and compares decoding on the GPU
with what the compiled reference returned here.  Nothing of hmm.py is copied into the repository.

pypore_<name>.npz : replay log (JSON), the reference's baked state names, dense transition matrix, and for sampled
                    events its log_probability, viterbi score + path and forward_backward outputs
model_write.txt   : Model.write() of the reference for the pro_001 model (identities replaced by first-appearance
                    ordinals: the reference's are id()-based), for the byte comparison with our Model.write
"""
import io
import itertools
import json
import os
import random
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_loader  # noqa: E402
import protpore_b200.yahmm as ours  # noqa: E402
from protpore_b200 import peptide_hmm  # noqa: E402
import util  # noqa: E402

ref = ref_loader.load_reference_yahmm()
HERE = os.path.dirname(os.path.abspath(__file__))
HMM_PY = "/root/reference/PyPore/hmm.py"


def load_builders(y):
    """Execute the reference's PyPore/hmm.py against yahmm module `y` (Python-2 names shimmed)."""
    class It(object):
        izip = staticmethod(zip)
        def __getattr__(self, k):
            return getattr(itertools, k)
    ns = {k: getattr(y, k) for k in dir(y) if not k.startswith("_")}
    ns.update({"xrange": range, "it": It(), "np": np, "__name__": "pypore_hmm"})
    src = open(HMM_PY).read().replace("from yahmm import *", "")
    exec(compile(src, HMM_PY, "exec"), ns)
    return ns


def dist_spec(d):
    name = type(d).__name__
    p = list(d.parameters)
    if name in ("GaussianKernelDensity", "UniformKernelDensity", "TriangleKernelDensity"):
        return {"type": name, "points": [float(v) for v in p[0]], "bandwidth": float(p[1]), "weights": [float(v) for v in p[2]]}
    if name == "MixtureDistribution":
        return {"type": name, "parts": [dist_spec(x) for x in p[0]], "weights": [float(v) for v in p[1]]}
    return {"type": name, "parameters": [float(v) for v in p]}


def record(model_graph, start, end, name):
    """Replay log of a pre-bake graph of OUR yahmm: nodes / edges in insertion order, shared distribution objects kept."""
    nodes = model_graph.nodes()
    index = {s: i for i, s in enumerate(nodes)}
    dists, dist_index = [], {}
    out_nodes = []
    for s in nodes:
        di = None
        if s.distribution is not None:
            if id(s.distribution) not in dist_index:
                dist_index[id(s.distribution)] = len(dists)
                dists.append(dist_spec(s.distribution))
            di = dist_index[id(s.distribution)]
        out_nodes.append({"name": s.name, "weight": float(s.weight), "dist": di})
    edges = [[index[a], index[b], float(d["weight"]), float(d["pseudocount"]), d.get("group")] for a, b, d in model_graph.edges(data=True)]
    return {"name": name, "start": index[start], "end": index[end], "nodes": out_nodes, "dists": dists, "edges": edges}


def distributions_for(y, kind, n, seed):
    rs = np.random.RandomState(seed)
    means = np.sort(rs.uniform(20, 70, size=n))
    if kind == "normal":
        return [y.NormalDistribution(float(m), float(s)) for m, s in zip(means, rs.uniform(0.8, 2.5, size=n))]
    # kernel densities, as PyPore/alignment.py builds its profiles (alignment.py:444)
    return [y.GaussianKernelDensity([float(v) for v in rs.normal(m, 1.2, size=5)], 1.0) for m in means]


def cases():
    # name -> callable(ns, y) -> baked model
    def phi29(ns, y):
        return ns["ModularProfileModel"](ns["Phi29GlobalAlignmentModule"], distributions_for(y, "normal", 6, 1), "phi29mod",
                                         y.UniformDistribution(0, 90))
    def nanopore_fork(ns, y):
        d = distributions_for(y, "normal", 7, 2)
        seq = [d[0], d[1], {"a": d[2], "b": d[3]}, {"a": d[4], "b": d[5]}, d[6]]
        return ns["ModularProfileModel"](ns["NanoporeGlobalAlignmentModule"], seq, "npfork", y.UniformDistribution(0, 90))
    def global_kde(ns, y):
        return ns["ModularProfileModel"](ns["GlobalAlignmentModule"], distributions_for(y, "kde", 5, 3), "gkde",
                                         y.UniformDistribution(0, 90))
    def hel308(ns, y):
        return ns["Hel308ProfileHMM"](distributions_for(y, "normal", 6, 4), name="hel308")
    def phi29_profile(ns, y):
        return ns["Phi29ProfileHMM"](distributions_for(y, "normal", 5, 5), name="phi29prof")
    return {"phi29_module": phi29, "nanopore_fork": nanopore_fork, "global_kde": global_kde, "hel308": hel308,
            "phi29_profile": phi29_profile}


def main():
    ns_ref, ns_ours = load_builders(ref), load_builders(ours)
    for name, build in cases().items():
        # ---- the reference's model and results
        with redirect_stdout(io.StringIO()):
            m_ref = build(ns_ref, ref)
        # ---- our graph, recorded right before bake
        log = {}
        real_bake = ours.Model.bake
        def spy(self, *a, **k):
            if not log and self.name == m_ref.name:
                log.update(record(self.graph, self.start, self.end, self.name))
            return real_bake(self, *a, **k)
        ours.Model.bake = spy
        try:
            m_ours = build(ns_ours, ours)
        finally:
            ours.Model.bake = real_bake
        assert [s.name for s in m_ours.states] == [s.name for s in m_ref.states], name
        assert np.array_equal(np.asarray(m_ours.dense_transition_matrix()), np.asarray(m_ref.dense_transition_matrix())), name
        random.seed(7)
        evs = []
        while len(evs) < 40:
            x = np.asarray(m_ref.sample(), dtype=np.float32).astype(np.float64)
            if 1 <= len(x) <= 400:
                evs.append(x)
        obs = np.concatenate(evs)
        offsets = np.concatenate([[0], np.cumsum([len(e) for e in evs])]).astype(np.int64)
        logp = np.array([m_ref.log_probability(e) for e in evs])
        vscore, paths, poff = [], [], [0]
        for e in evs:
            s, p = m_ref.viterbi(e)
            vscore.append(s)
            idx = [] if p is None else [i for i, _ in p]
            paths.extend(idx)
            poff.append(len(paths))
        fb_t, fb_w = [], []
        for e in evs[:6]:
            with redirect_stdout(io.StringIO()):
                t, w = m_ref.forward_backward(e)
            fb_t.append(np.asarray(t))
            fb_w.append(np.asarray(w))
        np.savez_compressed(os.path.join(HERE, "pypore_%s.npz" % name), replay=json.dumps(log),
                            state_names=np.array([s.name for s in m_ref.states]), silent_start=m_ref.silent_start,
                            dense=np.asarray(m_ref.dense_transition_matrix()), obs=obs, offsets=offsets, logp=logp,
                            vscore=np.array(vscore), paths=np.array(paths, dtype=np.int32), path_offsets=np.array(poff, dtype=np.int64),
                            fb_transitions=np.stack(fb_t), fb_weights=np.concatenate([w.ravel() for w in fb_w]),
                            fb_weight_rows=np.array([w.shape[0] for w in fb_w]))
        print(name, len(m_ref.states), "states,", m_ref.edge_count(), "edges,", len(obs), "observations")

    # ---- Model.write of the reference for the pro_001 profile
    m = peptide_hmm.pro_001_model(yahmm_module=ref, py2_division=True)
    buf = io.StringIO()
    with redirect_stdout(io.StringIO()):
        m.write(buf)
    open(os.path.join(HERE, "model_write.txt"), "w").write(util.normalise_identities(buf.getvalue()))
    print("model_write.txt", len(buf.getvalue()), "bytes")


if __name__ == "__main__":
    main()
