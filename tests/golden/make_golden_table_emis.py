#!/usr/bin/env python
"""Generate tests/golden/table_emis.npz from the COMPILED, UNMODIFIED reference yahmm (oracle/_ref): a model whose
states use the families the kernels read from the emission table next to a native Normal state.

    python tests/golden/make_golden_table_emis.py

Stored: per-distribution log_probability on a grid, and for seeded events the reference's log_probability, Viterbi
score / path, forward / backward tables of the first event and forward_backward outputs of the first two.
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def bump(x):
    return -0.5 * (x - 4.0) ** 2 - 1.0


def build(y):
    """The same construction sequence for the reference module and for protpore_b200.yahmm."""
    m = y.Model(name="table_emis")
    dists = [y.GaussianKernelDensity([1.0, 1.5, 2.2, 3.0], bandwidth=0.5, weights=[1, 2, 1, 1]),
             y.LogNormalDistribution(0.5, 0.4),
             y.ExponentialDistribution(0.7),
             y.LambdaDistribution(bump),
             y.TriangleKernelDensity([2.0, 5.0], bandwidth=2.0),
             y.NormalDistribution(3.0, 1.0),
             y.UniformKernelDensity([1.0, 4.0], bandwidth=1.5)]
    st = [y.State(d, name="e%d" % i) for i, d in enumerate(dists)]
    sil = y.State(None, name="sil")
    for s in st + [sil]:
        m.add_state(s)
    n = len(st)
    m.add_transition(m.start, st[0], 0.4)
    m.add_transition(m.start, sil, 0.6)
    m.add_transition(sil, st[1], 0.5)
    m.add_transition(sil, st[5], 0.5)
    for i in range(n):
        m.add_transition(st[i], st[i], 0.2)
        m.add_transition(st[i], st[(i + 1) % n], 0.3)
        m.add_transition(st[i], st[(i + 3) % n], 0.2)
        m.add_transition(st[i], sil, 0.1)
        m.add_transition(st[i], m.end, 0.2)
    m.bake()
    return m, dists


def build_train(y):
    """Trainable families on the table path (Exponential, LogNormal) next to a native Normal state."""
    m = y.Model(name="table_train")
    st = [y.State(y.ExponentialDistribution(0.7), name="x0"), y.State(y.LogNormalDistribution(0.5, 0.4), name="x1"),
          y.State(y.NormalDistribution(3.0, 1.0), name="x2")]
    for s in st:
        m.add_state(s)
    for i in range(3):
        m.add_transition(m.start, st[i], 1 / 3.)
        for j in range(3):
            m.add_transition(st[i], st[j], 0.3)
        m.add_transition(st[i], m.end, 0.1)
    m.bake()
    return m


def main():
    ref = ref_loader.load_reference_yahmm()
    with redirect_stdout(io.StringIO()):
        model, dists = build(ref)
    rng = np.random.default_rng(77)
    grid = np.round(rng.uniform(0.05, 7.0, size=64), 3)
    blob = {"grid": grid}
    for i, d in enumerate(dists):
        blob["dist%d" % i] = np.array([d.log_probability(float(x)) for x in grid])
    events = [np.round(rng.uniform(0.2, 6.0, size=int(n)), 3) for n in (5, 12, 1, 30, 60, 8)]
    blob["obs"] = np.concatenate(events)
    blob["offsets"] = np.concatenate([[0], np.cumsum([len(e) for e in events])]).astype(np.int64)
    blob["logp"] = np.array([model.log_probability(e) for e in events])
    vs, paths = [], []
    for e in events:
        s, p = model.viterbi(e)
        vs.append(s)
        paths.append(np.array([k for k, _ in p], dtype=np.int64) if p is not None else np.zeros(0, dtype=np.int64))
    blob["vscore"] = np.array(vs)
    blob["paths"] = np.concatenate(paths)
    blob["path_offsets"] = np.concatenate([[0], np.cumsum([len(p) for p in paths])]).astype(np.int64)
    blob["forward0"] = np.asarray(model.forward(events[0]))
    blob["backward0"] = np.asarray(model.backward(events[0]))
    for j in (0, 1):
        with redirect_stdout(io.StringIO()):
            t, w = model.forward_backward(events[j])
        blob["fb_trans%d" % j] = np.asarray(t)
        blob["fb_emis%d" % j] = np.asarray(w)
    blob["state_names"] = np.array([s.name for s in model.states])
    # one Baum-Welch iteration (yahmm.pyx:4075-4327) with table-path families
    tr_events = [np.round(rng.uniform(0.2, 6.0, size=int(n)), 3) for n in (9, 14, 6, 25)]
    with redirect_stdout(io.StringIO()):
        tm = build_train(ref)
        seqs = np.empty(len(tr_events), dtype=object)          # numpy 2.x: ragged input must be an object array (3998)
        for i, e in enumerate(tr_events):
            seqs[i] = e
        blob["train_improvement"] = np.float64(tm.train(seqs, max_iterations=1, verbose=False))
    blob["train_obs"] = np.concatenate(tr_events)
    blob["train_offsets"] = np.concatenate([[0], np.cumsum([len(e) for e in tr_events])]).astype(np.int64)
    blob["train_params"] = np.array([float(v) for s in tm.states[:3] for v in s.distribution.parameters])
    blob["train_dense"] = np.asarray(tm.dense_transition_matrix())
    np.savez_compressed(os.path.join(HERE, "table_emis.npz"), **blob)
    print("events", len(events), "logp", blob["logp"], "vscore", blob["vscore"])


if __name__ == "__main__":
    main()
