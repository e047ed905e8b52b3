#!/usr/bin/env python
"""Golden vector for UniformDistribution.summarize (yahmm.pyx:304-326) at a tiny posterior weight.

    python tests/golden/make_golden_tiny_weight.py      (build container: needs oracle/_ref)

A Uniform state that every event reaches with probability 1e-200: its posterior weight is ~1e-203 per row -- far from
zero for the reference's log-space arithmetic, which therefore re-estimates the state's range from the observed minimum
/ maximum of ALL events in one Baum-Welch iteration.  A scaled linear-space sweep that multiplies the factors in the
wrong order underflows on the way and would leave the range untouched."""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tiny_weight(y):
    m = y.Model(name="tiny")
    a = y.State(y.NormalDistribution(0.0, 1.0), name="a")
    b = y.State(y.NormalDistribution(3.0, 1.0), name="b")
    u = y.State(y.UniformDistribution(-50.0, 50.0), name="u")
    for s in (a, b, u):
        m.add_state(s)
    m.add_transition(m.start, a, 0.6); m.add_transition(m.start, b, 0.4); m.add_transition(m.start, u, 1e-200)
    m.add_transition(a, a, 0.5); m.add_transition(a, b, 0.3); m.add_transition(a, m.end, 0.2)
    m.add_transition(b, b, 0.5); m.add_transition(b, a, 0.3); m.add_transition(b, m.end, 0.2)
    m.add_transition(u, a, 0.5); m.add_transition(u, b, 0.5)
    m.bake()
    return m


def main():
    ref = ref_loader.load_reference_yahmm()
    rs = np.random.RandomState(3)
    evs = [np.asarray(rs.normal(rs.choice([0.0, 3.0]), 1.0, size=rs.randint(3, 40)), dtype=np.float32).astype(np.float64)
           for _ in range(40)]
    m = tiny_weight(ref)
    seqs = np.empty(len(evs), dtype=object)
    for i, e in enumerate(evs):
        seqs[i] = e
    with redirect_stdout(io.StringIO()):
        imp = m.train(seqs, max_iterations=1, verbose=False)
    params = np.array([[float(p) for p in s.distribution.parameters[:2]] for s in m.states[:m.silent_start]])
    names = [s.name for s in m.states[:m.silent_start]]
    obs = np.concatenate(evs)
    offsets = np.concatenate([[0], np.cumsum([len(e) for e in evs])]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "tiny_weight.npz"), obs=obs, offsets=offsets, params=params, names=np.array(names),
                        improvement=imp, dense=np.asarray(m.dense_transition_matrix()))
    print(dict(zip(names, params.tolist())), "min/max of all observations:", obs.min(), obs.max())


if __name__ == "__main__":
    main()
