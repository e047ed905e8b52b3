#!/usr/bin/env python
"""Generate tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref, built by
oracle/build_ref.sh from /root/reference/yahmm/yahmm.pyx).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds, for one model: the reference's baked structure (state names in baked order, dense
transition matrix, indices) and, for a seeded set of events (float32-rounded, stored as float64), the
reference's log_probability, viterbi score and viterbi path, plus full forward/backward tables,
forward_backward outputs and one Baum-Welch iteration on a few of them.
"""
import io
import os
import random
import sys
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from protpore_b200 import peptide_hmm, events  # noqa: E402
import protpore_b200.yahmm as ours  # noqa: E402

ref = ref_loader.load_reference_yahmm()
HERE = os.path.dirname(os.path.abspath(__file__))


def toy(y):
    """SURVEY 8c toy model."""
    m = y.Model(name="toy")
    a = y.State(y.NormalDistribution(0.0, 1.0), name="a")
    b = y.State(y.NormalDistribution(5.0, 1.0), name="b")
    m.add_state(a); m.add_state(b)
    m.add_transition(m.start, a, 0.5); m.add_transition(m.start, b, 0.5)
    m.add_transition(a, a, 0.4); m.add_transition(a, b, 0.5); m.add_transition(a, m.end, 0.1)
    m.add_transition(b, b, 0.4); m.add_transition(b, a, 0.5); m.add_transition(b, m.end, 0.1)
    m.bake()
    return m


def toy_uniform(y):
    """Silent chain + Uniform emissions with bounded support: impossible events exist."""
    m = y.Model(name="toyu")
    u1 = y.State(y.UniformDistribution(0.0, 1.0), name="u1")
    u2 = y.State(y.UniformDistribution(0.5, 2.0), name="u2")
    n1 = y.State(y.NormalDistribution(1.0, 0.0), name="p1")
    s1 = y.State(None, name="s1"); s2 = y.State(None, name="s2")
    for s in (u1, u2, n1, s1, s2):
        m.add_state(s)
    m.add_transition(m.start, s1, 0.6); m.add_transition(m.start, u1, 0.4)
    m.add_transition(s1, s2, 0.5); m.add_transition(s1, u2, 0.5)
    m.add_transition(s2, u1, 0.3); m.add_transition(s2, n1, 0.7)
    m.add_transition(u1, u1, 0.3); m.add_transition(u1, s1, 0.3); m.add_transition(u1, m.end, 0.4)
    m.add_transition(u2, u2, 0.2); m.add_transition(u2, u1, 0.4); m.add_transition(u2, m.end, 0.4)
    m.add_transition(n1, u2, 0.5); m.add_transition(n1, m.end, 0.5)
    m.bake()
    return m


def toy_infinite(y):
    """No edges into end: infinite model (yahmm.pyx:2557-2560)."""
    m = y.Model(name="toyinf")
    a = y.State(y.NormalDistribution(0.0, 1.0), name="a")
    b = y.State(y.NormalDistribution(3.0, 0.5), name="b")
    s = y.State(None, name="s")
    for st in (a, b, s):
        m.add_state(st)
    m.add_transition(m.start, a, 0.5); m.add_transition(m.start, b, 0.5)
    m.add_transition(a, a, 0.6); m.add_transition(a, s, 0.4)
    m.add_transition(s, b, 0.7); m.add_transition(s, a, 0.3)
    m.add_transition(b, b, 0.5); m.add_transition(b, a, 0.5)
    m.bake()
    return m


BUILDERS = {
    "pro_001_py2": lambda y: peptide_hmm.pro_001_model(yahmm_module=y, py2_division=True),
    "pro_001_py3": lambda y: peptide_hmm.pro_001_model(yahmm_module=y, py2_division=False),
    "synth_12": lambda y: peptide_hmm.model_maker(peptide_hmm.synthetic_profile(12, seed=3), "synth12", y, True),
    "toy": toy,
    "toy_uniform": toy_uniform,
    "toy_infinite": toy_infinite,
}


def make_events(name, rmodel, omodel):
    mu = [float(l.split('\t')[1]) for l in open(peptide_hmm.PRO_001)]
    ev = []
    if name.startswith("pro_001"):
        ev += [mu, mu[::2], [v for v in mu for _ in (0, 1)], mu[:10] + mu[7:10] + mu[10:20],
               mu[:5] + [100.0] + mu[5:9], mu[:5] + [200.0] + mu[5:9], [mu[0]], [35.0], mu[20:30]]
        random.seed(20261019)
        n_free = 120 if name.endswith("py2") else 40
        ev += [rmodel.sample() for _ in range(n_free)]
        lens = events.uniform_lengths(24 if name.endswith("py2") else 8, 50, 500, seed=7)
        ev += [list(x) for x in events.sample_events(omodel, lens, seed=11)]
        ev += [list(x) for x in events.sample_events(omodel, [1500], seed=5)]
    elif name == "synth_12":
        lens = events.uniform_lengths(30, 5, 120, seed=2)
        ev += [list(x) for x in events.sample_events(omodel, lens, seed=3)]
    elif name == "toy":
        ev += [[0.1, 0.2, 4.9, 5.2, 0.0], [5.0], [0.0, 0.0, 0.0], [2.5] * 40]
        rng = np.random.RandomState(1)
        ev += [list(rng.normal(2.5, 3.0, size=n)) for n in rng.randint(1, 60, size=40)]
    elif name == "toy_uniform":
        ev += [[0.5], [0.7, 1.0, 1.5], [3.0], [0.2, 3.0, 0.4], [1.0, 1.0], [0.75, 0.75, 1.0, 0.1]]
        rng = np.random.RandomState(2)
        ev += [list(np.round(rng.uniform(0.0, 2.2, size=n), 2)) for n in rng.randint(1, 12, size=60)]
    elif name == "toy_infinite":
        rng = np.random.RandomState(4)
        ev += [list(rng.normal(1.5, 2.0, size=n)) for n in rng.randint(1, 40, size=30)]
    # the GPU path stores float32 observations: feed the reference the same rounded values
    return [np.asarray(np.asarray(e, dtype=np.float32), dtype=np.float64) for e in ev if len(e) > 0]


def main():
    for name, build in BUILDERS.items():
        rmodel = build(ref)
        omodel = build(ours)
        evs = make_events(name, rmodel, omodel)
        out = {}
        out["state_names"] = np.array([s.name for s in rmodel.states])
        out["dense"] = rmodel.dense_transition_matrix()
        out["meta"] = np.array([len(rmodel.states), rmodel.silent_start, rmodel.start_index, rmodel.end_index,
                                rmodel.edge_count(), 0 if rmodel.is_infinite() else 1], dtype=np.int64)
        out["obs"] = np.concatenate(evs)
        out["offsets"] = np.concatenate([[0], np.cumsum([len(e) for e in evs])]).astype(np.int64)
        logp, vscore, paths, poff = [], [], [], [0]
        for e in evs:
            logp.append(rmodel.log_probability(e))
            s, p = rmodel.viterbi(e)
            vscore.append(s)
            if p is not None:
                paths.extend(int(i) for i, _ in p)
                poff.append(poff[-1] + len(p))
            else:
                poff.append(poff[-1])
        out["logp"] = np.array(logp)
        out["vscore"] = np.array(vscore)
        out["paths"] = np.array(paths, dtype=np.int32)
        out["path_offsets"] = np.array(poff, dtype=np.int64)
        # tables / forward-backward on the first few short events
        small = [i for i, e in enumerate(evs) if len(e) <= 90][:4]
        out["table_events"] = np.array(small, dtype=np.int64)
        for j, i in enumerate(small):
            out["fwd_%d" % j] = rmodel.forward(evs[i])
            out["bwd_%d" % j] = rmodel.backward(evs[i])
            with redirect_stdout(io.StringIO()):
                et, ew = rmodel.forward_backward(evs[i])
            if et is not None:
                out["fb_et_%d" % j] = et
                out["fb_ew_%d" % j] = ew
        # posterior decoding (yahmm.pyx:3654-3742) on the table events
        for j, i in enumerate(small):
            with redirect_stdout(io.StringIO()):
                ms, mp = rmodel.maximum_a_posteriori(evs[i])
            if mp is not None:
                out["map_score_%d" % j] = np.array([ms])
                out["map_path_%d" % j] = np.array([k for k, _ in mp], dtype=np.int32)
        # one Baum-Welch iteration on the possible events among the first 30 (fresh model: train mutates)
        if name in ("pro_001_py2", "toy", "synth_12"):
            tmodel = build(ref)
            idx = [i for i in range(min(30, len(evs))) if np.isfinite(logp[i])]
            seqs = np.empty(len(idx), dtype=object)
            for j, i in enumerate(idx):
                seqs[j] = evs[i]
            with redirect_stdout(io.StringIO()):
                imp = tmodel.train(seqs, max_iterations=1, verbose=False)
            out["bw_events"] = np.array(idx, dtype=np.int64)
            out["bw_improvement"] = np.array([imp])
            out["bw_dense"] = tmodel.dense_transition_matrix()
            out["bw_params"] = np.array([[float(p) for p in s.distribution.parameters[:2]]
                                         for s in tmodel.states[:tmodel.silent_start]])
            # Viterbi (hard-assignment) training on the same events (yahmm.pyx:4329-4499)
            vmodel = build(ref)
            with redirect_stdout(io.StringIO()):
                vimp = vmodel.train(seqs, algorithm='viterbi', verbose=False)
            out["vt_improvement"] = np.array([vimp])
            out["vt_dense"] = vmodel.dense_transition_matrix()
            out["vt_params"] = np.array([[float(p) for p in s.distribution.parameters[:2]]
                                         for s in vmodel.states[:vmodel.silent_start]])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "events", len(evs), "obs", len(out["obs"]), "impossible", int(np.sum(~np.isfinite(out["vscore"]))))


if __name__ == "__main__":
    main()
