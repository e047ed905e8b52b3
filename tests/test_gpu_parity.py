"""GPU parity tests proper: the CUDA path, called through the C ABI, against the golden vectors of the
compiled reference and against the CPU oracle on fresh seeded events.

Bars (BASELINE north star): Viterbi state paths identical (an exact tie in the reference's own
arithmetic is the only documented exception -- the oracle reports it as margin 0); Viterbi scores
BIT-identical (same float64 operations in the same order); forward log-probabilities within
1e-3 nats absolute or 1e-6 relative (util.close) -- in practice ~1e-9 relative.
"""
import os

import numpy as np
import pytest

import util
from protpore_b200 import _engine, events

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models():
    return {n: util.build_model(n) for n in util.MODEL_NAMES}


@pytest.mark.parametrize("name", util.MODEL_NAMES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_viterbi_batch_matches_golden(models, name, dtype):
    g = util.load_golden(name)
    model = models[name]
    evs = [np.asarray(e, dtype=dtype) for e in util.golden_events(g)]
    res = model.viterbi_batch(evs, dtype=dtype, states=False)
    for (score, path), gs, gp in zip(res, g["vscore"], util.golden_paths(g)):
        assert score == gs                                  # bit-identical
        if gp is None:
            assert path is None
        else:
            assert np.array_equal(path, gp)


@pytest.mark.parametrize("name", util.MODEL_NAMES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_forward_batch_matches_golden(models, name, dtype):
    g = util.load_golden(name)
    model = models[name]
    evs = [np.asarray(e, dtype=dtype) for e in util.golden_events(g)]
    lp = model.log_probability_batch(evs, dtype=dtype)
    assert util.close(lp, g["logp"]).all()
    fin = np.isfinite(g["logp"])
    assert np.max(np.abs(lp[fin] - g["logp"][fin]) / np.maximum(1.0, np.abs(g["logp"][fin]))) < 1e-6


@pytest.mark.parametrize("name", util.MODEL_NAMES)
def test_single_sequence_api_matches_golden(models, name):
    """Model.viterbi / log_probability / forward / backward / forward_backward, reference signatures."""
    g = util.load_golden(name)
    model = models[name]
    evs = util.golden_events(g)
    gp = util.golden_paths(g)
    for i in list(range(0, len(evs), max(1, len(evs) // 12)))[:12]:
        score, path = model.viterbi(evs[i])
        assert score == g["vscore"][i]
        if gp[i] is None:
            assert path is None
        else:
            assert [k for k, _ in path] == list(gp[i])
            assert all(s is model.states[k] for k, s in path)
        lp = model.log_probability(evs[i])
        assert util.close([lp], [g["logp"][i]], abs_tol=1e-9, rel_tol=1e-12).all()
    for j, i in enumerate(g["table_events"]):
        f, b = model.forward(evs[i]), model.backward(evs[i])
        gf, gb = g["fwd_%d" % j], g["bwd_%d" % j]
        assert f.shape == gf.shape and b.shape == gb.shape
        assert util.close(f, gf, abs_tol=1e-9, rel_tol=1e-12).all()
        assert util.close(b, gb, abs_tol=1e-9, rel_tol=1e-12).all()
        et, ew = model.forward_backward(evs[i])
        if "fb_et_%d" % j in g:
            assert np.allclose(et, g["fb_et_%d" % j], rtol=1e-9, atol=1e-12)
            assert util.close(ew, g["fb_ew_%d" % j], abs_tol=1e-8, rel_tol=1e-10).all()
        else:
            assert et is None and ew is None


@pytest.mark.parametrize("name", ["pro_001_py2", "toy", "synth_12"])
def test_baum_welch_iteration_matches_golden(name):
    g = util.load_golden(name)
    model = util.build_model(name)          # fresh: training mutates
    evs = util.golden_events(g)
    seqs = [evs[i] for i in g["bw_events"]]
    imp = model.train(seqs, max_iterations=1, verbose=False)
    assert imp == pytest.approx(float(g["bw_improvement"][0]), rel=1e-7, abs=1e-7)
    dense = model.dense_transition_matrix()
    fin = np.isfinite(g["bw_dense"])
    assert np.array_equal(np.isfinite(dense), fin)
    assert np.allclose(dense[fin], g["bw_dense"][fin], rtol=1e-8, atol=1e-9)
    params = np.array([[float(p) for p in s.distribution.parameters[:2]] for s in model.states[:model.silent_start]])
    assert np.allclose(params, g["bw_params"], rtol=1e-7, atol=1e-9)


def test_fresh_events_against_oracle(models):
    """2,000 length-conditioned events from the CUDA sampler: Viterbi bit-exact, forward within tolerance."""
    model = models["pro_001_py2"]
    orc = util.oracle_for(model)
    lens = events.uniform_lengths(2000, 50, 500, seed=42)
    obs, offsets = model.engine().sample_events(lens, seed=1234)
    assert obs.dtype == np.float32 and np.isfinite(obs).all()
    scores, plen, paths = model.engine().viterbi_batch(obs, offsets)
    lp = model.engine().forward_batch(obs, offsets)
    poff = np.concatenate([[0], np.cumsum(np.maximum(plen, 0))])
    check = np.random.RandomState(0).choice(len(lens), size=300, replace=False)
    worst = 0.0
    for e in check:
        x = obs[offsets[e]:offsets[e + 1]].astype(np.float64)
        s, p, margin = orc.viterbi(x, with_margin=True)
        assert scores[e] == s
        got = paths[poff[e]:poff[e + 1]]
        assert np.array_equal(got, p) or margin == 0.0     # documented exception: exact tie in the reference
        ref_lp = orc.log_probability(x)
        assert util.close([lp[e]], [ref_lp]).all()
        worst = max(worst, abs(lp[e] - ref_lp) / max(1.0, abs(ref_lp)))
    assert worst < 1e-7
    assert (scores <= lp + 1e-6).all()                     # best path <= sum over paths


def test_batch_composition_invariance(models):
    """Per-event results do not depend on batch order, batch size, chunking or observation dtype."""
    model = models["pro_001_py2"]
    lens = events.log_uniform_lengths(3000, 1, 900, seed=5)
    obs, offsets = model.engine().sample_events(lens, seed=77)
    evs = [obs[offsets[i]:offsets[i + 1]] for i in range(len(lens))]
    s0, l0, p0 = model.engine().viterbi_batch(obs, offsets)
    f0 = model.engine().forward_batch(obs, offsets)
    perm = np.random.RandomState(3).permutation(len(evs))
    o2, off2 = _engine.pack_events([evs[i] for i in perm], np.float32)
    model.engine().set_workspace_limit(64 << 20)           # forces several chunks
    try:
        s2, l2, p2 = model.engine().viterbi_batch(o2, off2)
        f2 = model.engine().forward_batch(o2, off2)
    finally:
        model.engine().set_workspace_limit(32 << 30)
    assert np.array_equal(s2, s0[perm]) and np.array_equal(l2, l0[perm]) and np.array_equal(f2, f0[perm])
    a = np.concatenate([[0], np.cumsum(l0)])
    b = np.concatenate([[0], np.cumsum(l2)])
    for j in range(0, len(perm), 37):
        assert np.array_equal(p2[b[j]:b[j + 1]], p0[a[perm[j]]:a[perm[j] + 1]])
    o3, off3 = _engine.pack_events(evs[:500], np.float64)
    s3, l3, _ = model.engine().viterbi_batch(o3, off3)
    assert np.array_equal(s3, s0[:500]) and np.array_equal(l3, l0[:500])


def test_path_is_a_valid_walk_with_the_reported_score(models):
    """Size-independent property: re-scoring the decoded path along the model's edges gives the score."""
    model = models["pro_001_py2"]
    lens = events.uniform_lengths(200, 50, 500, seed=9)
    obs, offsets = model.engine().sample_events(lens, seed=5)
    evs = [obs[offsets[i]:offsets[i + 1]].astype(np.float64) for i in range(len(lens))]
    for x, (score, path) in list(zip(evs, model.viterbi_batch(evs, dtype=np.float64)))[::10]:
        states = [s for _, s in path]
        assert states[0] is model.start and states[-1] is model.end
        assert sum(1 for s in states if not s.is_silent()) == len(x)
        assert model.log_probability(x, path=states) == pytest.approx(score, rel=1e-12, abs=1e-9)


def test_edge_cases(models):
    model = models["toy_uniform"]
    # impossible events: (-inf, None) like the reference (yahmm.pyx:3621-3623); forward -inf
    res = model.viterbi_batch([np.array([3.0]), np.array([0.5]), np.array([0.2, 3.0, 0.4])], dtype=np.float64)
    assert res[0] == (float("-inf"), None) and res[2] == (float("-inf"), None) and res[1][1] is not None
    lp = model.log_probability_batch([np.array([3.0]), np.array([0.5])], dtype=np.float64)
    assert lp[0] == float("-inf") and np.isfinite(lp[1])
    assert model.forward_backward(np.array([3.0])) == (None, None)
    # empty sequence: ValueError like the reference's cvarray (yahmm.pyx:2836)
    with pytest.raises(ValueError):
        model.viterbi(np.array([]))
    with pytest.raises(ValueError):
        model.log_probability_batch([np.array([0.5]), np.array([])])
    # zero events
    assert model.viterbi_batch([]) == []
    # too-small path buffer is reported, then retried with the size the library asks for
    big = models["pro_001_py2"]
    obs, offsets = big.engine().sample_events(np.array([300, 20, 77]), seed=1)
    s1, l1, p1 = big.engine().viterbi_batch(obs, offsets, capacity=5)
    s2, l2, p2 = big.engine().viterbi_batch(obs, offsets)
    assert np.array_equal(p1, p2) and np.array_equal(l1, l2) and len(p1) == l1.sum()


def test_outlier_event_goes_through_exact_rerun(models):
    """x = 1e6 makes every emission underflow the scaled linear-space sweep; the exact log-space kernel re-runs it."""
    model = models["pro_001_py2"]
    orc = util.oracle_for(model)
    mu = util.golden_events(util.load_golden("pro_001_py2"))[0]
    x = np.concatenate([mu[:5], [1.0e4], mu[5:9]])
    lp = model.log_probability_batch([x, mu], dtype=np.float64)
    assert util.close(lp, [orc.log_probability(x), orc.log_probability(mu)]).all()


def test_pypore_apply_hmm(models):
    from protpore_b200 import pypore
    model = models["pro_001_py2"]
    g = util.load_golden("pro_001_py2")
    evs = util.golden_events(g)[:6]
    pev = [pypore.Event(segments=[pypore.Segment(mean=float(v)) for v in x]) for x in evs]
    one = pev[0].apply_hmm(model)                        # the reference's call (DataTypes.py:349-355)
    assert one[0] == g["vscore"][0] and one[1][0][1].name.endswith("-start")
    many = pypore.apply_hmm_batch(pev, model, 'viterbi')
    assert [m[0] for m in many] == list(g["vscore"][:6])


def test_distributed_baum_welch_single_process_equals_train():
    """protpore_b200.parallel.train_baum_welch without a process group == Model.train (reference stopping rule)."""
    from protpore_b200 import parallel
    g = util.load_golden("toy")
    evs = util.golden_events(g)
    seqs = [evs[i] for i in g["bw_events"]]
    a, b = util.build_model("toy"), util.build_model("toy")
    ia = a.train(seqs, max_iterations=2, verbose=False)
    ib = parallel.train_baum_welch(b, seqs, max_iterations=2, dtype=np.float64)
    assert ia == pytest.approx(ib, rel=1e-9, abs=1e-9)
    assert np.allclose(a.dense_transition_matrix(), b.dense_transition_matrix(), rtol=1e-10, atol=1e-12, equal_nan=True)


def test_large_model_config4_global_column_kernels():
    """BASELINE config 4 shape: a ~2,000-state synthetic profile (skip / back-slip chains 334 long).  Its column (3,004
    slots x 256 B) does not fit shared memory, so the lane-per-event kernels run with the column in a per-CTA global
    scratch (k_*_lane<..., CG = true>) -- same bars as the shared-memory kernels: Viterbi bit-exact, paths identical,
    forward within the north-star tolerance.  Several groups of 32 events, ragged lengths."""
    from protpore_b200 import peptide_hmm
    import protpore_b200.yahmm as y
    model = peptide_hmm.model_maker(peptide_hmm.synthetic_profile(334, seed=11), "synth334", y, True)
    assert len(model.states) == 6 * 334 - 1
    assert "lane_ok 1" in _engine.describe_schedule(model)          # in-degrees fit, the column (3,004 slots) does not
    orc = util.oracle_for(model)
    lens = np.concatenate([[40, 700, 1, 333, 2500, 90], np.random.RandomState(3).randint(1, 60, size=70)])
    evs = [np.asarray(x, dtype=np.float64) for x in events.sample_events(model, lens, seed=4)]
    res = model.viterbi_batch(evs, dtype=np.float64, states=False)
    lp = model.log_probability_batch(evs, dtype=np.float64)
    check = list(range(8)) + [20, 41, 75]
    for e in check:
        x, (score, path), l = evs[e], res[e], lp[e]
        s, p, margin = orc.viterbi(x, with_margin=True)
        assert score == s
        assert np.array_equal(path, p) or margin == 0.0
        assert util.close([l], [orc.log_probability(x)]).all()
    # float32 observations give the same results as their float64 images
    res32 = model.viterbi_batch([e.astype(np.float32) for e in evs], dtype=np.float32, states=False)
    assert [r[0] for r in res32] == [r[0] for r in res]
    # the exact state-parallel kernels (full tables) agree with the sweep's score
    f = model.forward(evs[0])
    assert util.close([f[len(evs[0]), model.end_index]], [lp[0]]).all()


def test_mixed_length_batch_config5(models):
    """BASELINE config 5 shape: 20-20,000 segments, log-uniform, length-bucketed: spot-check against the oracle."""
    model = models["pro_001_py2"]
    orc = util.oracle_for(model)
    lens = events.log_uniform_lengths(400, 20, 20000, seed=8)
    obs, offsets = model.engine().sample_events(lens, seed=21)
    scores, plen, paths = model.engine().viterbi_batch(obs, offsets)
    lp = model.engine().forward_batch(obs, offsets)
    poff = np.concatenate([[0], np.cumsum(np.maximum(plen, 0))])
    for e in list(np.argsort(lens)[-3:]) + list(np.argsort(lens)[:3]) + [7, 100, 250]:
        x = obs[offsets[e]:offsets[e + 1]].astype(np.float64)
        s, p, margin = orc.viterbi(x, with_margin=True)
        assert scores[e] == s
        assert np.array_equal(paths[poff[e]:poff[e + 1]], p) or margin == 0.0
        assert util.close([lp[e]], [orc.log_probability(x)]).all()


def _random_model(rng, idx):
    """Random small HMM: emitting states with Normal / point-mass / Uniform emissions, a silent DAG with chains,
    sometimes infinite (no edge into end), random state weights -- the topologies PyPore/hmm.py-style builders make."""
    import protpore_b200.yahmm as y
    m = y.Model(name="rnd%d" % idx)
    n_e, n_s = int(rng.randint(2, 9)), int(rng.randint(0, 7))
    emit = []
    for k in range(n_e):
        kind = rng.randint(0, 4)
        if kind <= 1:
            d = y.NormalDistribution(float(rng.uniform(-3, 3)), float(rng.uniform(0.3, 2.0)))
        elif kind == 2:
            d = y.NormalDistribution(float(np.round(rng.uniform(-2, 2), 1)), 0.0)
        else:
            lo = float(rng.uniform(-4, 0))
            d = y.UniformDistribution(lo, lo + float(rng.uniform(1, 6)))
        emit.append(y.State(d, name="e%d" % k, weight=float(rng.choice([1.0, 1.0, 0.5, 2.0]))))
    sil = [y.State(None, name="s%d" % k) for k in range(n_s)]
    for s in emit + sil:
        m.add_state(s)
    finite = rng.rand() < 0.75
    nodes = emit + sil
    for a in [m.start] + nodes:
        targets = []
        for b in nodes + ([m.end] if finite else []):
            if a is b and a.is_silent():
                continue
            if a.is_silent() and b is not m.end and b.is_silent():
                # keep the silent sub-graph acyclic: start -> s_i -> s_j only for i < j
                if a is not m.start and sil.index(a) >= sil.index(b):
                    continue
            if rng.rand() < (0.45 if b is not m.end else 0.6):
                targets.append(b)
        if not targets:
            targets = [emit[int(rng.randint(n_e))]]
        p = rng.dirichlet(np.ones(len(targets)))
        for b, pb in zip(targets, p):
            m.add_transition(a, b, float(pb))
    m.bake()
    return m


def test_random_models_against_oracle():
    """60 random HMMs x 12 random events each: Viterbi bit-exact (paths identical unless the oracle reports an exact
    tie), forward / tables within tolerance, impossible events in-band."""
    rng = np.random.RandomState(2026)
    checked = 0
    for idx in range(60):
        model = _random_model(rng, idx)
        if model.silent_start == 0:
            continue
        orc = util.oracle_for(model)
        evs = []
        for _ in range(12):
            n = int(rng.randint(1, 40))
            x = rng.uniform(-4, 4, size=n)
            evs.append(np.round(x, 1) if rng.rand() < 0.4 else x)      # rounded values can hit the point masses
        res = model.viterbi_batch(evs, dtype=np.float64, states=False)
        lp = model.log_probability_batch(evs, dtype=np.float64)
        for x, (score, path), l in zip(evs, res, lp):
            s, p, margin = orc.viterbi(x, with_margin=True)
            assert score == s, (idx, score, s)
            assert (path is None) == (p is None)
            if p is not None:
                assert np.array_equal(path, p) or margin == 0.0, idx
            assert util.close([l], [orc.log_probability(x)], abs_tol=1e-5, rel_tol=1e-7).all(), (idx, l, orc.log_probability(x))
            checked += 1
        x = evs[0]
        assert util.close(model.forward(x), orc.forward(x), abs_tol=1e-9, rel_tol=1e-12).all()
        assert util.close(model.backward(x), orc.backward(x), abs_tol=1e-9, rel_tol=1e-12).all()
    assert checked > 500


def test_silent_runs_whose_items_read_each_other():
    """Regression (found by tools/parity_soak.py, 4 of 8,000 random models): two consecutive silent states of one shape
    formed a run although the second read the first through an edge that is not the forwarded chain edge; the kernels
    load all sources of a group before storing any, so it saw the previous row's value.  The lowering now only builds
    runs of independent items or of pure chain links.  Replays the soak's random stream up to two of the models that
    failed (Viterbi score off by up to 59 nats) and checks them against the oracle."""
    rng = np.random.RandomState(777)
    checked = 0
    for idx in range(2570):
        model = _random_model(rng, idx)
        if model.silent_start == 0:
            continue
        evs = []
        for _ in range(10):
            x = rng.uniform(-4, 4, size=int(rng.randint(1, 120)))
            evs.append(np.round(x, 1) if rng.rand() < 0.4 else x)
        if idx not in (1259, 2569):
            continue
        assert (len(model.states), model.silent_start) == {1259: (13, 5), 2569: (11, 4)}[idx]
        orc = util.oracle_for(model)
        res = model.viterbi_batch(evs, dtype=np.float64, states=False)
        lp = model.log_probability_batch(evs, dtype=np.float64)
        for x, (score, path), l in zip(evs, res, lp):
            s, p, margin = orc.viterbi(x, with_margin=True)
            assert score == s, idx
            assert (path is None) == (p is None)
            if p is not None:
                assert np.array_equal(path, p) or margin == 0.0, idx
            assert util.close([l], [orc.log_probability(x)]).all(), idx
            checked += 1
    assert checked == 20


@pytest.mark.parametrize("name", ["pro_001_py2", "toy", "synth_12"])
def test_viterbi_training_matches_golden(name):
    """Model.train(algorithm='viterbi') (yahmm.pyx:4329-4499): paths from ONE batched Viterbi call, host M-step."""
    g = util.load_golden(name)
    model = util.build_model(name)
    evs = util.golden_events(g)
    seqs = [evs[i] for i in g["bw_events"]]
    imp = model.train(seqs, algorithm='viterbi', verbose=False)
    assert imp == pytest.approx(float(g["vt_improvement"][0]), rel=1e-7, abs=1e-6)
    dense = model.dense_transition_matrix()
    fin = np.isfinite(g["vt_dense"])
    assert np.array_equal(np.isfinite(dense), fin)
    assert np.allclose(dense[fin], g["vt_dense"][fin], rtol=1e-10, atol=1e-12)
    params = np.array([[float(p) for p in s.distribution.parameters[:2]] for s in model.states[:model.silent_start]])
    assert np.allclose(params, g["vt_params"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", util.MODEL_NAMES)
def test_maximum_a_posteriori_matches_golden(models, name):
    g = util.load_golden(name)
    model = models[name]
    evs = util.golden_events(g)
    for j, i in enumerate(g["table_events"]):
        score, path = model.maximum_a_posteriori(evs[i])
        if "map_score_%d" % j not in g:
            assert score is None and path is None
            continue
        assert score == pytest.approx(float(g["map_score_%d" % j][0]), rel=1e-9, abs=1e-9)
        assert [k for k, _ in path] == list(g["map_path_%d" % j])


def test_config1_thousand_free_running_events(models):
    """BASELINE config 1: 1,000 events by free-running Model.sample() (mean ~15 segments): Viterbi + forward
    log-probability, full parity against the oracle (pinned to the compiled reference) on every event."""
    import random
    model = models["pro_001_py2"]
    orc = util.oracle_for(model)
    random.seed(20261019)
    evs = []
    while len(evs) < 1000:
        x = np.asarray(np.asarray(model.sample(), dtype=np.float32), dtype=np.float64)
        if len(x):
            evs.append(x)
    res = model.viterbi_batch(evs, dtype=np.float64, states=False)
    lp = model.log_probability_batch(evs, dtype=np.float64)
    n_seg = 0
    for x, (score, path), l in zip(evs, res, lp):
        s, p, margin = orc.viterbi(x, with_margin=True)
        assert score == s
        assert np.array_equal(path, p) or margin == 0.0
        assert util.close([l], [orc.log_probability(x)]).all()
        n_seg += len(x)
    assert 5 < n_seg / 1000.0 < 40


def test_tied_states_and_labelled_training():
    """Tied emission distributions (shared object) and tied edge groups: forward_backward(tie=True) pools
    posteriors (yahmm.pyx:3325-3361); labelled training follows yahmm.pyx:4359-4499.  Checked against a direct
    numpy evaluation of the same formulas from the oracle's tables."""
    import protpore_b200.yahmm as y
    shared = y.NormalDistribution(1.0, 0.7)
    m = y.Model(name="tied")
    a = y.State(shared, name="a"); b = y.State(shared, name="b"); c = y.State(y.NormalDistribution(4.0, 1.0), name="c")
    for s in (a, b, c):
        m.add_state(s)
    m.add_transition(m.start, a, 0.6); m.add_transition(m.start, c, 0.4)
    m.add_transition(a, b, 0.5, group="g"); m.add_transition(a, a, 0.3); m.add_transition(a, m.end, 0.2)
    m.add_transition(b, c, 0.5, group="g"); m.add_transition(b, b, 0.3); m.add_transition(b, m.end, 0.2)
    m.add_transition(c, a, 0.5); m.add_transition(c, c, 0.3); m.add_transition(c, m.end, 0.2)
    m.bake()
    x = np.array([0.8, 1.3, 3.9, 4.2, 1.1, 0.7])
    et, ew = m.forward_backward(x, tie=False)
    et2, ew2 = m.forward_backward(x, tie=True)
    ia, ib = [s.name for s in m.states].index("a"), [s.name for s in m.states].index("b")
    pooled = np.logaddexp(ew[:, ia], ew[:, ib])
    assert np.allclose(ew2[:, ia], pooled, rtol=1e-12, atol=1e-12) and np.allclose(ew2[:, ib], pooled, rtol=1e-12, atol=1e-12)
    assert np.allclose(et, et2)
    # labelled training: counts along the given path (plus start -> first, last -> end), tied edges pooled
    path = [m.start, a, a, c, c, a, b, m.end]
    before = m.dense_transition_matrix().copy()
    m.train([(x, path)], algorithm='labelled', verbose=False)
    dense = m.dense_transition_matrix()
    names = [s.name for s in m.states]
    i = {n: names.index(n) for n in names}
    # group "g" = {a->b, b->c}: counts 1 and 0 -> both 1; a's out-edges: a->b 1, a->a 1, a->end 0 (path ends in b->end)
    assert np.exp(dense[i["a"], i["b"]]) == pytest.approx(0.5) and np.exp(dense[i["a"], i["a"]]) == pytest.approx(0.5)
    assert np.isneginf(dense[i["a"], m.end_index])
    # b: b->c pooled count 1, b->end 1 -> 0.5 each
    assert np.exp(dense[i["b"], i["c"]]) == pytest.approx(0.5) and np.exp(dense[i["b"], m.end_index]) == pytest.approx(0.5)
    # the shared distribution was re-estimated once from the symbols of a AND b
    assert shared.parameters[0] == pytest.approx(np.mean([0.8, 1.3, 1.1, 0.7]))
    assert before.shape == dense.shape


def test_lane_forward_backward_statistics(models):
    """Baum-Welch E-step on the lane-per-event kernel (pp_lane_fb.cuh) against the oracle's bw_accumulate
    (yahmm.pyx:4107-4236) and against the exact state-parallel kernel, including impossible events and an outlier
    that the scaled sweep hands to the exact kernel."""
    model = models["pro_001_py2"]
    eng = model.engine()
    assert "bw_ok 1" in _engine.describe_schedule(model)
    lens = events.uniform_lengths(300, 1, 400, seed=7)
    obs, offsets = eng.sample_events(lens, seed=99)
    obs = obs.astype(np.float64)
    obs[offsets[5]:offsets[5] + 1] = 250.0                      # underflows every linear-space factor: exact re-run
    obs[offsets[9] + 2] = 119.5                                  # far outlier the scaled sweep still resolves
    logp, edge, mom, mm, _ = eng.forward_backward_batch(obs, offsets)
    info = eng.forward_backward_info()
    assert info["path"] in ("lane", "profile") and info["rejects"] >= 1 and info["failed"] == 0
    # exact kernel: kernel path 2 keeps the batch on the state-parallel path
    eng.set_kernel_path(2)
    try:
        logp_x, edge_x, mom_x, mm_x, _ = eng.forward_backward_batch(obs, offsets, posteriors=True)
        assert eng.forward_backward_info()["path"] == "exact"
    finally:
        eng.set_kernel_path(int(os.environ.get("PP_KERNEL_PATH", "0")))
    assert np.array_equal(np.isfinite(logp), np.isfinite(logp_x))
    fin = np.isfinite(logp)
    assert np.allclose(logp[fin], logp_x[fin], rtol=1e-12, atol=1e-9)
    assert np.allclose(edge, edge_x, rtol=1e-9, atol=1e-12)
    assert np.allclose(mom, mom_x, rtol=1e-9, atol=1e-11)
    assert np.array_equal(mm, mm_x)
    # oracle on a subset (float64 CPU restatement of the reference)
    orc = util.oracle_for(model)
    sub = list(range(0, 300, 6))
    seqs = [obs[offsets[e]:offsets[e + 1]] for e in sub]
    so, oo = np.concatenate(seqs), np.concatenate([[0], np.cumsum([len(s) for s in seqs])])
    lp_s, edge_s, mom_s, mm_s, _ = eng.forward_backward_batch(so, oo)
    exp_t, mom_o, mm_o, lz = orc.bw_accumulate(seqs)
    b = model._baked
    src = np.repeat(np.arange(len(model.states)), np.diff(b.out_ptr))
    assert np.allclose(edge_s, exp_t[src, b.out_idx], rtol=1e-9, atol=1e-12)
    assert np.allclose(mom_s, mom_o, rtol=1e-9, atol=1e-11)
    assert np.array_equal(mm_s, mm_o)
    assert np.allclose(lp_s, lz, rtol=1e-12, atol=1e-9)


def test_lane_forward_backward_accumulates_and_is_composition_invariant(models):
    """Accumulators are read-modify-write; statistics of a batch = sum over its halves (different bucketing)."""
    model = models["pro_001_py2"]
    eng = model.engine()
    lens = events.log_uniform_lengths(500, 1, 700, seed=3)
    obs, offsets = eng.sample_events(lens, seed=5)
    _, edge, mom, mm, _ = eng.forward_backward_batch(obs, offsets)
    assert eng.forward_backward_info()["path"] in ("lane", "profile")
    h = 250
    o1, f1 = obs[:offsets[h]], offsets[:h + 1]
    o2, f2 = obs[offsets[h]:], offsets[h:] - offsets[h]
    _, e1, m1, mm1, _ = eng.forward_backward_batch(o1, f1)
    acc = (e1, m1, mm1)
    _, e2, m2, mm2, _ = eng.forward_backward_batch(o2, f2, accum=acc)
    assert np.allclose(e2, edge, rtol=1e-11, atol=1e-13) and np.allclose(m2, mom, rtol=1e-11, atol=1e-12)
    assert np.array_equal(mm2, mm)
    # posterior mass: sum_k gamma_r[k] = 1 for every observation, so sum of `sum w` = number of observations
    assert abs(mom[:, 0].sum() - len(obs)) < 1e-6 * len(obs)


def test_kernels_follow_a_changed_emission_kind():
    """After one Baum-Welch iteration on sampled events the sigma == 0 insert states get the min_std clamp
    (yahmm.pyx:513) and leave the point-mass branch: the run schedule must be lowered again.  The lane kernels of
    the updated model are checked against the exact state-parallel tables and the oracle, and a second iteration
    still runs on the lane forward-backward kernel."""
    model = util.build_model("pro_001_py2")
    eng = model.engine()
    lens = events.uniform_lengths(400, 20, 200, seed=11)
    obs, offsets = eng.sample_events(lens, seed=12)
    seqs = [obs[offsets[i]:offsets[i + 1]].astype(np.float64) for i in range(len(lens))]
    from protpore_b200 import parallel
    parallel.baum_welch_iteration(model, seqs)
    sig = np.array([float(s.distribution.parameters[1]) for s in model.states[:model.silent_start]
                    if s.distribution.name == "NormalDistribution"])
    assert (sig > 0).all() and (sig == 0.01).any()              # some insert state moved off sigma == 0
    eng = model.engine()
    orc = util.oracle_for(model)
    lp = model.log_probability_batch(seqs[:40], dtype=np.float64)
    vit = model.viterbi_batch(seqs[:40], dtype=np.float64, states=False)
    for i in range(40):
        assert util.close([lp[i]], [orc.log_probability(seqs[i])]).all()
        s, p, margin = orc.viterbi(seqs[i], with_margin=True)
        assert vit[i][0] == s and (np.array_equal(vit[i][1], p) or margin == 0.0)
    logp, edge, mom, mm, _ = eng.forward_backward_batch(obs.astype(np.float64), offsets)
    assert eng.forward_backward_info()["path"] in ("lane", "profile")
    exp_t, mom_o, mm_o, lz = orc.bw_accumulate(seqs[:60])
    so = np.concatenate(seqs[:60])
    oo = np.concatenate([[0], np.cumsum([len(s) for s in seqs[:60]])])
    _, edge_s, mom_s, mm_s, _ = eng.forward_backward_batch(so, oo)
    b = model._baked
    src = np.repeat(np.arange(len(model.states)), np.diff(b.out_ptr))
    assert np.allclose(edge_s, exp_t[src, b.out_idx], rtol=1e-9, atol=1e-12)
    assert np.allclose(mom_s, mom_o, rtol=1e-9, atol=1e-11)


def test_score_batch_equals_the_two_single_calls(models):
    """pp_score_batch (one upload, one GPU-side bucketing, both sweeps) gives exactly what pp_viterbi_batch and
    pp_forward_batch give -- host buffers, device buffers, several chunks, and a model the profile kernels do not serve."""
    import torch
    for name in ("pro_001_py2", "pro_001_py3", "toy_uniform"):
        model = models[name]
        eng = model.engine()
        lens = events.log_uniform_lengths(1500, 1, 700, seed=11)
        if name == "toy_uniform":
            g = util.load_golden(name)
            obs, offsets = _engine.pack_events([np.asarray(e, dtype=np.float32) for e in util.golden_events(g)], np.float32)
        else:
            obs, offsets = eng.sample_events(lens, seed=12)
        s0, l0, p0 = eng.viterbi_batch(obs, offsets)
        f0 = eng.forward_batch(obs, offsets)
        s1, l1, p1, f1 = eng.score_batch(obs, offsets)
        assert np.array_equal(s0, s1) and np.array_equal(l0, l1) and np.array_equal(p0, p1)
        assert np.array_equal(f0, f1, equal_nan=True)
        eng.set_workspace_limit(48 << 20)                    # several chunks
        try:
            s2, l2, p2, f2 = eng.score_batch(obs, offsets)
        finally:
            eng.set_workspace_limit(32 << 30)
        assert np.array_equal(s0, s2) and np.array_equal(l0, l2) and np.array_equal(p0, p2) and np.array_equal(f0, f2, equal_nan=True)
        dev = torch.device("cuda", eng.device)
        s3, l3, p3, f3 = eng.score_batch(torch.from_numpy(obs).to(dev), torch.from_numpy(offsets).to(dev))
        assert np.array_equal(s0, s3.cpu().numpy()) and np.array_equal(l0, l3.cpu().numpy())
        assert np.array_equal(p0, p3.cpu().numpy()) and np.array_equal(f0, f3.cpu().numpy(), equal_nan=True)


def test_profile_kernels_equal_run_based_kernels(models):
    """The register-resident linear-profile kernels and the run-based lane kernels are two implementations of the same
    arithmetic: bit-identical Viterbi scores, identical paths, forward within 1e-7 relative of each other -- on the
    untrained profile (point-mass inserts), the Python-3 variant (Normal inserts) and a small synthetic profile."""
    for name in ("pro_001_py2", "pro_001_py3", "synth_12"):
        model = models[name]
        eng = model.engine()
        assert eng.kernel_info()["family"] == "profile", name
        lens = events.log_uniform_lengths(2500, 1, 600, seed=21)
        obs, offsets = eng.sample_events(lens, seed=22)
        obs[offsets[7]] = 300.0                              # an outlier far from every state
        s0, l0, p0, f0 = eng.score_batch(obs, offsets)
        eng.set_kernel_path(1)
        try:
            assert eng.kernel_info()["family"] == "lane"
            s1, l1, p1, f1 = eng.score_batch(obs, offsets)
        finally:
            eng.set_kernel_path(0)
        assert np.array_equal(s0, s1) and np.array_equal(l0, l1) and np.array_equal(p0, p1)
        fin = np.isfinite(f1)
        assert np.array_equal(fin, np.isfinite(f0))
        assert np.max(np.abs(f0[fin] - f1[fin]) / np.maximum(1.0, np.abs(f1[fin]))) < 1e-7   # two MUFU-based sweeps, each < 1e-7 of the reference


def test_profile_forward_backward_equals_run_based_and_exact_kernels(models):
    """k_profile_fb (pp_profile_fb.cuh) against k_fb_lane and the exact state-parallel kernel: log-probabilities, expected
    transition counts, emission moments and observed ranges of the Baum-Welch E-step (yahmm.pyx:4107-4236), with an event
    the scaled sweep must hand to the exact kernel, a far outlier, and lengths from 1 to 600 in one batch."""
    for name in ("pro_001_py2", "pro_001_py3", "synth_12"):
        model = models[name]
        eng = model.engine()
        assert "profile_fb ok" in _engine.describe_schedule(model), name
        lens = events.log_uniform_lengths(700, 1, 600, seed=41)
        obs, offsets = eng.sample_events(lens, seed=42)
        obs = obs.astype(np.float64)
        obs[offsets[5]:offsets[5] + 1] = 250.0
        obs[offsets[9] + 2] = 119.5
        lp0, e0, m0, mm0, _ = eng.forward_backward_batch(obs, offsets)
        info = eng.forward_backward_info()
        assert info["path"] == "profile" and info["failed"] == 0 and info["rejects"] >= 1, (name, info)
        eng.set_kernel_path(1)
        try:
            lp1, e1, m1, mm1, _ = eng.forward_backward_batch(obs, offsets)
            # synth_12's batch holds a 2-observation event with log P = -845: k_fb_lane fails its range check on it and the
            # exact kernel redoes its batch; k_profile_fb hands over that event alone
            assert eng.forward_backward_info()["path"] in ("lane", "lane-failed")
        finally:
            eng.set_kernel_path(0)
        # posteriors (the emission weights forward_backward() returns, 3304-3321) from the profile kernel ...
        eng.set_kernel_path(3)
        try:
            lpp, ep, mp, mmp, post0 = eng.forward_backward_batch(obs, offsets, posteriors=True)
            assert eng.forward_backward_info()["path"] == "profile", name
        finally:
            eng.set_kernel_path(0)
        assert np.allclose(ep, e0, rtol=1e-12, atol=1e-12) and np.array_equal(mmp, mm0)
        # ... and from the exact state-parallel kernel
        eng.set_kernel_path(2)
        try:
            lpx, ex, mx, mmx, postx = eng.forward_backward_batch(obs, offsets, posteriors=True)
            assert eng.forward_backward_info()["path"] == "exact"
        finally:
            eng.set_kernel_path(0)
        rows = np.repeat(np.isfinite(lpx), np.diff(offsets))            # impossible events have no posteriors
        a, b = post0[rows], postx[rows]
        # every posterior above e^-250 within 1e-8 (measured: 2e-11; the largest posterior a soak found lost: e^-296); below that a scaled linear-space sweep -- each spans
        # ~2^1500 under its row maximum -- loses precision or flushes to -inf where the log-space kernel stays finite
        big = np.isfinite(b) & (b > -250.0)
        assert np.all(np.isfinite(a[big])), name
        assert np.max(np.abs(a[big] - b[big])) < 1e-8, (name, np.max(np.abs(a[big] - b[big])))
        assert not np.any(np.isfinite(a) & ~np.isfinite(b)), name


def test_profile_kernels_after_baum_welch(models):
    """Training changes every weight and turns the point-mass inserts into Normals (min_std clamp, yahmm.pyx:513): the
    plan is rebuilt, the profile kernels keep serving the model and keep agreeing with the run-based kernels."""
    model = util.build_model("pro_001_py2")
    eng = model.engine()
    lens = events.uniform_lengths(400, 30, 200, seed=31)
    obs, offsets = eng.sample_events(lens, seed=32)
    seqs = [obs[offsets[i]:offsets[i + 1]].astype(np.float64) for i in range(len(lens))]
    model.train(seqs, max_iterations=1, verbose=False)
    eng = model.engine()
    assert eng.kernel_info()["family"] == "profile"
    s0, l0, p0, f0 = eng.score_batch(obs, offsets)
    eng.set_kernel_path(1)
    try:
        s1, l1, p1, f1 = eng.score_batch(obs, offsets)
    finally:
        eng.set_kernel_path(0)
    assert np.array_equal(s0, s1) and np.array_equal(l0, l1) and np.array_equal(p0, p1)
    assert np.allclose(f0, f1, rtol=1e-9, atol=1e-9)
    orc = util.oracle_for(model)
    for e in range(0, len(lens), 40):
        s, p, margin = orc.viterbi(seqs[e], with_margin=True)
        assert s0[e] == s


def test_uniform_range_at_a_tiny_posterior_weight():
    """UniformDistribution.summarize (yahmm.pyx:304-326) keeps an event's observed range whenever the state's posterior
    weight is not exactly zero.  Here it is ~1e-203 per row: the compiled reference re-estimates the range from every event
    (tests/golden/make_golden_tiny_weight.py); the scaled linear-space statistics kernel must decide "non-zero" from the
    weight itself, not from a product that underflowed on the way."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("mk_tiny", os.path.join(util.GOLDEN, "make_golden_tiny_weight.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import protpore_b200.yahmm as y
    g = np.load(os.path.join(util.GOLDEN, "tiny_weight.npz"))
    model = mod.tiny_weight(y)
    evs = [g["obs"][g["offsets"][i]:g["offsets"][i + 1]] for i in range(len(g["offsets"]) - 1)]
    imp = model.train(evs, max_iterations=1, verbose=False)
    assert model.engine().forward_backward_info()["path"] in ("lane", "profile")
    params = np.array([[float(p) for p in s.distribution.parameters[:2]] for s in model.states[:model.silent_start]])
    assert [s.name for s in model.states[:model.silent_start]] == list(g["names"])
    iu = list(g["names"]).index("u")
    assert tuple(params[iu]) == tuple(g["params"][iu]) == (g["obs"].min(), g["obs"].max())
    assert np.allclose(params, g["params"], rtol=1e-7, atol=1e-9)
    assert imp == pytest.approx(float(g["improvement"]), rel=1e-7, abs=1e-7)
    assert np.allclose(model.dense_transition_matrix(), g["dense"], rtol=1e-8, atol=1e-12, equal_nan=True)


@pytest.mark.gpu
def test_guard_mode_catches_writes_outside_a_buffer():
    """PP_GUARD=1 (tests/conftest.py checks the guards after every GPU test): the check must actually see a write one byte
    behind and a write in front of a buffer.  Without PP_GUARD the entry point says so and the test is skipped."""
    lib = _engine.load_library()
    if os.environ.get("PP_GUARD", "0") in ("", "0"):
        assert lib.pp_debug_check_guards() == -1
        pytest.skip("guard mode off")
    assert lib.pp_debug_guard_selftest() == 0, lib.pp_last_error().decode()
