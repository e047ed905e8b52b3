"""States whose emission family the kernels do not evaluate themselves: the host fills the emission
matrix column, the sweeps run on the GPU.  Golden vectors come from the compiled, unmodified reference
(tests/golden/make_golden_table_emis.py)."""
import importlib.util
import os

import numpy as np
import pytest

import protpore_b200.yahmm as y
from protpore_b200 import _engine

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("mk_table", os.path.join(HERE, "golden", "make_golden_table_emis.py"))


def _builder(name="build"):
    import ast
    src = open(os.path.join(HERE, "golden", "make_golden_table_emis.py")).read()
    ns = {}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("bump", "build", "build_train"):
            exec(compile(ast.Module([node], []), "make_golden_table_emis.py", "exec"), ns)
    return ns[name]


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "table_emis.npz"))


@pytest.fixture(scope="module")
def model_and_dists():
    return _builder()(y)


def test_distribution_log_probabilities_match_reference(golden, model_and_dists):
    model, dists = model_and_dists
    for i, d in enumerate(dists):
        want = golden["dist%d" % i]
        got = np.array([d.log_probability(float(x)) for x in golden["grid"]])
        many = d.log_probability_many(golden["grid"])
        inf = ~np.isfinite(want)
        assert np.array_equal(np.isfinite(got), ~inf) and np.array_equal(np.isfinite(many), ~inf), d.name
        assert np.allclose(got[~inf], want[~inf], rtol=1e-13, atol=1e-13), d.name
        assert np.allclose(many[~inf], want[~inf], rtol=1e-13, atol=1e-13), d.name


def test_bake_and_lowering_of_table_states(golden, model_and_dists):
    model, dists = model_and_dists
    assert [s.name for s in model.states] == [str(n) for n in golden["state_names"]]
    kinds = model._device_params()[0]
    assert kinds == [2, 2, 2, 2, 2, 0, 2]
    text = _engine.describe_schedule(model)                    # lowers without a GPU
    assert "lane_ok 0" in text


def _events(g):
    off = g["offsets"]
    return [g["obs"][off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.mark.gpu
def test_gpu_sweeps_match_reference(golden, model_and_dists):
    model, _ = model_and_dists
    evs = _events(golden)
    lp = model.log_probability_batch(evs, dtype=np.float64)
    assert np.allclose(lp, golden["logp"], rtol=1e-10, atol=1e-10)
    res = model.viterbi_batch(evs, dtype=np.float64, states=False)
    poff = golden["path_offsets"]
    for e, (score, path) in enumerate(res):
        assert score == pytest.approx(golden["vscore"][e], rel=1e-12, abs=1e-12)
        assert np.array_equal(path, golden["paths"][poff[e]:poff[e + 1]])
    # single-sequence API
    assert model.log_probability(evs[1]) == pytest.approx(golden["logp"][1], rel=1e-10)
    s, p = model.viterbi(evs[0])
    assert [k for k, _ in p] == list(golden["paths"][poff[0]:poff[1]])
    f, b = model.forward(evs[0]), model.backward(evs[0])
    for got, want in ((f, golden["forward0"]), (b, golden["backward0"])):
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)
        assert np.allclose(got[fin], want[fin], rtol=1e-10, atol=1e-10)
    for j in (0, 1):
        t, w = model.forward_backward(evs[j])
        assert np.allclose(t, golden["fb_trans%d" % j], rtol=1e-8, atol=1e-10)
        fin = np.isfinite(golden["fb_emis%d" % j])
        assert np.allclose(np.asarray(w)[fin], golden["fb_emis%d" % j][fin], rtol=1e-8, atol=1e-8)


@pytest.mark.gpu
def test_gpu_table_state_equals_native_state():
    """A Normal density supplied as a LambdaDistribution (table path) decodes like the native Normal state."""
    def build(use_lambda):
        m = y.Model(name="t")
        mk = (lambda mu, sd: y.LambdaDistribution(y.NormalDistribution(mu, sd).log_probability)) if use_lambda else y.NormalDistribution
        a, b = y.State(mk(0.0, 1.0), name="a"), y.State(mk(5.0, 1.5), name="b")
        m.add_state(a); m.add_state(b)
        m.add_transition(m.start, a, 0.5); m.add_transition(m.start, b, 0.5)
        m.add_transition(a, a, 0.4); m.add_transition(a, b, 0.5); m.add_transition(a, m.end, 0.1)
        m.add_transition(b, b, 0.4); m.add_transition(b, a, 0.5); m.add_transition(b, m.end, 0.1)
        m.bake()
        return m
    rng = np.random.default_rng(4)
    evs = [rng.normal(2.5, 3.0, size=int(n)) for n in (3, 40, 17, 200)]
    native, table = build(False), build(True)
    rn = native.viterbi_batch(evs, dtype=np.float64, states=False)
    rt = table.viterbi_batch(evs, dtype=np.float64, states=False)
    for (sn, pn), (st, pt) in zip(rn, rt):
        assert st == pytest.approx(sn, rel=1e-12)
        assert np.array_equal(pn, pt)
    assert np.allclose(native.log_probability_batch(evs, dtype=np.float64), table.log_probability_batch(evs, dtype=np.float64),
                       rtol=1e-7, atol=1e-7)


@pytest.mark.gpu
def test_gpu_table_missing_is_an_error(model_and_dists):
    model, _ = model_and_dists
    eng = model.engine()
    eng.set_emission_table(None)
    with pytest.raises(_engine.EngineError):
        eng.forward_batch(np.ones(4), np.array([0, 4]))
    eng.set_emission_table(np.zeros((2, model.silent_start)))
    with pytest.raises(_engine.EngineError):
        eng.viterbi_batch(np.ones(4), np.array([0, 4]))


@pytest.mark.gpu
def test_gpu_baum_welch_with_table_families(golden):
    """One Baum-Welch iteration where two of three states are re-estimated through Distribution.summarize /
    from_sample (yahmm.pyx:4215-4236) from GPU posteriors, the third from GPU moments."""
    model = _builder("build_train")(y)
    off = golden["train_offsets"]
    evs = [golden["train_obs"][off[i]:off[i + 1]] for i in range(len(off) - 1)]
    improvement = model.train(evs, max_iterations=1, verbose=False)
    assert improvement == pytest.approx(float(golden["train_improvement"]), rel=1e-8)
    got = np.array([float(v) for s in model.states[:3] for v in s.distribution.parameters])
    assert np.allclose(got, golden["train_params"], rtol=1e-8, atol=1e-10)
    want = golden["train_dense"]
    fin = np.isfinite(want)
    dense = model.dense_transition_matrix()
    assert np.array_equal(np.isfinite(dense), fin)
    assert np.allclose(dense[fin], want[fin], rtol=1e-8, atol=1e-10)
