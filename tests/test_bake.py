"""`Model.bake` parity: our insertion-ordered restatement of yahmm.pyx:2280-2655 against the baked
structure the compiled reference produced for the same construction sequence (golden fixtures)."""
import numpy as np
import pytest

import util


@pytest.mark.parametrize("name", util.MODEL_NAMES)
def test_bake_matches_reference(name):
    g = util.load_golden(name)
    model = util.build_model(name)
    m, ss, si, ei, ne, finite = [int(v) for v in g["meta"]]
    assert len(model.states) == m and model.silent_start == ss
    assert model.start_index == si and model.end_index == ei
    assert model.edge_count() == ne
    assert (0 if model.is_infinite() else 1) == finite
    assert [s.name for s in model.states] == list(g["state_names"])
    assert np.array_equal(model.dense_transition_matrix(), g["dense"])      # bit-identical log weights


def test_pro_001_shape():
    model = util.build_model("pro_001_py2")
    b = model._baked
    assert (len(model.states), model.silent_start, model.edge_count()) == (257, 128, 765)
    assert (model.start_index, model.end_index) == (212, 256)
    indeg = np.diff(b.in_ptr)
    assert indeg[model.end_index] == 87 and indeg[:128].max() == 7
    # the Python-2 division trap: every insert state is a point mass (peptide_hmm_v0.0.1.py:194-195)
    ins = [s for s in model.states[:128] if s.name.startswith("I_INS")]
    assert len(ins) == 42 and all(s.distribution.parameters[1] == 0.0 for s in ins)
    ins3 = [s for s in util.build_model("pro_001_py3").states[:128] if s.name.startswith("I_INS")]
    assert all(s.distribution.parameters[1] > 0.0 for s in ins3)


def test_in_edge_order_is_source_insertion_order():
    model = util.build_model("pro_001_py2")
    b = model._baked
    k = [s.name for s in model.states].index("M_5_5")
    names = [model.states[i].name for i in b.in_idx[b.in_ptr[k]:b.in_ptr[k + 1]]]
    assert names == ["M_4_4", "S_SKIP_4_4", "M_5_5", "S_DROPOFF_5_5", "I_BLIP_5_5", "I_INS_5_5", "B_BACK_SHORT_6_6"]


def test_model_text_roundtrip(tmp_path):
    import protpore_b200.yahmm as y
    import io
    from contextlib import redirect_stdout
    model = util.build_model("toy")
    buf = io.StringIO()
    with redirect_stdout(io.StringIO()):
        model.write(buf)
    buf.seek(0)
    again = y.Model.read(buf)
    assert sorted(s.name for s in again.states) == sorted(s.name for s in model.states)
    a = {(model.states[i].name, model.states[j].name): v for (i, j), v in np.ndenumerate(model.dense_transition_matrix())}
    b = {(again.states[i].name, again.states[j].name): v for (i, j), v in np.ndenumerate(again.dense_transition_matrix())}
    for key, v in a.items():
        assert b[key] == pytest.approx(v, abs=1e-12) or (np.isinf(v) and np.isinf(b[key]))


@pytest.mark.reference
def test_from_matrix_matches_reference():
    """Model.from_matrix (yahmm.pyx:3884-3936), including its quirk of wiring every `ends` entry to the LAST state."""
    import numpy as np
    import protpore_b200.yahmm as ours
    from oracle import ref_loader
    ref = ref_loader.load_reference_yahmm()
    T = np.array([[0.5, 0.3, 0.2], [0.1, 0.6, 0.3], [0.2, 0.2, 0.6]]) * 0.9

    def make(mod):
        d = [mod.NormalDistribution(0, 1), mod.NormalDistribution(3, 1), mod.UniformDistribution(0, 5)]
        return mod.Model.from_matrix(T, d, starts=[0.5, 0.25, 0.25], ends=[0.1, 0.1, 0.1], state_names=["a", "b", "c"], name="fm")

    a, b = make(ours), make(ref)
    assert [s.name for s in a.states] == [s.name for s in b.states]
    assert np.array_equal(a.dense_transition_matrix(), np.asarray(b.dense_transition_matrix()))
    assert (a.start_index, a.end_index, a.silent_start) == (b.start_index, b.end_index, b.silent_start)


def test_profile_plan_recognises_linear_profiles_only():
    """pp_profile_plan.cpp classifies states from the baked graph alone: the peptide profiles take the register-resident
    kernels, everything else keeps the run-based ones (the text report needs no GPU)."""
    from protpore_b200 import _engine
    def line(name, prefix):
        hits = [l for l in _engine.describe_schedule(util.build_model(name)).splitlines() if l.startswith(prefix)]
        assert len(hits) == 1, (name, prefix, hits)
        return hits[0]
    for name in ("pro_001_py2", "pro_001_py3", "synth_12"):
        assert line(name, "profile ").startswith("profile ok"), name
        # the forward-backward kernel of the profile: every out-edge of the model is counted by exactly one of its terms
        assert line(name, "profile_fb ").startswith("profile_fb ok"), (name, line(name, "profile_fb "))
    last = line("pro_001_py2", "profile ")
    assert "positions 43 " in last and "end_edges 87" in last and "aliases 1" in last
    for name in ("toy", "toy_uniform", "toy_infinite"):
        last = _engine.describe_schedule(util.build_model(name)).splitlines()[-1]
        assert last.startswith("profile no:"), (name, last)


def test_profile_plans_over_profile_lengths():
    """Every profile length the kernels serve (1 warp ... 11 warps, lengths that do and do not fill the last warp), both division
    variants: the plan is recognised, and the forward-backward plan counts every out-edge exactly once (CPU: text report)."""
    from protpore_b200 import _engine, peptide_hmm
    import protpore_b200.yahmm as y
    for K in (2, 3, 4, 5, 8, 9, 17, 31, 40, 43, 44):
        for py2 in (True, False):
            model = peptide_hmm.model_maker(peptide_hmm.synthetic_profile(K, seed=K), "plan_%d" % K, y, py2)
            text = _engine.describe_schedule(model)
            assert "profile ok positions %d " % K in text, (K, py2, text.splitlines()[-2:])
            assert "profile_fb ok" in text, (K, py2, text.splitlines()[-2:])
    # one position more than the kernels hold in registers: the run-based kernels keep the model
    text = _engine.describe_schedule(peptide_hmm.model_maker(peptide_hmm.synthetic_profile(45, seed=1), "plan_45", y, True))
    assert "profile no:" in text
