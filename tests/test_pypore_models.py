"""Models built by the reference's PyPore/hmm.py and the model text format (8f-3).

tests/golden/make_golden_pypore.py ran the reference's own builders on the compiled reference (results) and on this
repository's yahmm (a replay log of the graph, insertion order included); here the models are rebuilt from the log --
/root/reference does not exist on the GPU box -- and compared with what the compiled reference returned."""
import io
import json
import os
from contextlib import redirect_stdout

import numpy as np
import pytest

import util

CASES = ["phi29_module", "nanopore_fork", "global_kde", "hel308", "phi29_profile"]


def _load(name):
    g = np.load(os.path.join(util.GOLDEN, "pypore_%s.npz" % name))
    return g, util.replay_model(json.loads(str(g["replay"])))


@pytest.mark.parametrize("name", CASES)
def test_replayed_model_bakes_like_the_reference(name):
    g, model = _load(name)
    assert [s.name for s in model.states] == list(g["state_names"])
    assert model.silent_start == int(g["silent_start"])
    assert np.array_equal(np.asarray(model.dense_transition_matrix()), g["dense"])


@pytest.mark.parametrize("name", ["phi29_module", "nanopore_fork", "hel308", "phi29_profile"])
def test_oracle_decodes_pypore_models_like_the_reference(name):
    """The C restatement against the compiled reference on these topologies (tied states, ports merged by bake)."""
    g, model = _load(name)
    orc = util.oracle_for(model)
    off = g["offsets"]
    for i in range(len(off) - 1):
        x = g["obs"][off[i]:off[i + 1]]
        s, p, margin = orc.viterbi(x, with_margin=True)
        assert s == g["vscore"][i]
        assert np.array_equal(p, g["paths"][g["path_offsets"][i]:g["path_offsets"][i + 1]]) or margin == 0.0
        assert orc.log_probability(x) == g["logp"][i]


def test_model_write_matches_the_reference_byte_for_byte():
    """Model.write (yahmm.pyx:3744-3807) of the pro_001 profile: same bytes as the compiled reference wrote (state
    identities, id()-based on both sides, replaced by first-appearance ordinals)."""
    model = util.build_model("pro_001_py2")
    buf = io.StringIO()
    with redirect_stdout(io.StringIO()):
        model.write(buf)
    want = open(os.path.join(util.GOLDEN, "model_write.txt")).read()
    got = util.normalise_identities(buf.getvalue())
    assert got == want


def test_model_read_of_the_reference_text():
    """Model.read (3809-3877) accepts what the reference wrote and bakes (merge=None) the same transition structure."""
    import protpore_b200.yahmm as y
    text = open(os.path.join(util.GOLDEN, "model_write.txt")).read()
    with redirect_stdout(io.StringIO()):
        model = y.Model.read(io.StringIO(text))
    ref = util.build_model("pro_001_py2")
    assert model.edge_count() >= ref.edge_count() and model.name == ref.name
    a = sorted(s.name for s in ref.states)
    b = sorted(s.name for s in model.states)
    assert set(a) <= set(b)                      # merge=None keeps the silent state `bake` merges away by default


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_decodes_pypore_models_like_the_reference(name):
    """Viterbi scores bit-identical and paths identical to the compiled reference, forward within the north-star
    tolerance, forward-backward transition counts within 1e-8 -- tied states, kernel-density emissions (table path),
    forks, and the silent port states `bake` merges."""
    g, model = _load(name)
    evs = util.golden_events(g)
    res = model.viterbi_batch(evs, dtype=np.float64, states=False)
    for i, (score, path) in enumerate(res):
        assert score == g["vscore"][i] or (name == "global_kde" and abs(score - g["vscore"][i]) <= 1e-9 * abs(score))
        want = g["paths"][g["path_offsets"][i]:g["path_offsets"][i + 1]]
        assert np.array_equal(path, want)
    lp = model.log_probability_batch(evs, dtype=np.float64)
    assert util.close(lp, g["logp"]).all()
    rows = g["fb_weight_rows"]
    w_off = np.concatenate([[0], np.cumsum(rows * int(g["silent_start"]))])
    for i in range(len(rows)):
        with redirect_stdout(io.StringIO()):
            t, w = model.forward_backward(evs[i])
        assert np.allclose(np.asarray(t), g["fb_transitions"][i], rtol=1e-8, atol=1e-10)
        want_w = g["fb_weights"][w_off[i]:w_off[i + 1]].reshape(rows[i], -1)
        fin = np.isfinite(want_w)
        assert np.array_equal(fin, np.isfinite(np.asarray(w)))
        assert np.allclose(np.asarray(w)[fin], want_w[fin], rtol=1e-8, atol=1e-8)
