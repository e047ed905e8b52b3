"""The kernels' exact Normal log-density (pp_device_common.cuh emission_log, pp_profile_kernels.cuh prof_log_normal).

NormalDistribution._log_probability is  _log(1 / (sigma * 2.50662827463)) - (x - mu) ** 2 / (2 * sigma ** 2)
(yahmm.pyx:389-390): an IEEE float64 division per emission.  The kernels evaluate the quotient a / b as a Markstein
sequence with y = RN(1 / b):  q0 = RN(a y), r = fma(-q0, b, a), q = fma(r, y, q0)  -- three flops instead of a ~30
instruction divide -- which returns the correctly rounded quotient.  This test pins that claim on the GPU: a model with
ONE Normal state scores a one-observation event as exactly its log-density (0 + 0 + e + 0), so a batch of single-
observation events evaluates the device function on millions of operands, compared bit for bit with numpy's division.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SQRT_2_PI = 2.50662827463          # the reference's truncated constant (yahmm.pyx:33)


def _one_state_model(mu, sigma):
    import protpore_b200.yahmm as y
    m = y.Model(name="one")
    s = y.State(y.NormalDistribution(mu, sigma), name="s")
    m.add_state(s)
    m.add_transition(m.start, s, 1.0)
    m.add_transition(s, m.end, 1.0)
    m.bake()
    return m


def _reference_logpdf(x, mu, sigma):
    import math
    with np.errstate(divide="ignore", over="ignore", under="ignore"):
        c = math.log(1.0 / (sigma * SQRT_2_PI))     # libm, as the reference's clog (numpy's vectorised log is not the same function)
        d = x - mu
        return c - (d * d) / (2.0 * sigma ** 2)


@pytest.mark.parametrize("mu,sigma", [(36.7, 1.37), (27.6, 0.37), (46.8, 2.54), (0.0, 1.0), (1e-3, 0.01), (-12.5, 17.25),
                                      (3.0, 1e-8), (40.0, 3e5)])
def test_normal_log_density_is_bit_identical_to_ieee_division(mu, sigma):
    rs = np.random.RandomState(int(abs(mu) * 1000) % 2 ** 31 + 1)
    n = 400000
    x = np.concatenate([
        rs.normal(mu, sigma, size=n),                               # where the mass is
        rs.uniform(mu - 40 * sigma, mu + 40 * sigma, size=n),       # far tails
        mu + rs.choice([-1.0, 1.0], size=n // 4) * 10.0 ** rs.uniform(-170, -1, size=n // 4),   # (x - mu)^2 subnormal or 0
        mu + rs.choice([-1.0, 1.0], size=n // 4) * 10.0 ** rs.uniform(0, 150, size=n // 4),      # huge deviations
        np.array([mu, np.nextafter(mu, np.inf), np.nextafter(mu, -np.inf), mu + sigma, mu - sigma]),
    ])
    model = _one_state_model(mu, sigma)
    eng = model.engine()
    offsets = np.arange(len(x) + 1, dtype=np.int64)
    # the run-based kernels (a one-state model is no linear profile) ...
    scores, plen, _ = eng.viterbi_batch(x, offsets, with_paths=False)
    want = _reference_logpdf(x, mu, sigma)
    want = np.where(np.isnan(want), -np.inf, want)
    ok = np.isfinite(want) | np.isneginf(want)
    assert ok.all()
    same = (scores == want) | (np.isneginf(scores) & np.isneginf(want))
    assert same.all(), (x[~same][:5], scores[~same][:5], want[~same][:5])


def test_profile_kernels_use_the_same_density():
    """The register-resident profile kernels carry their own copy of the sequence (prof_log_normal): a profile scored with
    them and with the run-based kernels gives bit-identical Viterbi scores on observations that exercise tiny and huge
    deviations from every match state's mean."""
    import util
    from protpore_b200 import events
    model = util.build_model("pro_001_py2")
    eng = model.engine()
    lens = events.uniform_lengths(4000, 1, 60, seed=8)
    obs, offsets = eng.sample_events(lens, seed=9)
    rs = np.random.RandomState(1)
    obs = obs.astype(np.float64)
    k = rs.choice(len(obs), size=len(obs) // 3, replace=False)
    obs[k] += rs.choice([-1.0, 1.0], size=len(k)) * 10.0 ** rs.uniform(-160, 2, size=len(k))
    s0, l0, p0 = eng.viterbi_batch(obs, offsets)
    eng.set_kernel_path(1)
    try:
        s1, l1, p1 = eng.viterbi_batch(obs, offsets)
    finally:
        eng.set_kernel_path(0)
    assert np.array_equal(s0, s1) and np.array_equal(l0, l1) and np.array_equal(p0, p1)

