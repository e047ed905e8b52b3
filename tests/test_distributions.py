"""Training of the emission families that have no fused kernel, against the compiled reference's classes
(CPU, build container only: needs oracle/_ref)."""
import numpy as np
import pytest


@pytest.mark.reference
def test_table_family_training_matches_the_compiled_reference():
    """from_sample and summarize + from_summaries (with inertia) of the families without fused kernels, against the compiled
    reference's classes on the same weighted samples (yahmm.pyx:560-688 LogNormal, 748-842 Exponential): the Baum-Welch
    M-step calls exactly these (4253-4327)."""
    from oracle import ref_loader
    import protpore_b200.yahmm as ours
    ref = ref_loader.load_reference_yahmm()
    rs = np.random.RandomState(7)
    for trial in range(20):
        n = int(rs.randint(1, 40))
        items = rs.lognormal(1.0, 0.6, size=n)
        weights = rs.uniform(0, 1, size=n) * (rs.uniform(size=n) > 0.2)
        if trial == 3:
            weights[:] = 0.0
            weights[0] = 0.7                                  # a single non-zero weight: the std is kept
        inertia = float(rs.choice([0.0, 0.25, 0.9]))
        for cls, params in (("LogNormalDistribution", (0.8, 0.5)), ("ExponentialDistribution", (0.4,))):
            a, b = getattr(ours, cls)(*params), getattr(ref, cls)(*params)
            a.from_sample(items, weights=weights, inertia=inertia)
            b.from_sample(items, weights=weights, inertia=inertia)
            assert np.allclose(a.parameters, b.parameters, rtol=1e-13, atol=0), (cls, trial, a.parameters, b.parameters)
            a, b = getattr(ours, cls)(*params), getattr(ref, cls)(*params)
            for lo in range(0, n, 7):                         # several summaries, merged by from_summaries
                a.summarize(items[lo:lo + 7], weights[lo:lo + 7])
                b.summarize(items[lo:lo + 7], weights[lo:lo + 7])
            a.from_summaries(inertia=inertia)
            b.from_summaries(inertia=inertia)
            assert np.allclose(a.parameters, b.parameters, rtol=1e-13, atol=0), (cls, trial, a.parameters, b.parameters)
