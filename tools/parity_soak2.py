"""Second parity soak (GPU box): (a) the global-column lane kernels on synthetic profiles of several sizes against the
C oracle; (b) every random model of the first soak's generator against its TWIN whose emissions go through the
emission-table path (state-parallel kernels) -- two independent CUDA implementations of the same sweeps.
usage: python tools/parity_soak2.py [n_twin_models]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import test_gpu_parity as T
    import util
    import protpore_b200.yahmm as y
    from protpore_b200 import peptide_hmm, events
    n_twins = int(sys.argv[1]) if len(sys.argv) > 1 else 400

    # (a) large profiles: column in global memory (rows >= ~150) or shared memory (smaller)
    t0 = time.time()
    st = dict(models=0, events=0, score_mismatch=0, path_mismatch=0, exact_ties=0, logp_out_of_tolerance=0)
    for rows, py2, seed in ((100, True, 1), (160, True, 2), (200, False, 3), (250, True, 4), (334, False, 5), (400, True, 6)):
        model = peptide_hmm.model_maker(peptide_hmm.synthetic_profile(rows, seed=seed), "synth%d" % rows, y, py2)
        orc = util.oracle_for(model)
        rng = np.random.RandomState(rows)
        lens = rng.randint(1, 500, size=40)
        evs = [np.asarray(x, dtype=np.float64) for x in events.sample_events(model, lens, seed=seed + 10)]
        evs[3] = evs[3] + rng.normal(0, 3.0, size=len(evs[3]))                  # an off-model event
        res = model.viterbi_batch(evs, dtype=np.float64, states=False)
        lp = model.log_probability_batch(evs, dtype=np.float64)
        for x, (score, path), l in zip(evs, res, lp):
            s, p, margin = orc.viterbi(x, with_margin=True)
            st["events"] += 1
            st["score_mismatch"] += int(score != s)
            if (p is None) != (path is None):
                st["path_mismatch"] += 1
            elif p is not None and not np.array_equal(path, p):
                st["exact_ties" if margin == 0.0 else "path_mismatch"] += 1
            st["logp_out_of_tolerance"] += int(not util.close([l], [orc.log_probability(x)]).all())
        st["models"] += 1
        model._drop_engine()
    print("synthetic profiles of 100-400 rows (600-2,400 states):", st, "(%.0f s)" % (time.time() - t0), flush=True)

    # (b) twins through the emission table
    t0 = time.time()
    rng = np.random.RandomState(99)
    st = dict(models=0, events=0, score_mismatch_over_1e9=0, path_mismatch=0, logp_out_of_tolerance=0, impossible_disagree=0)
    for idx in range(n_twins):
        native = T._random_model(rng, idx)
        if native.silent_start == 0:
            continue
        evs = [rng.uniform(-4, 4, size=int(rng.randint(1, 90))) for _ in range(6)]
        # the twin: same graph object re-baked with every distribution wrapped in a LambdaDistribution
        import copy
        twin = copy.deepcopy(native)
        twin._drop_engine() if hasattr(twin, "_drop_engine") else None
        for s in twin.states[:twin.silent_start]:
            d = s.distribution
            s.distribution = y.LambdaDistribution(d.log_probability)
        twin._engine = None
        rn = native.viterbi_batch(evs, dtype=np.float64, states=False)
        rt = twin.viterbi_batch(evs, dtype=np.float64, states=False)
        ln = native.log_probability_batch(evs, dtype=np.float64)
        lt = twin.log_probability_batch(evs, dtype=np.float64)
        for (sn, pn), (stw, pt), a, b in zip(rn, rt, ln, lt):
            st["events"] += 1
            if (pn is None) != (pt is None):
                st["impossible_disagree"] += 1
                continue
            if pn is not None:
                st["score_mismatch_over_1e9"] += int(abs(sn - stw) > 1e-9 * max(1.0, abs(sn)))
                st["path_mismatch"] += int(not np.array_equal(pn, pt))
            st["logp_out_of_tolerance"] += int(not util.close([a], [b]).all())
        st["models"] += 1
        native._drop_engine(); twin._drop_engine()
    print("lane kernels vs emission-table twins (state-parallel kernels):", st, "(%.0f s)" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
