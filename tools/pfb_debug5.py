import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from protpore_b200 import peptide_hmm, events
import protpore_b200.yahmm as y
rng = np.random.RandomState(2026)
# replay the soak's random stream up to the failing model
hit = None
for K in (2, 3, 4, 5, 7, 8, 9, 12):
    for seed in range(2):
        for py2 in (True, False):
            model = peptide_hmm.model_maker(peptide_hmm.synthetic_profile(K, seed=100 * K + seed), "soak", y, py2)
            eng = model.engine()
            for stage in ("built", "trained"):
                a1, a2 = int(rng.randint(1 << 30)), int(rng.randint(1 << 30))
                if (K, seed, py2, stage) == (12, 0, False, "built"):
                    lens = events.log_uniform_lengths(320, 1, 300, seed=a1)
                    obs, off = eng.sample_events(lens, seed=a2)
                    o64 = obs.astype(np.float64)
                    eng.set_kernel_path(3); lp0, _, _, _, p0 = eng.forward_backward_batch(o64, off, posteriors=True)
                    eng.set_kernel_path(2); lpx, _, _, _, px = eng.forward_backward_batch(o64, off, posteriors=True)
                    rows = np.repeat(np.isfinite(lpx), np.diff(off))
                    ev_of_row = np.repeat(np.arange(len(lens)), np.diff(off))
                    big = rows[:, None] & np.isfinite(px) & (px > -300)
                    d = np.where(big, np.abs(np.where(np.isfinite(p0), p0, -1e9) - px), 0)
                    idx = np.argwhere(d > 1e-8)
                    print("bad entries", len(idx))
                    for r, k in idx[:12]:
                        e = ev_of_row[r]
                        print("  event", e, "len", lens[e], "row", r - off[e], "state", model.states[k].name, "ours", p0[r, k], "exact", px[r, k], "x", o64[r], "logp", lpx[e])
                    sys.exit(0)
                if stage == "built":
                    b1, b2 = int(rng.randint(1 << 30)), int(rng.randint(1 << 30))
                    lens = events.uniform_lengths(200, 5, 120, seed=b1)
                    obs, off = eng.sample_events(lens, seed=b2)
                    seqs = [obs[off[e]:off[e + 1]].astype(np.float64) for e in range(len(lens))]
                    model.train(seqs, max_iterations=1, verbose=False)
                    eng = model.engine()
