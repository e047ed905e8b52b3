"""Soak of the linear-profile kernel family (run on the GPU box; output kept under profiles/): synthetic profiles of many
lengths, both division variants, before and after one Baum-Welch iteration (all weights change, point-mass inserts become
Normals) -- k_profile_viterbi / k_profile_forward against the run-based lane kernels (bit-identical scores, identical
paths), k_profile_fb against the exact state-parallel kernel (counts, moments, ranges), batched posteriors against it.
usage: python tools/profile_soak.py [seeds per length]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from protpore_b200 import peptide_hmm, events, _engine  # noqa: E402
import protpore_b200.yahmm as y  # noqa: E402


def compare(model, tag, rng, st):
    eng = model.engine()
    if eng.kernel_info()["family"] != "profile" or "profile_fb ok" not in _engine.describe_schedule(model):
        st["not_profile"].append(tag)
        return
    lens = events.log_uniform_lengths(320, 1, 300, seed=int(rng.randint(1 << 30)))
    obs, off = eng.sample_events(lens, seed=int(rng.randint(1 << 30)))
    s0, l0, p0, f0 = eng.score_batch(obs, off)
    eng.set_kernel_path(1)
    s1, l1, p1, f1 = eng.score_batch(obs, off)
    eng.set_kernel_path(0)
    st["events"] += len(lens)
    st["score_mismatch"] += int(np.sum(s0 != s1))
    st["path_mismatch"] += int(not (np.array_equal(l0, l1) and np.array_equal(p0, p1)))
    fin = np.isfinite(f1)
    if not np.array_equal(fin, np.isfinite(f0)) or np.max(np.abs(f0[fin] - f1[fin]) / np.maximum(1.0, np.abs(f1[fin])), initial=0.0) > 1e-7:
        st["forward_mismatch"].append(tag)
    o64 = obs.astype(np.float64)
    lp0, e0, m0, mm0, _ = eng.forward_backward_batch(o64, off)
    info = eng.forward_backward_info()
    st["fb_" + info["path"]] = st.get("fb_" + info["path"], 0) + 1
    st["fb_rejects"] += info["rejects"]
    eng.set_kernel_path(3)
    _, _, _, _, post0 = eng.forward_backward_batch(o64, off, posteriors=True)
    eng.set_kernel_path(2)
    lpx, ex, mx, mmx, postx = eng.forward_backward_batch(o64, off, posteriors=True)
    eng.set_kernel_path(0)
    finx = np.isfinite(lpx)
    ok = (np.array_equal(finx, np.isfinite(lp0)) and np.allclose(lp0[finx], lpx[finx], rtol=1e-12, atol=1e-9) and
          np.allclose(e0, ex, rtol=1e-9, atol=1e-12) and np.allclose(m0, mx, rtol=1e-9, atol=1e-11) and np.array_equal(mm0, mmx))
    if not ok:
        st["fb_mismatch"].append((tag, info))
    st["fb_edge_worst"] = max(st["fb_edge_worst"], float(np.max(np.abs(e0 - ex) / np.maximum(1e-3, np.abs(ex)))))
    rows = np.repeat(finx, np.diff(off))
    a, b = post0[rows], postx[rows]
    big = np.isfinite(b) & (b > -250.0)
    if big.any():
        if not np.all(np.isfinite(a[big])) or np.max(np.abs(a[big] - b[big])) > 1e-8 or np.any(np.isfinite(a) & ~np.isfinite(b)):
            st["posterior_mismatch"].append(tag)
        else:
            st["posterior_worst"] = max(st["posterior_worst"], float(np.max(np.abs(a[big] - b[big]))))


def main():
    seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    rng = np.random.RandomState(2026)
    st = dict(models=0, events=0, score_mismatch=0, path_mismatch=0, forward_mismatch=[], fb_mismatch=[], fb_rejects=0, fb_edge_worst=0.0,
              posterior_mismatch=[], posterior_worst=0.0, not_profile=[])
    t0 = time.time()
    for K in (2, 3, 4, 5, 7, 8, 9, 12, 13, 16, 17, 23, 31, 32, 33, 40, 41, 42, 43, 44):
        for seed in range(seeds):
            for py2 in (True, False):
                model = peptide_hmm.model_maker(peptide_hmm.synthetic_profile(K, seed=100 * K + seed), "soak_%d_%d" % (K, seed), y, py2)
                tag = (K, seed, py2)
                compare(model, tag + ("built",), rng, st)
                # one Baum-Welch iteration on sampled events: every weight changes, the inserts become Normals
                eng = model.engine()
                lens = events.uniform_lengths(200, 5, 120, seed=int(rng.randint(1 << 30)))
                obs, off = eng.sample_events(lens, seed=int(rng.randint(1 << 30)))
                seqs = [obs[off[e]:off[e + 1]].astype(np.float64) for e in range(len(lens))]
                model.train(seqs, max_iterations=1, verbose=False)
                compare(model, tag + ("trained",), rng, st)
                st["models"] += 2
                model._drop_engine()
        print("K", K, {k: v for k, v in st.items() if not isinstance(v, list) or v}, "(%.0f s)" % (time.time() - t0), flush=True)
    print("profile soak:", st, "(%.0f s)" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
