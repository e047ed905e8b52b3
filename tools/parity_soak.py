"""Parity soak (run once per round on the GPU box; output kept under profiles/): many more cases than the test-suite
runs -- random HMMs and sampled pro_001 events, CUDA path against the CPU oracle (oracle/hmm_oracle.c).
usage: python tools/parity_soak.py [n_random_models] [n_pro001_events]"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _oracle_shard(args):
    py2, shard = args
    from protpore_b200 import peptide_hmm
    import protpore_b200.yahmm as y
    from oracle import cpu_oracle
    orc = cpu_oracle.OracleModel.from_model(peptide_hmm.pro_001_model(yahmm_module=y, py2_division=py2))
    out = []
    for x in shard:
        s, p, margin = orc.viterbi(x, with_margin=True)
        out.append((s, p, margin, orc.log_probability(x)))
    return out


def main():
    n_models = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n_events = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
    import test_gpu_parity as T
    import util
    from protpore_b200 import peptide_hmm, events
    t0 = time.time()
    rng = np.random.RandomState(777)
    stats = dict(models=0, events=0, score_mismatch=0, path_mismatch=0, exact_ties=0, logp_out_of_tolerance=0, impossible=0)
    worst = 0.0
    bad_models = set()
    only = set(int(v) for v in os.environ.get("SOAK_ONLY", "").split(",") if v)
    for idx in range(n_models):
        model = T._random_model(rng, idx)
        if model.silent_start == 0:
            continue
        orc = util.oracle_for(model)
        evs = []
        for _ in range(10):
            x = rng.uniform(-4, 4, size=int(rng.randint(1, 120)))
            evs.append(np.round(x, 1) if rng.rand() < 0.4 else x)
        if only and idx not in only:
            continue
        res = model.viterbi_batch(evs, dtype=np.float64, states=False)
        lp = model.log_probability_batch(evs, dtype=np.float64)
        for x, (score, path), l in zip(evs, res, lp):
            s, p, margin = orc.viterbi(x, with_margin=True)
            stats["events"] += 1
            stats["score_mismatch"] += int(score != s)
            if score != s and idx not in bad_models:
                bad_models.add(idx)
                from protpore_b200 import _engine
                kinds = model._device_params()
                print("MISMATCH model", idx, "states", len(model.states), "ne", model.silent_start, "finite", model._baked.finite,
                      "n", len(x), "gpu", score, "oracle", s, "logp gpu", l, "oracle", orc.log_probability(x), flush=True)
                print(_engine.describe_schedule(model).split("backward")[0], flush=True)
                print("kinds", kinds[0], "p0", kinds[1], "p1", kinds[2], "logw", list(model._baked.state_weights), flush=True)
                print("in_ptr", list(model._baked.in_ptr), "in_idx", list(model._baked.in_idx), flush=True)
            if p is None or path is None:
                stats["impossible"] += 1
                stats["path_mismatch"] += int((p is None) != (path is None))
            elif not np.array_equal(path, p):
                if margin == 0.0:
                    stats["exact_ties"] += 1
                else:
                    stats["path_mismatch"] += 1
            ol = orc.log_probability(x)
            if np.isfinite(ol):
                worst = max(worst, abs(l - ol) / max(1.0, abs(ol)))
            stats["logp_out_of_tolerance"] += int(not util.close([l], [ol]).all())
        # Baum-Welch sufficient statistics (k_fb_lane where the backward lowering applies, else the exact kernel)
        eng = model.engine()
        from protpore_b200 import _engine as _E
        obs_p, off_p = _E.pack_events(evs, dtype=np.float64)
        lp_s, edge_s, mom_s, mm_s, _ = eng.forward_backward_batch(obs_p, off_p)
        stats["fb_" + eng.forward_backward_info()["path"]] = stats.get("fb_" + eng.forward_backward_info()["path"], 0) + 1
        exp_t, mom_o, mm_o, lz = orc.bw_accumulate(evs)
        b = model._baked
        src = np.repeat(np.arange(len(model.states)), np.diff(b.out_ptr))
        uni = np.array(model._device_params()[0]) == 1           # the observed range only feeds Uniform re-estimation (304-348)
        ok = (np.allclose(edge_s, exp_t[src, b.out_idx], rtol=1e-8, atol=1e-10) and np.allclose(mom_s, mom_o, rtol=1e-8, atol=1e-9)
              and np.array_equal(mm_s[uni], mm_o[uni]) and np.array_equal(np.isfinite(lp_s), np.isfinite(lz))
              and np.allclose(lp_s[np.isfinite(lz)], lz[np.isfinite(lz)], rtol=1e-10, atol=1e-8))
        if not ok:
            stats["fb_mismatch"] = stats.get("fb_mismatch", 0) + 1
            if stats["fb_mismatch"] <= 3:
                ref_e = exp_t[src, b.out_idx]
                print("FB MISMATCH model", idx, "path", eng.forward_backward_info(), "states", len(model.states), "ne", model.silent_start,
                      "nan ours/oracle edge", int(np.isnan(edge_s).sum()), int(np.isnan(ref_e).sum()),
                      "mom", int(np.isnan(mom_s).sum()), int(np.isnan(mom_o).sum()),
                      "edge ok", bool(np.allclose(edge_s, ref_e, rtol=1e-8, atol=1e-10)),
                      "mom ok", bool(np.allclose(mom_s, mom_o, rtol=1e-8, atol=1e-9)), "mm ok", bool(np.array_equal(mm_s, mm_o)),
                      "lp", list(np.round(lp_s, 6)), "lz", list(np.round(lz, 6)), flush=True)
                print("  mom ours", mom_s.tolist(), "oracle", mom_o.tolist(), "mm ours", mm_s.tolist(), "oracle", mm_o.tolist(), flush=True)
                print("  kinds", model._device_params(), "lens", [len(e) for e in evs], flush=True)
        stats["models"] += 1
        model._drop_engine()
    print("random models:", stats, "worst relative logp error %.3g" % worst, "(%.0f s)" % (time.time() - t0), flush=True)

    for py2 in ((True, False) if n_events > 0 else ()):
        t0 = time.time()
        model = peptide_hmm.pro_001_model(py2_division=py2)
        eng = model.engine()
        lens = events.uniform_lengths(n_events, 20, 600, seed=31 + py2)
        obs, off = eng.sample_events(lens, seed=32 + py2)
        scores, plen, paths, logp = eng.score_batch(obs, off)            # both sweeps in one call (round 2)
        family = eng.kernel_info()["family"]
        poff = np.concatenate([[0], np.cumsum(np.maximum(plen, 0))])
        xs = [obs[off[e]:off[e + 1]].astype(np.float64) for e in range(n_events)]
        cores = os.cpu_count() or 1
        with mp.get_context("fork").Pool(cores) as pool:
            parts = pool.map(_oracle_shard, [(py2, xs[w::cores]) for w in range(cores)])
        st = dict(events=n_events, score_mismatch=0, path_mismatch=0, exact_ties=0, logp_out_of_tolerance=0)
        worst = 0.0
        for w, part in enumerate(parts):
            for k, (s, p, margin, ol) in enumerate(part):
                e = w + k * cores
                st["score_mismatch"] += int(scores[e] != s)
                if not np.array_equal(paths[poff[e]:poff[e + 1]], p):
                    st["exact_ties" if margin == 0.0 else "path_mismatch"] += 1
                worst = max(worst, abs(logp[e] - ol) / max(1.0, abs(ol)))
                st["logp_out_of_tolerance"] += int(not util.close([logp[e]], [ol]).all())
        print("pro_001 py2_division=%s (%s kernels):" % (py2, family), st, "worst relative logp error %.3g" % worst,
              "(%.0f s, %d cores)" % (time.time() - t0, cores), flush=True)
        # Baum-Welch E-step (round 2: k_profile_fb): a 30,000-event subset against the exact state-parallel kernel,
        # 480 of them against the oracle
        t0 = time.time()
        n_fb = min(n_events, 30000)
        o_fb, f_fb = obs[:off[n_fb]].astype(np.float64), off[:n_fb + 1]
        lp0, e0, m0, mm0, _ = eng.forward_backward_batch(o_fb, f_fb)
        info = eng.forward_backward_info()
        eng.set_kernel_path(2)
        lpx, ex, mx, mmx, _ = eng.forward_backward_batch(o_fb, f_fb)
        eng.set_kernel_path(0)
        fin = np.isfinite(lpx)
        fb = dict(events=n_fb, path=info["path"], rejects=info["rejects"], failed=info["failed"],
                  logp_ok=bool(np.array_equal(fin, np.isfinite(lp0)) and np.allclose(lp0[fin], lpx[fin], rtol=1e-12, atol=1e-9)),
                  edge_max_rel=float(np.max(np.abs(e0 - ex) / np.maximum(1e-3, np.abs(ex)))),
                  moment_max_rel=float(np.max(np.abs(m0 - mx) / np.maximum(1e-3, np.abs(mx)))), ranges_equal=bool(np.array_equal(mm0, mmx)))
        sub = list(range(0, n_fb, max(1, n_fb // 480)))[:480]
        seqs = [xs[e] for e in sub]
        so, oo = np.concatenate(seqs), np.concatenate([[0], np.cumsum([len(x) for x in seqs])])
        _, e_s, m_s, mm_s, _ = eng.forward_backward_batch(so, oo)
        from oracle import cpu_oracle
        orc = cpu_oracle.OracleModel.from_model(model)
        exp_t, mom_o, mm_o, lz = orc.bw_accumulate(seqs)
        b = model._baked
        src = np.repeat(np.arange(len(model.states)), np.diff(b.out_ptr))
        fb["oracle_edges_ok"] = bool(np.allclose(e_s, exp_t[src, b.out_idx], rtol=1e-9, atol=1e-12))
        fb["oracle_moments_ok"] = bool(np.allclose(m_s, mom_o, rtol=1e-9, atol=1e-11))
        fb["oracle_ranges_equal"] = bool(np.array_equal(mm_s, mm_o))
        print("pro_001 py2_division=%s E-step:" % py2, fb, "(%.0f s)" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
