"""FastStatSplit segmenter: the CPU oracle against golden vectors of the compiled reference, the host
logic, and -- on a GPU -- the CUDA path through the C ABI against both."""
import json
import os

import numpy as np
import pytest

from oracle import statsplit_oracle as so, ref_loader
from protpore_b200 import _engine, parsers

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "statsplit.npz")


def golden_cases():
    g = np.load(GOLDEN)
    for name in g["names"]:
        name = str(name)
        kw = json.loads(str(g[name + "__kw"]))
        yield (name, kw, g[name + "__current"].astype(np.float64), float(g[name + "__min_gain"]),
               g[name + "__starts"], g[name + "__means"], g[name + "__stds"])


def staircase(rng, n_levels, lo, hi):
    return np.concatenate([rng.normal(rng.uniform(25, 50), rng.uniform(0.4, 2.5), size=int(rng.integers(lo, hi)))
                           for _ in range(n_levels)])


CLI = dict(min_width=250, max_width=1000, window_width=2000, min_gain_per_sample=0.05)


# ---------------------------------------------------------------------------------------- CPU: oracle, host logic
def test_oracle_matches_golden_bit_for_bit():
    n = 0
    for name, kw, cur, gain, starts, means, stds in golden_cases():
        gkw = {k: v for k, v in kw.items() if k not in ("min_width", "max_width")}
        assert so.min_gain(**gkw) == gain, name
        bp = so.parse(cur, kw["min_width"], kw["max_width"], kw["window_width"], gain=gain)
        assert np.array_equal(np.concatenate([[0], bp]), starts), name
        assert np.array_equal(so.segment_means(cur, bp), means), name
        n += 1
    assert n == 9


def test_host_threshold_matches_golden():
    """pp_statsplit_min_gain is host arithmetic: callable without a GPU, and bit-identical to the reference."""
    for name, kw, cur, gain, starts, means, stds in golden_cases():
        p = parsers.FastStatSplit(**kw)
        assert p.min_gain == gain, name
    with pytest.raises(AssertionError):
        parsers.FastStatSplit(min_width=100, max_width=50)                     # cparsers.pyx:69-70
    with pytest.raises(AssertionError):
        parsers.SpeedyStatSplit(min_width=100, window_width=150)               # cparsers.pyx:71-72


@pytest.mark.reference
def test_oracle_matches_compiled_reference_on_fresh_events():
    if not ref_loader.cparsers_available():
        pytest.skip("oracle/_ref/cparsers*.so not built")
    cp = ref_loader.load_reference_cparsers()
    rng = np.random.default_rng(5)
    for trial in range(12):
        kw = CLI if trial % 2 == 0 else dict(min_width=60, max_width=3000, window_width=600,
                                             prior_segments_per_second=25., false_positive_rate=10.)
        cur = staircase(rng, int(rng.integers(2, 30)), 80, 1300)
        ref = cp.FastStatSplit(**kw)
        segs = ref.parse(cur)
        bp = so.parse(cur, kw["min_width"], kw["max_width"], kw["window_width"], gain=ref.min_gain)
        assert list(bp) == [s.start for s in segs][1:]
        assert np.array_equal(so.segment_means(cur, bp), [s.mean for s in segs])


@pytest.mark.reference
def test_oracle_matches_compiled_reference_over_random_parameters():
    """400 random parameter sets x 8 signals, including the regime where forced splits leave segments shorter than
    min_width: the C restatement and the compiled reference give the same breakpoints, bit for bit."""
    if not ref_loader.cparsers_available():
        pytest.skip("oracle/_ref/cparsers*.so not built")
    cp = ref_loader.load_reference_cparsers()
    rng = np.random.default_rng(424242)
    checked = short = 0
    for trial in range(400):
        mw = int(rng.integers(1, 300))
        kw = dict(min_width=mw, max_width=int(mw * rng.integers(1, 30) + rng.integers(0, 50)),
                  window_width=int(2 * mw + rng.integers(0, 6 * mw + 20)))
        if trial % 3 == 0:
            kw["min_gain_per_sample"] = float(rng.uniform(0.001, 0.5))
        elif trial % 3 == 1:
            kw.update(prior_segments_per_second=float(rng.uniform(1, 1000)), false_positive_rate=float(rng.uniform(0.1, 100)))
        ref = cp.FastStatSplit(**kw)
        for _ in range(8):
            n = int(rng.integers(1, 6000))
            kind = int(rng.integers(0, 3))
            if kind == 0:
                cur = staircase(rng, max(1, n // 400), 50, 700)[:max(1, n)]
            elif kind == 1:
                cur = 40 + np.cumsum(rng.normal(0, 0.05, size=n)) + rng.normal(0, 0.5, size=n)
            else:
                cur = rng.normal(35, 0.8, size=n)
                cur[rng.integers(0, n, size=max(1, n // 200))] += 30
            want = [s.start for s in ref.parse(cur)][1:]
            got = list(so.parse(cur, kw["min_width"], kw["max_width"], kw["window_width"], gain=ref.min_gain))
            assert got == want, (trial, kw)
            checked += 1
            short += int(len(want) > 0 and min(np.diff([0] + want + [len(cur)])) < kw["min_width"])
    assert checked == 3200 and short > 20


def test_segmenter_fails_loudly_without_gpu():
    if _engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_engine.EngineError):
        parsers.SpeedyStatSplit(**CLI).parse(np.ones(1000))


# ---------------------------------------------------------------------------------------- GPU: parity through the ABI
def _check_event(cur, kw, gain, starts_gpu, means_gpu, stds_gpu, label):
    bp, stats = so.parse(cur, kw["min_width"], kw["max_width"], kw["window_width"], gain=gain, with_stats=True)
    want = np.concatenate([[0], bp])
    if not np.array_equal(starts_gpu, want):
        # the only accepted difference: a window whose two best gains tie within the accuracy of log()
        assert stats[0] < 1e-9, (label, stats[0], starts_gpu, want)
        return False
    edges = np.append(want, len(cur))
    m = np.array([np.mean(cur[a:b]) for a, b in zip(edges[:-1], edges[1:])])
    s = np.array([np.std(cur[a:b]) for a, b in zip(edges[:-1], edges[1:])])
    assert np.allclose(means_gpu, m, rtol=1e-12, atol=0), label
    assert np.allclose(stds_gpu, s, rtol=1e-9, atol=1e-12), label
    return True


@pytest.mark.gpu
def test_gpu_matches_golden():
    for name, kw, cur, gain, starts, means, stds in golden_cases():
        segs = parsers.SpeedyStatSplit(**kw).parse(cur)
        got = np.array([s.start for s in segs])
        assert np.array_equal(got, starts), name
        assert np.allclose([s.mean for s in segs], means, rtol=1e-12, atol=0), name
        assert np.allclose([s.std for s in segs], stds, rtol=1e-9, atol=1e-12), name
        assert [s.end for s in segs] == list(starts[1:]) + [len(cur)]
        assert all(len(s.current) == s.duration for s in segs)


@pytest.mark.gpu
def test_gpu_batch_matches_oracle_ragged():
    rng = np.random.default_rng(11)
    currents = [staircase(rng, int(rng.integers(1, 25)), 60, 1400) for _ in range(70)]
    currents += [np.array([3.0]), rng.normal(40, 1, size=499), rng.normal(40, 1, size=501), rng.normal(40, 1, size=2)]
    for kw in (CLI, dict(min_width=40, max_width=800, window_width=300)):
        p = parsers.FastStatSplit(**kw)
        off, start, mean, std, f32 = p.segment_means_batch(currents, want_std=True, want_f32=True)
        assert off[0] == 0 and len(off) == len(currents) + 1
        same = 0
        for e, cur in enumerate(currents):
            a, b = off[e], off[e + 1]
            same += _check_event(cur, kw, p.min_gain, start[a:b], mean[a:b], std[a:b], (kw["min_width"], e))
        assert same >= len(currents) - 1
        assert np.array_equal(f32, mean.astype(np.float32))
        # the batch call and the one-event call agree exactly
        one = p.parse(currents[3])
        assert [s.start for s in one] == list(start[off[3]:off[4]])


@pytest.mark.gpu
def test_gpu_device_resident_feeds_the_hmm():
    """Device buffers in, device buffers out; seg_offsets / seg_mean_f32 go straight into the batched Viterbi."""
    import torch
    from protpore_b200 import peptide_hmm
    rng = np.random.default_rng(3)
    currents = [staircase(rng, int(rng.integers(3, 12)), 260, 700) for _ in range(40)]
    cur, off = _engine.pack_events(currents, dtype=np.float64)
    p = parsers.SpeedyStatSplit(**CLI)
    h = p.segment_means_batch((cur, off), want_std=False, want_f32=True)
    d = p.segment_means_batch((torch.from_numpy(cur).cuda(), torch.from_numpy(off).cuda()), want_std=False, want_f32=True)
    assert np.array_equal(d[0].cpu().numpy(), h[0]) and np.array_equal(d[1].cpu().numpy(), h[1])
    assert np.array_equal(d[2].cpu().numpy(), h[2]) and np.array_equal(d[4].cpu().numpy(), h[4])
    model = peptide_hmm.pro_001_model()
    eng = model.engine()
    s_dev, l_dev, _ = eng.viterbi_batch(d[4], d[0], with_paths=False)
    s_host, l_host, _ = eng.viterbi_batch(h[4], h[0], with_paths=False)
    assert np.array_equal(np.asarray(s_dev.cpu() if hasattr(s_dev, "cpu") else s_dev), np.asarray(s_host))


@pytest.mark.gpu
def test_gpu_capacity_and_errors():
    rng = np.random.default_rng(2)
    cur, off = _engine.pack_events([staircase(rng, 10, 300, 600) for _ in range(5)], dtype=np.float64)
    p = parsers.FastStatSplit(**CLI)
    full = _engine.statsplit_batch(cur, off, p.min_width, p.max_width, p.window_width, p.min_gain)
    small = _engine.statsplit_batch(cur, off, p.min_width, p.max_width, p.window_width, p.min_gain, capacity=2)   # retried
    assert np.array_equal(full[1], small[1])
    with pytest.raises(_engine.EngineError):
        _engine.statsplit_batch(cur, np.array([0, 0, len(cur)]), p.min_width, p.max_width, p.window_width, p.min_gain)
    with pytest.raises(_engine.EngineError):
        _engine.statsplit_batch(cur, off, 0, 10, 20, 1.0)


@pytest.mark.gpu
def test_gpu_division_by_reciprocal_is_exact():
    """k_statsplit divides by segment lengths through a reciprocal (3 instructions); it must be IEEE division bit for bit."""
    import ctypes
    lib = _engine.load_library()
    rng = np.random.default_rng(9)
    n = 1 << 22
    b = np.concatenate([np.arange(1, 1 << 20), rng.integers(1, 2 ** 31 - 1, size=n - (1 << 20) + 1)]).astype(np.int32)
    mant = rng.integers(0, 1 << 52, size=n, dtype=np.uint64)
    expo = rng.integers(1023 - 40, 1023 + 60, size=n, dtype=np.uint64)
    sign = rng.integers(0, 2, size=n, dtype=np.uint64)
    a = ((sign << np.uint64(63)) | (expo << np.uint64(52)) | mant).view(np.float64)
    a[:4096] = np.repeat(rng.normal(0, 1e6, size=64), 64) * b[:4096]          # exact and near-exact quotients
    a[4096:4100] = [0.0, -0.0, 1.0, 3.0]
    bad = (ctypes.c_int64 * 2)()                      # {mismatches, first mismatching index}; +-0 counts as equal
    lib.pp_debug_div_check.restype = ctypes.c_int
    lib.pp_debug_div_check.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    assert lib.pp_debug_div_check(0, a.ctypes.data, b.ctypes.data, n, bad) == 0
    assert bad[0] == 0, (bad[0], a[bad[1]], b[bad[1]])


@pytest.mark.gpu
def test_gpu_wide_window_takes_the_tableless_kernel():
    """window_width > 4096: reciprocals per position instead of the shared-memory table; same breakpoints."""
    rng = np.random.default_rng(21)
    currents = [staircase(rng, 12, 900, 4000) for _ in range(6)]
    kw = dict(min_width=100, max_width=1000000, window_width=10000)                 # the reference's defaults
    p = parsers.FastStatSplit(**kw)
    off, start, mean, std, _ = p.segment_means_batch(currents, want_std=True)
    for e, cur in enumerate(currents):
        assert _check_event(cur, kw, p.min_gain, start[off[e]:off[e + 1]], mean[off[e]:off[e + 1]], std[off[e]:off[e + 1]], e)


@pytest.mark.gpu
def test_gpu_segment_and_decode_pipeline():
    """pypore.segment_and_decode == the reference's per-event loop: parse, segment means, Viterbi (float32 means)."""
    from protpore_b200 import peptide_hmm, pypore
    rng = np.random.default_rng(13)
    currents = [staircase(rng, int(rng.integers(3, 10)), 260, 800) for _ in range(25)]
    model = peptide_hmm.pro_001_model()
    parser = parsers.SpeedyStatSplit(**CLI)
    for resident in (True, False):
        off, start, mean, (scores, plen, paths) = pypore.segment_and_decode(currents, model, parser, device_resident=resident)
        assert len(off) == len(currents) + 1
        poff = np.concatenate([[0], np.cumsum(np.maximum(plen, 0))])
        for e in (0, 7, 24):
            segs = parser.parse(currents[e])
            means32 = np.array([s.mean for s in segs]).astype(np.float32)
            assert np.array_equal(means32, mean[off[e]:off[e + 1]].astype(np.float32))
            s, p = model.viterbi(means32.astype(np.float64))
            assert s == scores[e]
            assert (p is None and plen[e] < 0) or [k for k, _ in p] == list(paths[poff[e]:poff[e + 1]])


def parameter_sweep(n_trials, seed, events_per_trial=24, mw_max=300, len_max=6000):
    """Seeded sweep over segmenter parameters and signal shapes (staircases, drifts, spikes, constant stretches, a NaN,
    a 300k-sample event).  Returns (events checked, events whose breakpoints differ from the oracle's, smallest oracle
    decision margin among those).  Also used by tools/statsplit_soak.py."""
    rng = np.random.default_rng(seed)

    def signal(kind, n):
        if kind == 0:
            return staircase(rng, max(1, n // 400), 100, 700)[:n] if n >= 700 else rng.normal(30, 1, size=n)
        if kind == 1:
            return 40 + np.cumsum(rng.normal(0, 0.05, size=n)) + rng.normal(0, 0.5, size=n)        # drift
        if kind == 2:
            x = rng.normal(35, 0.8, size=n)
            x[rng.integers(0, n, size=max(1, n // 200))] += rng.normal(0, 30, size=max(1, n // 200))   # spikes
            return x
        x = rng.normal(35, 1.0, size=n)
        a = int(rng.integers(0, max(1, n - 50)))
        x[a:a + 50] = 33.0                                                                         # a constant stretch
        return x

    checked, differ, worst_margin = 0, 0, np.inf
    for trial in range(n_trials):
        mw = int(rng.integers(1, mw_max))
        kw = dict(min_width=mw, max_width=int(mw * rng.integers(1, 30) + rng.integers(0, 50)),
                  window_width=int(2 * mw + rng.integers(0, 6 * mw + 20)))
        if trial % 3 == 0:
            kw["min_gain_per_sample"] = float(rng.uniform(0.001, 0.5))
        elif trial % 3 == 1:
            kw.update(prior_segments_per_second=float(rng.uniform(1, 1000)), false_positive_rate=float(rng.uniform(0.1, 100)))
        p = parsers.FastStatSplit(**kw)
        currents = [signal(int(rng.integers(0, 4)), int(rng.integers(1, len_max))) for _ in range(events_per_trial)]
        if trial == 0:
            currents.append(staircase(rng, 400, 500, 1000))                   # ~300k samples
        if trial == 1:
            x = rng.normal(30, 1, size=3000)
            x[1234] = np.nan                                                  # NaN poisons the sums from there on
            currents.append(x)
        off, start, mean, std, _ = p.segment_means_batch(currents, want_std=False)
        for e, cur in enumerate(currents):
            bp, stats = so.parse(cur, kw["min_width"], kw["max_width"], kw["window_width"], gain=p.min_gain, with_stats=True)
            got = start[off[e]:off[e + 1]]
            if not np.array_equal(got, np.concatenate([[0], bp])):
                differ += 1
                worst_margin = min(worst_margin, stats[0])
                assert stats[0] < 1e-9, (trial, e, kw, stats[0])    # only a tie within the accuracy of log() may differ
            checked += 1
    return checked, differ, worst_margin


@pytest.mark.gpu
def test_gpu_random_parameter_sweep():
    """Breakpoints identical to the oracle for every event of a seeded parameter / signal-shape sweep."""
    checked, differ, _ = parameter_sweep(14, 2026)
    assert checked >= 14 * 24 and differ == 0




@pytest.mark.gpu
def test_gpu_forced_splits_shorter_than_min_width():
    """Regression (found by tools/statsplit_soak.py): with max_width barely above min_width a forced split cuts off
    pieces shorter than min_width (cparsers.pyx:198-201), so an event can have up to 2 * (n / min_width) segments; the
    per-event split lists were sized for n / min_width + 1 and ran into the next event's.  Seed 99 reaches such a
    parameter set (min_width 275, max_width 313, window 1094) in its 11th trial."""
    checked, differ, _ = parameter_sweep(11, 99)
    assert checked >= 11 * 24 and differ == 0
    # the shape of the problem on its own: 910 samples, breakpoints 158 / 433 / 635 in the reference's terms
    rng = np.random.default_rng(1)
    cur = rng.normal(35, 1.0, size=910)
    kw = dict(min_width=275, max_width=313, window_width=1094, min_gain_per_sample=10.0)       # threshold never met: forced splits only
    p = parsers.FastStatSplit(**kw)
    batch = [cur[:400], cur, cur[:500], cur[:330], cur[:400]]        # 400 samples: forced split at min(313, 400 - 275) = 125
    off, start, mean, std, _ = p.segment_means_batch(batch, want_std=True)
    for e, c in enumerate(batch):
        assert _check_event(c, kw, p.min_gain, start[off[e]:off[e + 1]], mean[off[e]:off[e + 1]], std[off[e]:off[e + 1]], e)
    assert list(start[off[0]:off[1]]) == [0, 125] and list(start[off[4]:off[5]]) == [0, 125]
