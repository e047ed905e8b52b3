#!/usr/bin/env python
"""Generate tests/golden/statsplit.npz from the COMPILED, UNMODIFIED reference segmenter (oracle/_ref/cparsers*.so,
built by oracle/build_ref.sh from /root/reference/PyPore/cparsers.pyx).  Run in the build container only:

    python tests/golden/make_golden_statsplit.py

Per case: the FastStatSplit constructor arguments, a seeded synthetic current (float32-rounded so the fixture
stays small, stored as float32), and the reference's outputs: min_gain, segment starts, Segment.mean / .std.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def steps(rng, n_levels, lo, hi, sd_lo=0.4, sd_hi=2.5):
    """A staircase current like a translocation event: levels ~U(25, 50) pA with Gaussian noise."""
    parts = [rng.normal(rng.uniform(25, 50), rng.uniform(sd_lo, sd_hi), size=int(rng.integers(lo, hi)))
             for _ in range(n_levels)]
    return np.concatenate(parts).astype(np.float32).astype(np.float64)


def cases():
    rng = np.random.default_rng(20261020)
    cli = dict(min_width=250, max_width=1000, window_width=2000, min_gain_per_sample=0.05)     # peptide_hmm_v0.0.1.py:311-313, 426
    out = []
    out.append(("cli_staircase", cli, steps(rng, 14, 150, 900)))
    out.append(("cli_short_levels", cli, steps(rng, 30, 40, 400)))
    out.append(("cli_no_split", cli, steps(rng, 1, 480, 481)))                                  # n <= 2 * min_width
    out.append(("cli_flat_forced", cli, steps(rng, 1, 5200, 5201)))                             # max_width splits
    out.append(("default_threshold_zero", dict(min_width=50, max_width=2000, window_width=400), steps(rng, 8, 100, 600)))
    out.append(("right_only_path", dict(min_width=20, max_width=300, window_width=400, min_gain_per_sample=10.0),
                steps(rng, 1, 3000, 3001)))                                                     # cparsers.pyx:188-190
    out.append(("bayes_cutoff", dict(min_width=100, max_width=5000, window_width=1000, prior_segments_per_second=40.,
                                     false_positive_rate=5., sampling_freq=1.e5, cutoff_freq=5000.),
                steps(rng, 12, 200, 1200)))
    out.append(("constant_signal", dict(min_width=10, max_width=100000, window_width=200, min_gain_per_sample=0.05),
                np.full(900, 37.5)))
    out.append(("tiny", dict(min_width=1, max_width=4, window_width=2, min_gain_per_sample=0.5),
                np.array([1.0, 1.5, 9.0, 9.5, 9.25, 2.0, 2.5, 2.25, 2.0, 7.0, 7.5], dtype=np.float64)))
    return out


def main():
    cp = ref_loader.load_reference_cparsers()
    blob = {}
    names = []
    for name, kw, cur in cases():
        parser = cp.FastStatSplit(**kw)
        segs = parser.parse(cur)
        names.append(name)
        blob[name + "__kw"] = np.array(json.dumps(kw))
        blob[name + "__current"] = cur.astype(np.float32) if np.array_equal(cur.astype(np.float32), cur) else cur
        blob[name + "__min_gain"] = np.float64(parser.min_gain)
        blob[name + "__starts"] = np.array([s.start for s in segs], dtype=np.int64)
        blob[name + "__means"] = np.array([s.mean for s in segs], dtype=np.float64)
        blob[name + "__stds"] = np.array([s.std for s in segs], dtype=np.float64)
        print("{:24s} n={:6d} segments={:4d} min_gain={}".format(name, len(cur), len(segs), parser.min_gain))
    blob["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "statsplit.npz"), **blob)


if __name__ == "__main__":
    main()
