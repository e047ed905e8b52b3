"""ctypes front-end of oracle/statsplit_oracle.c (TEST INFRASTRUCTURE ONLY -- never imported by protpore_b200/).

CPU restatement of PyPore's FastStatSplit segmenter (/root/reference/PyPore/cparsers.pyx:45-203)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libstatsplit_oracle.so")
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)
_lib = None


def build(force=False):
    """Compile statsplit_oracle.c into oracle/_ref/libstatsplit_oracle.so (gcc only; a second)."""
    src = os.path.join(_HERE, "statsplit_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared",
                               src, "-o", _SO, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.statsplit_parse.restype = ctypes.c_int64
        lib.statsplit_parse.argtypes = [_f64p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_double, _i64p, ctypes.c_int64, _f64p]
        lib.statsplit_min_gain.restype = ctypes.c_double
        lib.statsplit_min_gain.argtypes = [ctypes.c_int64] + [ctypes.c_double] * 5
        _lib = lib
    return _lib


def min_gain(window_width=10000, min_gain_per_sample=None, false_positive_rate=None, prior_segments_per_second=None,
             sampling_freq=1.e5, cutoff_freq=None):
    """The threshold FastStatSplit.__init__ computes (cparsers.pyx:56-101)."""
    return _load().statsplit_min_gain(int(window_width), float(min_gain_per_sample or 0.0),
                                      float(false_positive_rate or 0.0), float(prior_segments_per_second or 0.0),
                                      float(sampling_freq), float(cutoff_freq or 0.0))


def parse(current, min_width=100, max_width=1000000, window_width=10000, gain=None, with_stats=False, **kw):
    """Breakpoints of FastStatSplit.parse(current) (cparsers.pyx:103-118, 157-203), ascending int64 array.
    with_stats: also (min decision margin, window scans, candidate positions evaluated)."""
    x = np.ascontiguousarray(current, dtype=np.float64)
    if gain is None:
        gain = min_gain(window_width=window_width, **kw)
    # a forced split (cparsers.pyx:198-201) may cut off a piece shorter than min_width: up to 2 * (n / min_width) segments
    cap = max(16, 2 * (len(x) // max(1, int(min_width))) + 16)
    stats = np.zeros(3, dtype=np.float64)
    while True:
        out = np.empty(cap, dtype=np.int64)
        n = _load().statsplit_parse(x.ctypes.data_as(_f64p), len(x), int(min_width), int(max_width), int(window_width),
                                    float(gain), out.ctypes.data_as(_i64p), cap, stats.ctypes.data_as(_f64p))
        if n <= cap:
            break
        cap = int(n)
    bp = out[:n].copy()
    return (bp, stats) if with_stats else bp


def segment_means(current, breakpoints):
    """Segment.mean of every segment FastStatSplit.parse builds (cparsers.pyx:115-116, core.py:209-211)."""
    x = np.asarray(current, dtype=np.float64)
    edges = np.concatenate([[0], np.asarray(breakpoints, dtype=np.int64), [len(x)]])
    return np.array([np.mean(x[a:b]) for a, b in zip(edges[:-1], edges[1:])], dtype=np.float64)
