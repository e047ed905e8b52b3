"""ctypes front-end of oracle/segalign_oracle.c (TEST INFRASTRUCTURE ONLY -- never imported by protpore_b200/).

CPU restatement of PyPore's cSegmentAligner (/root/reference/PyPore/calignment.pyx:20-101); `load_reference()` gives the
compiled, unmodified reference class (oracle/_ref/calignment*.so, built by oracle/build_ref.sh)."""
import ctypes
import glob
import importlib.machinery
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsegalign_oracle.so")
_f64p = ctypes.POINTER(ctypes.c_double)
_lib = None
_ref = None


def build(force=False):
    src = os.path.join(_HERE, "segalign_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared", src, "-o", _SO, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.segalign_align.restype = ctypes.c_int
        lib.segalign_align.argtypes = [_f64p, _f64p, _f64p, ctypes.c_int64, ctypes.c_double, ctypes.c_double, _f64p, _f64p, _f64p,
                                       ctypes.c_int64, _f64p, _f64p, ctypes.POINTER(ctypes.c_int32)]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(_f64p)


def align(model_means, model_stds, model_durs, skip_penalty, backslip_penalty, seq_means, seq_stds, seq_durs):
    """(score, order float64[s], status) of cSegmentAligner(...).align(...); status != 0 marks the cases where the
    reference itself is undefined (see segalign_oracle.c)."""
    mm, ms, md = (np.ascontiguousarray(a, dtype=np.float64) for a in (model_means, model_stds, model_durs))
    sm, ss, sd = (np.ascontiguousarray(a, dtype=np.float64) for a in (seq_means, seq_stds, seq_durs))
    score = ctypes.c_double(0.0)
    status = ctypes.c_int32(0)
    order = np.empty(len(sm), dtype=np.float64)
    rc = _load().segalign_align(_p(mm), _p(ms), _p(md), len(mm), float(skip_penalty), float(backslip_penalty), _p(sm), _p(ss), _p(sd),
                                len(sm), ctypes.byref(score), _p(order), ctypes.byref(status))
    if rc != 0:
        raise ValueError("segalign oracle: needs m >= 2 and s >= 1")
    return score.value, order, status.value


def reference_available():
    return bool(glob.glob(os.path.join(_HERE, "_ref", "calignment*.so")))


def load_reference():
    """The compiled, unmodified PyPore/calignment.pyx module (cSegmentAligner)."""
    global _ref
    if _ref is None:
        so = sorted(glob.glob(os.path.join(_HERE, "_ref", "calignment*.so")))[0]
        loader = importlib.machinery.ExtensionFileLoader("calignment", so)
        spec = importlib.util.spec_from_file_location("calignment", so, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _ref = mod
    return _ref
