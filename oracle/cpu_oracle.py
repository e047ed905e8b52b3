"""ctypes front-end of oracle/hmm_oracle.c (TEST INFRASTRUCTURE ONLY -- never imported by protpore_b200/).

`OracleModel` is built from baked CSR arrays (any object with the attribute names of
`protpore_b200.yahmm.bake.Baked`, or explicit arrays) and exposes the float64 restatements of the
reference sweeps: forward / backward / log_probability / viterbi / forward_backward / bw_accumulate.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libhmm_oracle.so")

_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)


class _CModel(ctypes.Structure):
    _fields_ = [("m", ctypes.c_int32), ("silent_start", ctypes.c_int32), ("start_index", ctypes.c_int32),
                ("end_index", ctypes.c_int32), ("finite", ctypes.c_int32), ("n_edges", ctypes.c_int32),
                ("in_ptr", _i32p), ("in_src", _i32p), ("in_logp", _f64p),
                ("out_ptr", _i32p), ("out_dst", _i32p), ("out_logp", _f64p),
                ("kind", _i32p), ("p0", _f64p), ("p1", _f64p), ("logw", _f64p)]


def build(force=False):
    """Compile hmm_oracle.c into oracle/_ref/libhmm_oracle.so (gcc only; seconds)."""
    src = os.path.join(_HERE, "hmm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared",
                               src, "-o", _SO, "-lm"])
    return _SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        mp = ctypes.POINTER(_CModel)
        lib.orc_emission.restype = ctypes.c_double
        lib.orc_emission.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]
        lib.orc_forward.restype = None
        lib.orc_forward.argtypes = [mp, _f64p, ctypes.c_int64, _f64p]
        lib.orc_backward.restype = None
        lib.orc_backward.argtypes = [mp, _f64p, ctypes.c_int64, _f64p]
        lib.orc_log_probability.restype = ctypes.c_double
        lib.orc_log_probability.argtypes = [mp, _f64p, ctypes.c_int64]
        lib.orc_viterbi.restype = ctypes.c_double
        lib.orc_viterbi.argtypes = [mp, _f64p, ctypes.c_int64, _i32p, ctypes.c_int64, _i64p, _f64p]
        lib.orc_forward_backward.restype = ctypes.c_double
        lib.orc_forward_backward.argtypes = [mp, _f64p, ctypes.c_int64, _f64p, _f64p]
        lib.orc_bw_accumulate.restype = ctypes.c_double
        lib.orc_bw_accumulate.argtypes = [mp, _f64p, ctypes.c_int64, _f64p, _f64p, _f64p]
        _lib = lib
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def emission(kind, p0, p1, x):
    return _load().orc_emission(int(kind), float(p0), float(p1), float(x))


class OracleModel(object):
    def __init__(self, baked, kind, p0, p1):
        c = np.ascontiguousarray
        self.m = len(baked.states) if hasattr(baked, "states") else int(baked.m)
        self.silent_start = int(baked.silent_start)
        self.start_index, self.end_index, self.finite = int(baked.start_index), int(baked.end_index), int(baked.finite)
        self._a = dict(
            in_ptr=c(baked.in_ptr, dtype=np.int32), in_src=c(baked.in_idx, dtype=np.int32),
            in_logp=c(baked.in_logp, dtype=np.float64),
            out_ptr=c(baked.out_ptr, dtype=np.int32), out_dst=c(baked.out_idx, dtype=np.int32),
            out_logp=c(baked.out_logp, dtype=np.float64),
            kind=c(kind, dtype=np.int32), p0=c(p0, dtype=np.float64), p1=c(p1, dtype=np.float64),
            logw=c(baked.state_weights, dtype=np.float64))
        a = self._a
        self._c = _CModel(self.m, self.silent_start, self.start_index, self.end_index, self.finite,
                          len(a["in_src"]), _p(a["in_ptr"], _i32p), _p(a["in_src"], _i32p),
                          _p(a["in_logp"], _f64p), _p(a["out_ptr"], _i32p), _p(a["out_dst"], _i32p),
                          _p(a["out_logp"], _f64p), _p(a["kind"], _i32p), _p(a["p0"], _f64p),
                          _p(a["p1"], _f64p), _p(a["logw"], _f64p))

    @classmethod
    def from_model(cls, model):
        """From a baked product `Model` (reads only its public baked arrays and distributions)."""
        b = model._baked
        trip = [s.distribution.device_params() for s in b.states[:b.silent_start]]
        return cls(b, [t[0] for t in trip], [t[1] for t in trip], [t[2] for t in trip])

    @staticmethod
    def _x(seq):
        x = np.ascontiguousarray(seq, dtype=np.float64)
        if x.ndim != 1 or len(x) == 0:
            raise ValueError("Invalid shape in axis 0: 0.")   # the reference's cvarray error (yahmm.pyx:2836)
        return x

    def forward(self, seq):
        x = self._x(seq)
        f = np.empty((len(x) + 1, self.m))
        _load().orc_forward(ctypes.byref(self._c), _p(x, _f64p), len(x), _p(f, _f64p))
        return f

    def backward(self, seq):
        x = self._x(seq)
        b = np.empty((len(x) + 1, self.m))
        _load().orc_backward(ctypes.byref(self._c), _p(x, _f64p), len(x), _p(b, _f64p))
        return b

    def log_probability(self, seq):
        x = self._x(seq)
        return _load().orc_log_probability(ctypes.byref(self._c), _p(x, _f64p), len(x))

    def viterbi(self, seq, with_margin=False):
        """(score, int32 path of state indices start..end) or (-inf, None); optionally the min decision margin."""
        x = self._x(seq)
        cap = 4 * (len(x) + 2) + 2 * self.m
        while True:
            path = np.empty(cap, dtype=np.int32)
            plen = ctypes.c_int64(0)
            margin = ctypes.c_double(0)
            score = _load().orc_viterbi(ctypes.byref(self._c), _p(x, _f64p), len(x), _p(path, _i32p), cap,
                                        ctypes.byref(plen), ctypes.byref(margin))
            if plen.value <= cap:
                break
            cap = plen.value
        out = (score, None) if plen.value < 0 else (score, path[:plen.value].copy())
        return out + (margin.value,) if with_margin else out

    def forward_backward(self, seq):
        """(expected m x m linear, emission_weights n x silent_start log) or (None, None)."""
        x = self._x(seq)
        exp_t = np.zeros((self.m, self.m))
        ew = np.zeros((len(x), self.silent_start))
        lsp = _load().orc_forward_backward(ctypes.byref(self._c), _p(x, _f64p), len(x), _p(exp_t, _f64p),
                                           _p(ew, _f64p))
        if lsp == -np.inf:
            return None, None
        return exp_t, ew

    def bw_accumulate(self, seqs):
        """E-step sufficient statistics over sequences: (expected m x m, moments ne x 3, minmax ne x 2, logZ list)."""
        exp_t = np.zeros((self.m, self.m))
        mom = np.zeros((self.silent_start, 3))
        mm = np.empty((self.silent_start, 2))
        mm[:, 0], mm[:, 1] = np.inf, -np.inf
        lz = []
        for s in seqs:
            x = self._x(s)
            lz.append(_load().orc_bw_accumulate(ctypes.byref(self._c), _p(x, _f64p), len(x), _p(exp_t, _f64p),
                                                _p(mom, _f64p), _p(mm, _f64p)))
        return exp_t, mom, mm, np.asarray(lz)
