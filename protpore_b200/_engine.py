"""ctypes binding of libprotpore_b200.so (the C ABI in include/protpore_b200.h).

`Engine` owns one `pp_model` handle (one baked model on one GPU) and exposes the batched entry
points with numpy (host) or torch-CUDA (device) buffers.  There is deliberately NO fallback: if the
shared library is missing, or no CUDA device is present, constructing an `Engine` raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PROTPORE_B200_LIB", os.path.join(_HERE, "libprotpore_b200.so"))

PP_OK, PP_ERR_ARG, PP_ERR_CUDA, PP_ERR_NOMEM, PP_ERR_UNSUPPORTED, PP_ERR_CAPACITY = 0, -1, -2, -3, -4, -5
PP_MEM_HOST, PP_MEM_DEVICE = 0, 1
PP_F32, PP_F64 = 0, 1
PROF_SLOTS = ("viterbi_ms", "traceback_ms", "forward_ms", "fb_ms", "sample_ms", "h2d_ms", "d2h_ms", "schedule_host_ms")

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)
_u16p = ctypes.POINTER(ctypes.c_uint16)
_vp = ctypes.c_void_p


class PPModelDesc(ctypes.Structure):
    """struct pp_model_desc (include/protpore_b200.h)."""
    _fields_ = [("n_states", ctypes.c_int32), ("silent_start", ctypes.c_int32), ("start_index", ctypes.c_int32),
                ("end_index", ctypes.c_int32), ("finite", ctypes.c_int32), ("n_edges", ctypes.c_int32),
                ("in_ptr", _i32p), ("in_src", _i32p), ("in_logp", _f64p),
                ("out_ptr", _i32p), ("out_dst", _i32p), ("out_logp", _f64p),
                ("emit_kind", _i32p), ("emit_p0", _f64p), ("emit_p1", _f64p), ("emit_logw", _f64p)]


class PPStatSplitParams(ctypes.Structure):
    """struct pp_statsplit_params (include/protpore_b200.h)."""
    _fields_ = [("min_width", ctypes.c_int64), ("max_width", ctypes.c_int64), ("window_width", ctypes.c_int64),
                ("min_gain", ctypes.c_double)]


class EngineError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "protpore_b200 error {}: {}".format(code, message))
        self.code = code


_lib = None

# every symbol include/protpore_b200.h declares: (name, restype, argtypes)
ABI = [
    ("pp_abi_version", ctypes.c_int, []),
    ("pp_last_error", ctypes.c_char_p, []),
    ("pp_device_count", ctypes.c_int, []),
    ("pp_host_alloc", _vp, [ctypes.c_int64]),
    ("pp_host_free", None, [_vp]),
    ("pp_model_create", ctypes.c_int, [ctypes.POINTER(PPModelDesc), ctypes.c_int, ctypes.POINTER(_vp)]),
    ("pp_model_destroy", None, [_vp]),
    ("pp_model_update_params", ctypes.c_int, [_vp, ctypes.POINTER(PPModelDesc)]),
    ("pp_model_set_workspace_limit", ctypes.c_int, [_vp, ctypes.c_int64]),
    ("pp_model_set_kernel_path", ctypes.c_int, [_vp, ctypes.c_int32]),
    ("pp_model_kernel_info", ctypes.c_int, [_vp, _vp]),
    ("pp_model_set_emission_table", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32]),
    ("pp_schedule_describe", ctypes.c_int64, [ctypes.POINTER(PPModelDesc), ctypes.c_char_p, ctypes.c_int64]),
    ("pp_viterbi_workspace_bytes", ctypes.c_int64, [_vp, _vp, ctypes.c_int64]),
    ("pp_forward_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, ctypes.c_int64, ctypes.c_int32, _vp]),
    ("pp_viterbi_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp, _vp,
                                        ctypes.c_int64, _i64p]),
    ("pp_score_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp, _vp,
                                      ctypes.c_int64, _i64p, _vp]),
    ("pp_forward_table", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp]),
    ("pp_backward_table", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp]),
    ("pp_forward_backward_info", ctypes.c_int, [_vp, _i64p]),
    ("pp_forward_backward_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int32, _vp, ctypes.c_int64, ctypes.c_int32, _vp,
                                                 _vp, _vp, _vp, _vp]),
    ("pp_sample_events", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int32, _vp]),
    ("pp_profile_get", ctypes.c_int, [_vp, _f64p, ctypes.c_int32]),
    ("pp_statsplit_min_gain", ctypes.c_double, [ctypes.c_int64] + [ctypes.c_double] * 5),
    ("pp_statsplit_batch", ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp, _vp, _vp,
                                          _vp, ctypes.c_int64, _i64p]),
    ("pp_statsplit_profile", ctypes.c_int, [ctypes.c_int, _f64p]),
    ("pp_segment_align_batch", ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp, _vp]),
    ("pp_device_peaks", ctypes.c_int, [ctypes.c_int, _f64p]),
    ("pp_launch_count", ctypes.c_int64, [_vp]),
    ("pp_segment_align_profile", ctypes.c_int, [_f64p]),
    ("pp_segment_align_trim", ctypes.c_int, [ctypes.c_int]),
    ("pp_debug_check_guards", ctypes.c_int, []),
    ("pp_debug_guard_count", ctypes.c_int, []),
    ("pp_debug_guard_selftest", ctypes.c_int, []),
    ("pp_model_stream", ctypes.c_uint64, [_vp]),
]


def load_library(path=None):
    """Load the C-ABI library and bind every declared symbol (raises OSError/AttributeError if absent)."""
    global _lib
    if _lib is None or path is not None:
        lib = ctypes.CDLL(path or LIB_PATH)
        for name, restype, argtypes in ABI:
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if path is not None:
            return lib
        _lib = lib
    return _lib


def make_desc(baked, kinds, p0, p1):
    """(pp_model_desc, keep-alive dict) for a baked model."""
    c = np.ascontiguousarray
    a = dict(in_ptr=c(baked.in_ptr, dtype=np.int32), in_src=c(baked.in_idx, dtype=np.int32),
             in_logp=c(baked.in_logp, dtype=np.float64), out_ptr=c(baked.out_ptr, dtype=np.int32),
             out_dst=c(baked.out_idx, dtype=np.int32), out_logp=c(baked.out_logp, dtype=np.float64),
             kind=c(kinds, dtype=np.int32), p0=c(p0, dtype=np.float64), p1=c(p1, dtype=np.float64),
             logw=c(baked.state_weights, dtype=np.float64))
    p = lambda k, t: a[k].ctypes.data_as(t)
    desc = PPModelDesc(len(baked.states), int(baked.silent_start), int(baked.start_index), int(baked.end_index),
                       int(baked.finite), len(a["in_src"]), p("in_ptr", _i32p), p("in_src", _i32p),
                       p("in_logp", _f64p), p("out_ptr", _i32p), p("out_dst", _i32p), p("out_logp", _f64p),
                       p("kind", _i32p), p("p0", _f64p), p("p1", _f64p), p("logw", _f64p))
    return desc, a


def describe_schedule(model):
    """Text dump of the per-warp run schedule of a baked `Model` (no GPU needed)."""
    lib = load_library()
    kinds, p0, p1 = model._device_params()
    desc, keep = make_desc(model._baked, kinds, p0, p1)
    buf = ctypes.create_string_buffer(1 << 16)
    rc = lib.pp_schedule_describe(ctypes.byref(desc), buf, 1 << 16)
    if rc < 0:
        raise EngineError(rc, lib.pp_last_error().decode())
    return buf.value.decode()


def device_peaks(device=0):
    """Measured float64 issue peaks: {'dadd_per_s', 'dfma_per_s', 'sms', 'clock_hz'} (pp_device_peaks)."""
    lib = load_library()
    out = (ctypes.c_double * 4)()
    rc = lib.pp_device_peaks(int(device), out)
    if rc != 0:
        raise EngineError(rc, lib.pp_last_error().decode())
    return {"dadd_per_s": out[0], "dfma_per_s": out[1], "sms": int(out[2]), "clock_hz": out[3]}


def check_guards():
    """Guard mode (PP_GUARD=1): number of device buffers whose guard bytes were overwritten; raises if any were."""
    lib = load_library()
    rc = lib.pp_debug_check_guards()
    if rc != 0:
        raise EngineError(rc, lib.pp_last_error().decode())
    return lib.pp_debug_guard_count()


def device_count():
    return load_library().pp_device_count()


def _ptr(a):
    """Raw address of a numpy array or a torch tensor."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def _torch_inputs_ready(*tensors):
    """The library enqueues on its own non-blocking stream: work torch still has in flight on ITS current stream (the
    cast / cat / kernel that produces `obs`) must be complete before a device pointer is handed over."""
    import torch
    for t in tensors:
        if t is not None and hasattr(t, "is_cuda") and t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
            return


def _is_device(a):
    return not isinstance(a, np.ndarray) and hasattr(a, "is_cuda") and a.is_cuda


class _Pinned(object):
    """Owner of one cudaHostAlloc block (freed when the last numpy view goes away)."""

    def __init__(self, lib, nbytes):
        self._lib, self.ptr = lib, lib.pp_host_alloc(int(nbytes))

    def __del__(self):
        if self.ptr:
            self._lib.pp_host_free(self.ptr)
            self.ptr = None


def pinned_empty(shape, dtype):
    """numpy array in page-locked host memory (pp_host_alloc): host<->device copies run at link speed.
    Falls back to pageable memory if the allocation fails (e.g. no CUDA runtime)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    if n == 0:
        return np.empty(shape, dtype=dtype)
    owner = _Pinned(load_library(), n * dtype.itemsize)
    if not owner.ptr:
        return np.empty(shape, dtype=dtype)
    buf = (ctypes.c_char * (n * dtype.itemsize)).from_address(owner.ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    _PINNED_OWNERS[id(buf)] = (owner, buf)        # keep the block alive as long as the buffer object is referenced
    import weakref
    weakref.finalize(arr.base if arr.base is not None else arr, _PINNED_OWNERS.pop, id(buf), None)
    return arr


_PINNED_OWNERS = {}


def pack_events(sequences, dtype=np.float32):
    """List of 1-D sequences -> (obs back to back, int64 offsets[n+1])."""
    lens = np.fromiter((len(s) for s in sequences), dtype=np.int64, count=len(sequences))
    offsets = np.zeros(len(sequences) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    obs = np.empty(int(offsets[-1]), dtype=dtype)
    for s, o in zip(sequences, offsets[:-1]):
        obs[o:o + len(s)] = s
    return obs, offsets


class Engine(object):
    """One baked model resident on one GPU."""

    def __init__(self, baked, kinds, p0, p1, device=0):
        self._lib = load_library()
        self._h = _vp()
        self.m = len(baked.states)
        self.ne = int(baked.silent_start)
        self.n_edges = len(baked.in_idx)
        self._keep = None
        desc = self._desc(baked, kinds, p0, p1)
        rc = self._lib.pp_model_create(ctypes.byref(desc), int(device), ctypes.byref(self._h))
        if rc != PP_OK:
            self._h = _vp()
            raise EngineError(rc, self._lib.pp_last_error().decode())
        self.device = int(device)

    def _desc(self, baked, kinds, p0, p1):
        desc, self._keep = make_desc(baked, kinds, p0, p1)
        return desc

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pp_model_destroy(self._h)
            self._h = _vp()

    __del__ = close

    def _check(self, rc):
        if rc != PP_OK:
            raise EngineError(rc, self._lib.pp_last_error().decode())

    def update_params(self, baked, kinds, p0, p1):
        desc = self._desc(baked, kinds, p0, p1)
        self._check(self._lib.pp_model_update_params(self._h, ctypes.byref(desc)))

    def set_emission_table(self, table):
        """pp_model_set_emission_table: float64 [rows, silent_start] log-densities of the table states for the `obs` of
        the calls that follow (numpy: copied to the device; torch CUDA tensor: used in place); None clears it."""
        if table is None:
            self._check(self._lib.pp_model_set_emission_table(self._h, None, 0, PP_MEM_HOST))
            self._table_keep = None
            return
        dev = _is_device(table)
        if not dev:
            table = np.ascontiguousarray(table, dtype=np.float64)
        self._table_keep = table if dev else None
        self._check(self._lib.pp_model_set_emission_table(self._h, _ptr(table), int(table.shape[0]),
                                                          PP_MEM_DEVICE if dev else PP_MEM_HOST))

    def set_kernel_path(self, path):
        """0: automatic (register-resident profile kernels where the model is a linear profile); 1: run-based lane kernels;
        2: like 1, and forward_backward_batch on the exact state-parallel kernel only; 3: like 0, and posteriors from the
        batched profile kernel (within 1e-10 above e^-250, imprecise or -inf below; the default keeps them on the exact kernel)."""
        self._check(self._lib.pp_model_set_kernel_path(self._h, int(path)))

    def kernel_info(self):
        out = (ctypes.c_int32 * 4)()
        self._check(self._lib.pp_model_kernel_info(self._h, out))
        return {"family": {2: "profile", 1: "lane", 0: "state"}[out[0]], "positions": out[1], "warps": out[2],
                "ptr_row_bytes": out[3]}

    def set_workspace_limit(self, nbytes):
        self._check(self._lib.pp_model_set_workspace_limit(self._h, int(nbytes)))

    def profile(self):
        out = (ctypes.c_double * 8)()
        self._check(self._lib.pp_profile_get(self._h, out, 8))
        return dict(zip(PROF_SLOTS, list(out)))

    def launch_count(self):
        return int(self._lib.pp_launch_count(self._h))

    def stream(self):
        return int(self._lib.pp_model_stream(self._h))

    # ---- batched entry points --------------------------------------------------------------
    @staticmethod
    def _obs_dtype(obs):
        name = str(obs.dtype)
        if name.endswith("float32"):
            return PP_F32
        if name.endswith("float64"):
            return PP_F64
        raise TypeError("observations must be float32 or float64, got {}".format(name))

    def forward_batch(self, obs, offsets, out=None):
        """log P(event) for every event.  numpy in -> numpy out; torch CUDA tensors in -> torch out."""
        n = len(offsets) - 1
        if _is_device(obs):
            import torch
            out = torch.empty(n, dtype=torch.float64, device=obs.device) if out is None else out
            mem = PP_MEM_DEVICE
            _torch_inputs_ready(obs)
        else:
            obs = np.ascontiguousarray(obs)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            out = np.empty(n, dtype=np.float64) if out is None else out
            mem = PP_MEM_HOST
        self._check(self._lib.pp_forward_batch(self._h, _ptr(obs), self._obs_dtype(obs), _ptr(offsets), n, mem,
                                               _ptr(out)))
        return out

    def viterbi_batch(self, obs, offsets, with_paths=True, capacity=None, out=None):
        """(scores f64[n], path_len i64[n], paths u16[total] back to back in event order).

        `out` = (scores, path_len, paths) lets a serving loop reuse its own (ideally pinned, see
        `pinned_empty`) host buffers; len(paths) is then the path capacity."""
        return self._decode(obs, offsets, with_paths, capacity, out, None)

    def score_batch(self, obs, offsets, with_paths=True, capacity=None, out=None):
        """Viterbi decode AND forward log-probability of every event in one call (pp_score_batch): one upload, one
        bucketing.  Returns (viterbi scores, path_len, paths, log P); `out` = (scores, path_len, paths, logp)."""
        n = len(offsets) - 1
        if out is not None:
            logp = out[3]
        elif _is_device(obs):
            import torch
            logp = torch.empty(n, dtype=torch.float64, device=obs.device)
        else:
            logp = np.empty(n, dtype=np.float64)
        scores, plen, paths = self._decode(obs, offsets, with_paths, capacity, out, logp)
        return scores, plen, paths, logp

    def _decode(self, obs, offsets, with_paths, capacity, out, logp):
        n = len(offsets) - 1
        dev = _is_device(obs)
        if dev:
            import torch
            total_obs = int(obs.numel())
            mk = lambda k, dt: torch.empty(k, dtype=dt, device=obs.device)
            scores, plen = (out[0], out[1]) if out is not None else (mk(n, torch.float64), mk(n, torch.int64))
            mem = PP_MEM_DEVICE
            _torch_inputs_ready(obs)
        else:
            obs = np.ascontiguousarray(obs)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            total_obs = len(obs)
            scores, plen = (out[0], out[1]) if out is not None else (np.empty(n, dtype=np.float64), np.empty(n, dtype=np.int64))
            mem = PP_MEM_HOST
        if out is not None:
            capacity = len(out[2])
        cap = int(capacity) if capacity is not None else 3 * total_obs + 8 * n + 64
        total = ctypes.c_int64(0)
        while True:
            if not with_paths:
                paths = None
            elif dev:
                import torch
                paths = out[2] if (out is not None and len(out[2]) >= cap) else torch.empty(cap, dtype=torch.uint16, device=obs.device)
            else:
                paths = out[2] if (out is not None and len(out[2]) >= cap) else np.empty(cap, dtype=np.uint16)
            if logp is None:
                rc = self._lib.pp_viterbi_batch(self._h, _ptr(obs), self._obs_dtype(obs), _ptr(offsets), n, mem,
                                                _ptr(scores), _ptr(plen), _ptr(paths), cap if with_paths else 0,
                                                ctypes.byref(total))
            else:
                rc = self._lib.pp_score_batch(self._h, _ptr(obs), self._obs_dtype(obs), _ptr(offsets), n, mem,
                                              _ptr(scores), _ptr(plen), _ptr(paths), cap if with_paths else 0,
                                              ctypes.byref(total), _ptr(logp))
            if rc == PP_ERR_CAPACITY:
                cap = max(int(total.value), 2 * cap)
                continue
            self._check(rc)
            break
        if with_paths:
            paths = paths[:total.value]
        return scores, plen, paths

    def forward_table(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty((len(x) + 1, self.m), dtype=np.float64)
        self._check(self._lib.pp_forward_table(self._h, _ptr(x), len(x), PP_MEM_HOST, _ptr(out)))
        return out

    def backward_table(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty((len(x) + 1, self.m), dtype=np.float64)
        self._check(self._lib.pp_backward_table(self._h, _ptr(x), len(x), PP_MEM_HOST, _ptr(out)))
        return out

    def forward_backward_batch(self, obs, offsets, posteriors=False, accum=None):
        """E-step statistics: (logp[n], edge_counts[E] (out-CSR order), moments[ne,3], minmax[ne,2], posteriors|None).

        `accum` = (edge_counts, moments, minmax) to keep accumulating into.  numpy in -> numpy out; torch CUDA
        tensors in -> torch out (the statistics stay on the device, ready for the NCCL allreduce)."""
        n = len(offsets) - 1
        if _is_device(obs):
            import torch
            mk = lambda shape, fill: torch.full(shape, fill, dtype=torch.float64, device=obs.device)
            logp = mk((n,), 0.0)
            if accum is None:
                edge, mom, mm = mk((self.n_edges,), 0.0), mk((self.ne, 3), 0.0), mk((self.ne, 2), 0.0)
                mm[:, 0], mm[:, 1] = float("inf"), float("-inf")
            else:
                edge, mom, mm = accum
            post = mk((int(obs.numel()), self.ne), 0.0) if posteriors else None
            mem = PP_MEM_DEVICE
            torch.cuda.current_stream(obs.device).synchronize()       # the handle enqueues on its own stream
        else:
            obs = np.ascontiguousarray(obs)
            offsets = np.ascontiguousarray(offsets, dtype=np.int64)
            logp = np.empty(n, dtype=np.float64)
            if accum is None:
                edge = np.zeros(self.n_edges, dtype=np.float64)
                mom = np.zeros((self.ne, 3), dtype=np.float64)
                mm = np.empty((self.ne, 2), dtype=np.float64)
                mm[:, 0], mm[:, 1] = np.inf, -np.inf
            else:
                edge, mom, mm = accum
            post = np.empty((len(obs), self.ne), dtype=np.float64) if posteriors else None
            mem = PP_MEM_HOST
        self._check(self._lib.pp_forward_backward_batch(self._h, _ptr(obs), self._obs_dtype(obs), _ptr(offsets), n,
                                                        mem, _ptr(logp), _ptr(edge), _ptr(mom), _ptr(mm), _ptr(post)))
        return logp, edge, mom, mm, post

    def forward_backward_info(self):
        """How the last forward_backward_batch was served: {'path': 'profile' | 'lane' | 'exact' | 'lane-failed' |
        'profile-failed', 'rejects', 'failed'}."""
        out = (ctypes.c_int64 * 3)()
        self._check(self._lib.pp_forward_backward_info(self._h, out))
        return {"path": ("exact", "lane", "lane-failed", "profile", "profile-failed")[int(out[0])], "rejects": int(out[1]), "failed": int(out[2])}

    def sample_events(self, lengths, seed=0, device_out=False):
        """Length-conditioned synthetic events: (obs float32, offsets int64)."""
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        offsets = np.zeros(len(lengths) + 1, dtype=np.int64)
        np.cumsum(lengths, out=offsets[1:])
        if device_out:
            import torch
            dev = torch.device("cuda", self.device)
            d_off = torch.from_numpy(offsets).to(dev)
            obs = torch.empty(int(offsets[-1]), dtype=torch.float32, device=dev)
            _torch_inputs_ready(d_off)
            self._check(self._lib.pp_sample_events(self._h, _ptr(d_off), len(lengths), int(seed), PP_MEM_DEVICE,
                                                   _ptr(obs)))
            return obs, d_off
        obs = np.empty(int(offsets[-1]), dtype=np.float32)
        self._check(self._lib.pp_sample_events(self._h, _ptr(offsets), len(lengths), int(seed), PP_MEM_HOST, _ptr(obs)))
        return obs, offsets


def statsplit_min_gain(window_width, min_gain_per_sample=None, false_positive_rate=None, prior_segments_per_second=None,
                       sampling_freq=1.e5, cutoff_freq=None):
    """pp_statsplit_min_gain: FastStatSplit.__init__'s threshold (cparsers.pyx:56-101).  Host arithmetic, no GPU."""
    return load_library().pp_statsplit_min_gain(int(window_width), float(min_gain_per_sample or 0.0),
                                                float(false_positive_rate or 0.0),
                                                float(prior_segments_per_second or 0.0), float(sampling_freq),
                                                float(cutoff_freq or 0.0))


def statsplit_batch(current, offsets, min_width, max_width, window_width, min_gain, device=0, want_std=True,
                    want_f32=False, capacity=None):
    """pp_statsplit_batch.  `current` / `offsets`: numpy (host) or torch-CUDA (device) buffers, float64 / int64.
    Returns (seg_offsets, seg_start, seg_mean, seg_std | None, seg_mean_f32 | None) of the same kind as the input."""
    lib = load_library()
    prm = PPStatSplitParams(int(min_width), int(max_width), int(window_width), float(min_gain))
    n_events = len(offsets) - 1
    total = ctypes.c_int64(0)
    on_device = _is_device(current)
    if on_device:
        import torch
        dev = current.device
        assert current.dtype == torch.float64 and offsets.dtype == torch.int64 and offsets.is_cuda
        _torch_inputs_ready(current)
        seg_offsets = torch.empty(n_events + 1, dtype=torch.int64, device=dev)
        mem = PP_MEM_DEVICE
        device = dev.index if dev.index is not None else device

        def alloc(n, dt):
            return torch.empty(n, dtype={"i4": torch.int32, "f8": torch.float64, "f4": torch.float32}[dt], device=dev)
    else:
        current = np.ascontiguousarray(current, dtype=np.float64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        seg_offsets = np.empty(n_events + 1, dtype=np.int64)
        mem = PP_MEM_HOST

        def alloc(n, dt):
            return np.empty(n, dtype=dt)

    def call(cap, bufs):
        rc = lib.pp_statsplit_batch(ctypes.byref(prm), int(device), _ptr(current), _ptr(offsets), n_events, mem,
                                    _ptr(seg_offsets), _ptr(bufs[0]), _ptr(bufs[1]), _ptr(bufs[2]), _ptr(bufs[3]), cap,
                                    ctypes.byref(total))
        return rc

    def make(cap):
        return [alloc(cap, "i4"), alloc(cap, "f8"), alloc(cap, "f8") if want_std else None,
                alloc(cap, "f4") if want_f32 else None]

    if capacity is None:                      # usually n // min_width + 1 segments per event (forced splits can double that:
        lens = (offsets[1:] - offsets[:-1])   # the call is repeated with the exact size if this guess is too small)
        capacity = int((lens // max(1, int(min_width))).sum()) + n_events
    bufs = make(int(capacity))
    rc = call(int(capacity), bufs)
    if rc == PP_ERR_CAPACITY:
        bufs = make(total.value)
        rc = call(total.value, bufs)
    if rc != PP_OK:
        raise EngineError(rc, lib.pp_last_error().decode())
    n = total.value
    return (seg_offsets, bufs[0][:n], bufs[1][:n], None if bufs[2] is None else bufs[2][:n],
            None if bufs[3] is None else bufs[3][:n])


def statsplit_profile(device=0):
    out = (ctypes.c_double * 4)()
    load_library().pp_statsplit_profile(int(device), out)
    return {"statsplit_ms": out[0], "segments_ms": out[1], "launches": int(out[2]), "cumsum_ms": out[3]}


class PPSegmentModel(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("means", _f64p), ("stds", _f64p), ("durs", _f64p),
                ("skip_penalty", ctypes.c_double), ("backslip_penalty", ctypes.c_double)]


def segment_align_trim(device=0):
    """Frees the device buffers pp_segment_align_batch keeps between calls."""
    load_library().pp_segment_align_trim(int(device))


def segment_align_profile():
    """{'kernel_ms', 'launches', 'workspace_bytes', 'groups'} of the last segment_align_batch call."""
    out = (ctypes.c_double * 4)()
    load_library().pp_segment_align_profile(out)
    return {"kernel_ms": out[0], "launches": int(out[1]), "workspace_bytes": out[2], "groups": int(out[3])}


def segment_align_batch(model_means, model_stds, model_durs, skip_penalty, backslip_penalty, seq_means, seq_stds, seq_durs,
                        offsets, device=0):
    """pp_segment_align_batch: (score f64[n], order i32[total], status i32[n]).  Host numpy arrays, or float64 / int64 torch
    tensors resident on the device (results are then device tensors too)."""
    lib = load_library()
    mm, ms, md = (np.ascontiguousarray(a, dtype=np.float64) for a in (model_means, model_stds, model_durs))
    if hasattr(seq_means, "is_cuda") and seq_means.is_cuda:
        import torch
        _torch_inputs_ready(seq_means, seq_stds, seq_durs, offsets)
        n = int(offsets.numel()) - 1
        model = PPSegmentModel(len(mm), mm.ctypes.data_as(_f64p), ms.ctypes.data_as(_f64p), md.ctypes.data_as(_f64p),
                               float(skip_penalty), float(backslip_penalty))
        dev = seq_means.device
        score = torch.empty(n, dtype=torch.float64, device=dev)
        order = torch.empty(int(seq_means.numel()), dtype=torch.int32, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        vp = lambda t: ctypes.c_void_p(t.data_ptr())
        rc = lib.pp_segment_align_batch(ctypes.byref(model), int(dev.index or 0), vp(seq_means), vp(seq_stds), vp(seq_durs), vp(offsets), n,
                                        PP_MEM_DEVICE, vp(score), vp(order), vp(status))
        if rc != 0:
            raise EngineError(rc, lib.pp_last_error().decode())
        return score, order, status
    sm, ss, sd = (np.ascontiguousarray(a, dtype=np.float64) for a in (seq_means, seq_stds, seq_durs))
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    model = PPSegmentModel(len(mm), mm.ctypes.data_as(_f64p), ms.ctypes.data_as(_f64p), md.ctypes.data_as(_f64p),
                           float(skip_penalty), float(backslip_penalty))
    score = np.empty(n, dtype=np.float64)
    order = np.empty(len(sm), dtype=np.int32)
    status = np.empty(n, dtype=np.int32)
    rc = lib.pp_segment_align_batch(ctypes.byref(model), int(device), _ptr(sm), _ptr(ss), _ptr(sd), _ptr(offsets), n, PP_MEM_HOST,
                                    _ptr(score), _ptr(order), _ptr(status))
    if rc != 0:
        raise EngineError(rc, lib.pp_last_error().decode())
    return score, order, status
