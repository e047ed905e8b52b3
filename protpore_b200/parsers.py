"""The segmenters that feed the HMM path (PyPore/parsers.py:505-534, PyPore/cparsers.pyx:45-203), on the GPU.

`SpeedyStatSplit` / `FastStatSplit` keep the reference's constructor arguments and `parse(current)` ->
list of `Segment`; `parse_batch(currents)` is the added batched entry point (one GPU call for many events) and
`segment_means_batch` returns the CSR arrays the batched HMM calls take, without building Python objects.
There is no CPU fallback: without libprotpore_b200.so or a CUDA device `parse` raises.
"""
import numpy as np

from . import _engine
from .pypore import Segment


class FastStatSplit(object):
    """Kevin Karplus' recursive maximum-log-variance-gain segmenter (cparsers.pyx:45-118)."""

    def __init__(self, min_width=100, max_width=1000000, window_width=10000, min_gain_per_sample=None,
                 false_positive_rate=None, prior_segments_per_second=None, sampling_freq=1.e5, cutoff_freq=None,
                 device=0):
        self.min_width = int(min_width)
        self.max_width = int(max_width)
        self.window_width = int(window_width)
        self.sampling_freq = sampling_freq
        # the reference's own assertions (cparsers.pyx:69-76)
        assert self.max_width >= self.min_width, "Maximum width must be greater than minimum width."
        assert self.window_width >= 2 * self.min_width, "Window width must be greater than twice the minimum width."
        if cutoff_freq:
            assert cutoff_freq <= 0.5 * sampling_freq, "Cutoff freq must be less than half the sampling frequency."
        self.min_gain = _engine.statsplit_min_gain(self.window_width, min_gain_per_sample, false_positive_rate,
                                                   prior_segments_per_second, sampling_freq, cutoff_freq)
        self.device = device

    # ---- batched entry points (added) -------------------------------------------------------------------
    def segment_means_batch(self, currents, want_std=False, want_f32=False):
        """(seg_offsets, seg_start, seg_mean, seg_std, seg_mean_f32) for a list of 1-D currents, or for an
        already packed (current, offsets) pair of numpy / torch-CUDA arrays."""
        if isinstance(currents, tuple):
            cur, off = currents
        else:
            cur, off = _engine.pack_events([np.asarray(c, dtype=np.float64) for c in currents], dtype=np.float64)
        return _engine.statsplit_batch(cur, off, self.min_width, self.max_width, self.window_width, self.min_gain,
                                       device=self.device, want_std=want_std, want_f32=want_f32)

    def parse_batch(self, currents):
        """[self.parse(c) for c in currents] in one GPU call."""
        currents = [np.asarray(c, dtype=np.float64) for c in currents]
        seg_off, seg_start, seg_mean, seg_std, _ = self.segment_means_batch(currents, want_std=True)
        out = []
        for e, cur in enumerate(currents):
            a, b = int(seg_off[e]), int(seg_off[e + 1])
            starts = np.append(seg_start[a:b], len(cur)).astype(np.int64)
            out.append([Segment(current=cur[starts[j]:starts[j + 1]], start=int(starts[j]),
                                duration=int(starts[j + 1] - starts[j]), end=int(starts[j + 1]),
                                mean=float(seg_mean[a + j]), std=float(seg_std[a + j]))
                        for j in range(b - a)])
        return out

    # ---- the reference's interface -----------------------------------------------------------------------
    def parse(self, current):
        """Segments of one event (cparsers.pyx:103-118): Segment(current=..., start=, duration=, end=)."""
        return self.parse_batch([current])[0]


class SpeedyStatSplit(FastStatSplit):
    """PyPore/parsers.py:505-534: a thin wrapper that builds a FastStatSplit per parse() call."""
