"""The PyPore caller side of the hot path (PyPore/core.py:115-261, PyPore/DataTypes.py:49-68, 349-355).

Only what feeds the HMM is mirrored: a `Segment` with a `.mean`, and an `Event` holding segments whose
`apply_hmm(hmm, algorithm)` is the reference's one-liner.  `apply_hmm_batch` is the added batched
sibling: one GPU call for many events.
"""
import numpy as np


class Segment(object):
    """A run of ionic current (PyPore/core.py:115-261): `.mean`, `.std`, `.duration`, `.start`."""

    def __init__(self, current=None, **kwargs):
        self.current = None if current is None else np.asarray(current, dtype=np.float64)
        for key, value in kwargs.items():
            setattr(self, "_" + key if key in ("mean", "std", "min", "max", "n") else key, value)

    @property
    def mean(self):
        if getattr(self, "_mean", None) is not None:
            return self._mean
        return float(np.mean(self.current))

    @property
    def std(self):
        if getattr(self, "_std", None) is not None:
            return self._std
        return float(np.std(self.current))

    @property
    def n(self):
        if getattr(self, "_n", None) is not None:
            return self._n
        return len(self.current)


class Event(Segment):
    """An event = a list of segments (PyPore/DataTypes.py:49-68)."""

    def __init__(self, current=None, segments=None, **kwargs):
        Segment.__init__(self, current, **kwargs)
        self.segments = list(segments) if segments is not None else []

    def segment_means(self):
        return np.array([seg.mean for seg in self.segments])

    def parse(self, parser, hmm=None):
        """Split the event's current into segments (PyPore/DataTypes.py:292-332 without the hmm-merging branch)."""
        self.segments = parser.parse(self.current)
        self.n = len(self.segments)
        return self.segments

    def apply_hmm(self, hmm, algorithm='viterbi'):
        """`getattr(hmm, algorithm)(segment means)` (PyPore/DataTypes.py:349-355, 62-68)."""
        return getattr(hmm, algorithm)(self.segment_means())


def apply_hmm_batch(events, hmm, algorithm='viterbi'):
    """What `[event.apply_hmm(hmm, algorithm) for event in events]` returns, in one batched GPU call."""
    means = [ev.segment_means() for ev in events]
    if algorithm == 'viterbi':
        return hmm.viterbi_batch(means, dtype=np.float64)
    if algorithm == 'log_probability':
        return [float(v) for v in hmm.log_probability_batch(means, dtype=np.float64)]
    if algorithm == 'forward_backward':
        return hmm.forward_backward_batch(means, dtype=np.float64)
    return [getattr(hmm, algorithm)(x) for x in means]


def segment_and_decode(currents, hmm, parser, algorithm='viterbi', device_resident=True):
    """Raw current -> segments -> HMM, for many events, in two batched GPU calls.

    What peptide_hmm_v0.0.1.py:443-472 does per event -- `event.parse(SpeedyStatSplit(...))`, collect the segment means,
    `prediction(models, [means], 'viterbi')` -- for a list of 1-D currents.  With `device_resident` the segment offsets
    and float32 means produced by `pp_statsplit_batch` stay on the GPU and are the `offsets` / `obs` of the batched
    HMM call (needs torch); otherwise they round-trip through host numpy arrays.

    Returns (seg_offsets, seg_start, seg_mean, result) with `result` = (scores, path_len, paths) for 'viterbi' or the
    log-probabilities for 'log_probability'."""
    from . import _engine
    cur, off = _engine.pack_events([np.asarray(c, dtype=np.float64) for c in currents], dtype=np.float64)
    if device_resident:
        import torch
        dev = torch.device("cuda", parser.device)
        cur, off = torch.from_numpy(cur).to(dev), torch.from_numpy(off).to(dev)
    seg_off, seg_start, seg_mean, _, mean32 = parser.segment_means_batch((cur, off), want_std=False, want_f32=True)
    eng = hmm.engine()
    if algorithm == 'viterbi':
        result = eng.viterbi_batch(mean32, seg_off)
    elif algorithm == 'log_probability':
        result = eng.forward_batch(mean32, seg_off)
    else:
        raise ValueError("algorithm must be 'viterbi' or 'log_probability'")
    host = lambda a: a.cpu().numpy() if hasattr(a, "cpu") else a
    result = tuple(host(r) for r in result) if isinstance(result, tuple) else host(result)
    return host(seg_off), host(seg_start), host(seg_mean), result
