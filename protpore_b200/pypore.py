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
