"""Drop-in for the segment aligner of PyPore/alignment.py:26-47 (`SegmentAligner`, the Python face of the Cython
`cSegmentAligner`, PyPore/calignment.pyx:20-101), on the GPU, with a batched entry point.

    aligner = SegmentAligner(model_means, model_stds, model_durs, skip_penalty, backslip_penalty)
    score, order = aligner.align(seq_means, seq_stds, seq_durs)          # the reference's call and return values
    results = aligner.align_batch([(means, stds, durs), ...])              # many sequences in one GPU call

`order[i]` is the model position segment i is aligned to (a float64 array, as the reference returns it).  Two cases in
which the reference is undefined -- it reads an uninitialised `argmax` when no final score exceeds -1, and it raises
IndexError when the backtrace stands at model position 0 after the first row -- are defined here (first maximum of the
last row; the out-of-range test is skipped); `last_status` tells which sequences were affected.
"""
import numpy as np

from . import _engine


class SegmentAligner(object):
    def __init__(self, model_means, model_stds, model_durs, skip_penalty, backslip_penalty, device=0):
        self.model_means = np.ascontiguousarray(model_means, dtype=np.float64)
        self.model_stds = np.ascontiguousarray(model_stds, dtype=np.float64)
        self.model_durs = np.ascontiguousarray(model_durs, dtype=np.float64)
        self.skip_penalty, self.backslip_penalty = float(skip_penalty), float(backslip_penalty)
        self.device = device
        self.last_status = None

    def align_batch(self, sequences):
        """sequences: iterable of (means, stds, durs).  Returns [(score, order float64[s]), ...]."""
        sequences = [tuple(np.asarray(a, dtype=np.float64) for a in seq) for seq in sequences]
        if not sequences:
            return []
        offsets = np.concatenate([[0], np.cumsum([len(s[0]) for s in sequences])]).astype(np.int64)
        cat = [np.concatenate([s[k] for s in sequences]) for k in range(3)]
        score, order, status = _engine.segment_align_batch(self.model_means, self.model_stds, self.model_durs, self.skip_penalty,
                                                           self.backslip_penalty, cat[0], cat[1], cat[2], offsets, self.device)
        self.last_status = status
        return [(float(score[i]), order[offsets[i]:offsets[i + 1]].astype(np.float64)) for i in range(len(sequences))]

    def align(self, seq_means, seq_stds, seq_durs):
        try:
            return self.align_batch([(seq_means, seq_stds, seq_durs)])[0]
        except ValueError:
            return None, None
