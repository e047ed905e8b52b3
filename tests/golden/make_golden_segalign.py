#!/usr/bin/env python
"""Golden vectors of the segment aligner from the COMPILED, UNMODIFIED PyPore/calignment.pyx
(oracle/_ref, built by oracle/build_ref.sh).  Build container only:

    python tests/golden/make_golden_segalign.py

Sequences walk a model with stays, skips and back-slips; only cases the reference itself handles (it raises IndexError
when a path visits model position 0 after the first row, and its argmax is uninitialised when every final score is
below -1) are kept -- the oracle's status word says which."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import segalign_oracle as so  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make_case(rs, m, s):
    mm = np.sort(rs.uniform(20, 70, m)); ms = rs.uniform(0.5, 2, m); md = rs.uniform(0.001, 0.05, m)
    path = [rs.randint(0, min(3, m))]
    while len(path) < s:
        r, j = rs.rand(), path[-1]
        j = j + 1 if r < 0.6 else (j if r < 0.8 else (j + 2 if r < 0.9 else j - 1))
        path.append(int(np.clip(j, 0, m - 1)))
    sm = mm[path] + rs.normal(0, 0.05, s); ss = rs.uniform(0.5, 2, s); sd = rs.uniform(0.0005, 0.01, s)
    return mm, ms, md, sm, ss, sd


def main():
    ref = so.load_reference()
    rs = np.random.RandomState(11)
    m = 24
    mm = np.sort(rs.uniform(20, 70, m)); ms = rs.uniform(0.5, 2, m); md = rs.uniform(0.001, 0.05, m)
    sp, bp = 0.7, 0.4
    aligner = ref.cSegmentAligner(mm, ms, md, sp, bp)
    seqs, scores, orders = [], [], []
    while len(seqs) < 300:
        s = rs.randint(1, 120)
        path = [rs.randint(0, 3)]
        while len(path) < s:
            r, j = rs.rand(), path[-1]
            j = j + 1 if r < 0.55 else (j if r < 0.8 else (j + 2 if r < 0.9 else j - 1))
            path.append(int(np.clip(j, 0, m - 1)))
        sm = mm[path] + rs.normal(0, 0.05, s); ss = rs.uniform(0.5, 2, s); sd = rs.uniform(0.0005, 0.01, s)
        _, _, status = so.align(mm, ms, md, sp, bp, sm, ss, sd)
        if status != 0:
            continue
        score, order = aligner.align(sm, ss, sd)
        seqs.append((sm, ss, sd)); scores.append(score); orders.append(order)
    offsets = np.concatenate([[0], np.cumsum([len(x[0]) for x in seqs])]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "segalign.npz"), model_means=mm, model_stds=ms, model_durs=md, skip_penalty=sp,
                        backslip_penalty=bp, seq_means=np.concatenate([x[0] for x in seqs]), seq_stds=np.concatenate([x[1] for x in seqs]),
                        seq_durs=np.concatenate([x[2] for x in seqs]), offsets=offsets, scores=np.array(scores),
                        orders=np.concatenate(orders))
    print(len(seqs), "sequences,", int(offsets[-1]), "segments")


if __name__ == "__main__":
    main()
