"""The segment aligner: oracle pinned to the compiled reference, GPU against both."""
import os

import numpy as np
import pytest

import util
from oracle import segalign_oracle as so


def _golden():
    return np.load(os.path.join(util.GOLDEN, "segalign.npz"))


def _seqs(g):
    off = g["offsets"]
    return [(g["seq_means"][off[i]:off[i + 1]], g["seq_stds"][off[i]:off[i + 1]], g["seq_durs"][off[i]:off[i + 1]])
            for i in range(len(off) - 1)]


def test_oracle_matches_golden_bit_for_bit():
    g = _golden()
    off = g["offsets"]
    for i, (sm, ss, sd) in enumerate(_seqs(g)):
        score, order, status = so.align(g["model_means"], g["model_stds"], g["model_durs"], float(g["skip_penalty"]),
                                        float(g["backslip_penalty"]), sm, ss, sd)
        assert status == 0
        assert score == g["scores"][i]
        assert np.array_equal(order, g["orders"][off[i]:off[i + 1]])


def test_oracle_matches_compiled_reference_on_fresh_cases():
    if not so.reference_available():
        pytest.skip("compiled reference not present (GPU box)")
    ref = so.load_reference()
    rs = np.random.RandomState(5)
    tested = 0
    for _ in range(600):
        m, s = rs.randint(2, 40), rs.randint(1, 80)
        mm = np.sort(rs.uniform(20, 70, m)); ms = rs.uniform(0.5, 2, m); md = rs.uniform(0.001, 0.05, m)
        path = [rs.randint(0, min(3, m))]
        while len(path) < s:
            r, j = rs.rand(), path[-1]
            j = j + 1 if r < 0.6 else (j if r < 0.8 else (j + 2 if r < 0.9 else j - 1))
            path.append(int(np.clip(j, 0, m - 1)))
        sm = mm[path] + rs.normal(0, 0.05, s); ss = rs.uniform(0.5, 2, s); sd = rs.uniform(0.0005, 0.01, s)
        sp, bp = rs.uniform(0.05, 2), rs.uniform(0.05, 2)
        score, order, status = so.align(mm, ms, md, sp, bp, sm, ss, sd)
        if status != 0:
            with pytest.raises(IndexError) if status & 2 else _noraise():
                ref.cSegmentAligner(mm, ms, md, sp, bp).align(sm, ss, sd)
            continue
        rscore, rorder = ref.cSegmentAligner(mm, ms, md, sp, bp).align(sm, ss, sd)
        assert rscore == score and np.array_equal(rorder, order)
        tested += 1
    assert tested > 300


class _noraise(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return a[0] is not None and issubclass(a[0], IndexError)      # bit 0 alone: the reference may or may not survive


@pytest.mark.gpu
def test_gpu_matches_golden_and_oracle():
    """pp_segment_align_batch: scores bit-identical and alignments identical to the compiled reference on the golden batch,
    and to the oracle -- status words included -- on a ragged random batch with undefined-in-the-reference cases."""
    from protpore_b200.alignment import SegmentAligner
    g = _golden()
    off = g["offsets"]
    al = SegmentAligner(g["model_means"], g["model_stds"], g["model_durs"], float(g["skip_penalty"]), float(g["backslip_penalty"]))
    res = al.align_batch(_seqs(g))
    assert (al.last_status == 0).all()
    for i, (score, order) in enumerate(res):
        assert score == g["scores"][i]
        assert np.array_equal(order, g["orders"][off[i]:off[i + 1]])
    one = al.align(*_seqs(g)[3])
    assert one[0] == g["scores"][3] and np.array_equal(one[1], g["orders"][off[3]:off[4]])
    rs = np.random.RandomState(9)
    m = 31
    mm = np.sort(rs.uniform(20, 70, m)); ms = rs.uniform(0.5, 2, m); md = rs.uniform(0.001, 0.05, m)
    seqs = []
    for _ in range(700):
        s = rs.randint(1, 200)
        seqs.append((rs.uniform(15, 75, s), rs.uniform(0.5, 2, s), rs.uniform(0.0005, 0.01, s)))   # arbitrary data: many undefined cases
    al = SegmentAligner(mm, ms, md, 0.3, 1.1)
    res = al.align_batch(seqs)
    n_undef = 0
    for i, (sm, ss, sd) in enumerate(seqs):
        score, order, status = so.align(mm, ms, md, 0.3, 1.1, sm, ss, sd)
        assert al.last_status[i] == status
        assert res[i][0] == score and np.array_equal(res[i][1], order)
        n_undef += status != 0
    assert n_undef > 50
