"""Peptide profile-HMM builder and per-event driver (mirror of peptide_hmm_v0.0.1.py:37-262).

Same public names and argument meaning as the reference script (`kmer_current_map`,
`HMM_Constructor.HMM_linear_model`, `model_maker`, `prediction`) so it reads as a drop-in; the
fast5 reading / segmentation / plotting around it (299-494) is out of scope.

Two additions:
  * `yahmm_module=` lets the same builder construct the model on any module exposing the yahmm API
    (the product's `protpore_b200.yahmm`, or the compiled reference module used for the CPU
    baseline), which is how tests prove the two bakes are identical;
  * `py2_division=` makes the Python-2 integer division at peptide_hmm_v0.0.1.py:194-195 explicit:
    the reference README mandates Python 2.7, where `1/4` and `1/2` are 0, so every insert state is
    NormalDistribution(mu, 0.0).  True (default) reproduces that; False gives the mixture sd.

The order of `add_state` / `add_transition` calls is part of the contract: it fixes the baked state
order and the Viterbi tie order (see yahmm/bake.py), so it follows the reference call for call.
"""
import os

import numpy

PRO_001 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "pro_001.txt")


def _default_yahmm():
    from . import yahmm
    return yahmm


def kmer_current_map(file):
    """Read `index<TAB>mean<TAB>sd` rows into {key: [mean, sd]} (peptide_hmm_v0.0.1.py:37-52)."""
    table = {}
    with open(file, 'r') as handle:
        for line in handle:
            cols = line.strip().split('\t')
            table[cols[0].strip()] = [float(cols[1].strip()), float(cols[2].strip())]
    return table


class HMM_Constructor():
    """Builds the linear match / drop-off / blip / back-slip / skip / insert profile HMM."""

    # transition constants, verbatim from peptide_hmm_v0.0.1.py:90-118
    SELF_LOOP = 0.1
    END = 0.05
    DROP = 0.05
    BLIP = 0.05
    BLIP_SELF = 0.05
    STEP_BACK = 0.05
    SKIP = 0.1
    LONG_SKIP = 0.05
    INS_SELF = 0.05
    BLIP_RANGE = (15.0, 120.0)

    def __init__(self, yahmm_module=None, py2_division=True):
        self.yahmm = yahmm_module or _default_yahmm()
        self.py2_division = py2_division

    def _insert_sd(self, prev_mean, prev_sd, mean, sd):
        # reference: numpy.sqrt(1/4 * (dm ** 2) + 1/2 * (prev_sd ** 2 + sd ** 2))   (194-195)
        quarter, half = (0, 0) if self.py2_division else (0.25, 0.5)
        return numpy.sqrt(quarter * ((prev_mean - mean) ** 2) + half * (prev_sd ** 2 + sd ** 2))

    def HMM_linear_model(self, kmer_list, kmer_current_dict, model_name=None):
        y = self.yahmm
        model = y.Model(name=model_name if model_name is not None else 'HMM_linear_model')
        last = len(kmer_list) - 1
        matches = []
        prev_skip = prev_back = None
        prev_mean = prev_sd = 0

        for index, kmer in enumerate(kmer_list):
            mean, sd = kmer_current_dict[kmer]
            tag = kmer + '_' + str(index)
            slip = 0.05 if index > 0 else 0.00
            insert = 0.05 if index > 0 else 0.00
            forward = 1 - (self.SELF_LOOP + self.END + self.BLIP + slip + self.SKIP + insert)

            match = y.State(y.NormalDistribution(mean, sd), name='M_' + tag)
            model.add_state(match)
            model.add_transition(match, match, self.SELF_LOOP)
            if index < last:
                model.add_transition(match, model.end, self.END)

            drop_off = y.State(None, name='S_DROPOFF_' + tag)
            model.add_state(drop_off)
            model.add_transition(match, drop_off, self.DROP)
            model.add_transition(drop_off, match, 1.0 - self.BLIP_SELF)
            model.add_transition(drop_off, model.end, 1.00)

            blip = y.State(y.UniformDistribution(*self.BLIP_RANGE), name='I_BLIP_' + tag)
            model.add_state(blip)
            model.add_transition(blip, blip, self.BLIP_SELF)
            model.add_transition(match, blip, self.BLIP)
            model.add_transition(blip, match, 1.0 - self.BLIP_SELF)

            back = None
            if index >= 1:
                back = y.State(None, name='B_BACK_SHORT_' + tag)
                model.add_state(back)
                model.add_transition(match, back, slip)
                if index >= 2:
                    model.add_transition(back, matches[index - 1], self.STEP_BACK)
                    model.add_transition(back, prev_back, 1 - self.STEP_BACK)
                else:
                    model.add_transition(back, matches[index - 1], 1.00)

            skip = y.State(None, name='S_SKIP_' + tag)
            model.add_state(skip)
            if prev_skip is not None:
                model.add_transition(prev_skip, skip, self.LONG_SKIP)
                model.add_transition(prev_skip, match, 1 - self.LONG_SKIP)
            if index == 0:
                model.add_transition(model.start, skip, 1.0 - forward)
            else:
                model.add_transition(matches[index - 1], skip, self.SKIP)

            if index > 0:
                ins = y.State(y.NormalDistribution((prev_mean + mean) / 2.0,
                                                   self._insert_sd(prev_mean, prev_sd, mean, sd)),
                              name='I_INS_' + tag)
                model.add_state(ins)
                model.add_transition(ins, ins, self.INS_SELF)
                model.add_transition(matches[index - 1], ins, insert)
                model.add_transition(ins, match, 1.0 - self.INS_SELF)

            if index == 0:
                model.add_transition(model.start, match, forward)
            elif index == 1:
                model.add_transition(matches[0], match, forward + slip)
            else:
                model.add_transition(matches[index - 1], match, forward)

            matches.append(match)
            prev_skip = skip
            prev_back = back          # None at index 0, as in the reference
            prev_mean, prev_sd = mean, sd

            if index == last:
                # skip = insert = 0 for the closing match state (234-240)
                closing = 1 - (self.SELF_LOOP + self.END + self.BLIP + slip + 0.0 + 0.0)
                model.add_transition(match, model.end, closing + self.END)
                model.add_transition(prev_skip, model.end, 1.00)

        model.bake()
        return model


def model_maker(kmer_current_dict, model_name=None, yahmm_module=None, py2_division=True):
    """peptide_hmm_v0.0.1.py:245-249 (the profile rows are keyed '0'..'K-1')."""
    kmer_list = [str(i) for i in range(len(kmer_current_dict))]
    return HMM_Constructor(yahmm_module, py2_division).HMM_linear_model(
        kmer_list, kmer_current_dict, model_name)


def pro_001_model(yahmm_module=None, py2_division=True, model_name='pro_001'):
    """The benchmark model: profile HMM of the packaged pro_001 table (257 states, 765 edges)."""
    return model_maker(kmer_current_map(PRO_001), model_name, yahmm_module, py2_division)


def synthetic_profile(rows, seed=0):
    """Synthetic profile table for the scaled model of BASELINE config 4 (means~U(25,50), sds~U(0.4,2.5))."""
    rng = numpy.random.RandomState(seed)
    means = rng.uniform(25.0, 50.0, size=rows)
    sds = rng.uniform(0.4, 2.5, size=rows)
    return {str(i): [float(means[i]), float(sds[i])] for i in range(rows)}


def prediction(models, sequences, algorithm='forward-backward'):
    """Per-event loop of the reference driver (peptide_hmm_v0.0.1.py:251-262), batched underneath.

    Returns, in the reference's order (sequence-major, model-minor), what `model.viterbi(seq)` or
    `model.forward_backward(seq)` would return; models exposing `viterbi_batch` /
    `forward_backward_batch` get ONE batched GPU call per model instead of one call per event.
    """
    per_model = []
    for model in models:
        if algorithm == 'viterbi':
            batch = getattr(model, 'viterbi_batch', None)
            per_model.append(batch(sequences) if batch else [model.viterbi(s) for s in sequences])
        elif algorithm == 'forward-backward':
            batch = getattr(model, 'forward_backward_batch', None)
            per_model.append(batch(sequences) if batch else [model.forward_backward(s) for s in sequences])
        else:
            per_model.append(None)
    out = []
    for i in range(len(sequences)):
        for res in per_model:
            if res is not None:
                out.append(res[i])
    return out
