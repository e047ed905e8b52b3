"""One Viterbi + forward batch of BASELINE config 4 (2,003-state profile, 10,000-segment events) -- the ncu target for
the global-column lane kernels.  usage: python tools/config4_run.py [n_events] [reps]"""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from protpore_b200 import peptide_hmm
import protpore_b200.yahmm as y
n_ev = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
big = peptide_hmm.model_maker(peptide_hmm.synthetic_profile(334, seed=11), "synth334", y, True)
eng = big.engine()
obs, off = eng.sample_events(np.full(n_ev, 10000), seed=7, device_out=True)
for _ in range(reps):
    eng.viterbi_batch(obs, off, with_paths=False)
    pv = eng.profile()
    eng.forward_batch(obs, off)
    pf = eng.profile()
torch.cuda.synchronize()
cells = n_ev * 10000.0 * len(big.states)
print("events", n_ev, "viterbi_ms", pv["viterbi_ms"], "GCUPS", cells / pv["viterbi_ms"] / 1e6, "forward_ms", pf["forward_ms"], "GCUPS", cells / pf["forward_ms"] / 1e6)
