"""Throughput of BASELINE configs 4 and 5 (parity for both is in tests/test_gpu_parity.py)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import torch
from protpore_b200 import peptide_hmm, events
import protpore_b200.yahmm as y

def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps

out = {}
# config 5: pro_001, mixed lengths 20-20,000 (log-uniform), lane kernels
model = peptide_hmm.pro_001_model(); eng = model.engine(); m = len(model.states)
for n_ev in (20000,):
    lens = events.log_uniform_lengths(n_ev, 20, 20000, seed=5)
    obs, off = eng.sample_events(lens, seed=6, device_out=True)
    tv = timed(lambda: eng.viterbi_batch(obs, off)); pv = eng.profile()
    tf = timed(lambda: eng.forward_batch(obs, off)); pf = eng.profile()
    cells = float(lens.sum()) * m
    out["config5_%d" % n_ev] = dict(events=n_ev, observations=int(lens.sum()), max_len=int(lens.max()),
        viterbi_gcups_wall=cells / tv / 1e9, forward_gcups_wall=cells / tf / 1e9,
        viterbi_kernel_ms=pv["viterbi_ms"], traceback_ms=pv["traceback_ms"], forward_kernel_ms=pf["forward_ms"],
        fwd_plus_viterbi_gcups=2 * cells / (tv + tf) / 1e9)
    del obs, off
# config 4: ~2,000-state synthetic profile, 10k-segment events, state-parallel exact kernels
big = peptide_hmm.model_maker(peptide_hmm.synthetic_profile(334, seed=11), "synth334", y, True)
eb = big.engine(); mb = len(big.states)
for n_ev in (96, 1024, 4736):
    lens = np.full(n_ev, 10000)
    obs, off = eb.sample_events(lens, seed=7, device_out=True)
    tv = timed(lambda: eb.viterbi_batch(obs, off, with_paths=(n_ev <= 1024)), reps=1); pv = eb.profile()
    tf = timed(lambda: eb.forward_batch(obs, off), reps=1); pf = eb.profile()
    cells = float(lens.sum()) * mb
    out["config4_%d" % n_ev] = dict(states=mb, events=n_ev, segments_per_event=10000, viterbi_gcups=cells / tv / 1e9,
                                    forward_gcups=cells / tf / 1e9, viterbi_s=tv, forward_s=tf,
                                    viterbi_kernel_ms=pv["viterbi_ms"], traceback_ms=pv["traceback_ms"], forward_kernel_ms=pf["forward_ms"])
    del obs, off
print(json.dumps(out, indent=1))
