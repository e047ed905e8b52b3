"""Where the end-to-end call spends its time: host buffers in, host results out (pp_score_batch, PP_MEM_HOST)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from protpore_b200 import peptide_hmm, events
from protpore_b200._engine import pinned_empty
model = peptide_hmm.pro_001_model(); eng = model.engine()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
lens = events.uniform_lengths(n, 50, 500, seed=2026)
obs_d, off_d = eng.sample_events(lens, seed=3026, device_out=True)
s, l, p, f = eng.score_batch(obs_d, off_d)
print("device:", eng.profile())
path_elems = int(p.numel())
obs_h = obs_d.cpu().pin_memory().numpy(); off_h = off_d.cpu().numpy()
outs = (pinned_empty(n, np.float64), pinned_empty(n, np.int64), pinned_empty(path_elems + 1024, np.uint16), pinned_empty(n, np.float64))
eng.score_batch(obs_h, off_h, out=outs)
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    eng.score_batch(obs_h, off_h, out=outs)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("e2e %.1f ms" % (dt * 1e3), eng.profile())
