"""Throughput probe of the GPU segmenter: N synthetic events of ~18k samples (the CLI's slice, peptide_hmm_v0.0.1.py:399-400)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from protpore_b200 import parsers, _engine
import torch

n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rng = np.random.default_rng(1)
def ev():
    return np.concatenate([rng.normal(rng.uniform(25, 50), rng.uniform(0.4, 2.5), size=int(rng.integers(260, 1500))) for _ in range(21)])
base = [ev() for _ in range(64)]
currents = [base[i % 64] for i in range(n_events)]
cur, off = _engine.pack_events(currents, dtype=np.float64)
p = parsers.SpeedyStatSplit(min_width=250, max_width=1000, window_width=2000, min_gain_per_sample=0.05)
dc, do = torch.from_numpy(cur).cuda(), torch.from_numpy(off).cuda()
for it in range(3):
    torch.cuda.synchronize(); t = time.time()
    out = p.segment_means_batch((dc, do), want_std=False, want_f32=True)
    torch.cuda.synchronize(); dt = time.time() - t
    prof = _engine.statsplit_profile(0)
    print("events", n_events, "samples", len(cur), "segments", int(out[0][-1]), "wall_ms", dt * 1e3, prof,
          "Msamples/s (kernels)", len(cur) / (prof["statsplit_ms"] + prof["segments_ms"]) / 1e3)
from oracle import statsplit_oracle as so
t = time.time(); n = 0
for c in base[:16]:
    bp, st = so.parse(c, 250, 1000, 2000, gain=p.min_gain, with_stats=True); n += len(c)
dt = time.time() - t
print("oracle 1 core: Msamples/s", n / dt / 1e6, "positions/sample", st[2] / len(c))
