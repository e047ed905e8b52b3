"""Debug aid: run the lane forward-backward kernel on mixed-length events and report failures (PP_DEBUG_FB=1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import util
from protpore_b200 import events

model = util.build_model("pro_001_py2")
eng = model.engine()
for lo, hi, n, seed in [(1, 1, 64, 1), (2, 2, 64, 2), (1, 8, 256, 3), (1, 700, 500, 3), (50, 500, 2000, 4)]:
    lens = events.log_uniform_lengths(n, lo, hi, seed=seed) if hi > lo else np.full(n, lo, dtype=np.int64)
    obs, offsets = eng.sample_events(lens, seed=5)
    eng.forward_backward_batch(obs, offsets)
    print(lo, hi, n, eng.forward_backward_info(), flush=True)
