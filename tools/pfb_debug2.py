import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import util
from protpore_b200 import events
for name in ("pro_001_py2", "pro_001_py3", "synth_12"):
    model = util.build_model(name)
    eng = model.engine()
    lens = events.log_uniform_lengths(700, 1, 600, seed=41)
    obs, offsets = eng.sample_events(lens, seed=42)
    obs = obs.astype(np.float64)
    obs[offsets[5]:offsets[5] + 1] = 250.0
    obs[offsets[9] + 2] = 119.5
    print(name, "len5", lens[5], "len9", lens[9], flush=True)
    lp0, e0, m0, mm0, _ = eng.forward_backward_batch(obs, offsets)
    print(name, eng.forward_backward_info(), flush=True)
    # without the outliers
    obs2, _ = eng.sample_events(lens, seed=42)
    eng.forward_backward_batch(obs2.astype(np.float64), offsets)
    print(name, "clean:", eng.forward_backward_info(), flush=True)
    if name == "synth_12":
        o2 = obs2.astype(np.float64)
        for e in (15,):
            print("event", e, "obs", o2[offsets[e]:offsets[e+1]])
        order = np.argsort(-lens, kind="stable")
        print("last group events", order[-28:], "lens", lens[order[-28:]])
        eng.set_kernel_path(1)
        lp = eng.forward_backward_batch(o2, offsets)[0]
        print("lane logp of 15:", lp[15], eng.forward_backward_info())
