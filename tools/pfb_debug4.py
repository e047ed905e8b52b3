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
    eng.set_kernel_path(3)
    lp0, _, _, _, post0 = eng.forward_backward_batch(obs, offsets, posteriors=True)
    print(name, eng.forward_backward_info())
    eng.set_kernel_path(2)
    lpx, _, _, _, postx = eng.forward_backward_batch(obs, offsets, posteriors=True)
    rows = np.repeat(np.isfinite(lpx), np.diff(offsets))
    a, b = post0[rows], postx[rows]
    both = np.isfinite(a) & np.isfinite(b)
    d = np.abs(a - b); d[~both] = 0
    for thr in (-100, -300, -500, -600, -650, -700, -2000):
        sel = both & (b > thr)
        print("  ref >", thr, "max diff", d[sel].max() if sel.any() else None, "count", int(sel.sum()))
    w = np.argsort(-d.ravel())[:8]
    for i in w:
        r, k = np.unravel_index(i, d.shape)
        print("   worst: ours", a[r, k], "ref", b[r, k], "state", model.states[k].name)
    print("  ours finite, ref -inf:", int((np.isfinite(a) & ~np.isfinite(b)).sum()), " ours -inf, ref finite:", int((~np.isfinite(a) & np.isfinite(b)).sum()),
          "max ref among lost", b[~np.isfinite(a) & np.isfinite(b)].max() if (~np.isfinite(a) & np.isfinite(b)).any() else None)
