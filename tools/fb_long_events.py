import sys, time, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import util
from protpore_b200 import events
m = util.build_model("pro_001_py2"); eng = m.engine()
lens = events.log_uniform_lengths(600, 20, 20000, seed=5)
obs, off = eng.sample_events(lens, seed=6)
t0 = time.perf_counter(); lp0, e0, m0, mm0, _ = eng.forward_backward_batch(obs, off); t1 = time.perf_counter()
print("profile:", eng.forward_backward_info(), "%.0f ms" % ((t1 - t0) * 1e3), "rows", int(lens.sum()))
eng.set_kernel_path(1)
t0 = time.perf_counter(); lp1, e1, m1, mm1, _ = eng.forward_backward_batch(obs, off); t1 = time.perf_counter()
print("lane:", eng.forward_backward_info(), "%.0f ms" % ((t1 - t0) * 1e3))
fin = np.isfinite(lp1) & np.isfinite(lp0)
print("logp", np.max(np.abs(lp0[fin] - lp1[fin]) / np.abs(lp1[fin])), "edge rel", np.max(np.abs(e0 - e1) / np.maximum(1e-3, np.abs(e1))),
      "mom rel", np.max(np.abs(m0 - m1) / np.maximum(1e-3, np.abs(m1))), "mm", np.array_equal(mm0, mm1), "nonfinite", int((~np.isfinite(lp0)).sum()), int((~np.isfinite(lp1)).sum()))
