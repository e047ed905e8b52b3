"""Debug helper: profile forward-backward kernel against the exact state-parallel kernel on a small batch."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import util
from protpore_b200 import events

name = sys.argv[1] if len(sys.argv) > 1 else "pro_001_py2"
ne = int(sys.argv[2]) if len(sys.argv) > 2 else 64
model = util.build_model(name)
eng = model.engine()
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 5
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 60
lens = events.uniform_lengths(ne, lo, hi, seed=7)
obs, offsets = eng.sample_events(lens, seed=99)
obs = obs.astype(np.float64)
logp, edge, mom, mm, _ = eng.forward_backward_batch(obs, offsets)
print("info", eng.forward_backward_info())
eng.set_kernel_path(1)
logp_l, edge_l, mom_l, mm_l, _ = eng.forward_backward_batch(obs, offsets)
print("lane info", eng.forward_backward_info())
print("logp max diff", np.nanmax(np.abs(logp - logp_l)))
d = np.abs(edge - edge_l); print("edge max abs diff", d.max(), "at", d.argmax(), edge[d.argmax()], edge_l[d.argmax()])
d = np.abs(mom - mom_l); print("mom max abs diff", d.max(), "at", d.argmax())
print("mm equal", np.array_equal(mm, mm_l))
b = model._baked
src = np.repeat(np.arange(len(model.states)), np.diff(b.out_ptr))
bad = np.where(np.abs(edge - edge_l) > 1e-9 * (1 + np.abs(edge_l)))[0]
names = [s.name for s in model.states]
for e in bad[:40]:
    print("  edge", e, names[src[e]], "->", names[b.out_idx[e]], edge[e], edge_l[e])
