import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import util
for name in ("synth_12", "pro_001_py2"):
    g = util.load_golden(name)
    model = util.build_model(name)
    evs = util.golden_events(g)
    for j, i in enumerate(g["table_events"]):
        if "fb_ew_%d" % j not in g:
            continue
        et, ew = model.forward_backward(evs[i])
        print(name, j, "len", len(evs[i]), model.engine().forward_backward_info())
        ref = g["fb_ew_%d" % j]
        ok = util.close(ew, ref, abs_tol=1e-8, rel_tol=1e-10)
        bad = np.argwhere(~ok)
        for r, k in bad[:12]:
            print("   row", r, "state", k, model.states[k].name, "ours", ew[r, k], "ref", ref[r, k], "x", evs[i][r])
