import sys, time; sys.path.insert(0,'.')
import numpy as np
from protpore_b200 import peptide_hmm, events
model = peptide_hmm.pro_001_model(); eng = model.engine()
lens = events.uniform_lengths(20000, 50, 500, seed=3)
obs, off = eng.sample_events(lens, seed=4)
eng.forward_backward_batch(obs[:off[200]], off[:201])
t=time.time(); out = eng.forward_backward_batch(obs, off); dt=time.time()-t
cells = lens.sum()*257
print("FB 20000 events: %.3f s -> %.2f GCUPS (2 passes counted: %.2f)" % (dt, cells/dt/1e9, 2*cells/dt/1e9), eng.profile()["fb_ms"])
