import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
from protpore_b200 import _engine
model, (sm, ss, sd), off = bench.align_data(200000, 64, 1)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in (sm, ss, sd, off)]
for _ in range(3):
    _engine.segment_align_batch(model[0], model[1], model[2], 0.3, 1.1, d[0], d[1], d[2], d[3], 0)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    _engine.segment_align_batch(model[0], model[1], model[2], 0.3, 1.1, d[0], d[1], d[2], d[3], 0)
torch.cuda.synchronize()
print("step %.1f ms, kernel %.1f ms" % ((time.perf_counter() - t0) / 5 * 1e3, _engine.segment_align_profile()["kernel_ms"]))
