#!/usr/bin/env python
"""Per-warp phase clocks of the linear-profile Viterbi kernel (debug build with -DPP_PHASE_TIMING, never shipped).

    python tools/profile_phase_timing.py --build     # CPU container: builds protpore_b200/libprotpore_b200_timing.so
    python tools/profile_phase_timing.py [events]    # on a GPU box

Per warp, average cycles per row: phase B work (chain warps: their chain; position warps: the two neighbour loads),
wait at the barrier after it, (end-state step), phase A work, wait at the barrier after it."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
csrc = os.path.join(ROOT, "protpore_b200", "csrc")
out = os.path.join(ROOT, "protpore_b200", "libprotpore_b200_timing.so")
if "--build" in sys.argv:
    # only pp_profile.cu carries the clocks; everything else is the regular build
    subprocess.check_call(["make", "-j8", "-C", csrc])
    flags = "-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr".split()
    obj = os.path.join(ROOT, "protpore_b200", "_build", "pp_profile_timing.o")
    subprocess.check_call(["nvcc"] + flags + ["-DPP_PHASE_TIMING", "-c", "-o", obj, os.path.join(csrc, "pp_profile.cu")])
    objs = [os.path.join(ROOT, "protpore_b200", "_build", f) for f in os.listdir(os.path.join(ROOT, "protpore_b200", "_build"))
            if f.endswith(".o") and f not in ("pp_profile.o", "pp_profile_timing.o") and not f.startswith("pp_profile_alt")]
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + [obj, "-lcudart"])
    sys.exit(0)
from protpore_b200 import _engine, peptide_hmm, events
_engine.LIB_PATH = out
lib = _engine.load_library()
lib.pp_debug_profile_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
model = peptide_hmm.pro_001_model()
eng = model.engine()
n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
lens = events.uniform_lengths(n_events, 50, 500, seed=1)
obs, offsets = eng.sample_events(lens, seed=2, device_out=True)
eng.viterbi_batch(obs, offsets, with_paths=False)
lib.pp_debug_profile_phase_cycles(None, 1)
eng.viterbi_batch(obs, offsets, with_paths=False)
print("viterbi kernel ms:", eng.profile())
buf = np.zeros(512, dtype=np.uint64)
lib.pp_debug_profile_phase_cycles(buf.ctypes.data, 0)
buf = buf.reshape(2, 32, 8).astype(np.float64)[0]
print("warp  phaseB  waitB   end  phaseA  waitA   (cycles per row; %d warp-rows)" % buf[0, 6])
for w in range(17):
    if buf[w, 6] == 0:
        continue
    c = buf[w] / buf[w, 6]
    print("%4d %7.0f %6.0f %5.0f %7.0f %6.0f" % (w, c[0], c[1], c[2], c[3], c[4]))
print("row total (warp 1): %.0f cycles" % (buf[1, :5].sum() / buf[1, 6]))
