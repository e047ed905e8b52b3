#!/usr/bin/env python
"""Per-warp phase clocks of the Viterbi lane kernel (debug build with -DPP_PHASE_TIMING, never shipped).

    python tools/phase_timing.py            # on a GPU box; builds protpore_b200/libprotpore_b200_timing.so

Prints, per warp, the average cycles per row spent: working on silent runs, working on early-emitting runs,
waiting at phase B's barrier, working on late-emitting runs (phase A), waiting at phase A's barrier."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
csrc = os.path.join(ROOT, "protpore_b200", "csrc")
out = os.path.join(ROOT, "protpore_b200", "libprotpore_b200_timing.so")
if not os.path.exists(out) or "--build" in sys.argv:        # build in the CPU container: nvcc cross-compiles, the .so travels
    subprocess.check_call(["make", "-j6", "-C", csrc, "OUT=" + out, "OBJDIR=../_build_timing", "NVFLAGS=-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a "
                           "-Xcompiler -fPIC --expt-relaxed-constexpr -DPP_PHASE_TIMING"])
    if "--build" in sys.argv:
        sys.exit(0)
from protpore_b200 import _engine, peptide_hmm, events
_engine.LIB_PATH = out
lib = _engine.load_library()
lib.pp_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
model = peptide_hmm.pro_001_model()
eng = model.engine()
lens = events.uniform_lengths(20000, 50, 500, seed=1)
obs, offsets = eng.sample_events(lens, seed=2)
eng.viterbi_batch(obs, offsets, with_paths=False)
lib.pp_debug_phase_cycles(None, 1)
eng.viterbi_batch(obs, offsets, with_paths=False)
buf = np.zeros(128, dtype=np.uint64)
lib.pp_debug_phase_cycles(buf.ctypes.data, 0)
buf = buf.reshape(16, 8).astype(np.float64)
rows = buf[:, 6]
print("warp  silent  early  waitB  finalgap  late(A)  waitA   (cycles per row, %d warp-rows)" % rows[0])
for w in range(8):
    c = buf[w] / max(rows[w], 1)
    print("%4d %7.0f %6.0f %6.0f %9.0f %8.0f %6.0f" % (w, c[0], c[1], c[2], c[5], c[3], c[4]))
print("row total (warp 0): %.0f cycles" % (buf[0, :6].sum() / rows[0]))
