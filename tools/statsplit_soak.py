"""Segmenter parity soak (GPU box): tests/test_statsplit.py's parameter sweep with many more trials.
usage: python tools/statsplit_soak.py [n_trials] [seed]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_statsplit as T
n, seed = (int(sys.argv[1]) if len(sys.argv) > 1 else 400), (int(sys.argv[2]) if len(sys.argv) > 2 else 99)
mw_max = int(sys.argv[3]) if len(sys.argv) > 3 else 300            # > 680: windows beyond 4,096 samples take the tableless kernel
len_max = int(sys.argv[4]) if len(sys.argv) > 4 else 6000
t0 = time.time()
checked, differ, margin = T.parameter_sweep(n, seed, mw_max=mw_max, len_max=len_max)
print("segmenter soak (min_width < %d, events < %d samples): %d parameter sets, %d events checked against the oracle, %d with different breakpoints "
      "(all within the log() tie margin; smallest margin %s), %.0f s" % (mw_max, len_max, n, checked, differ, margin, time.time() - t0))
