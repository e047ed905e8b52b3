#!/usr/bin/env python
"""Static SASS instruction mix of the sweep kernels in the built library (cuobjdump -sass): which opcodes the compiler
emitted for each kernel, so that the executed mix of an ncu capture (tools/ncu_opcodes.py) can be attributed.

    python tools/sass_histogram.py [protpore_b200/libprotpore_b200.so] > profiles/r2_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "protpore_b200/libprotpore_b200.so"
want = ("k_profile_viterbi", "k_profile_forward", "k_profile_fb", "k_viterbi_lane", "k_forward_lane", "k_fb_lane", "k_traceback", "k_statsplit", "k_segment_align")
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, {}
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = name if any(w in name for w in want) else None
        if kern:
            hist[kern] = collections.Counter()
        continue
    if kern:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m:
            op = m.group(1)
            if op in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR"):
                op += (m.group(2).split(".")[1:2] or [""])[0] and "." + m.group(2).split(".")[1]
            hist[kern][op] += 1
for k, h in hist.items():
    total = sum(h.values())
    print("== %s\n   %d instructions; " % (k[:110], total) + "  ".join("%s %d" % kv for kv in h.most_common(30)))
    fp64 = sum(v for o, v in h.items() if o in ("DADD", "DMUL", "DFMA", "DSETP"))
    print("   float64 pipe %d (%.0f%%), shared-memory loads/stores %d, local (spill) %d, barriers %d" % (
        fp64, 100.0 * fp64 / total, sum(v for o, v in h.items() if o.startswith(("LDS", "STS"))),
        sum(v for o, v in h.items() if o.startswith(("LDL", "STL"))), sum(v for o, v in h.items() if o.startswith("BAR"))))
