#!/usr/bin/env python
"""Condense an .ncu-rep (via `ncu -i ... --page raw --csv`) into the handful of numbers we track per kernel."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]          # the second row of the raw page carries each metric's unit
unit_of = dict(zip(hdr, units))
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__warps_eligible.avg.per_cycle_active", "idc__request_hit_rate.pct", "sm__icc_requests_lookup_hit.sum", "sm__icc_requests_lookup_miss.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:60])
    for k in keys:
        if k in d and d[k] not in ("", "n/a"):
            print("   %-70s %s %s" % (k, d[k], unit_of.get(k, "")))
    st = [(h, float(v.replace(",", ""))) for h, v in d.items()
          if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
    st.sort(key=lambda t: -t[1])
    print("   stall cycles per issued instruction (sum %.1f):" % sum(v for _, v in st))
    for h, v in st[:8]:
        print("      %-40s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
