#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep: headline counters, stall reasons, and the EXECUTED instruction mix per CTA-row
(ncu source page, SASS view): python tools/ncu_opcodes.py report.ncu-rep [cta_rows]."""
import csv, sys, subprocess, collections
rep=sys.argv[1]; nrows=float(sys.argv[2]) if len(sys.argv)>2 else 258000
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","launch__registers_per_thread","smsp__warps_eligible.avg.per_cycle_active"]
for r in rows[2:]:
    d=dict(zip(hdr,r)); u=dict(zip(hdr,units))
    print("==", d["Kernel Name"][:50])
    print("  "+"; ".join("%s=%s"%(k.split(".")[0].replace("sm__inst_executed_pipe_","pipe_").replace("smsp__",""),d[k][:8]) for k in keys if k in d))
    st=[(h,float(v.replace(",",""))) for h,v in d.items() if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio") and v not in ("","n/a")]
    st.sort(key=lambda t:-t[1])
    print("  stalls per issue (sum %.1f):"%sum(v for _,v in st), ", ".join("%s %.2f"%(h.replace("smsp__average_warps_issue_stalled_","").replace("_per_issue_active.ratio",""),v) for h,v in st[:9]))
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
kern=None; data={}
for r in rows:
    if r and r[0]=="Kernel Name": kern=r[1]; data[kern]=[]; continue
    if r and r[0]=="Address": hdr=r; continue
    if kern and len(r)>6: data[kern].append(r)
for k,rs in data.items():
    iS=hdr.index("Source"); iE=hdr.index("Instructions Executed")
    tot=sum(int(r[iE]) for r in rs); agg=collections.Counter()
    for r in rs:
        op=r[iS].split(); o=op[1] if op[0].startswith('@') else op[0]
        o=o.split('.')[0] + ('.'+o.split('.')[1] if o.startswith(('LDS','STS','LDL','STL','LDG','STG','IMAD','BAR')) and '.' in o else '')
        agg[o]+=int(r[iE])
    print("==",k[:40]," instr per CTA-row %.0f"%(tot/nrows))
    print("   "+"  ".join("%s %.0f"%(o,c/nrows) for o,c in agg.most_common(36)))
