#!/usr/bin/env python
"""Key metrics + stall reasons + top source lines of one-kernel ncu exports: python tools/ncu_quick.py <prefix> (reads <prefix>_raw.csv, <prefix>_src.csv)"""
import csv, sys
from collections import defaultdict
pre = sys.argv[1]
rows = list(csv.reader(open(pre + "_raw.csv")))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d["Kernel Name"][:70])
    for k in ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
              "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.per_cycle_active",
              "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]:
        if k in d: print("  ", k, d[k])
    st = {}
    for k, v in d.items():
        if "pcsamp_warps_issue_stalled" in k and not k.endswith("_not_issued"):
            try: st[k.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v.replace(",", ""))
            except ValueError: pass
    tot = sum(st.values()) or 1
    print("   stalls:", ", ".join(f"{k} {v / tot:.0%}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
if len(sys.argv) > 2 and sys.argv[2] == "nosrc": sys.exit()
hdr = None; fpath = None
lines = defaultdict(lambda: [0, 0, ""])
for row in csv.reader(open(pre + "_src.csv", errors="replace")):
    if not row: continue
    if row[0] in ("File Path", "File Name"): fpath = row[1].split("/")[-1]; continue
    if row[0] == "Line No":
        hdr = row; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr is None or not row[0].isdigit(): continue
    try: inst = int(row[ii]); smp = int(row[si])
    except (ValueError, IndexError): continue
    v = lines[(fpath, int(row[0]))]; v[0] += inst; v[1] += smp; v[2] = row[1].strip()
ti = sum(v[0] for v in lines.values()) or 1; ts = sum(v[1] for v in lines.values()) or 1
print("total inst", ti, "samples", ts)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{ln:<5d} inst {v[0] / ti:6.1%} smp {v[1] / ts:6.1%} | {v[2][:110]}")
