#!/usr/bin/env python
"""Compact JSON digest of one-kernel-per-row ncu exports: python tools/ncu_digest.py <prefix> [top-lines] > profiles/<name>.json
Reads <prefix>_raw.csv (ncu -i rep --page raw --csv) and, if present, <prefix>_src.csv (--page source --csv --print-source cuda,sass).
Per kernel: key metrics with units, the stall-reason mix of the sampled warps, and the source lines with the most stall samples."""
import csv
import json
import os
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed_op_shared_atom.sum", "sm__cycles_elapsed.avg"]


def main():
    pre = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rows = list(csv.reader(open(pre + "_raw.csv")))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        k = {"kernel": d["Kernel Name"][:110]}
        for key in KEYS:
            if key in d and d[key] != "":
                k[key] = f"{d[key]} {u.get(key, '')}".strip()
        st = {}
        for key, v in d.items():
            if "pcsamp_warps_issue_stalled" in key and not key.endswith("_not_issued"):
                try:
                    st[key.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v.replace(",", ""))
                except ValueError:
                    pass
        tot = sum(st.values()) or 1.0
        k["stall_mix_pct"] = {a: round(100 * b / tot, 1) for a, b in sorted(st.items(), key=lambda kv: -kv[1])[:8]}
        out.append(k)
    if os.path.exists(pre + "_src.csv"):
        kernel = fpath = hdr = None
        data = defaultdict(lambda: defaultdict(lambda: [0, 0, ""]))
        for row in csv.reader(open(pre + "_src.csv", errors="replace")):
            if not row:
                continue
            if row[0] == "Function Name":
                kernel = row[1]; continue
            if row[0] in ("File Path", "File Name"):
                fpath = row[1].split("/")[-1]; continue
            if row[0] == "Line No":
                hdr = row; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
            if hdr is None or not row[0].isdigit():
                continue
            try:
                inst, smp = int(row[ii]), int(row[si])
            except (ValueError, IndexError):
                continue
            v = data[kernel][(fpath, int(row[0]))]
            v[0] += inst; v[1] += smp; v[2] = row[1].strip()
        for i, (kernel, lines) in enumerate(data.items()):
            ti = sum(v[0] for v in lines.values()) or 1
            ts = sum(v[1] for v in lines.values()) or 1
            tl = [{"line": f"{f}:{ln}", "inst_pct": round(100 * v[0] / ti, 1), "samples_pct": round(100 * v[1] / ts, 1), "src": v[2][:100]}
                  for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]]
            tgt = out[i] if i < len(out) else {"kernel": kernel}
            tgt["top_lines_by_stall_samples"] = tl
    json.dump(out, sys.stdout, indent=1)


main()
