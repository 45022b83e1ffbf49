#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo at compile time and --import-source on).

  ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv
  python tools/ncu_lines.py src.csv [kernel-regex] [top-N]

Prints, per kernel, the lines with the most warp-level instructions executed and the most stall samples,
plus totals per function group (a crude "which phase costs what" table)."""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    kernel = fpath = None
    hdr = None
    data = defaultdict(lambda: defaultdict(lambda: [0, 0, 0, ""]))   # kernel -> (file, line) -> [inst, samples, thread_inst, src]
    for row in csv.reader(open(path, errors="replace")):
        if not row:
            continue
        if row[0] == "Function Name":
            kernel = row[1]
            continue
        if row[0] in ("File Path", "File Name"):
            fpath = row[1].split("/")[-1]
            continue
        if row[0] == "Kernel Name":
            continue
        if row[0] == "Line No":
            hdr = row
            i_inst = hdr.index("Instructions Executed")
            i_smp = hdr.index("# Samples")
            i_tinst = hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or not row[0] or not row[0].isdigit():
            continue
        try:
            inst = int(row[i_inst]); smp = int(row[i_smp]); tinst = int(row[i_tinst])
        except (ValueError, IndexError):
            continue
        d = data[kernel][(fpath, int(row[0]))]
        d[0] += inst; d[1] += smp; d[2] += tinst; d[3] = row[1].strip()
    for k, lines in data.items():
        if pat and not pat.search(k):
            continue
        tot_i = sum(v[0] for v in lines.values()) or 1
        tot_s = sum(v[1] for v in lines.values()) or 1
        print(f"==== {k[:100]}\n  total warp-inst {tot_i:,}  samples {tot_s:,}")
        print("  -- by instructions executed")
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"  {f}:{ln:<5d} inst {v[0] / tot_i:6.1%} smp {v[1] / tot_s:6.1%} thr/inst {v[2] / max(v[0], 1):5.1f} | {v[3][:110]}")
        print("  -- by stall samples")
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
            print(f"  {f}:{ln:<5d} inst {v[0] / tot_i:6.1%} smp {v[1] / tot_s:6.1%} thr/inst {v[2] / max(v[0], 1):5.1f} | {v[3][:110]}")


if __name__ == "__main__":
    main()
