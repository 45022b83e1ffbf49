#!/usr/bin/env python
"""Instruction / stall-sample share per phase of the compact POA tiers from an ncu source-page export:
python tools/poa_phases.py <prefix>_src.csv > profiles/rNN_poa_phases.json   (line ranges of consent_b200/csrc/k_poa2.cuh as committed with it)"""
import csv
import json
import sys
from collections import defaultdict

PHASES = [  # (name, file, first line, last line) — k_poa2.cuh of this commit
    ("exact DFS order (lane 0; on row ties and once before the vote)", "k_poa2.cuh", 183, 330),
    ("traceback (whole warp)", "k_poa2.cuh", 331, 452),
    ("score matrix (rows, scan, per-row maximum)", "k_poa2.cuh", 453, 676),
    ("max search after the DP", "k_poa2.cuh", 677, 705),
    ("job setup: segments, staging, repeat check", "k_poa2.cuh", 721, 790),
    ("tier dispatch, winning cell, row ties", "k_poa2.cuh", 791, 881),
    ("graph update U1-U2 (all lanes)", "k_poa2.cuh", 882, 1013),
    ("order splice + row descriptors U3-U4 (all lanes)", "k_poa2.cuh", 1014, 1123),
    ("vote", "k_poa2.cuh", 1124, 1170),
    ("queue", "k_poa2.cuh", 1171, 1210),
    ("id packing helpers (inlined everywhere)", "k_poa2.cuh", 37, 182),
]


def main():
    kernel = fpath = hdr = None
    data = defaultdict(lambda: defaultdict(lambda: [0, 0, 0]))
    for row in csv.reader(open(sys.argv[1], errors="replace")):
        if not row:
            continue
        if row[0] == "Function Name":
            kernel = row[1]; continue
        if row[0] in ("File Path", "File Name"):
            fpath = row[1].split("/")[-1]; continue
        if row[0] == "Line No":
            hdr = row; ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); ti = hdr.index("Thread Instructions Executed"); continue
        if hdr is None or not row[0].isdigit():
            continue
        try:
            inst, smp, tin = int(row[ii]), int(row[si]), int(row[ti])
        except (ValueError, IndexError):
            continue
        ln = int(row[0])
        name = "other (intrinsics headers, cg_common.cuh)"
        for pn, pf, a, b in PHASES:
            if fpath == pf and a <= ln <= b:
                name = pn; break
        v = data[kernel][name]
        v[0] += inst; v[1] += smp; v[2] += tin
    out = {}
    for kernel, ph in data.items():
        ti = sum(v[0] for v in ph.values()) or 1
        ts = sum(v[1] for v in ph.values()) or 1
        tier = "C1" if "(int)0" in kernel else "G" if "(int)1" in kernel else kernel[:40]
        out[tier] = {n: {"inst_pct": round(100 * v[0] / ti, 1), "samples_pct": round(100 * v[1] / ts, 1), "threads_per_inst": round(v[2] / max(1, v[0]), 1)}
                     for n, v in sorted(ph.items(), key=lambda kv: -kv[1][0])}
        single = sum(v[0] for v in ph.values() if v[0] and v[2] / v[0] < 4.0)
        out[tier]["_single_lane_share_of_instructions_pct"] = round(100 * single / ti, 1)
    json.dump(out, sys.stdout, indent=1)


main()
