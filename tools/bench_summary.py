#!/usr/bin/env python
"""One-line summary of bench.py JSON lines read from stdin (tag = argv[1])."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    cpu = d.get("cpu_baseline") or {}
    rf = d.get("roofline") or {}
    print(tag, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2),
          "stages", rf.get("stage_ms_per_step"), "roof", rf.get("kernel"), round(rf.get("frac", 0), 4),
          "cpu", cpu.get("value") and round(cpu["value"]), cpu.get("kind"), "parity", cpu.get("parity_spot_check"),
          "clk", d.get("clocks", {}).get("sm_mhz"), flush=True)
