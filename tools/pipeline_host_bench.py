#!/usr/bin/env python
"""BASELINE configs 4 and 5 through the product's host programs, timed per GPU count, with the md5 of the FASTA checked against the
unmodified reference's (tests/golden/pipeline_md5.json).

  python tools/pipeline_host_bench.py config4|config5 [--gpus 1,2,4,8] [--work DIR] [--repeat 2]

Generates the seeded input (tools/make_pipeline_data.py), runs the vendored minimap2 once (a CPU program, outside the timed part —
as the reference wrapper does before its own C++ binary), then bin/CONSENT-correction / bin/CONSENT-polishing -v on 1..N GPUs.
One JSON line per run: the binary's own timing (file loading, first batch, processing = first batch taken .. last record written)
plus the md5 verdict."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(ROOT, "consent_b200", "bin")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "pipeline_md5.json")))
MM = ["--dual=yes", "-PD", "--no-long-join", "-w5", "-g1000", "-m30", "-n1", "-I1G"]

ap = argparse.ArgumentParser()
ap.add_argument("config", choices=("config4", "config5"))
ap.add_argument("--gpus", default="1")
ap.add_argument("--work", default="/tmp/consent_pipeline_bench")
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--batch-mb", type=int, default=0)
a = ap.parse_args()
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
g.build_cuda(); g.build_bin()
d = os.path.join(a.work, a.config)
os.makedirs(d, exist_ok=True)
mm2 = os.path.join(REF, "minimap2")
nthreads = str(min(os.cpu_count() or 4, 32))
t0 = time.time()
subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_pipeline_data.py"), a.config, d], check=True)
t_gen = time.time() - t0
t0 = time.time()
paf = os.path.join(d, "aln.paf")
if a.config == "config4":
    reads = os.path.join(d, "reads.fasta")
    with open(paf, "wb") as out:
        subprocess.run([mm2] + MM + ["-t" + nthreads, reads, reads], check=True, stdout=out, stderr=subprocess.DEVNULL)
    cmd = [os.path.join(BIN, "CONSENT-correction"), "-a", paf, "-r", reads, "-s", "3", "-S", "150", "-l", "500", "-k", "9", "-c", "8", "-A", "2",
           "-f", "4", "-m", "50", "-M", "150"]
else:
    contigs, reads = os.path.join(d, "contigs.fasta"), os.path.join(d, "reads.fasta")
    raw = os.path.join(d, "raw.paf")
    with open(raw, "wb") as out:
        subprocess.run([mm2] + MM + ["-t" + nthreads, contigs, reads], check=True, stdout=out, stderr=subprocess.DEVNULL)
    srt = os.path.join(d, "sorted.paf")
    with open(srt, "wb") as out:
        subprocess.run(["sort", "-k6,6", raw], check=True, stdout=out, env=dict(os.environ, LC_COLLATE="C"))
    subprocess.run([os.path.join(BIN, "reformatPAF"), srt, paf], check=True)
    cmd = [os.path.join(BIN, "CONSENT-polishing"), "-a", paf, "-r", contigs, "-R", reads, "-s", "1", "-S", "20000", "-l", "500", "-k", "9", "-c", "8",
           "-A", "2", "-f", "4", "-m", "50", "-M", "150"]
t_mm = time.time() - t0
if a.batch_mb:
    cmd += ["-B", str(a.batch_mb)]
print(json.dumps({"config": a.config, "generate_s": round(t_gen, 1), "minimap2_and_paf_prep_s": round(t_mm, 1), "minimap2_threads": int(nthreads),
                  "paf_bytes": os.path.getsize(paf), "paf_md5_matches_golden": hashlib.md5(open(paf, "rb").read()).hexdigest() == GOLD[a.config]["paf_md5"]}), flush=True)
for n in [int(x) for x in a.gpus.split(",")]:
    for rep in range(a.repeat):
        t0 = time.time()
        r = subprocess.run(cmd + ["-v", "-g", ",".join(str(i) for i in range(n))], capture_output=True)
        wall = time.time() - t0
        if r.returncode != 0:
            print(json.dumps({"config": a.config, "gpus": n, "error": r.stderr.decode()[-400:]}), flush=True)
            break
        line = [ln for ln in r.stderr.decode().splitlines() if ln.startswith("{")][-1]
        row = json.loads(line)
        row.update(config=a.config, rep=rep, wall_s=round(wall, 3), records=r.stdout.count(b">"),
                   fasta_md5_equals_reference=hashlib.md5(r.stdout).hexdigest() == GOLD[a.config]["fasta_md5"])
        print(json.dumps(row), flush=True)
