"""PAF ingest + post-filters at scale (one GPU).  python tools/ingest_bench.py [n_reads] [coverage]"""
import json
import os
import sys
import time

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from consent_b200.engine import Corrector  # noqa: E402
from consent_b200.synth import synth_paf, synth_piles  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 150.0
read_len = 8000
t = time.time()
p = synth_piles(n_reads, genome_len=int(n_reads * read_len / cov), read_len=read_len, seed=42, max_support=4000)
text, names = synth_paf(p, seed=42, tie_range=60)
print(f"generated {n_reads} reads, {len(p.overlaps)} overlaps, {len(text)} bytes of PAF in {time.time() - t:.1f}s", flush=True)
cor = Corrector(device=0)
for i in range(4):
    t = time.time()
    ps = cor.ingest_paf(text, names, 150)
    dt = time.time() - t
    st = cor.ingest_stats()
    print(json.dumps({"step": "ingest_paf", "wall_s": dt, **st, "lines": ps.n_lines, "piles": ps.n_piles, "kept": len(ps.overlaps),
                      "parse_GBps_text": st["paf_bytes"] / max(st["parse_ms"], 1e-6) / 1e6,
                      "kernels_GBps_text": st["paf_bytes"] / max(st["kernel_ms"], 1e-6) / 1e6}), flush=True)
for i in range(2):
    t = time.time()
    ps = cor.ingest_paf(text, names, 150)
    cor.upload_piles(ps.piles(p.store_off, p.store_bases)); cor.run(); res = cor.download()
    batch, reads, _ = cor.download_windows(with_bases=False)
    got = cor.finish_reads(batch, res, reads, 1)
    dt = time.time() - t
    print(json.dumps({"chain_s": dt, "windows": batch.n_windows, "windows_per_s": batch.n_windows / dt, **cor.finish_stats(),
                      "records": int((got.read_off[1:] != got.read_off[:-1]).sum())}), flush=True)
