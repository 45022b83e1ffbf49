"""Window extraction + the whole chain piles -> corrected reads at scale (one GPU).  python tools/extract_bench.py [n_reads] [coverage]"""
import json
import os
import sys
import time

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from consent_b200.engine import Corrector  # noqa: E402
from consent_b200.synth import synth_piles  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
read_len = 8000
genome = int(n_reads * read_len / cov)
t = time.time()
p = synth_piles(n_reads, genome_len=genome, read_len=read_len, seed=42, max_support=150)
print(f"generated {n_reads} reads, {len(p.overlaps)} overlaps, genome {genome} in {time.time() - t:.1f}s", flush=True)
cor = Corrector(device=0)
for i in range(3):
    t = time.time()
    cor.upload_piles(p)
    dt = time.time() - t
    st = cor.extract_stats()
    W = cor.chunk_count()
    print(json.dumps({"step": "upload_piles", "wall_s": dt, **st, "copy_GBps_rw": 2 * st["pile_bytes"] / max(st["copy_ms"], 1e-6) / 1e6}), flush=True)
t = time.time(); cor.run(); t_run = time.time() - t
res = cor.download()
batch, reads, _ = cor.download_windows(with_bases=False)
t = time.time(); got = cor.reanchor_reads(batch, res, reads); t_ra = time.time() - t
print(json.dumps({"windows": batch.n_windows, "seqs": batch.n_seqs, "mean_depth": batch.n_seqs / max(1, batch.n_windows), "run_s": t_run,
                  "run_ms_device": cor.run_ms(), "reanchor_s": t_ra, **cor.reanchor_stats()}), flush=True)
# whole chain, host buffers in (store + overlaps) -> corrected reads out
for i in range(2):
    t = time.time()
    cor.upload_piles(p); cor.run(); res = cor.download(); batch, reads, _ = cor.download_windows(with_bases=False); got = cor.reanchor_reads(batch, res, reads)
    dt = time.time() - t
    print(json.dumps({"chain_s": dt, "windows_per_s": batch.n_windows / dt, "h2d_bytes": int(p.store_bases.nbytes + p.overlaps.nbytes)}), flush=True)
