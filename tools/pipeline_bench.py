"""BASELINE config 4 shape: the whole CONSENT-correct pipeline behind the overlapper — PAF text + reads in, corrected FASTA lines on
rank 0 — with the read piles sharded over the GPUs (contiguous blocks of piles in PAF order, no data-path collective,
one ordered gather of the corrected reads).  Strong scaling: the job is fixed, ranks split it.

    python tools/pipeline_bench.py [n_reads] [coverage]                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pipeline_bench.py ...
"""
import json
import os
import sys
import time

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")          # stdout = the JSON line only

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from consent_b200.engine import Corrector  # noqa: E402
from consent_b200.shard import gather_corrected, shard_piles  # noqa: E402
from consent_b200.synth import synth_paf, synth_piles  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
HOST_TAIL = len(sys.argv) > 4 and sys.argv[4] == "host-tail"
read_len = 8000
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p = synth_piles(n_reads, genome_len=int(n_reads * read_len / cov), read_len=read_len, seed=42, max_support=4000)
text, names = synth_paf(p, seed=42, tie_range=60)
cor = Corrector(device=local)
if os.environ.get("CG_CHUNK_WINDOWS"):
    cor.set_option("chunk_max_windows", int(os.environ["CG_CHUNK_WINDOWS"]))


phase = {}


def step():
    t = [time.perf_counter()]
    ps = cor.ingest_paf(text, names, 150)                          # every rank parses the text (1 ms per 26 MB), then owns a block
    p0, p1 = shard_piles(ps.pile_qlen, ps.pile_ov_begin, world)[rank]
    t.append(time.perf_counter())
    cor.upload_piles(ps.piles(p.store_off, p.store_bases, p0, p1))
    t.append(time.perf_counter())
    cor.run()
    t.append(time.perf_counter())
    if HOST_TAIL:                                                  # the tail through the host: results, windows and reads come down first
        res = cor.download()
        batch, reads, _ = cor.download_windows(with_bases=False)
        got = cor.finish_reads(batch, res, reads, 1)
    else:                                                          # cg_finish_resident: only the corrected reads leave the device
        got = cor.finish_resident(1)
    t.append(time.perf_counter())
    full = gather_corrected(got) if world > 1 else got
    t.append(time.perf_counter())
    for k, name in enumerate(("ingest", "upload_piles", "run", "finish", "gather")):
        phase[name] = round(t[k + 1] - t[k], 4)
    return int(cor.counters()["windows"]), p1 - p0, full


for _ in range(2):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(steps):
    W, P, full = step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = torch.tensor([(time.perf_counter() - t0) / steps, float(W), float(P)], dtype=torch.float64, device="cuda")
if world > 1:
    allv = [torch.zeros_like(dt) for _ in range(world)]
    dist.all_gather(allv, dt)
else:
    allv = [dt]
if rank == 0:
    sec = max(float(v[0]) for v in allv)
    Wt = int(sum(float(v[1]) for v in allv))
    import hashlib
    print(json.dumps({"pipeline": ("cg_ingest_paf -> cg_upload_piles -> cg_run -> cg_download -> cg_download_windows -> cg_finish_reads -> gather" if HOST_TAIL
                                   else "cg_ingest_paf -> cg_upload_piles -> cg_run -> cg_finish_resident -> gather"),
                      "n_gpus": world, "reads": n_reads, "coverage": cov, "paf_bytes": len(text), "windows": Wt,
                      "windows_per_rank": [int(v[1]) for v in allv], "s_per_step": sec, "reads_per_s": n_reads / sec, "windows_per_s": Wt / sec,
                      "fasta_records": int((np.diff(full.read_off) > 0).sum()), "corrected_bases": int(full.read_off[-1]),
                      "digest": hashlib.sha256(np.ascontiguousarray(full.bases).tobytes()).hexdigest()[:16], "scaling": "strong", "phase_s_rank0_last_step": phase, "run_ms_device": cor.run_ms(), "reanchor": cor.reanchor_stats(), "finish": cor.finish_stats()}), flush=True)
if world > 1:
    dist.destroy_process_group()
