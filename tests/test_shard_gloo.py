"""CPU suite, part 4: the multi-GPU host logic (window sharding + the final ordered gather), world_size 2 over gloo.
Each rank computes its block with the oracle standing in for its GPU (this test checks the plumbing, not the kernels)."""
import os
import subprocess
import sys

import numpy as np

from consent_b200.shard import shard_by_bases, shard_range
from consent_b200.synth import synth_windows
from tests.cases import concat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch.distributed as dist
from consent_b200.shard import gather_results, shard_range
from consent_b200.synth import synth_windows
from tests.refs import Oracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
batch = synth_windows(11, 5, seed=71)
w0, w1 = shard_range(batch.n_windows, rank, 2)
orc = Oracle()
local, _ = orc.correct_windows(batch.slice(w0, w1))
full = gather_results(local)
if rank == 0:
    want, _ = orc.correct_windows(batch)
    assert full.equals(want), "gathered results differ from the single-process run"
    print("GATHER_OK", full.n_windows)
else:
    assert full is None
parts = gather_results(local, with_solid=False, concat=False)
if rank == 0:
    assert parts.n_windows == want.n_windows
    assert all(parts.consensus(w) == want.consensus(w) and parts.status(w) == int(want.status[w]) for w in range(want.n_windows))
    print("PARTS_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 100, 100001):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_shard_by_bases_balances_mixed_depths():
    batch = concat([synth_windows(20, 3, seed=1), synth_windows(4, 60, seed=2)])
    blocks = shard_by_bases(batch, 4)
    assert blocks[0][0] == 0 and blocks[-1][1] == batch.n_windows
    assert all(blocks[i][1] == blocks[i + 1][0] for i in range(3))
    loads = [int(batch.seq_off[int(batch.win_seq_begin[b])]) - int(batch.seq_off[int(batch.win_seq_begin[a])]) for a, b in blocks]
    assert max(loads) <= 0.5 * sum(loads)


def test_gather_world_size_2_gloo(entry, tmp_path):
    port = 29500 + os.getpid() % 2000
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "GATHER_OK 11" in outs[0] and "PARTS_OK" in outs[0]


PIPE_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from consent_b200.shard import gather_corrected, shard_piles
from consent_b200.synth import synth_paf, synth_piles
from tests.refs import Oracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
orc = Oracle()
p = synth_piles(n_reads=16, genome_len=6000, read_len=1800, seed=9, max_support=4000)
text, names = synth_paf(p, seed=3, tie_range=2)
ps = orc.ingest_paf(text, names, 8)                       # every rank parses the (small) PAF, then owns a block of piles

def run(p0, p1):
    b, rd, _ = orc.extract_windows(ps.piles(p.store_off, p.store_bases, p0, p1))
    res, _ = orc.correct_windows(b, threads=4)
    return orc.finish_reads(orc.reanchor_reads(b, res, rd, threads=2)[0], 1)

blocks = shard_piles(ps.pile_qlen, ps.pile_ov_begin, 2)
local = run(*blocks[rank])
full = gather_corrected(local)
if rank == 0:
    want = run(0, ps.n_piles)
    assert blocks[0][1] > 0 and blocks[1][1] == ps.n_piles and blocks[0][1] == blocks[1][0]
    assert full.equals(want), "gathered corrected reads differ from the single-process run"
    print("PIPELINE_OK", full.n_reads, int((np.diff(full.read_off) > 0).sum()))
else:
    assert full is None
dist.barrier()
dist.destroy_process_group()
"""


def test_pipeline_sharded_by_piles_world_size_2_gloo(entry, tmp_path):
    """BASELINE config 4 shape, small: PAF -> piles sharded over two ranks -> windows -> consensuses -> re-anchored, trimmed reads,
    gathered to rank 0 in PAF order = the single-process result (oracle standing in for the GPUs)."""
    port = 31500 + os.getpid() % 2000
    script = tmp_path / "pipe_worker.py"
    script.write_text(PIPE_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "PIPELINE_OK 16" in outs[0]


POLISH_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from consent_b200.shard import gather_results, shard_by_bases
from consent_b200.synth import synth_paf, synth_piles
from tests.refs import Oracle
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
orc = Oracle()
p = synth_piles(n_reads=40, genome_len=6000, read_len=3000, n_piles=2, seed=19, profile="ONT", max_support=4000)   # 2 "contigs"
text, names = synth_paf(p, seed=4, tie_range=3)
ps = orc.ingest_paf(text, names, 30)
batch, reads, _ = orc.extract_windows(ps.piles(p.store_off, p.store_bases))          # every rank cuts all windows (cheap) ...
w0, w1 = shard_by_bases(batch, 2)[rank]                                             # ... and corrects its block of them
local, _ = orc.correct_windows(batch.slice(w0, w1), threads=4)
full = gather_results(local)                                                         # consensuses + solid k-mers to rank 0
if rank == 0:
    want_res, _ = orc.correct_windows(batch, threads=4)
    assert full.equals(want_res)
    got = orc.finish_reads(orc.reanchor_reads(batch, full, reads, threads=2)[0], 0)  # polishing never trims
    want = orc.finish_reads(orc.reanchor_reads(batch, want_res, reads, threads=2)[0], 0)
    assert got.equals(want) and 0 < w1 < batch.n_windows
    print("POLISH_OK", batch.n_windows, got.n_reads)
dist.barrier()
dist.destroy_process_group()
"""


def test_polishing_sharded_by_windows_world_size_2_gloo(entry, tmp_path):
    """BASELINE config 5 shape, small: a contig's windows are split over the ranks (one contig can dominate), the
    window results are gathered and rank 0 re-anchors them on the contigs."""
    port = 33500 + os.getpid() % 2000
    script = tmp_path / "polish_worker.py"
    script.write_text(POLISH_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "POLISH_OK" in outs[0]
