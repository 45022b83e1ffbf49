#!/usr/bin/env python
"""bench.py — corrected windows/s of the CONSENT per-window hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--windows W] [--seqs N]

A step = one pass of the hot path over one batch of synthetic windows: by default the configuration the metric is
quoted on, BASELINE.json configs[2] = 100 000 windows of 500 bases x 150 sequences, PacBio 15 % profile, seed 42
(per GPU: weak scaling; --windows takes a slice of that stream).  Prints ONE JSON line (rank 0).

  value        windows/s with the batch resident in HBM (cg_run timed with CUDA events on the library's stream)
  e2e          the same through the reference-facing call cg_correct_windows with HOST (pinned) buffers:
               H2D of the piles, every kernel, D2H of consensus + solid k-mers inside the timed region
  roofline     dominant kernel: algorithmic bytes per launch / its CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline the reference's own CPU code (oracle/_ref) on this box's host cores, bounded sample of the same stream
  config2      (N=1) BASELINE.json configs[1]: 10 000 windows x 20 sequences (the shape real coverage has: few anchors, whole-window
               POA on the wide tier k_poa2<W1>) — resident windows/s, e2e, its own roofline row and the reference on the host cores
  reanchor     (N=1) the next row of SURVEY §8f on its own bounded workload: alignConsensus for every read
               (cg_reanchor_reads) — kernel windows/s and GCUPS from CUDA events, the two-stage call chain
               cg_correct_windows -> cg_reanchor_reads with host buffers, and the reference's alignConsensus on the host cores
  ingest       (N=1) SURVEY §8f rank 3-4 on their own bounded workload: PAF text parsed, grouped, sorted and cut on the device
               (k_paf_parse against the HBM roofline), next to the reference's getNextReadPile on one host thread; chain = PAF
               text + read store in, trimmed / filtered FASTA sequence lines out
  extract      (N=1) SURVEY §8f rank 2 on its own bounded workload: windows cut on the device from a read store + overlap
               tuples (cg_upload_piles); the copy kernel against the HBM roofline, and the whole chain overlap tuples ->
               corrected reads (extraction, window path, re-anchoring) with only the read store and the tuples uploaded
  --impl reference : times that CPU implementation as the arm itself.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line.  Libraries print there too (NCCL's "NCCL version ..." banner is a bare printf at
# NCCL_DEBUG=VERSION), so file descriptor 1 is pointed at stderr for the whole run and the line is written to the saved descriptor.
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: str) -> None:
    os.write(_REAL_STDOUT, (line + "\n").encode())


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = "config3: synthetic 500 bp windows x 150 seqs/pile (maxMSA cap), PB 15% error, seed 42"
METRIC = "corrected windows/sec"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(n_windows: int, n_seqs: int, first_window: int, pinned: bool, world: int = 1):
    """Slice [first_window, first_window + n_windows) of the seed-42 PB stream; arrays in pinned host memory if asked."""
    from consent_b200._ffi import Batch
    from consent_b200.synth import synth_windows
    b = synth_windows(n_windows, n_seqs, seed=42, profile="PB", first_window=first_window,
                      threads=max(1, min(host_cores() // max(world, 1), 64)))
    if pinned:
        import torch
        keep = []
        arrs = []
        for a in (b.win_seq_begin, b.seq_off, b.bases):
            t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).pin_memory()
            keep.append(t)
            arrs.append(t.numpy().view(a.dtype))
        nb = Batch.__new__(Batch)
        nb.win_seq_begin, nb.seq_off, nb.bases = arrs
        nb._keep = keep
        return nb
    return b


def cpu_reference(batch, threads: int):
    """-> (checker, kind): oracle/_ref (the reference itself) when it loads on this host, else the oracle port."""
    from tests.refs import Oracle, Reference, ref_library_path
    import __graft_entry__ as g
    g.build_oracle()
    if ref_library_path() is not None:
        return Reference(), "reference"
    return Oracle(), "port"


def time_cpu(checker, batch, threads: int, budget_s: float = 15.0):
    """Bounded sample: grow the sample until one timed call takes >= budget/3, report windows/s of the last call."""
    n = min(batch.n_windows, max(threads, 8))
    best = None
    while True:
        sample = batch.slice(0, n)
        _, sec = checker.correct_windows(sample, threads=threads, with_status=False)
        best = (n, sec)
        if sec >= budget_s / 3 or n >= batch.n_windows:
            break
        n = min(batch.n_windows, max(n * 2, int(n * (budget_s / 2) / max(sec, 1e-3))))
    n, sec = best
    return n / sec, n, sec


def algorithmic_bytes(counters: dict, n_occ: int) -> dict:
    """Per-stage algorithmic bytes of one pass."""
    b_in = (counters["bases"] + 3) // 4 + 8 * counters["sequences"]
    b_out = counters["consensus_bytes"] + 8 * counters["solid_kmers"]
    b_dp = 2 * (counters["dp_cells"] + counters["dp_pred_cells"])
    b_idx = 8 * n_occ
    return {"in": b_in, "out": b_out, "poa": b_dp, "index": b_idx, "window_total": b_in + b_out + b_dp}


def bench_reanchor(cor, n_reads: int, cores: int, steps: int) -> dict:
    """Consensus re-anchoring on its own bounded workload: seeded 8 kb PB reads, 20 sequences per
    window.  Kernel time from CUDA events (cg_reanchor_stats); the chain cg_correct_windows -> cg_reanchor_reads with
    host buffers; the reference's own alignConsensus (oracle/_ref) on a sample of the same reads, all host threads."""
    import torch
    from consent_b200._ffi import Reads, Results
    from consent_b200.synth import synth_reads
    batch, reads = synth_reads(n_reads, 20, truth_len=8000, seed=42)
    live = cor.correct_windows(batch)
    cor.reanchor_reads(batch, live, reads)                                  # warm-up (allocations)
    k_ms = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        live = None
        live = cor.correct_windows(batch)
        got = cor.reanchor_reads(batch, live, reads)
        k_ms.append(cor.reanchor_stats()["kernel_ms"])
    torch.cuda.synchronize()
    chain_s = (time.perf_counter() - t0) / steps
    st = cor.reanchor_stats()
    kernel_ms = float(np.mean(k_ms))
    out = {"workload": f"{n_reads} synthetic 8 kb PB reads (seed 42), {batch.n_windows} windows x 20 seqs, window 500 / overlap 50",
           "windows": batch.n_windows, "reads": n_reads,
           "kernel_ms": kernel_ms, "kernel_windows_per_s": batch.n_windows / (kernel_ms / 1e3),
           "dp_cells": st["dp_cells"], "gcups": st["dp_cells"] / (kernel_ms / 1e3) / 1e9,
           "chain_windows_per_s": batch.n_windows / chain_s,
           "chain": "cg_correct_windows + cg_reanchor_reads, host buffers in / corrected reads out"}
    try:
        checker, kind = cpu_reference(batch, cores)
        n = min(n_reads, max(64 * cores, 1024))
        w1 = int(reads.read_win_begin[n])
        sub_batch = batch.slice(0, w1)
        sub_reads = Reads(reads.read_win_begin[:n + 1], reads.read_off[:n + 1], reads.read_bases[:int(reads.read_off[n])],
                          reads.win_pos[:w1])
        host = Results(live._r)
        sub_res = Results.__new__(Results)
        sub_res.n_windows = w1
        sub_res._r = sub_res._free = None
        c1, s1 = int(host.cons_off[w1]), int(host.solid_off[w1])
        sub_res.cons_off, sub_res.cons, sub_res.status = host.cons_off[:w1 + 1], host.cons[:c1], host.status[:w1]
        sub_res.solid_off, sub_res.solid_kmer, sub_res.solid_count = host.solid_off[:w1 + 1], host.solid_kmer[:s1], host.solid_count[:s1]
        want, sec = checker.reanchor_reads(sub_batch, sub_res, sub_reads, threads=cores)
        out["cpu_reference"] = {"value": w1 / sec, "unit": "windows/s", "cores": cores, "kind": kind,
                                "sample": f"first {n} reads ({w1} windows), {sec:.2f} s, all host threads"}
        out["parity_spot_check"] = all(got.read(r) == want.read(r) for r in range(n))
    except Exception as e:
        out["cpu_reference"] = {"value": None, "kind": "unavailable", "sample": repr(e)}
    return out


def bench_extract(cor, n_reads: int, cores: int, steps: int, hbm_peak: float) -> dict:
    """Window extraction on the device at the config-3 depth: seeded 8 kb PB reads over a random genome
    at 150x, piles capped at 150 overlaps.  k_ex_copy reads and writes one byte per pile base: a plain HBM roofline row.
    chain = cg_upload_piles -> cg_run -> cg_download -> cg_download_windows -> cg_reanchor_reads, i.e. overlap tuples and the
    read store in host memory to corrected reads in host memory."""
    import torch
    from consent_b200.synth import synth_piles
    read_len = 8000
    piles = synth_piles(n_reads, genome_len=int(n_reads * read_len / 150), read_len=read_len, seed=42, max_support=150)
    cor.upload_piles(piles)                                   # warm-up (allocations)
    k_ms, c_ms = [], []
    for _ in range(max(steps, 1)):
        cor.upload_piles(piles)
        st = cor.extract_stats()
        k_ms.append(st["kernel_ms"]); c_ms.append(st["copy_ms"])
    batch, reads, _ = cor.download_windows(with_bases=False)
    W = batch.n_windows
    copy_ms, kernel_ms = float(np.mean(c_ms)), float(np.mean(k_ms))
    achieved = 2 * st["pile_bytes"] / (copy_ms / 1e3) / 1e9
    out = {"workload": f"{n_reads} synthetic 8 kb PB reads at 150x (seed 42), {len(piles.overlaps)} overlap tuples -> {W} windows, "
                       f"{batch.n_seqs / max(W, 1):.0f} seqs/window",
           "windows": W, "pile_bytes": st["pile_bytes"], "kernel_ms": kernel_ms,
           "kernel_windows_per_s": W / (kernel_ms / 1e3),
           "roofline": {"bound": "hbm", "kernel": "k_ex_copy", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "algorithmic_bytes_per_launch": 2 * st["pile_bytes"], "ms_per_launch": copy_ms}}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(max(steps, 1)):
        cor.upload_piles(piles)
        cor.run()
        res = cor.download()
        b, rd, _ = cor.download_windows(with_bases=False)
        got = cor.reanchor_reads(b, res, rd)
        res = None
    torch.cuda.synchronize()
    chain_s = (time.perf_counter() - t0) / max(steps, 1)
    out["chain_windows_per_s"] = W / chain_s
    out["chain"] = "cg_upload_piles + cg_run + cg_download + cg_download_windows + cg_reanchor_reads: overlap tuples + read store in, corrected reads out"
    out["chain_h2d_bytes"] = int(piles.store_bases.nbytes + piles.overlaps.nbytes + piles.store_off.nbytes)
    try:
        checker, kind = cpu_reference(None, cores)
        n = min(piles.n_piles, 64)
        from consent_b200._ffi import Piles
        sub = Piles(piles.store_off, piles.store_bases, piles.pile_read[:n], piles.pile_qlen[:n], piles.pile_ov_begin[:n + 1],
                    piles.overlaps[:int(piles.pile_ov_begin[n])], piles.min_support, piles.window_size, piles.window_overlap)
        t0 = time.perf_counter()
        wb, wr, _ = checker.extract_windows(sub)
        sec = time.perf_counter() - t0
        out["cpu_reference"] = {"value": wb.n_windows / sec, "unit": "windows/s", "cores": 1, "kind": kind,
                                "sample": f"first {n} piles ({wb.n_windows} windows), {sec:.2f} s, one thread (the harness is serial)"}
        w1 = wb.n_windows
        out["parity_spot_check"] = bool(np.array_equal(wb.seq_off, b.seq_off[:len(wb.seq_off)]) and np.array_equal(wr.win_pos, rd.win_pos[:w1]))
    except Exception as e:
        out["cpu_reference"] = {"value": None, "kind": "unavailable", "sample": repr(e)}
    return out


def bench_ingest(cor, n_reads: int, cores: int, steps: int, hbm_peak: float) -> dict:
    """PAF ingest on the device and the post-filters (rank 4): the PAF text of seeded 8 kb PB reads at 150x
    (every true overlap, resMatches with ties) -> piles cut to 150 overlaps.  k_paf_parse reads every byte of the text once and
    writes one 40-byte record per line: an HBM roofline row.  chain = cg_ingest_paf -> cg_upload_piles -> cg_run -> cg_download ->
    cg_download_windows -> cg_finish_reads, i.e. PAF text + read store in host memory to FASTA sequence lines in host memory."""
    import torch
    from consent_b200.synth import synth_paf, synth_piles
    read_len = 8000
    piles = synth_piles(n_reads, genome_len=int(n_reads * read_len / 150), read_len=read_len, seed=42, max_support=4000)
    text, names = synth_paf(piles, seed=42, tie_range=60)
    ps = cor.ingest_paf(text, names, 150)                      # warm-up (allocations)
    k_ms, p_ms = [], []
    for _ in range(max(steps, 1)):
        ps = cor.ingest_paf(text, names, 150)
        st = cor.ingest_stats()
        k_ms.append(st["kernel_ms"]); p_ms.append(st["parse_ms"])
    kernel_ms, parse_ms = float(np.mean(k_ms)), float(np.mean(p_ms))
    alg = len(text) + 40 * ps.n_lines + 8 * ps.n_lines
    achieved = alg / (parse_ms / 1e3) / 1e9
    out = {"workload": f"PAF text of {n_reads} synthetic 8 kb PB reads at 150x (seed 42): {len(text)} bytes, {ps.n_lines} lines -> "
                       f"{ps.n_piles} piles, {len(ps.overlaps)} overlaps kept (maxSupport 150)",
           "paf_bytes": len(text), "lines": ps.n_lines, "kernel_ms": kernel_ms, "kernel_lines_per_s": ps.n_lines / (kernel_ms / 1e3),
           "kernel_GBps_text": len(text) / (kernel_ms / 1e3) / 1e9,
           "roofline": {"bound": "hbm", "kernel": "k_paf_parse", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                        "algorithmic_bytes_per_launch": alg, "ms_per_launch": parse_ms,
                        "model": "text bytes + 8 B newline position read + 40 B record written per line"}}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(max(steps, 1)):
        ps = cor.ingest_paf(text, names, 150)
    torch.cuda.synchronize()
    out["e2e_lines_per_s"] = ps.n_lines / ((time.perf_counter() - t0) / max(steps, 1))
    t0 = time.perf_counter()
    for _ in range(max(steps, 1)):
        ps = cor.ingest_paf(text, names, 150)
        cor.upload_piles(ps.piles(piles.store_off, piles.store_bases))
        cor.run()
        got = cor.finish_resident(1)
    torch.cuda.synchronize()
    chain_s = (time.perf_counter() - t0) / max(steps, 1)
    n_win = int(cor.counters()["windows"])
    out["chain_windows_per_s"] = n_win / chain_s
    out["chain"] = ("cg_ingest_paf + cg_upload_piles + cg_run + cg_finish_resident: "
                    "PAF text + read store in, trimmed / filtered FASTA sequence lines out (nothing else crosses the bus)")
    out["chain_h2d_bytes"] = int(len(text) + piles.store_bases.nbytes + piles.store_off.nbytes + ps.overlaps.nbytes)
    out["finish_kernel_ms"] = cor.finish_stats()["kernel_ms"]
    out["fasta_records"] = int((got.read_off[1:] != got.read_off[:-1]).sum())
    try:
        checker, kind = cpu_reference(None, cores)
        cut = text[:text.index(b"\n", min(len(text) - 1, 8 << 20)) + 1]             # a bounded sample: the first ~8 MB of lines
        t0 = time.perf_counter()
        want = checker.ingest_paf(cut, names, 150)
        sec = time.perf_counter() - t0
        out["cpu_reference"] = {"value": want.n_lines / sec, "unit": "lines/s", "cores": 1, "kind": kind,
                                "sample": f"first {len(cut)} bytes ({want.n_lines} lines), {sec:.2f} s, one thread (getNextReadPile runs on the "
                                          "reference's main thread, src/CONSENT-correction.cpp:87,107)"}
        n = max(want.n_piles - 1, 0)                                                 # the sample's last pile may be cut short
        out["parity_spot_check"] = bool(np.array_equal(want.overlaps[:int(want.pile_ov_begin[n])], ps.overlaps[:int(want.pile_ov_begin[n])]))
    except Exception as e:
        out["cpu_reference"] = {"value": None, "kind": "unavailable", "sample": repr(e)}
    return out


def full_stream_parity(res, name: str, n_windows: int, n_seqs: int) -> dict:
    """The WHOLE output of a run against the unmodified reference's: stream digests committed under tests/golden/stream_digests.json
    (made once with oracle/_ref by tests/golden/make_stream_digests.py) — every window of the batch, not a sample."""
    try:
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "stream_digests.json")))[name]
    except Exception as e:
        return {"checked": False, "why": f"no golden digest: {e!r}"}
    if gold["windows"] != n_windows or gold["seqs_per_window"] != n_seqs:
        return {"checked": False, "why": f"golden digest is for {gold['windows']} x {gold['seqs_per_window']}, this run is {n_windows} x {n_seqs}"}
    cons, solid = res.stream_digest_pair(res.stream_digests())
    return {"checked": True, "windows": n_windows, "consensus_equal_to_reference": cons == gold["consensus_sha256"],
            "solid_lists_equal_to_reference": solid == gold["solid_sha256"], "golden": "tests/golden/stream_digests.json"}


def kernel_table(acc: dict, steps: int, ab_by_kernel: dict) -> dict:
    """Per kernel: ms per step (sum of its launches' own CUDA-event durations), launches per step, algorithmic GB/s."""
    out = {}
    for name, v in acc.items():
        if not v["launches"]:
            continue
        ms = v["ms"] / steps
        row = {"ms_per_step": round(ms, 3), "launches_per_step": v["launches"] / steps, "ms_per_launch": round(v["ms"] / v["launches"], 4)}
        if name in ab_by_kernel and ms > 0:
            row["algorithmic_bytes_per_step"] = int(ab_by_kernel[name])
            row["algorithmic_GBps"] = round(ab_by_kernel[name] / (ms / 1e3) / 1e9, 1)
        out[name] = row
    return out


def acc_kernels(acc: dict, ks: dict) -> None:
    for name, v in ks.items():
        a = acc.setdefault(name, {"ms": 0.0, "launches": 0, "dp_cells": 0, "dp_pred_cells": 0})
        a["ms"] += v["ms"]; a["launches"] += v["launches"]
        a["dp_cells"] = v.get("dp_cells", 0); a["dp_pred_cells"] = v.get("dp_pred_cells", 0)      # per run, identical every step


def roofline_row(ktab: dict, peak: float, peak_src: str, traffic_json: dict, ms_per_step: float, windows_per_step: int) -> dict:
    """The dominant kernel = the one that does most of the path's algorithmic work per step (the wide POA tier's few long jobs run
    for a long time in the background of a 150-sequence step without doing much of its work, so the longest-running kernel would
    be the wrong pick there)."""
    cand = {k: v for k, v in ktab.items() if "algorithmic_GBps" in v}
    dom = max(cand, key=lambda k: cand[k]["algorithmic_bytes_per_step"])
    v = cand[dom]
    per_launch = v["algorithmic_bytes_per_step"] / v["launches_per_step"]
    achieved = per_launch / (v["ms_per_launch"] / 1e3) / 1e9
    tr = traffic_json.get(dom, {})
    traffic = None
    if tr.get("dram_bytes_per_window") is not None:       # per launch like `achieved`: bytes per window x windows of a step / launches per step
        traffic = int(tr["dram_bytes_per_window"] * windows_per_step / v["launches_per_step"])
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic,
            "traffic_source": (tr["source"] + f"; {tr['dram_bytes_per_window']:.0f} DRAM bytes per window ({tr.get('shape', '')}) x the step's windows / "
                               "launches per step — a constant from an ncu capture, not measured in this run") if traffic is not None
            else "none: no ncu capture of this kernel committed",
            "peak_source": peak_src, "algorithmic_bytes_per_launch": per_launch, "launches_per_step": v["launches_per_step"],
            "ms_per_launch": v["ms_per_launch"],
            "timing": "each launch bracketed by its own CUDA events on the stream it runs on (cg_get_kernel_stats); kernels of "
                      "different tiers / chunks overlap, so the per-kernel sums may exceed ms_per_step but no kernel's own sum does",
            "kernel_ms_sum_le_step": bool(v["ms_per_step"] <= ms_per_step * 1.001),
            "kernels": ktab}


def bench_config2(cor, cores: int, steps: int, warmup: int, peak: float, peak_src: str, traffic_json: dict, cpu_budget: float) -> dict:
    """BASELINE.json configs[1]: 10 000 synthetic 500 bp windows x 20 sequences, PB 15 %, seed 42, one B200."""
    import torch
    W, N = 10000, 20
    batch = make_batch(W, N, 0, pinned=True)
    n_occ = int(np.maximum(np.diff(batch.seq_off.astype(np.int64)) - 8, 0).sum())
    cor.upload(batch)
    for _ in range(max(warmup, 1)):
        cor.run()
    dev_ms, acc = 0.0, {}
    for _ in range(steps):
        cor.run()
        dev_ms += cor.run_ms()
        acc_kernels(acc, cor.kernel_stats())
    counters = cor.counters()
    ab = {name: 2 * (acc[name]["dp_cells"] + acc[name]["dp_pred_cells"]) for name in ("k_poa2<C1>", "k_poa2<G>", "k_poa2<W1>", "k_poa2<W2>")}
    ab["k_index"] = 8 * n_occ
    ktab = kernel_table(acc, steps, ab)
    res = cor.correct_windows(batch)                          # warm-up of the pipelined path
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = None
        res = cor.correct_windows(batch)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / steps
    parity = full_stream_parity(res, "config2", W, N)
    out = {"workload": "config2: 10000 synthetic 500 bp windows x 20 seqs/pile, PB 15% error, seed 42 (steps re-run the same 100 MB batch: it fits L2 "
                       "once, the 3.6 MB score matrices per window do not)",
           "windows": W, "value": W * steps / (dev_ms / 1e3), "unit": "windows/s", "ms_per_step": dev_ms / steps,
           "e2e": {"value": W / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": int(batch.n_bases + batch.seq_off.nbytes + batch.win_seq_begin.nbytes),
                   "d2h_bytes_per_step": int(res.cons.nbytes + res.status.nbytes + res.cons_off.nbytes + res.solid_off.nbytes + res.solid_kmer.nbytes + res.solid_count.nbytes)},
           "roofline": roofline_row(ktab, peak, peak_src, traffic_json, dev_ms / steps, W),
           "parity_full_stream": parity,
           "counters_per_step": counters}
    try:
        checker, kind = cpu_reference(batch, cores)
        wps, n, sec = time_cpu(checker, batch, cores, cpu_budget)
        want, _ = checker.correct_windows(batch.slice(0, min(n, 256)), threads=cores)
        out["cpu_reference"] = {"value": wps, "unit": "windows/s", "cores": cores, "kind": kind,
                                "sample": f"first {n} windows of the same batch, {sec:.1f} s, all host threads",
                                "parity_spot_check": all(res.consensus(w) == want.consensus(w) for w in range(want.n_windows))}
        out["speedup_vs_cpu_reference"] = {"resident": out["value"] / wps, "e2e": out["e2e"]["value"] / wps}
    except Exception as e:
        out["cpu_reference"] = {"value": None, "kind": "unavailable", "sample": repr(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=("b200", "reference"))
    ap.add_argument("--windows", type=int, default=int(os.environ.get("CG_BENCH_WINDOWS", "100000")), help="windows per step and per GPU")
    ap.add_argument("--seqs", type=int, default=150)
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--chunk-windows", type=int, default=0, help="windows per chunk (0: library default); never changes results")
    ap.add_argument("--lanes", type=int, default=0, help="chunks in flight (0: library default = 2); never changes results")
    ap.add_argument("--ingest-reads", type=int, default=int(os.environ.get("CG_BENCH_INGEST_READS", "1000")),
                    help="reads of the PAF-ingest / post-filter measurement (0: skip it)")
    ap.add_argument("--extract-reads", type=int, default=int(os.environ.get("CG_BENCH_EXTRACT_READS", "2100")),
                    help="reads of the window-extraction measurement (0: skip it)")
    ap.add_argument("--no-config2", action="store_true", help="skip the config-2 (10 000 x 20) row")
    ap.add_argument("--reanchor-reads", type=int, default=int(os.environ.get("CG_BENCH_REANCHOR_READS", "2600")),
                    help="reads of the re-anchoring measurement (0: skip it)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = host_cores()
    config = {"workload": WORKLOAD, "windows_per_step_per_gpu": args.windows, "seqs_per_window": args.seqs, "window_len": 500,
              "k": 9, "solid": 4, "commonKMers": 8, "minAnchors": 2,
              "l2": "inputs larger than L2 (>= 1.5 GB of piles per step, 126 MB L2)", "parallelism": f"windows sharded x{world}"}

    import __graft_entry__ as g
    g.build_host()

    if args.impl == "reference":
        if rank != 0:
            return
        batch = make_batch(min(args.windows, 8192), args.seqs, 0, pinned=False)
        checker, kind = cpu_reference(batch, cores)
        for _ in range(max(args.warmup, 1)):
            checker.correct_windows(batch.slice(0, min(batch.n_windows, cores)), threads=cores, with_status=False)
        per_step_budget = max(2.0, min(20.0, 90.0 / max(args.steps, 1)))
        tot_w, tot_s, n = 0, 0.0, 0
        for _ in range(args.steps):
            wps, n, sec = time_cpu(checker, batch, cores, per_step_budget)
            tot_w += n; tot_s += sec
        value = tot_w / tot_s
        config = dict(config, reference_sample_windows_per_step=n,
                      reference_sample_note="same stream and per-window shape as the b200 arm; each step times a bounded sample "
                                            f"(the first {n} windows), not windows_per_step_per_gpu windows: the metric is a rate")
        out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": args.gpus, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "int16/u8", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": value, "unit": "windows/s", "cores": cores, "kind": kind,
                                "sample": f"{n} windows x {args.seqs} seqs per step (first windows of the same stream)"},
               "e2e": {"value": value, "unit": "windows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g.build_cuda()
    from consent_b200.engine import Corrector
    from consent_b200.shard import gather_results

    t0 = time.time()
    batch = make_batch(args.windows, args.seqs, rank * args.windows, pinned=True, world=world)   # weak scaling: every rank its own slice
    n_occ = int(np.maximum(np.diff(batch.seq_off.astype(np.int64)) - 8, 0).sum())
    log(f"[rank {rank}] generated {batch.n_windows} windows, {batch.n_bases / 1e9:.2f} GB of bases in {time.time() - t0:.1f}s")

    cor = Corrector(device=local_rank)
    if args.chunk_windows:
        cor.set_option("chunk_max_windows", args.chunk_windows)
    if args.lanes:
        cor.set_option("lanes", args.lanes)
    if os.environ.get("CG_CHUNK_BUDGET_GB"):
        cor.set_option("chunk_budget_bytes", int(float(os.environ["CG_CHUNK_BUDGET_GB"]) * (1 << 30)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- HBM-resident arm
    cor.upload(batch)
    for _ in range(args.warmup):
        cor.run()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, stage_acc, launches, kacc = 0.0, {}, 0, {}
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        cor.run()
        dev_ms += cor.run_ms()
        acc_kernels(kacc, cor.kernel_stats())
        for k, v in cor.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v["ms"]
            launches += v["launches"]
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    clocks = sampler.stop()
    counters = cor.counters()
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(t[0]), float(t[1])
    value = world * args.windows * args.steps / (dev_ms / 1e3)

    # ---- end-to-end arms: host buffers in, host buffers out (+ the ordered gather to rank 0 when sharded).
    # "e2e"            the batch as a host that keeps reads the way the reference does holds it: 2 bits per base (the reference's read
    #                  index, src/utils.cpp:21-54; packed once, outside the timed region, like generating the batch), consensus + status +
    #                  offsets back; the solid k-mer lists stay in HBM, where their only consumer (re-anchoring) reads them
    # "e2e_ascii_full" ASCII piles in, consensus AND solid k-mer lists out: round 1's e2e, 5x the bytes over PCIe
    def e2e_arm(in_batch):
        res = None
        for _ in range(max(1, args.warmup - 1)):
            res = None
            res = cor.correct_windows(in_batch)
            if world > 1:
                gather_results(res, with_solid=False, concat=False)
        barrier()
        e0 = time.perf_counter()
        for _ in range(args.steps):
            res = None
            res = cor.correct_windows(in_batch)
            if world > 1:
                gather_results(res, with_solid=False, concat=False)   # the corrected windows, in input order, to rank 0 (NCCL)
        barrier()
        sec = time.perf_counter() - e0
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return res, float(t[0])

    res, e2e_ascii_s = e2e_arm(batch)
    h2d_ascii = int(batch.n_bases + batch.seq_off.nbytes + batch.win_seq_begin.nbytes)
    d2h_ascii = int(res.cons.nbytes + res.status.nbytes + res.cons_off.nbytes + res.solid_off.nbytes + res.solid_kmer.nbytes + res.solid_count.nbytes)
    res_ascii = res
    parity_full = full_stream_parity(res_ascii, "config3", args.windows, args.seqs) if rank == 0 else None
    packed = cor.pack_2bit(batch, threads=max(1, min(cores // max(world, 1), 32)), pinned=True)
    cor.set_option("input_2bit", 1)
    cor.set_option("results_with_solid", 0)
    res, e2e_s = e2e_arm(packed)
    h2d = int((batch.n_bases + 3) // 4 + batch.seq_off.nbytes + batch.win_seq_begin.nbytes)
    d2h = int(res.cons.nbytes + res.status.nbytes + res.cons_off.nbytes + res.solid_off.nbytes)
    same_cons = bool(np.array_equal(res.cons, res_ascii.cons) and np.array_equal(res.cons_off, res_ascii.cons_off))
    cor.set_option("input_2bit", 0)
    cor.set_option("results_with_solid", 1)
    res = res_ascii

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    ab = algorithmic_bytes(counters, n_occ)
    stage_avg = {k: v / args.steps for k, v in stage_acc.items()}
    try:
        traffic_json = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic_json = {}
    abk = {name: 2 * (kacc[name]["dp_cells"] + kacc[name]["dp_pred_cells"]) for name in ("k_poa2<C1>", "k_poa2<G>", "k_poa2<W1>", "k_poa2<W2>") if name in kacc}
    abk["k_index"] = ab["index"]
    ktab = kernel_table(kacc, args.steps, abk)
    roofline = roofline_row(ktab, peak, peak_src, traffic_json, dev_ms / args.steps, args.windows)
    roofline["whole_path_GBps"] = ab["window_total"] / (dev_ms / args.steps / 1e3) / 1e9
    roofline["poa_all_tiers"] = {"algorithmic_bytes_per_step": ab["poa"], "note": "tiers run side by side; see kernels[*] for each"}
    roofline["stage_ms_per_step"] = {k: round(v, 3) for k, v in stage_avg.items()}

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only)
    cpu = None
    if world == 1:
        try:
            checker, kind = cpu_reference(batch, cores)
            wps, n, sec = time_cpu(checker, batch, cores, args.cpu_budget)
            cpu = {"value": wps, "unit": "windows/s", "cores": cores, "kind": kind,
                   "sample": f"first {n} windows of the same batch, {sec:.1f} s, all host threads"}
            # the sample doubles as a parity spot-check of this very run
            want, _ = checker.correct_windows(batch.slice(0, min(n, 64)), threads=cores)
            got = [res.consensus(w) for w in range(want.n_windows)]
            cpu["parity_spot_check"] = all(got[w] == want.consensus(w) for w in range(want.n_windows))
        except Exception as e:  # the checker is optional for the bench line
            cpu = {"value": None, "unit": "windows/s", "cores": cores, "kind": "unavailable", "sample": repr(e)}

    config2 = None
    if world == 1 and not args.no_config2 and args.seqs == 150:
        try:
            res = None
            config2 = bench_config2(cor, cores, args.steps, args.warmup, peak, peak_src, traffic_json, min(args.cpu_budget, 10.0))
        except Exception as e:
            config2 = {"error": repr(e)}

    reanchor = None
    if world == 1 and args.reanchor_reads > 0:
        try:
            reanchor = bench_reanchor(cor, args.reanchor_reads, cores, args.steps)
        except Exception as e:
            reanchor = {"error": repr(e)}

    extract = None
    if world == 1 and args.extract_reads > 0:
        try:
            extract = bench_extract(cor, args.extract_reads, cores, args.steps, peak)
        except Exception as e:
            extract = {"error": repr(e)}

    ingest = None
    if world == 1 and args.ingest_reads > 0:
        try:
            ingest = bench_ingest(cor, args.ingest_reads, cores, args.steps, peak)
        except Exception as e:
            ingest = {"error": repr(e)}

    out = {"metric": METRIC, "value": value, "unit": "windows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "int16/u8", "data": "synthetic", "config": config, "clocks": clocks,
           "e2e": {"value": world * args.windows * args.steps / e2e_s, "unit": "windows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "input": "host batch with 2-bit packed bases (the reference's read-index form, src/utils.cpp:21-54; cg_set_option input_2bit), pinned",
                   "output": "consensus bytes + status + offsets to the host; solid k-mer lists stay resident for cg_reanchor_reads (results_with_solid 0)",
                   "consensus_equal_to_ascii_arm": same_cons},
           "e2e_ascii_full": {"value": world * args.windows * args.steps / e2e_ascii_s, "unit": "windows/s", "h2d_bytes_per_step": h2d_ascii,
                              "d2h_bytes_per_step": d2h_ascii, "input": "ASCII piles (1 byte per base), pinned",
                              "output": "consensus + solid k-mer lists (round 1's e2e)"},
           "gpu_launches": launches, "wall_ms_per_step": wall_ms / args.steps,
           "roofline": roofline, "cpu_baseline": cpu, "parity_full_stream": parity_full,
           "counters_per_step": counters, "config2": config2, "reanchor": reanchor, "extract": extract, "ingest": ingest}
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
