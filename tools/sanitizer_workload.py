"""A small pass over every kernel family for compute-sanitizer (memcheck / racecheck / initcheck): window path at 150, 20 and 8 sequences
per window (tiers C1 / G / W1 / W2 re-queues), k = 11 (hashed index), error windows, PAF ingest -> extraction -> run -> finish (re-anchoring +
post-filters), 2-bit input.  python tools/sanitizer_workload.py [scale]"""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from consent_b200.engine import Corrector  # noqa: E402
from consent_b200._ffi import Batch, Params  # noqa: E402
from consent_b200.synth import synth_windows, synth_piles, synth_paf  # noqa: E402
from tests.cases import concat  # noqa: E402

s = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n = lambda x: max(1, int(x * s))
cor = Corrector(device=0)
b = concat([synth_windows(n(24), 150, seed=1), synth_windows(n(40), 20, seed=2), synth_windows(n(40), 8, seed=3),
            Batch.from_piles([["ACGTTGCA" * 5] + ["ACGTACGTACGTAAC"] * 4199])])
r = cor.correct_windows(b)
print("windows", r.n_windows, "errors", int((r.status == 2).sum()), r.digest()[:12])
cor.set_option("input_2bit", 1)
r2 = cor.correct_windows(cor.pack_2bit(b))
print("2bit same", r2.digest() == r.digest())
cor.set_option("input_2bit", 0)
r = r2 = None
piles = synth_piles(n(60), genome_len=int(n(60) * 4000 / 30), read_len=4000, seed=4, max_support=4000)
text, names = synth_paf(piles, seed=4, tie_range=20)
ps = cor.ingest_paf(text, names, 150)
cor.upload_piles(ps.piles(piles.store_off, piles.store_bases))
cor.run()
fin = cor.finish_resident(1)
print("pipeline reads", fin.n_reads, fin.digest()[:12])
cor.close()
k11 = Corrector(Params(mer_size=11), device=0)
r = k11.correct_windows(concat([synth_windows(n(8), 150, seed=5), synth_windows(n(16), 20, seed=6)]))
print("k11", r.n_windows, r.digest()[:12])
