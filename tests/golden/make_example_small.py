"""Regenerates the real-data end-to-end fixture (build container only: needs /root/reference):

  example_small.paf.gz              the overlaps of the first 20 read piles of the shipped example, computed by the vendored
                                    minimap2 with the PacBio flags of CONSENT-correct:185 (first 12 PAF columns; reads renamed
                                    R0, R1, ... — names only label records: Overlap ordering is by residue matches, src/Overlap.h:91-97)
  example_small_reads.fasta.gz      the 507 reads those overlaps mention
  example_small_corrected.fasta.gz  what the UNMODIFIED reference binary (oracle/_ref/consent_correction_ref = src/main.cpp and
                                    everything bin/CONSENT-correction links, `make -C oracle dropin`) prints for them with the
                                    defaults of the CONSENT-correct wrapper (:42-50)

    python tests/golden/make_example_small.py
"""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
N_PILES = 20
FLAGS = ["-s", "3", "-S", "150", "-l", "500", "-k", "9", "-c", "8", "-A", "2", "-f", "4", "-m", "50", "-M", "150"]


def main():
    t = tempfile.mkdtemp()
    paf = os.path.join(t, "example.paf")
    subprocess.run(f"cp -r {REF}/minimap2 {t}/mm2 && make -C {t}/mm2 -j8 > {t}/mm2.log 2>&1 && "
                   f"{t}/mm2/minimap2 --dual=yes -PD --no-long-join -w5 -g1000 -m30 -n1 -t8 {REF}/example/reads.fasta "
                   f"{REF}/example/reads.fasta > {paf} 2> {t}/mm2.err", shell=True, check=True)
    lines, names, seen, cur, piles = [], [], set(), None, 0
    for ln in open(paf):
        f = ln.rstrip("\n").split("\t")
        if f[0] != cur:
            if piles == N_PILES:
                break
            cur, piles = f[0], piles + 1
        lines.append(f)
        for nm in (f[0], f[5]):
            if nm not in seen:
                seen.add(nm)
                names.append(nm)
    ids = {nm: f"R{i}" for i, nm in enumerate(names)}
    seqs, name = {}, None
    for ln in open(os.path.join(REF, "example", "reads.fasta")):
        ln = ln.rstrip("\n")
        if ln.startswith(">"):
            name = ln[1:].split()[0]
            name = name if name in ids else None
            if name:
                seqs[name] = []
        elif name:
            seqs[name].append(ln)
    with gzip.open(os.path.join(HERE, "example_small_reads.fasta.gz"), "wt", compresslevel=9) as f:
        for nm in names:
            f.write(">%s\n%s\n" % (ids[nm], "".join(seqs[nm])))
    with gzip.open(os.path.join(HERE, "example_small.paf.gz"), "wt", compresslevel=9) as f:
        for x in lines:
            x = list(x[:12])
            x[0], x[5] = ids[x[0]], ids[x[5]]
            f.write("\t".join(x) + "\n")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "dropin"], check=True)
    small_paf, small_fa = os.path.join(t, "small.paf"), os.path.join(t, "small.fasta")
    open(small_paf, "wb").write(gzip.open(os.path.join(HERE, "example_small.paf.gz")).read())
    open(small_fa, "wb").write(gzip.open(os.path.join(HERE, "example_small_reads.fasta.gz")).read())
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "consent_correction_ref"), "-a", small_paf, "-r", small_fa,
                          "-j", "8", "-p", REF] + FLAGS, check=True, capture_output=True).stdout
    with gzip.open(os.path.join(HERE, "example_small_corrected.fasta.gz"), "wb", compresslevel=9) as f:
        f.write(out)
    print(len(lines), "PAF lines,", len(names), "reads,", out.count(b">"), "corrected reads")


if __name__ == "__main__":
    sys.exit(main())
