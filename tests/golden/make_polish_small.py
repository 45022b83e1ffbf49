"""Regenerates the polishing end-to-end fixture (BASELINE config 5 shape, small; build container only: needs /root/reference):

  polish_small_contigs.fasta.gz   4 synthetic "contigs" (6 kb, ONT error profile over a seeded 24 kb genome)
  polish_small_reads.fasta.gz     the 116 synthetic ONT reads (30x) that overlap them
  polish_small.paf.gz             the reads-on-contigs overlaps as CONSENT-polish hands them to its binary (query = contig, sorted by
                                  contig: CONSENT-polish:189-193), resMatches with ties, every other line with SAM-like tags
  polish_small_polished.fasta.gz  what the UNMODIFIED reference polisher (oracle/_ref/consent_polishing_ref = src/main.cpp +
                                  src/CONSENT-polishing.cpp and everything bin/CONSENT-polishing links, `make -C oracle dropin`)
                                  prints for them with -j 4 and the flags below

    python tests/golden/make_polish_small.py
"""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
FLAGS = ["-s", "3", "-S", "150", "-l", "500", "-k", "9", "-c", "8", "-A", "2", "-f", "4", "-m", "50", "-M", "150"]
N_CONTIGS = 4


def build():
    """-> (paf text, contigs fasta text, reads fasta text)"""
    from consent_b200.synth import synth_paf, synth_piles
    p = synth_piles(120, genome_len=24000, read_len=6000, n_piles=N_CONTIGS, seed=77, profile="ONT", max_support=4000)
    text, names = synth_paf(p, seed=77, tie_range=12)
    nm = [names.names[int(names.name_off[i]):int(names.name_off[i + 1])].tobytes().decode() for i in range(names.n_reads)]
    seq = [p.store_bases[int(p.store_off[i]):int(p.store_off[i + 1])].tobytes().decode() for i in range(p.n_store)]
    contigs = "".join(f">{nm[i]} contig\n{seq[i]}\n" for i in range(N_CONTIGS))
    reads = "".join(f">{nm[i]}\n{seq[i]}\n" for i in range(N_CONTIGS, p.n_store))
    return text, contigs, reads


def main():
    text, contigs, reads = build()
    t = tempfile.mkdtemp()
    for name, data in (("p.paf", text), ("c.fasta", contigs.encode()), ("r.fasta", reads.encode())):
        open(os.path.join(t, name), "wb").write(data)
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "consent_polishing_ref"), "-a", f"{t}/p.paf", "-r", f"{t}/c.fasta", "-R", f"{t}/r.fasta",
                          "-j", "4", "-p", "/nonexistent"] + FLAGS, check=True, capture_output=True).stdout
    assert out.count(b">") == N_CONTIGS, out[:200]
    for name, data in (("polish_small.paf.gz", text), ("polish_small_contigs.fasta.gz", contigs.encode()),
                       ("polish_small_reads.fasta.gz", reads.encode()), ("polish_small_polished.fasta.gz", out)):
        with gzip.GzipFile(os.path.join(HERE, name), "wb", compresslevel=9, mtime=0) as f:
            f.write(data)
    up = sum(c.isupper() for c in out.decode() if c in "ACGTacgt")
    print(f"{len(text)} bytes of PAF, {out.count(b'>')} polished contigs, {len(out)} bytes, {up} upper-case bases")


if __name__ == "__main__":
    main()
