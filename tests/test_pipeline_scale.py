"""Pipeline-scale parity (BASELINE configs 1, 4, 5): the product's wrappers and host programs against the UNMODIFIED reference
binaries, by md5 of the whole FASTA.

  config 1  the shipped example (example/reads.fasta, 5 653 reads): vendored minimap2 2.17 with the flags of CONSENT-correct:185, then
            bin/CONSENT-correction with the wrapper's flags (:202).  Golden md5s: 912f11b9... for the whole file (126 877 PAF lines,
            5 604 corrected reads) and 6971c8c1... for small.paf = its first 2 000 lines (94 piles) — the two values SURVEY §6 / §8c
            record for the reference, reproduced here with oracle/_ref/consent_correction_ref before they were committed.
  config 4  30x PacBio over a 5 Mb genome (tools/make_pipeline_data.py config4), through consent_b200/CONSENT-correct
  config 5  50 contigs x 100 kb polished with 30x ONT reads (config5), through consent_b200/CONSENT-polish with ITS defaults
            (minSupport 1, maxSupport 20000: reference CONSENT-polish:42-43)

tests/golden/pipeline_md5.json holds the md5s of what the unmodified reference printed for these inputs (generator:
tests/golden/make_pipeline_golden.sh).  minimap2's output does not depend on its thread count (checked: -t3 and -t8 give the same file),
so the PAF is regenerated wherever the test runs; its md5 is compared first.  Needs oracle/_ref/minimap2 and oracle/_ref/example/
(built / copied by `make -C oracle pipeline` where /root/reference exists; they travel with the snapshot)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
PKG = os.path.join(ROOT, "consent_b200")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "pipeline_md5.json")))
CORRECT_FLAGS = ["-s", "3", "-S", "150", "-l", "500", "-k", "9", "-c", "8", "-A", "2", "-f", "4", "-m", "50", "-M", "150"]   # CONSENT-correct:42-50


def md5(path_or_bytes):
    h = hashlib.md5()
    if isinstance(path_or_bytes, bytes):
        h.update(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            for blk in iter(lambda: f.read(1 << 22), b""):
                h.update(blk)
    return h.hexdigest()


@pytest.fixture(scope="module")
def mm2():
    p = os.path.join(REFDIR, "minimap2")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/minimap2 not built (make -C oracle pipeline needs /root/reference)")
    return p


@pytest.fixture(scope="module")
def example_paf(mm2, tmp_path_factory):
    reads = os.path.join(REFDIR, "example", "reads.fasta")
    if not os.path.exists(reads):
        pytest.skip("oracle/_ref/example/reads.fasta not present (make -C oracle pipeline)")
    d = tmp_path_factory.mktemp("example_full")
    paf = str(d / "example.paf")
    with open(paf, "wb") as out:
        subprocess.run([mm2, "--dual=yes", "-PD", "--no-long-join", "-w5", "-g1000", "-m30", "-n1", f"-t{min(os.cpu_count() or 4, 16)}", "-I1G", reads, reads],
                       check=True, stdout=out, stderr=subprocess.DEVNULL)
    assert md5(paf) == GOLD["example"]["paf_md5"], "minimap2 wrote a different PAF here than where the golden md5s were made"
    return d, paf, reads


@pytest.mark.gpu
def test_example_small_paf_md5_is_the_reference_s(example_paf, gpu_lib, entry):
    d, paf, reads = example_paf
    small = str(d / "small.paf")
    with open(paf, "rb") as f, open(small, "wb") as g:
        for _ in range(2000):
            g.write(f.readline())
    exe = entry.build_bin()
    out = subprocess.run([exe, "-a", small, "-r", reads, "-j", "4"] + CORRECT_FLAGS, check=True, capture_output=True).stdout
    assert out.count(b">") == GOLD["example"]["small_records"]
    assert md5(out) == GOLD["example"]["small_fasta_md5"] == "6971c8c11ec660409beb6642fb78fc5c"


@pytest.mark.gpu
def test_example_full_md5_is_the_reference_s(example_paf, gpu_lib, entry):
    d, paf, reads = example_paf
    exe = entry.build_bin()
    out = subprocess.run([exe, "-a", paf, "-r", reads, "-j", "4", "-B", "8"] + CORRECT_FLAGS, check=True, capture_output=True).stdout
    assert out.count(b">") == GOLD["example"]["records"]
    assert md5(out) == GOLD["example"]["fasta_md5"] == "912f11b9df1c5cf926c2fb8d4959b0bd"


def _generate(cfg, d):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_pipeline_data.py"), cfg, str(d)], check=True)


@pytest.mark.gpu
def test_config4_through_the_correct_wrapper_md5_is_the_reference_s(mm2, gpu_lib, entry, tmp_path):
    entry.build_bin()
    _generate("config4", tmp_path)
    reads, out = str(tmp_path / "reads.fasta"), str(tmp_path / "corrected.fasta")
    assert md5(reads) == GOLD["config4"]["reads_md5"]
    subprocess.run([os.path.join(PKG, "CONSENT-correct"), "--in", reads, "--out", out, "--type", "PB", "--tmpdir", str(tmp_path / "tmp")],
                   check=True, env=dict(os.environ, MINIMAP2=mm2), stdout=subprocess.DEVNULL)
    assert open(out, "rb").read().count(b">") == GOLD["config4"]["records"]
    assert md5(out) == GOLD["config4"]["fasta_md5"]


@pytest.mark.gpu
def test_config5_through_the_polish_wrapper_md5_is_the_reference_s(mm2, gpu_lib, entry, tmp_path):
    entry.build_bin()
    _generate("config5", tmp_path)
    contigs, reads, out = str(tmp_path / "contigs.fasta"), str(tmp_path / "reads.fasta"), str(tmp_path / "polished.fasta")
    assert md5(reads) == GOLD["config5"]["reads_md5"] and md5(contigs) == GOLD["config5"]["contigs_md5"]
    subprocess.run([os.path.join(PKG, "CONSENT-polish"), "--contigs", contigs, "--reads", reads, "--out", out, "--tmpdir", str(tmp_path / "tmp")],
                   check=True, env=dict(os.environ, MINIMAP2=mm2), stdout=subprocess.DEVNULL)
    assert open(out, "rb").read().count(b">") == GOLD["config5"]["records"]
    assert md5(out) == GOLD["config5"]["fasta_md5"]


# ------------------------------------------------------------------------------------------------ CPU: the PAF helpers
def _ref_tool(name):
    p = os.path.join(REFDIR, name)
    if not os.path.exists(p):
        pytest.skip(f"oracle/_ref/{name} not built (make -C oracle pipeline needs /root/reference)")
    return p


def _tiny_paf(n_reads=40, seed=3, comeback=False):
    """A PAF-like text: consecutive lines per query; with comeback=True queries return after other queries (a split minimap2 index)."""
    import random
    rng = random.Random(seed)
    passes = 3 if comeback else 1
    lines = []
    for ps in range(passes):
        for q in range(n_reads):
            for _ in range(rng.randint(1, 5)):
                t = rng.randrange(n_reads)
                lines.append("\t".join([f"r{q}", "5000", str(rng.randrange(100)), str(4000 + rng.randrange(900)), "+-"[rng.randrange(2)], f"r{t}",
                                        "6000", str(rng.randrange(100)), str(4500 + rng.randrange(900)), str(rng.randrange(3000)), "4800", "60",
                                        f"tp:A:{'PS'[ps % 2]}", "cm:i:33"]))
    return "\n".join(lines) + "\n"


def test_reformat_paf_equals_the_reference_tool(entry, tmp_path):
    entry.build_bin()
    src = tmp_path / "in.paf"
    src.write_text(_tiny_paf())
    subprocess.run([os.path.join(PKG, "bin", "reformatPAF"), str(src), str(tmp_path / "ours.paf")], check=True)
    subprocess.run([_ref_tool("reformatPAF_ref"), str(src), str(tmp_path / "ref.paf")], check=True)
    assert (tmp_path / "ours.paf").read_bytes() == (tmp_path / "ref.paf").read_bytes()


def test_explode_and_merge_equal_the_reference_tools(entry, tmp_path):
    entry.build_bin()
    src = tmp_path / "in.paf"
    src.write_text(_tiny_paf(comeback=True))
    hdr = tmp_path / "headers"
    hdr.write_text("".join(f">r{q}\n" for q in range(40)))
    for who, exp, mrg in (("ours", os.path.join(PKG, "bin", "explode"), os.path.join(PKG, "bin", "merge")),
                          ("ref", _ref_tool("explode_ref"), _ref_tool("merge_ref"))):
        d = tmp_path / who
        d.mkdir()
        subprocess.run([exp, str(src), str(d / "x")], check=True)
        chunks = sorted(str(p) for p in d.iterdir())
        assert len(chunks) == 3
        subprocess.run([mrg, str(d / "merged.paf"), str(hdr)] + chunks, check=True)
    assert (tmp_path / "ours" / "merged.paf").read_bytes() == (tmp_path / "ref" / "merged.paf").read_bytes()
    for i in (1, 2, 3):
        assert (tmp_path / "ours" / f"x_{i}").read_bytes() == (tmp_path / "ref" / f"x_{i}").read_bytes()
