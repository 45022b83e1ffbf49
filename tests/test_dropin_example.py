"""BASELINE config 1 shape, end to end: PAF piles of the shipped example -> corrected FASTA.

oracle/_ref/consent_correction_b200 (oracle/dropin_correction.cpp) is the reference's own host code — PAF pile reading,
window extraction, trimming, FASTA output, compiled unmodified from /root/reference — with the hot path replaced by the two
calls INTEGRATION.md describes (cg_correct_windows, cg_reanchor_reads).  Its output must be byte-identical to what the
unmodified reference binary printed for the same input (tests/golden/example_small_corrected.fasta.gz, written by
tests/golden/make_example_small.py)."""
import gzip
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REFDIR = os.path.join(ROOT, "oracle", "_ref")
FLAGS = ["-s", "3", "-S", "150", "-l", "500", "-k", "9", "-c", "8", "-A", "2", "-f", "4", "-m", "50", "-M", "150"]   # CONSENT-correct:42-50


@pytest.fixture(scope="module")
def small(tmp_path_factory):
    d = tmp_path_factory.mktemp("example_small")
    paf, fa = str(d / "small.paf"), str(d / "small.fasta")
    open(paf, "wb").write(gzip.open(os.path.join(GOLD, "example_small.paf.gz")).read())
    open(fa, "wb").write(gzip.open(os.path.join(GOLD, "example_small_reads.fasta.gz")).read())
    want = gzip.open(os.path.join(GOLD, "example_small_corrected.fasta.gz")).read()
    return d, paf, fa, want


def _binary(name):
    p = os.path.join(REFDIR, name)
    if not os.path.exists(p):
        pytest.skip(f"oracle/_ref/{name} not built (make -C oracle dropin needs /root/reference)")
    return p


def test_golden_fasta_is_what_the_unmodified_reference_prints(small):
    """Pins the golden file against the reference binary itself (where it was built)."""
    d, paf, fa, want = small
    out = subprocess.run([_binary("consent_correction_ref"), "-a", paf, "-r", fa, "-j", "4", "-p", "/nonexistent"] + FLAGS,
                         check=True, capture_output=True).stdout
    assert out.count(b">") == 20
    assert out == want


def test_dropin_binary_on_emulated_kernels_prints_the_reference_fasta(small, entry):
    """The reference's host code + this repo's kernel sources (SIMT emulator) = the reference's FASTA, byte for byte."""
    d, paf, fa, want = small
    exe = _binary("consent_correction_b200")
    emu = entry.build_emu()
    libdir = d / "emulib"
    libdir.mkdir(exist_ok=True)
    link = libdir / "libconsent_b200.so"
    if not link.exists():
        os.symlink(emu, link)
    env = dict(os.environ, LD_LIBRARY_PATH=str(libdir))
    out = subprocess.run([exe, "-a", paf, "-r", fa] + FLAGS, check=True, capture_output=True, env=env).stdout
    assert out == want


def test_dropin_binary_with_device_side_extraction_on_emulated_kernels(small, entry):
    """-x: the window cutting of phase A runs on the device too (cg_upload_piles); still the reference's FASTA."""
    d, paf, fa, want = small
    exe = _binary("consent_correction_b200")
    emu = entry.build_emu()
    libdir = d / "emulib"
    libdir.mkdir(exist_ok=True)
    link = libdir / "libconsent_b200.so"
    if not link.exists():
        os.symlink(emu, link)
    env = dict(os.environ, LD_LIBRARY_PATH=str(libdir))
    out = subprocess.run([exe, "-x", "-a", paf, "-r", fa] + FLAGS, check=True, capture_output=True, env=env).stdout
    assert out == want


def test_dropin_binary_with_device_side_ingest_and_post_filters_on_emulated_kernels(small, entry):
    """-X: PAF parsing / sorting / top-maxSupport (cg_ingest_paf) and trimRead / dropRead (cg_finish_reads) run on the device as
    well — the host reads two files and prints records; still the reference's FASTA."""
    d, paf, fa, want = small
    exe = _binary("consent_correction_b200")
    emu = entry.build_emu()
    libdir = d / "emulib"
    libdir.mkdir(exist_ok=True)
    link = libdir / "libconsent_b200.so"
    if not link.exists():
        os.symlink(emu, link)
    env = dict(os.environ, LD_LIBRARY_PATH=str(libdir))
    out = subprocess.run([exe, "-X", "-a", paf, "-r", fa] + FLAGS, check=True, capture_output=True, env=env).stdout
    assert out == want


@pytest.mark.gpu
def test_dropin_binary_with_device_side_ingest_and_post_filters_on_the_gpu(small, gpu_lib):
    d, paf, fa, want = small
    out = subprocess.run([_binary("consent_correction_b200"), "-X", "-a", paf, "-r", fa] + FLAGS, check=True, capture_output=True).stdout
    assert out == want


@pytest.mark.gpu
def test_dropin_binary_with_device_side_extraction_on_the_gpu(small, gpu_lib):
    d, paf, fa, want = small
    out = subprocess.run([_binary("consent_correction_b200"), "-x", "-a", paf, "-r", fa] + FLAGS, check=True, capture_output=True).stdout
    assert out == want


@pytest.mark.gpu
def test_dropin_binary_on_the_gpu_prints_the_reference_fasta(small, gpu_lib):
    """The same through libconsent_b200.so on a B200: bit-exact corrected FASTA for the real-data example."""
    d, paf, fa, want = small
    out = subprocess.run([_binary("consent_correction_b200"), "-a", paf, "-r", fa] + FLAGS, check=True, capture_output=True).stdout
    assert out.count(b">") == 20
    assert out == want


# ---------------------------------------------------------------------------------------------------- CONSENT-polish (config 5 shape)
# bin/CONSENT-polishing = the same path with contigs as queries, windows on a thread pool and no trimming
# (src/CONSENT-polishing.cpp:19-111).  Fixture: tests/golden/make_polish_small.py (synthetic contigs + 30x ONT reads, golden =
# what the unmodified reference polisher prints).  The drop-in binary is the polisher when given -R.
@pytest.fixture(scope="module")
def polish(tmp_path_factory):
    d = tmp_path_factory.mktemp("polish_small")
    files = {}
    for name in ("polish_small.paf", "polish_small_contigs.fasta", "polish_small_reads.fasta"):
        files[name] = str(d / name)
        open(files[name], "wb").write(gzip.open(os.path.join(GOLD, name + ".gz")).read())
    want = gzip.open(os.path.join(GOLD, "polish_small_polished.fasta.gz")).read()
    args = ["-a", files["polish_small.paf"], "-r", files["polish_small_contigs.fasta"], "-R", files["polish_small_reads.fasta"]]
    return d, args, want


def test_polish_golden_is_what_the_unmodified_reference_polisher_prints(polish):
    d, args, want = polish
    out = subprocess.run([_binary("consent_polishing_ref")] + args + ["-j", "4", "-p", "/nonexistent"] + FLAGS, check=True, capture_output=True).stdout
    assert out.count(b">") == 4
    assert out == want


@pytest.mark.parametrize("mode", [[], ["-x"], ["-X"]])
def test_dropin_binary_polishes_like_the_reference_on_emulated_kernels(polish, entry, mode):
    d, args, want = polish
    exe = _binary("consent_correction_b200")
    emu = entry.build_emu()
    libdir = d / "emulib"
    libdir.mkdir(exist_ok=True)
    link = libdir / "libconsent_b200.so"
    if not link.exists():
        os.symlink(emu, link)
    env = dict(os.environ, LD_LIBRARY_PATH=str(libdir))
    out = subprocess.run([exe] + mode + args + FLAGS, check=True, capture_output=True, env=env).stdout
    assert out == want


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [[], ["-X"]])
def test_dropin_binary_polishes_like_the_reference_on_the_gpu(polish, gpu_lib, mode):
    d, args, want = polish
    out = subprocess.run([_binary("consent_correction_b200")] + mode + args + FLAGS, check=True, capture_output=True).stdout
    assert out == want


# ---------------------------------------------------------------------------------------------------- the product's own host programs
# consent_b200/bin/CONSENT-correction / CONSENT-polishing (consent_b200/host/*.cpp): own FASTA reader, PAF streamed in bounded
# batches of whole piles, one host thread + handle per GPU, records written in input order.  No reference code is linked.
def _product(entry, name="CONSENT-correction"):
    entry.build_bin()
    return os.path.join(ROOT, "consent_b200", "bin", name)


def _emu_env(d, entry, **extra):
    emu = entry.build_emu()
    libdir = d / "emulib"
    libdir.mkdir(exist_ok=True)
    link = libdir / "libconsent_b200.so"
    if not link.exists():
        os.symlink(emu, link)
    return dict(os.environ, LD_LIBRARY_PATH=str(libdir), **extra)


@pytest.mark.parametrize("batch_bytes", ["100000000", "20000", "1"])
def test_product_corrector_on_emulated_kernels_prints_the_reference_fasta(small, entry, batch_bytes):
    """One batch, a handful of batches, one pile per batch: always the unmodified reference binary's FASTA, byte for byte."""
    d, paf, fa, want = small
    out = subprocess.run([_product(entry), "-a", paf, "-r", fa, "-j", "4", "-p", "/nonexistent"] + FLAGS, check=True, capture_output=True,
                         env=_emu_env(d, entry, CONSENT_BATCH_BYTES=batch_bytes)).stdout
    assert out == want


def test_product_polisher_on_emulated_kernels_prints_the_reference_fasta(polish, entry):
    d, args, want = polish
    out = subprocess.run([_product(entry, "CONSENT-polishing")] + args + FLAGS, check=True, capture_output=True,
                         env=_emu_env(d, entry, CONSENT_BATCH_BYTES="30000")).stdout
    assert out == want


@pytest.mark.gpu
@pytest.mark.parametrize("batch_bytes", ["100000000", "20000"])
def test_product_corrector_on_the_gpu_prints_the_reference_fasta(small, gpu_lib, entry, batch_bytes):
    d, paf, fa, want = small
    out = subprocess.run([_product(entry), "-a", paf, "-r", fa] + FLAGS, check=True, capture_output=True,
                         env=dict(os.environ, CONSENT_BATCH_BYTES=batch_bytes)).stdout
    assert out == want


@pytest.mark.gpu
def test_product_polisher_on_the_gpu_prints_the_reference_fasta(polish, gpu_lib, entry):
    d, args, want = polish
    out = subprocess.run([_product(entry, "CONSENT-polishing")] + args + FLAGS, check=True, capture_output=True).stdout
    assert out == want
