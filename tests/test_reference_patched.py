"""the reference itself, with the libpbk seam applied to counter.h (integration/patch_reference.py,
built by `make -C oracle ref_patched`), runs the FULL `assemble` -- first-k counting, graph construction, bubble crush, branch
cut, the iterative k rounds (k = 32 -> 42 -> ... -> 75: one-, two- and three-word keys through makeKmerReadDistributionMT,
pickupReadMatchedEdgeKmer and makeKmerReadDistributionConsideringPreviousGraph) -- and must give the byte-identical contig FASTA,
kmerFrq.tsv and stderr log as the unmodified reference at -t 1.

The same body runs twice: on CPU against the host-emulated C ABI (LD_LIBRARY_PATH puts tests/cpu_emul's libpbk.so in front),
and with `-m gpu` against the product library on the B200."""
import dataclasses
import os
import subprocess

import pytest

from platanus_b_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "platanus_b")
PATCHED = os.path.join(ROOT, "oracle", "_ref_patched", "platanus_b_pbk")
needs_bins = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(PATCHED)),
                                reason="oracle/_ref/platanus_b and oracle/_ref_patched/platanus_b_pbk are built where /root/reference exists")


def _log(text):
    # the first lines echo the command line, the last ones report memory use
    return [ln for ln in text.splitlines() if not ln.startswith(("Vm", "/")) and "platanus_b" not in ln]


def _run(binary, workdir, files, threads, extra=(), env=None):
    os.makedirs(workdir, exist_ok=True)
    p = subprocess.run([binary, "assemble", "-t", str(threads), "-m", "1", "-tmp", workdir, "-o", os.path.join(workdir, "out"), "-f", *files, *extra],
                       capture_output=True, text=True, cwd=workdir, env=dict(os.environ, **(env or {})), timeout=1500)
    return p


def _full_assemble_is_identical(tmp_path, genome_bp, coverage, env, extra=(), config="C1"):
    base = synth.config(config)
    spec = dataclasses.replace(synth.config(config, scale=genome_bp / base.total_genome), coverage=coverage)
    rs = synth.make_reads(spec)
    files = synth.write_fastq(rs, str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq"))
    ref = _run(REF, str(tmp_path / "ref"), files, 1, extra)
    assert ref.returncode == 0, ref.stderr[-2000:]
    got = _run(PATCHED, str(tmp_path / "pbk"), files, 1, extra, env)
    assert got.returncode == 0, got.stderr[-2000:]
    assert "saving additional kmers(not found in contigs)" in ref.stderr          # the iterative-k rounds really ran
    assert _log(got.stderr) == _log(ref.stderr)
    for name in ("out_contig.fa", "out_32merFrq.tsv"):
        a, b = open(tmp_path / "ref" / name, "rb").read(), open(tmp_path / "pbk" / name, "rb").read()
        assert len(a) > 0 and a == b, name
    return ref.stderr


@needs_bins
def test_patched_reference_full_assemble_on_the_emulated_abi(tmp_path):
    import emul_helper
    emul_helper.abi_cli_path()                                        # builds tests/cpu_emul/_build/cli/libpbk.so
    env = {"LD_LIBRARY_PATH": os.path.join(ROOT, "tests", "cpu_emul", "_build", "cli")}
    # the dynamic loader must really pick the emulated library, otherwise this test would need a GPU
    ldd = subprocess.run(["ldd", PATCHED], capture_output=True, text=True, env=dict(os.environ, **env)).stdout
    assert "cpu_emul/_build/cli/libpbk.so" in ldd, ldd
    _full_assemble_is_identical(tmp_path, 50_000, 40.0, env)


@needs_bins
def test_patched_reference_in_hash_range_passes_on_the_emulated_abi(tmp_path):
    """The same full assemble with every count of the iterative schedule done in THREE hash-range passes over the SEQ temp files
    (PBK_NUM_PASSES: what pbk::Counter does by itself when a table does not fit the HBM budget -- the counterpart of the
    reference's temp-file rounds, counter.h:340-364): first-k counting, the seeded counts of the later k, the entries published
    to kmerFP from the passes' temporary file -- byte-identical contigs, .tsv and log."""
    import emul_helper
    emul_helper.abi_cli_path()
    env = {"LD_LIBRARY_PATH": os.path.join(ROOT, "tests", "cpu_emul", "_build", "cli"), "PBK_NUM_PASSES": "3"}
    _full_assemble_is_identical(tmp_path, 30_000, 40.0, env)


@needs_bins
def test_patched_reference_long_reads_up_to_four_word_keys_on_the_emulated_abi(tmp_path):
    """2x250 bp reads (config C3 scaled down): the reference's own schedule runs k = 32, 42, ... 122, 125, i.e. Kmer31, Binstr63,
    Binstr95 and Binstr127 keys, with coverage cutoffs 3 -> 1 along the way."""
    import emul_helper
    emul_helper.abi_cli_path()
    env = {"LD_LIBRARY_PATH": os.path.join(ROOT, "tests", "cpu_emul", "_build", "cli")}
    log = _full_assemble_is_identical(tmp_path, 40_000, 60.0, env, config="C3")
    assert "K = 122, saving additional kmers" in log


@needs_bins
@pytest.mark.gpu
@pytest.mark.skipif(bool(os.environ.get("PBK_TEST_EMULATED_ABI")), reason="covered by the CPU tests above")
def test_patched_reference_full_assemble_on_the_b200(tmp_path):
    ldd = subprocess.run(["ldd", PATCHED], capture_output=True, text=True).stdout
    assert "platanus_b_b200/_lib/libpbk.so" in ldd, ldd
    _full_assemble_is_identical(tmp_path, 400_000, 60.0, None)
