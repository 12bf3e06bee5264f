"""The host C++ side of the drop-in (platanus_b_b200/host: pbk::Counter + the pbk_assemble program) against the
UNMODIFIED reference program on the same input files: stderr markers, exit code, PREFIX_<k>merFrq.tsv bytes,
PREFIX_kmer_occ.bin as a sorted (key, count) dump + header, and the reference's own `kmer_divide` consuming our .bin
(checks 2 and 3).  Run on a B200 with `pytest -m gpu`."""
import os
import re
import subprocess

import numpy as np
import pytest

import golden_cases as G
from platanus_b_b200 import build as pbuild
from platanus_b_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "platanus_b")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built")


@pytest.fixture(scope="module")
def cli():
    return pbuild.build_cli()


def run_ours(cli, files, k, workdir, prefix="gpu", threads=2, mem_gb=1, n_opt=0, repeat=False, extra=(), env=None):
    cmd = [cli, "assemble", "-kmer_occ_only", "-k", str(k), "-t", str(threads), "-m", str(mem_gb), "-tmp", workdir,
           "-o", os.path.join(workdir, prefix), "-f", *files, *extra]
    if n_opt:
        cmd += ["-n", str(n_opt)]
    if repeat:
        cmd += ["-repeat"]
    return subprocess.run(cmd, capture_output=True, text=True, cwd=workdir, env=dict(os.environ, **(env or {})))


def markers(stderr, k):
    """the lines scripts and `iterate` logs rely on (assemble.cpp:306, 334, 663-665, 189; counter.h:607)"""
    keep = []
    for ln in stderr.splitlines():
        if ln.startswith(("K = ", "AVE_READ_LEN=", "KMER_EXTENSION:", "K=", "loading kmers", "assemble completed")):
            keep.append(ln)
    return keep


@needs_ref
@pytest.mark.parametrize("k,repeat,seq_tmp", [(32, False, False), (75, False, False), (32, True, False), (32, False, True), (97, False, True)])
def test_pbk_assemble_matches_the_reference_program(oracle, cli, k, repeat, seq_tmp, tmp_path):
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 25))
    files = synth.write_fastq(rs, str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq"))
    ref = O.run_reference(files, k, str(tmp_path), threads=min(8, os.cpu_count() or 1), mem_gb=1, repeat=repeat)
    assert ref.returncode == 0, ref.stderr
    p = run_ours(cli, files, k, str(tmp_path), repeat=repeat, extra=("-seq_tmp",) if seq_tmp else ())
    assert p.returncode == 0, p.stderr
    assert markers(p.stderr, k) == markers(ref.stderr, k)
    assert open(tmp_path / f"gpu_{k}merFrq.tsv").read() == ref.tsv
    t = O.read_bin(str(tmp_path / "gpu_kmer_occ.bin"))
    assert t.reachable and t.k == k and t.index_size == ref.table.index_size
    gk, gc = t.sorted_dump()
    rk, rc = ref.table.sorted_dump()
    assert np.array_equal(gk, rk) and np.array_equal(gc, rc)


@needs_ref
@pytest.mark.parametrize("how", ["budget", "three_passes"])
def test_pbk_assemble_counts_in_hash_range_passes_when_the_table_does_not_fit(oracle, cli, how, tmp_path):
    """A table beyond the HBM budget (the reference's memory-limited mode writes the k-mers that found no room
    to temporary files and counts them in further rounds, counter.h:340-364, 442-449): pbk_assemble falls back from the streaming
    ingest to the SEQ temp files and pbk::Counter counts them in 2, 4, ... hash-range passes (pbk_config.n_passes), each pass's
    entries collected in a temporary file -- same .tsv, same table, same markers as the reference program.
    `budget`: PBK_HBM_BUDGET_MB leaves less room than the k = 40 table of this input needs next to the fixed buffers;
    `three_passes`: the pass count given outright (PBK_NUM_PASSES), through -seq_tmp."""
    O = oracle
    k = 40
    rs = synth.make_reads(synth.config("C1", scale=1 / 100))
    files = synth.write_fastq(rs, str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq"))
    ref = O.run_reference(files, k, str(tmp_path), threads=min(8, os.cpu_count() or 1), mem_gb=1)
    assert ref.returncode == 0, ref.stderr
    if how == "budget":
        p = run_ours(cli, files, k, str(tmp_path), env={"PBK_HBM_BUDGET_MB": "150", "PBK_TIMING": "1"})
    else:
        p = run_ours(cli, files, k, str(tmp_path), extra=("-seq_tmp",), env={"PBK_NUM_PASSES": "3", "PBK_TIMING": "1"})
    assert p.returncode == 0, p.stderr
    m = re.search(r"counted in (\d+) hash-range passes", p.stderr)
    assert m and int(m.group(1)) >= 2, p.stderr[-1500:]
    assert markers(p.stderr, k) == markers(ref.stderr, k)
    assert open(tmp_path / f"gpu_{k}merFrq.tsv").read() == ref.tsv
    t = O.read_bin(str(tmp_path / "gpu_kmer_occ.bin"))
    assert t.reachable and t.k == k and t.index_size == ref.table.index_size
    gk, gc = t.sorted_dump()
    rk, rc = ref.table.sorted_dump()
    assert np.array_equal(gk, rk) and np.array_equal(gc, rc)


@needs_ref
@pytest.mark.parametrize("k", [32, 75])
def test_reference_kmer_divide_reads_our_bin(oracle, cli, k, tmp_path):
    """Consumer test: the reference's `kmer_divide -k <bin> -f contigs` (kmer_divide.cpp:71-197) must produce the
    same output from our kmer_occ.bin as from its own."""
    O = oracle
    spec = synth.config("C1", scale=1 / 50)
    rs = synth.make_reads(spec)
    files = synth.write_fastq(rs, str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq"))
    ref = O.run_reference(files, k, str(tmp_path), threads=4, mem_gb=1, n_opt=1, parse_bin=False)
    assert ref.returncode == 0, ref.stderr
    p = run_ours(cli, files, k, str(tmp_path), n_opt=1)
    assert p.returncode == 0, p.stderr
    # contigs: slices of the genome, one with a foreign insert; header format of common.h:659-704
    g = synth.make_genome(spec.genome_lengths[0], spec.gc[0], spec.genome_seeds[0])
    genome = "".join("ACGT"[c] for c in g[:60000])
    rng = np.random.default_rng(7)
    junk = "".join("ACGT"[i] for i in rng.integers(0, 4, 400))
    contigs = [genome[1000:21000], genome[25000:33000] + junk + genome[40000:52000]]
    fa = tmp_path / "contigs.fa"
    with open(fa, "w") as fh:
        for i, c in enumerate(contigs):
            fh.write(f">seq{i + 1}_len{len(c)}_cov27_read150_maxK{k}\n{c}\n")
    outs = {}
    for tag in ("ref", "gpu"):
        q = subprocess.run([REF, "kmer_divide", "-k", str(tmp_path / f"{tag}_kmer_occ.bin"), "-f", str(fa), "-o",
                            str(tmp_path / f"div_{tag}")], capture_output=True, text=True, cwd=str(tmp_path))
        assert q.returncode == 0, q.stderr
        produced = sorted(f for f in os.listdir(tmp_path) if f.startswith(f"div_{tag}"))
        assert produced, q.stderr
        outs[tag] = [open(tmp_path / f).read() for f in produced]
    assert outs["ref"] == outs["gpu"]


@pytest.mark.parametrize("k", [32, 75])
def test_range_parallel_ingest_gives_the_same_outputs(oracle, cli, k, tmp_path):
    """The files cut into ~60 byte ranges parsed by 8 workers (pbk_ingest.hpp; PBK_INGEST_RANGE_BYTES forces small
    ranges) against one range per file: same .tsv bytes, same stderr markers, same sorted dump.  FASTQ and FASTA."""
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 50))
    files = list(synth.write_fastq(rs, str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq")))
    b, o = rs.flat()
    fa = tmp_path / "r_3.fa"
    with open(fa, "w") as fh:                                   # a FASTA file with multi-line records on top
        for i in range(0, 20000):
            s = bytes(b[int(o[i]):int(o[i + 1])]).decode()
            fh.write(f">r{i}\n{s[:70]}\n{s[70:]}\n")
    files.append(str(fa))
    outs = []
    for prefix, threads, env in (("one", 1, {"PBK_INGEST_RANGE_BYTES": str(1 << 40)}), ("many", 8, {"PBK_INGEST_RANGE_BYTES": str(1 << 20)})):
        p = run_ours(cli, files, k, str(tmp_path), prefix=prefix, threads=threads, env=env)
        assert p.returncode == 0, p.stderr
        t = O.read_bin(str(tmp_path / f"{prefix}_kmer_occ.bin"))
        assert t.reachable
        outs.append((markers(p.stderr, k), open(tmp_path / f"{prefix}_{k}merFrq.tsv").read(), t.sorted_dump()))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert np.array_equal(outs[0][2][0], outs[1][2][0]) and np.array_equal(outs[0][2][1], outs[1][2][1])


def test_pbk_assemble_on_the_golden_inputs(oracle, cli, tmp_path):
    """FASTA with multi-line records, lowercase, N, short reads (tests/golden/inputs) through the C++ parser."""
    O = oracle
    for name in ("kat_k4", "smallfa_k75", "multi_k32_n2", "tailhdr_k8"):
        case = G.CASE_BY_NAME[name]
        g = np.load(G.golden_path(case), allow_pickle=False)
        files = G.materialise(case, str(tmp_path))
        p = run_ours(cli, files, case.k, str(tmp_path), prefix=name, n_opt=case.n_opt, repeat=case.repeat)
        assert p.returncode == 0, p.stderr
        assert open(tmp_path / f"{name}_{case.k}merFrq.tsv").read() == str(g["tsv"])
        m = re.search(r"COVERAGE_CUTOFF=(\d+)", p.stderr)
        assert m and int(m.group(1)) == int(g["cutoff"])
        a = re.search(r"^AVE_READ_LEN=(\S+)", p.stderr, re.M)
        assert a and a.group(1) == str(g["ave_read_len"])
        t = O.read_bin(str(tmp_path / f"{name}_kmer_occ.bin"))
        k2, c2 = t.sorted_dump()
        assert t.reachable and t.index_size == int(g["index_size"])
        assert np.array_equal(k2, g["keys"]) and np.array_equal(c2, g["counts"])


def test_pbk_assemble_reads_gzip_and_bzip2_inputs(oracle, cli, tmp_path):
    """readFast[aq]Compressed (assemble.cpp:851-885, 945-986): a gzip FASTQ + a bzip2 FASTA give the golden outputs of the
    plain files (the reference itself cannot tell here: it asks the `file` utility, which this image lacks)."""
    import shutil
    if shutil.which("gzip") is None or shutil.which("bzip2") is None:
        pytest.skip("gzip/bzip2 not installed")
    O = oracle
    case = G.CASE_BY_NAME["multi_k32_n2"]
    g = np.load(G.golden_path(case), allow_pickle=False)
    packed = []
    for f, prog in zip(G.materialise(case, str(tmp_path)), ("gzip", "bzip2")):
        packed.append(str(tmp_path / (os.path.basename(f) + ".z")))
        with open(packed[-1], "wb") as out:
            subprocess.run([prog, "-c", f], stdout=out, check=True)
    p = run_ours(cli, packed, case.k, str(tmp_path), prefix="z", n_opt=case.n_opt)
    assert p.returncode == 0, p.stderr
    assert open(tmp_path / f"z_{case.k}merFrq.tsv").read() == str(g["tsv"])
    k2, c2 = O.read_bin(str(tmp_path / "z_kmer_occ.bin")).sorted_dump()
    assert np.array_equal(k2, g["keys"]) and np.array_equal(c2, g["counts"])


def test_pbk_assemble_error_codes(cli, tmp_path):
    """empty distribution -> KmerDistError, exit code 6 (counter.h:225-237, main.cpp:121-124); not FASTA/FASTQ -> ReadError (4)"""
    case = G.CASE_BY_NAME["empty_k8"]
    files = G.materialise(case, str(tmp_path))
    p = run_ours(cli, files, case.k, str(tmp_path))
    assert p.returncode == 6 and "kmer distribution" in p.stderr
    bad = tmp_path / "bad.txt"
    bad.write_text("hello\nworld\n")
    p = run_ours(cli, [str(bad)], 8, str(tmp_path))
    assert p.returncode == 4
    p = subprocess.run([cli, "assemble", "-k", "8"], capture_output=True, text=True)          # no -f: usage, exit 1
    assert p.returncode == 1 and "Usage" in p.stderr


@needs_ref
@pytest.mark.parametrize("k,n_gpus", [(32, 2), (75, 3)])
def test_pbk_assemble_on_several_gpus_gives_identical_files(oracle, cli, k, n_gpus, tmp_path):
    """The C++ host program on a pbk_group (include/pbk.h): PBK_NUM_GPUS devices -- logical shards when the box has fewer
    (PBK_GROUP_LOGICAL, a test switch; `gpurun --gpus 2` runs the k = 32 case on two real B200s) -- must write the same .tsv,
    print the same marker lines and dump the same sorted table as the unmodified reference program."""
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 25))
    files = synth.write_fastq(rs, str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq"))
    ref = O.run_reference(files, k, str(tmp_path), threads=min(8, os.cpu_count() or 1), mem_gb=1)
    assert ref.returncode == 0, ref.stderr
    p = run_ours(cli, files, k, str(tmp_path), env={"PBK_NUM_GPUS": str(n_gpus), "PBK_GROUP_LOGICAL": "1", "PBK_TIMING": "1"})
    assert p.returncode == 0, p.stderr
    assert f"counting on {n_gpus} GPUs" in p.stderr
    assert markers(p.stderr, k) == markers(ref.stderr, k)
    assert open(tmp_path / f"gpu_{k}merFrq.tsv").read() == ref.tsv
    t = O.read_bin(str(tmp_path / "gpu_kmer_occ.bin"))
    assert t.reachable and t.k == k and t.index_size == ref.table.index_size
    gk, gc = t.sorted_dump()
    rk, rc = ref.table.sorted_dump()
    assert np.array_equal(gk, rk) and np.array_equal(gc, rc)
