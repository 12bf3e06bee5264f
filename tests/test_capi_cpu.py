"""CPU-side checks of libpbk.so: it loads without a GPU, exports every symbol include/pbk.h declares,
and its host-side pieces (cutoff rule, table sizing, .tsv and kmer_occ.bin writers) agree with the
golden vectors of the reference binary.  No device compute happens here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import golden_cases as G
from platanus_b_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    return capi.load_library()


def test_library_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "pbk.h")).read()
    declared = set(re.findall(r"\b(pbk_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"pbk_ctx", "pbk_config", "pbk_stats", "pbk_status"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.pbk_abi_version() == 1


def test_no_torch_types_in_the_abi():
    hdr = open(os.path.join(ROOT, "include", "pbk.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)          # signatures only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code


def test_create_fails_loudly_without_a_device(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.PbkError) as e:
        capi.KmerCounter(32)
    assert e.value.status == -2           # PBK_E_NO_DEVICE: there is no CPU fallback


def test_unsupported_k_is_rejected_before_touching_the_device(L):
    ctx = C.c_void_p()
    for k in (0, 257):
        cfg = capi.PbkConfig(C.sizeof(capi.PbkConfig), k, -1, 0, 1, 0, 0, 0)
        assert L.pbk_create(C.byref(ctx), C.byref(cfg)) == -10
    cfg = capi.PbkConfig(4, 32, -1, 0, 1, 0, 0, 0)     # struct_size too small
    assert L.pbk_create(C.byref(ctx), C.byref(cfg)) == -1


def test_strerror_covers_all_codes(L):
    for code in capi.STATUS:
        assert L.pbk_strerror(code) not in (None, b"", b"unknown status")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("case", [c for c in G.CASES if not c.expect_fail], ids=lambda c: c.name)
def test_host_statistics_against_reference_golden(L, oracle, case, tmp_path):
    """cutoff, averages, .tsv and kmer_occ.bin produced by libpbk's host side from the oracle's
    histogram/table must equal what the reference binary printed and wrote."""
    O = oracle
    g = np.load(G.golden_path(case), allow_pickle=False)
    rd = O.Reads()
    for f in G.materialise(case, str(tmp_path)):
        rd.add_file(f)
    res = O.count(rd, case.k)
    occ = np.ascontiguousarray(res.occ_hist)
    cutoff = L.pbk_coverage_cutoff(_p(occ), res.max_occ, case.n_opt, int(case.repeat))
    assert cutoff == int(g["cutoff"])
    assert L.pbk_left_local_min(_p(occ), res.max_occ, 1) == O.left_local_min(occ, res.max_occ, 1)
    out = C.c_double()
    assert L.pbk_distribution_average(_p(res.len_hist), len(res.len_hist), 0, 500000, C.byref(out)) == 0
    assert "%g" % out.value == str(g["ave_read_len"])
    tsv = str(tmp_path / "x.tsv")
    assert L.pbk_write_frq_tsv(tsv.encode(), _p(occ), res.max_occ) == 0
    assert open(tsv).read() == str(g["tsv"])
    dh = L.pbk_double_hash_size(10 ** 9, case.k)
    assert dh == O.double_hash_size(10 ** 9, case.k)
    keep = res.counts >= cutoff
    keys = np.ascontiguousarray(res.keys[keep])
    counts = np.ascontiguousarray(res.counts[keep])
    path = str(tmp_path / "x_kmer_occ.bin")
    load = C.c_uint64()
    assert L.pbk_write_kmer_occ_bin(path.encode(), case.k, _p(keys), _p(counts), len(counts), dh, C.byref(load)) == 0
    assert load.value == O.load_size(len(counts))
    t = O.read_bin(path)
    assert t.reachable and t.k == case.k and t.index_size == int(g["index_size"])
    k2, c2 = t.sorted_dump()
    assert np.array_equal(k2, g["keys"]) and np.array_equal(c2, g["counts"])


def test_empty_distribution_is_kmer_dist_error(L):
    occ = np.zeros(65535, np.uint64)
    out = C.c_double()
    assert L.pbk_distribution_average(_p(occ), 65535, 2, 0, C.byref(out)) == -7
    assert L.pbk_distribution_average(_p(occ), 65535, 0, 10, C.byref(out)) == -7


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "platanus_b")), reason="reference binary not built")
@pytest.mark.parametrize("name", ["smallfa_k32", "smallfa_k75"])
def test_reference_kmer_divide_accepts_our_bin(L, oracle, name, tmp_path):
    """Integration oracle (check 3): the reference's own `kmer_divide` reads the
    kmer_occ.bin written by libpbk and produces the same output as with the reference's file."""
    import subprocess
    O = oracle
    case = G.CASE_BY_NAME[name]
    files = G.materialise(case, str(tmp_path))
    ref = O.run_reference(files, case.k, str(tmp_path), n_opt=1, prefix="ref", parse_bin=False)
    assert ref.returncode == 0
    rd = O.Reads()
    for f in files:
        rd.add_file(f)
    res = O.count(rd, case.k)
    ours = str(tmp_path / "ours_kmer_occ.bin")
    dh = L.pbk_double_hash_size(10 ** 9, case.k)
    keys = np.ascontiguousarray(res.keys)
    counts = np.ascontiguousarray(res.counts)
    assert L.pbk_write_kmer_occ_bin(ours.encode(), case.k, _p(keys), _p(counts), len(counts), dh, None) == 0
    # contigs: a few reads glued together, header in the format ContigDivider expects (common.h:659-704)
    bases, offs = rd.arrays()
    contig = str(tmp_path / "contigs.fa")
    with open(contig, "w") as f:
        for i in range(3):
            seq = bytes(bases[int(offs[5 * i]):int(offs[5 * i + 5])]).decode().upper().replace("N", "A")
            f.write(f">seq{i + 1}_len{len(seq)}_cov27_read250_maxK{case.k}\n{seq}\n")
    outs = []
    for tag, binf in (("r", str(tmp_path / "ref_kmer_occ.bin")), ("o", ours)):
        p = subprocess.run([O.REF_BINARY, "kmer_divide", "-k", binf, "-f", contig, "-o", str(tmp_path / tag)],
                           capture_output=True, text=True, cwd=str(tmp_path))
        assert p.returncode == 0, p.stderr
        produced = sorted(x for x in os.listdir(tmp_path) if x.startswith(tag + "_") or x.startswith(tag + "."))
        outs.append({x[1:]: open(tmp_path / x, "rb").read() for x in produced if not x.endswith(".bin")})
    assert outs[0] and outs[0] == outs[1]


def test_host_cpp_side_builds_and_fails_loudly_without_a_device(tmp_path):
    """platanus_b_b200/host (pbk::Counter + pbk_assemble) compiles with plain g++ against include/pbk.h; without a
    CUDA device the program must stop with pbk::GPUError (exit code 64) -- no CPU fallback, no output files."""
    import subprocess

    import torch

    from platanus_b_b200 import build as pbuild
    cli = pbuild.build_cli()
    assert os.access(cli, os.X_OK)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    fa = os.path.join(ROOT, "tests", "golden", "inputs", "kat.fa")
    p = subprocess.run([cli, "assemble", "-kmer_occ_only", "-k", "4", "-n", "1", "-o", str(tmp_path / "x"), "-f", fa],
                       capture_output=True, text=True)
    assert p.returncode == 64 and "GPU" in p.stderr
    assert not os.path.exists(tmp_path / "x_4merFrq.tsv") and not os.path.exists(tmp_path / "x_kmer_occ.bin")
    p = subprocess.run([cli, "assemble", "-k", "8"], capture_output=True, text=True)      # no -f: usage (baseCommand.cpp:56-136)
    assert p.returncode == 1 and "Usage" in p.stderr


def test_counter_shim_mirrors_the_reference_member_names():
    """pbk::Counter must offer the members Assemble::initialKmerAssemble calls on Counter<KMER> (assemble.cpp:303-350)."""
    hpp = open(os.path.join(ROOT, "platanus_b_b200", "host", "pbk_counter.hpp")).read()
    for name in ("makeKmerReadDistributionMT", "getLeftLocalMinimalValue", "calcOccurrenceDistributionAverage",
                 "calcLengthDistributionAverage", "getMaxOccurrence", "outputOccurrenceDistribution", "sortedKeyFromKmerFile",
                 "loadKmer", "outputOccurrenceTableBinary", "setKmerLength", "getKmerLength", "getLengthDistributionI", "kmerFP"):
        assert re.search(r"\b%s\b" % name, hpp), name
