"""Occurrence lookup (ContigDivider::getOccurrenceArray, kmer_divide.cpp:151-197) on CPU:
* the oracle's restatement against tests/golden/occ_k*.npz -- outputs of the UNMODIFIED reference's getOccurrenceArray /
  dumpKmerCoverage (oracle/ref_occ_harness.cpp, oracle/make_golden_occ.py);
* the product lookup kernel compiled for the host (tests/cpu_emul) against the oracle and the golden vectors;
* the product .bin reader (pbk_read_kmer_occ_bin) against files written by the reference program and by our writer."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import golden_cases as G
from platanus_b_b200 import capi

CASES = sorted(glob.glob(os.path.join(G.GOLDEN, "occ_k*.npz")), key=lambda p: int(os.path.basename(p)[5:-4]))


def parse_fasta(text):
    seqs, cur = [], None
    for ln in text.splitlines():
        if ln.startswith(">"):
            cur = []
            seqs.append(cur)
        elif cur is not None:
            cur.append(ln)
    return ["".join(s) for s in seqs]


def contig_reads(O, g):
    rd = O.Reads()
    seqs = parse_fasta(str(g["contigs_fa"]))
    for s in seqs:
        rd.add(s.encode())
    return rd, seqs


def expected_per_base(g, seqs, k):
    """golden: occ per window start, concatenated per contig -> one entry per base (0 where no window starts)"""
    out, at = [], 0
    for s, n in zip(seqs, g["occ_lens"]):
        a = np.zeros(len(s), np.uint16)
        a[:int(n)] = g["occ"][at:at + int(n)]
        at += int(n)
        out.append(a)
    return np.concatenate(out)


def test_there_are_golden_vectors():
    assert len(CASES) >= 6


@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_occurrence_array_matches_the_reference(oracle, path):
    O = oracle
    g = np.load(path, allow_pickle=False)
    k = int(g["k"])
    rd, seqs = contig_reads(O, g)
    want = expected_per_base(g, seqs, k)
    assert int((want > 0).sum()) > 0
    got = O.occurrence_array(rd, k, g["keys"], g["counts"])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_lookup_kernel_matches_the_reference(oracle, path):
    from emul_helper import emul_lookup
    O = oracle
    g = np.load(path, allow_pickle=False)
    k = int(g["k"])
    rd, seqs = contig_reads(O, g)
    bases, offs = rd.arrays()
    got = emul_lookup(bases, offs, k, g["keys"], g["counts"])
    assert np.array_equal(got, expected_per_base(g, seqs, k))


@pytest.mark.parametrize("k", [21, 32, 40, 75])
def test_lookup_kernel_on_the_reads_the_table_was_counted_from(oracle, k):
    """every usable window of the counted reads is found with its count; windows with N are 0"""
    from emul_helper import emul_lookup
    O = oracle
    rd = O.Reads()
    rd.add_file(os.path.join(G.INPUTS, "cov.fq"))
    want = O.count(rd, k)
    bases, offs = rd.arrays()
    got = emul_lookup(bases, offs, k, want.keys, want.counts)
    ref = O.occurrence_array(rd, k, want.keys, want.counts)
    assert np.array_equal(got, ref)
    assert int((got > 0).sum()) == want.n_instances
    # a table holding only every other key: the rest must read 0
    half = emul_lookup(bases, offs, k, want.keys[::2], want.counts[::2])
    assert np.array_equal(half, O.occurrence_array(rd, k, want.keys[::2], want.counts[::2]))


def _read_bin(L, path):
    k, idx, n = C.c_uint32(), C.c_uint64(), C.c_uint64()
    keys, counts = C.c_void_p(), C.c_void_p()
    rc = L.pbk_read_kmer_occ_bin(path.encode(), C.byref(k), C.byref(idx), C.byref(keys), C.byref(counts), C.byref(n))
    if rc:
        return rc, None
    W = (k.value + 31) // 32
    kk = np.ctypeslib.as_array(C.cast(keys, C.POINTER(C.c_uint64)), shape=(max(n.value, 1) * W,))[:n.value * W].copy().reshape(n.value, W)
    cc = np.ctypeslib.as_array(C.cast(counts, C.POINTER(C.c_uint16)), shape=(max(n.value, 1),))[:n.value].copy()
    L.pbk_free(keys); L.pbk_free(counts)
    return 0, (k.value, idx.value, kk, cc)


@pytest.mark.parametrize("k", [21, 32, 64, 75, 128, 161, 200])
def test_bin_reader_round_trip_with_our_writer(oracle, k, tmp_path):
    O = oracle
    L = capi.load_library()
    rd = O.Reads()
    rd.add_file(os.path.join(G.INPUTS, "small.fa"))
    want = O.count(rd, k)
    path = str(tmp_path / "t.bin")
    dh = O.double_hash_size(10 ** 8, k)
    assert L.pbk_write_kmer_occ_bin(path.encode(), k, want.keys.ctypes.data_as(C.c_void_p), want.counts.ctypes.data_as(C.c_void_p),
                                    len(want.counts), dh, None) == 0
    rc, got = _read_bin(L, path)
    assert rc == 0
    kk, idx, keys, counts = got
    assert kk == k and idx == O.read_bin(path).index_size
    W = (k + 31) // 32
    order = np.lexsort(tuple(keys[:, w] for w in range(W)))
    assert np.array_equal(keys[order], want.keys) and np.array_equal(counts[order], want.counts)
    # truncated file -> PBK_E_IO, not garbage
    data = open(path, "rb").read()
    open(path, "wb").write(data[:-3])
    assert _read_bin(L, path)[0] == -9


@pytest.mark.parametrize("k", [32, 75])
def test_bin_reader_on_a_file_written_by_the_reference_program(oracle, k, tmp_path):
    O = oracle
    if not O.have_ref_binary():
        pytest.skip("reference binary not built")
    L = capi.load_library()
    ref = O.run_reference([os.path.join(G.INPUTS, "cov.fq")], k, str(tmp_path), threads=2, mem_gb=1, n_opt=2)
    assert ref.returncode == 0, ref.stderr
    rc, got = _read_bin(L, str(tmp_path / "ref_kmer_occ.bin"))
    assert rc == 0
    kk, idx, keys, counts = got
    rk, rc_ = ref.table.sorted_dump()
    W = (k + 31) // 32
    order = np.lexsort(tuple(keys[:, w] for w in range(W)))
    assert kk == k and idx == ref.table.index_size
    assert np.array_equal(keys[order], rk) and np.array_equal(counts[order], rc_)


# ---- the two table-consuming steps of the iterative-k assembly ---------------------------
ITER_CASES = sorted(glob.glob(os.path.join(G.GOLDEN, "iter_k*.npz")), key=lambda p: int(os.path.basename(p)[6:-4]))


def iter_reads(O, g):
    rd = O.Reads()
    for r in str(g["reads"]).split("\n"):
        rd.add(r.encode())
    return rd


def test_there_are_iterative_k_golden_vectors():
    assert len(ITER_CASES) >= 5


@pytest.mark.parametrize("path", ITER_CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_iterative_k_steps_match_the_reference(oracle, path):
    """tests/golden/iter_k*.npz: what the UNMODIFIED reference's pickupReadMatchedEdgeKmer and
    makeKmerReadDistributionConsideringPreviousGraph produce (oracle/ref_iter_harness.cpp, oracle/make_golden_iter.py)."""
    O = oracle
    g = np.load(path, allow_pickle=False)
    k = int(g["k"])
    rd = iter_reads(O, g)
    assert np.array_equal(O.match_reads(rd, k, g["table_keys"], g["table_counts"]), g["kept"])
    res = O.count(rd, k, g["table_keys"], g["table_counts"])
    assert np.array_equal(res.keys, g["keys"]) and np.array_equal(res.counts, g["counts"])
    assert res.max_occ == int(g["max_occ"])
    # the seeded values differ from the plain counts somewhere, or the case would prove nothing
    plain = O.count(rd, k)
    assert not (len(plain.counts) == len(res.counts) and np.array_equal(plain.counts, res.counts))


@pytest.mark.parametrize("path", ITER_CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_kernels_iterative_k_steps_match_the_reference(oracle, path):
    from emul_helper import emul_match_reads, emul_seeded_count
    O = oracle
    g = np.load(path, allow_pickle=False)
    k = int(g["k"])
    rd = iter_reads(O, g)
    bases, offs = rd.arrays()
    assert np.array_equal(emul_match_reads(bases, offs, k, g["table_keys"], g["table_counts"]), g["kept"])
    keys, counts, n_inst = emul_seeded_count(bases, offs, k, g["table_keys"], g["table_counts"])
    assert np.array_equal(keys, g["keys"]) and np.array_equal(counts, g["counts"])


# ---- table from contigs (makeKmerReadDistributionFromContig, counter.h:511-593) ---------------------------------------------
CONTIG_CASES = sorted(glob.glob(os.path.join(G.GOLDEN, "contig_k*.npz")), key=lambda p: int(os.path.basename(p)[8:-4]))


def contig_seqs(O, g):
    rd = O.Reads()
    for s in parse_fasta(str(g["contigs_fa"])):
        rd.add(s.encode())
    return rd


@pytest.mark.parametrize("path", CONTIG_CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_contig_table_matches_the_reference(oracle, path):
    """tests/golden/contig_k*.npz: the kmerFP records the UNMODIFIED reference's makeKmerReadDistributionFromContig writes
    (oracle/ref_iter_harness.cpp mode contig, oracle/make_golden_contig.py) -- oracle restatement and the product kernel."""
    from emul_helper import emul_contigs
    O = oracle
    g = np.load(path, allow_pickle=False)
    k, min_occ = int(g["k"]), int(g["min_occ"])
    rd = contig_seqs(O, g)
    res = O.count_contigs(rd, k, g["coverage"], min_occ)
    assert np.array_equal(res.keys, g["keys"]) and np.array_equal(res.counts, g["counts"]) and res.max_occ == int(g["max_occ"])
    assert len(set(g["counts"].tolist())) >= 4                    # overlapping contigs: the larger coverage must have won somewhere
    bases, offs = rd.arrays()
    keys, counts = emul_contigs(bases, offs, k, g["coverage"], min_occ)
    assert np.array_equal(keys, g["keys"]) and np.array_equal(counts, g["counts"])
