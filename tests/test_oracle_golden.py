"""Pins the C restatement (oracle/kmer_oracle.c) to outputs of the unmodified reference binary.

CPU only.  The golden vectors were produced by oracle/make_golden.py from
oracle/_ref/platanus_b (`assemble -kmer_occ_only`, main.cpp:70 -> assemble.cpp:140-191).
"""
import os

import numpy as np
import pytest

import golden_cases as G


def _load(case):
    return np.load(G.golden_path(case), allow_pickle=False)


def _oracle_run(O, case, tmp_path):
    reads = O.Reads()
    for f in G.materialise(case, str(tmp_path)):
        reads.add_file(f)
    return reads, O.count(reads, case.k)


@pytest.mark.parametrize("case", [c for c in G.CASES if not c.expect_fail], ids=lambda c: c.name)
def test_oracle_matches_reference_binary(oracle, case, tmp_path):
    O = oracle
    g = _load(case)
    reads, res = _oracle_run(O, case, tmp_path)

    # PREFIX_<k>merFrq.tsv: byte-identical (counter.h:1000-1007)
    assert O.tsv_text(res.occ_hist, res.max_occ) == str(g["tsv"])

    # coverage cutoff (assemble.cpp:318-321)
    cutoff = O.coverage_cutoff(res.occ_hist, res.max_occ, case.n_opt, case.repeat)
    assert cutoff == int(g["cutoff"])

    # sorted (key, count) dump of kmer_occ.bin
    keep = res.counts >= cutoff
    assert np.array_equal(res.keys[keep], g["keys"])
    assert np.array_equal(res.counts[keep], g["counts"])

    # header: kmerLength and indexSize (counter.h:300-309, 621-627; doubleHash.h:268)
    dh = O.double_hash_size(10 ** 9, case.k)
    assert int(g["k"]) == case.k
    assert max(O.load_size(int(keep.sum())), dh) - 1 == int(g["index_size"])

    # stderr numbers (assemble.cpp:323-334, 664-665), ostream default formatting == %g
    ave_len = O.dist_average(res.len_hist, 0, O.MAX_READ_LEN)
    assert "%g" % ave_len == str(g["ave_read_len"])
    ave_cov = O.dist_average(res.occ_hist, cutoff, res.max_occ)
    ave_cov = ave_cov * ave_len / (ave_len - case.k + 1.0)
    assert "%g" % (ave_cov * (ave_len - case.k + 1.0) / ave_len) == str(g["kmer_coverage"])


def test_saturation_case_really_saturates():
    g = _load(G.CASE_BY_NAME["sat_k32"])
    assert int(g["counts"].max()) == 65534 and G.SAT_COPIES > 65534


def test_reference_failure_on_empty_distribution(oracle, tmp_path):
    """No k-mer at all: the reference throws KmerDistError (counter.h:225-237); exit code = error id."""
    O = oracle
    case = G.CASE_BY_NAME["empty_k8"]
    g = _load(case)
    assert bool(g["failed"]) and int(g["returncode"]) != 0
    reads, res = _oracle_run(O, case, tmp_path)
    assert res.n_distinct == 0 and res.max_occ == 0
    with pytest.raises(O.OracleError) as e:
        O.dist_average(res.occ_hist, O.coverage_cutoff(res.occ_hist, res.max_occ), res.max_occ)
    assert e.value.code == -4


@pytest.mark.parametrize("name", ["smallfq_k32", "smallfq_k75", "smallfa_k160", "smallfa_k200", "cov_k21_repeat"])
def test_bin_writer_roundtrip_and_probe_reachability(oracle, name, tmp_path):
    """pbo_write_bin emulates loadKmer + DoubleHash::writeTable; the file must parse back to the same
    sorted dump, carry the reference's header and keep every record where find_any looks for it."""
    O = oracle
    case = G.CASE_BY_NAME[name]
    g = _load(case)
    reads, res = _oracle_run(O, case, tmp_path)
    cutoff = int(g["cutoff"])
    path = str(tmp_path / "o_kmer_occ.bin")
    O.write_bin(path, case.k, res.keys, res.counts, cutoff, O.double_hash_size(10 ** 9, case.k))
    t = O.read_bin(path)
    assert t.reachable and t.k == case.k and t.index_size == int(g["index_size"])
    keys, counts = t.sorted_dump()
    assert np.array_equal(keys, g["keys"]) and np.array_equal(counts, g["counts"])
    rec = 8 + O.key_raw_size(case.k) + (8 * ((case.k + 31) // 32) if case.k > 160 else 0) + 2
    assert os.path.getsize(path) == 16 + rec * len(counts)


def test_brute_force_agrees_with_oracle(oracle):
    """Independent pure-Python definition (check 1): every N-free window,
    min(w, revcomp(w)) under A<C<G<T, clamp at 65534."""
    import random
    O = oracle
    rng = random.Random(5)
    comp = str.maketrans("ACGT", "TGCA")
    reads = ["".join(rng.choice("ACGTN" if rng.random() < 0.02 else "ACGT") for _ in range(rng.randrange(1, 90)))
             for _ in range(150)]
    for k in (1, 5, 16, 31, 32, 33, 40, 64, 70):
        want = {}
        for r in reads:
            for i in range(len(r) - k + 1):
                w = r[i:i + k]
                if "N" in w:
                    continue
                c = min(w, w.translate(comp)[::-1])
                want[c] = want.get(c, 0) + 1
        rd = O.Reads()
        for r in reads:
            rd.add(r.encode())
        res = O.count(rd, k)
        assert res.n_instances == sum(want.values())
        got = {}
        W = (k + 31) // 32
        for row, cnt in zip(res.keys, res.counts):
            v = sum(int(row[w]) << (64 * w) for w in range(W))
            s = "".join("ACGT"[(v >> (2 * (k - 1 - j))) & 3] for j in range(k))
            got[s] = int(cnt)
        assert got == want
        # keys come out in the reference's numeric order == lexicographic order under A<C<G<T
        assert list(got) == sorted(got)


def test_char2bin_table(oracle):
    O = oracle
    for ch, code in zip("ACGTNacgtn", [0, 1, 2, 3, 4] * 2):
        assert O.char2bin(ord(ch)) == code
    assert O.char2bin(ord("R")) == 46            # anything else is garbage in the reference (common.h:256)


def test_table_size_formulas(oracle):
    O = oracle
    # -m 16, k<=32: 2^29 slots; k=75: 2^28 slots of 56 bytes
    assert O.double_hash_size(16 * 10 ** 9, 32) == 1 << 29
    assert O.double_hash_size(16 * 10 ** 9, 75) == 1 << 28
    assert [O.pair_size(k) for k in (32, 64, 96, 128, 160, 161)] == [16, 48, 56, 64, 72, 32]
    assert O.load_size(4698849) == 1 << 23
