"""Differential fuzz on CPU: random small FASTA/FASTQ inputs (N, lowercase, reads shorter than k, multi-line records,
empty lines, quality lines starting with '@', several files) through (a) the UNMODIFIED reference program, (b) the
oracle restatement, (c) the product kernels compiled for the host (tests/cpu_emul).  All three must agree on the .tsv,
the cutoff, AVE_READ_LEN and the sorted (key, count) dump.  Seeds are fixed; the reference binary is needed."""
import os
import random

import numpy as np
import pytest

from emul_helper import emul_count

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "platanus_b")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference binary not built")


def _genome(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def _reads(rng, genome, n_reads, max_len):
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    out = []
    for _ in range(n_reads):
        L = rng.randint(1, max_len)
        st = rng.randrange(0, len(genome) - L + 1)
        r = genome[st:st + L]
        if rng.random() < 0.5:
            r = "".join(comp[c] for c in reversed(r))
        r = list(r)
        for i in range(L):
            x = rng.random()
            if x < 0.01:
                r[i] = "N"
            elif x < 0.03:
                r[i] = rng.choice("ACGT")
        r = "".join(r)
        if rng.random() < 0.2:
            r = r.lower()
        out.append(r)
    return out


def _write_fasta(rng, path, reads):
    with open(path, "w") as fh:
        for i, r in enumerate(reads):
            if i == 0:
                r = r.upper()                      # checkFileFormat: line 2 must be uppercase ACGTN (baseCommand.cpp:29-50)
                fh.write(f">r{i}\n{r}\n")
                continue
            fh.write(f">r{i} some text\n")
            w = rng.choice([len(r), 60, 17]) if r else 1
            for j in range(0, len(r), w):
                fh.write(r[j:j + w] + "\n")
            if rng.random() < 0.1:
                fh.write("\n")


def _write_fastq(rng, path, reads):
    with open(path, "w") as fh:
        for i, r in enumerate(reads):
            if i == 0:
                r = r.upper()
            q = "".join(rng.choice("@I+5") for _ in r)        # quality lines may start with '@' or '+'
            fh.write(f"@r{i}\n{r}\n+\n{q}\n")


@pytest.mark.parametrize("seed", range(8))
def test_reference_oracle_and_kernels_agree_on_random_inputs(oracle, seed, tmp_path):
    O = oracle
    rng = random.Random(1000 + seed)
    k = rng.choice([1, 2, 5, 13, 21, 31, 32, 33, 47, 64, 65, 96, 97, 130])
    genome = _genome(rng, rng.randint(300, 1500))
    files = []
    for f in range(rng.randint(1, 2)):
        reads = _reads(rng, genome, rng.randint(20, 250), max(k + 40, 60))
        reads[0] = reads[0] if len(reads[0]) >= 1 else "ACGT"
        path = str(tmp_path / f"in{f}.{'fq' if (seed + f) % 2 else 'fa'}")
        (_write_fastq if path.endswith("fq") else _write_fasta)(rng, path, reads)
        files.append(path)
    n_opt, repeat = rng.choice([(0, False), (1, False), (2, False), (0, True)])
    ref = O.run_reference(files, k, str(tmp_path), threads=rng.choice([1, 3]), mem_gb=1, n_opt=n_opt, repeat=repeat)

    rd = O.Reads()
    for f in files:
        rd.add_file(f)
    want = O.count(rd, k)
    if ref.returncode != 0:                                      # nothing at or above the cutoff: KmerDistError, exit code 6
        assert ref.returncode == 6
        cut = O.coverage_cutoff(want.occ_hist, want.max_occ, n_opt, repeat)
        assert want.n_distinct == 0 or cut > want.max_occ
        return
    cutoff = O.coverage_cutoff(want.occ_hist, want.max_occ, n_opt, repeat)
    assert cutoff == ref.cutoff
    assert O.tsv_text(want.occ_hist, want.max_occ) == ref.tsv
    rk, rc = ref.table.sorted_dump()
    sel = want.counts >= cutoff
    assert np.array_equal(want.keys[sel], rk) and np.array_equal(want.counts[sel], rc)

    bases, offs = rd.arrays()
    for partition in (False, True):
        got = emul_count(bases, offs, k, partition=partition)
        assert got["err"] == 0 and got["n_instances"] == want.n_instances
        assert np.array_equal(got["keys"], want.keys) and np.array_equal(got["counts"], want.counts)
        assert np.array_equal(got["occ_hist"], want.occ_hist) and np.array_equal(got["len_hist"], want.len_hist)
