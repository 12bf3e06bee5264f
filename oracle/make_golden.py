#!/usr/bin/env python
"""Generate tests/golden/ from the UNMODIFIED reference binary (oracle/_ref/platanus_b).

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference to have been
compiled by `make -C oracle ref`).  Usage:  python oracle/make_golden.py [--inputs]

  --inputs   (re)create the small input files under tests/golden/inputs/ first.  They are
             committed, so the RNG used here never has to be reproducible elsewhere.
"""
from __future__ import annotations

import os
import random
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT                      # not oracle/: `oracle` must resolve to the package
sys.path.insert(1, os.path.join(ROOT, "tests"))

from oracle import oracle as O          # noqa: E402
import golden_cases as G                # noqa: E402

COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def revcomp(s):
    return "".join(COMP[c] for c in reversed(s))


def sample_reads(rng, genome, n, L, sub, nrate):
    reads = []
    for _ in range(n):
        st = rng.randrange(0, len(genome) - L + 1)
        r = genome[st:st + L]
        if rng.random() < 0.5:
            r = revcomp(r)
        r = list(r)
        for i in range(L):
            x = rng.random()
            if x < sub:
                r[i] = rng.choice([c for c in "ACGT" if c != r[i]])
            elif x < sub + nrate:
                r[i] = "N"
        reads.append("".join(r))
    return reads


def make_inputs():
    os.makedirs(G.INPUTS, exist_ok=True)
    rng = random.Random(20261018)
    with open(os.path.join(G.INPUTS, "kat.fa"), "w") as f:      # SURVEY.md section 4 known-answer vector
        f.write(">r1\nACGTACGTAC\n>r2 multi-line + lowercase\nacgtn\nGGCCAATT\n>r3 short\nACG\n"
                ">r4 palindrome-rich\nAATTAATT\n")

    genome = "".join(rng.choice("ACGT") for _ in range(3000))
    # plant palindromes (own reverse complement) of even lengths incl. 32 and 64
    pal32 = "".join(rng.choice("ACGT") for _ in range(16)); pal32 += revcomp(pal32)
    pal64 = "".join(rng.choice("ACGT") for _ in range(32)); pal64 += revcomp(pal64)
    genome = genome[:500] + pal32 + genome[532:1500] + pal64 + genome[1564:]
    reads = sample_reads(rng, genome, 400, 100, 0.005, 0.005)
    reads[0] = reads[0].replace("N", "A")                        # line 2 of the file must be ACGTN upper
    with open(os.path.join(G.INPUTS, "small.fq"), "w") as f:
        for i, r in enumerate(reads):
            if i % 7 == 3:
                r = r.lower()                                    # lowercase after the first record
            if i % 50 == 10:
                r = r[:20]                                       # shorter than most k
            if i % 61 == 5:
                r = "N" * 3 + r[3:50] + "NN" + r[52:]            # N at the start and adjacent Ns
            q = "I" * len(r)
            if i % 9 == 4:
                q = "@" + q[1:]                                  # quality line starting with '@'
            if i % 11 == 6:
                q = "+" + q[1:]                                  # quality line starting with '+'
            if i % 13 == 7 and len(r) > 60:
                f.write(f"@r{i}\n{r[:60]}\n{r[60:]}\n+\n{q}\n")  # multi-line sequence
            elif i % 17 == 8:
                f.write(f"@r{i}\n{r}\n+r{i}\n{q}\n\n")           # repeated name on '+', blank line
            else:
                f.write(f"@r{i}\n{r}\n+\n{q}\n")

    genome2 = "".join(rng.choice("ACGT") for _ in range(4000))
    reads2 = sample_reads(rng, genome2, 120, 250, 0.004, 0.002)
    reads2[0] = reads2[0].replace("N", "C")
    with open(os.path.join(G.INPUTS, "small.fa"), "w") as f:
        for i, r in enumerate(reads2):
            if i % 5 == 2:
                r = r.lower()
            if i % 40 == 9:
                r = r[:30]
            f.write(f">s{i} len={len(r)}\n")
            if i == 0:
                f.write(r + "\n")
            else:
                for j in range(0, len(r), 60):
                    f.write(r[j:j + 60] + "\n")
            if i % 33 == 4:
                f.write(f">empty{i}\n")                          # header without sequence

    genome3 = "".join(rng.choice("ACGT") for _ in range(5000))
    reads3 = sample_reads(rng, genome3, 2000, 100, 0.01, 0.0)
    with open(os.path.join(G.INPUTS, "cov.fq"), "w") as f:
        for i, r in enumerate(reads3):
            f.write(f"@c{i}\n{r}\n+\n{'I' * len(r)}\n")

    with open(os.path.join(G.INPUTS, "tail_header.fa"), "w") as f:
        f.write(">a\nACGTTGCAAGGCTTAACCGGTT\n>b\nTTGACCAGTTGACCAGGTTTACA\n>c only a header: a zero-length record is emitted\n")
    with open(os.path.join(G.INPUTS, "empty.fa"), "w") as f:
        f.write(">x\nACGT\n")                                    # one read shorter than k: no k-mers at all


def main():
    if "--inputs" in sys.argv:
        make_inputs()
    if not O.have_ref_binary():
        sys.exit("build the reference first: make -C oracle ref")
    for case in G.CASES:
        with tempfile.TemporaryDirectory() as wd:
            files = G.materialise(case, wd)
            r = O.run_reference(files, case.k, wd, threads=1, mem_gb=1, n_opt=case.n_opt, repeat=case.repeat)
            if case.expect_fail:
                assert r.table is None, case
                np.savez_compressed(G.golden_path(case), returncode=r.returncode, failed=True,
                                    stderr_tail=r.stderr[-400:])
                print(f"{case.name}: reference failed as expected (rc={r.returncode})")
                continue
            assert r.returncode == 0 and r.table is not None and r.table.reachable, (case, r.stderr)
            keys, counts = r.table.sorted_dump()
            np.savez_compressed(
                G.golden_path(case), k=r.table.k, index_size=r.table.index_size, keys=keys, counts=counts,
                tsv=r.tsv, cutoff=r.cutoff, ave_read_len=r.ave_read_len, kmer_coverage=r.kmer_coverage,
                returncode=r.returncode, failed=False,
                ext_lines="\n".join(l for l in r.stderr.splitlines() if l.startswith("K=")))
            print(f"{case.name}: k={r.table.k} kept={len(counts)} cutoff={r.cutoff} ave_len={r.ave_read_len} "
                  f"index_size={r.table.index_size}")


if __name__ == "__main__":
    main()
