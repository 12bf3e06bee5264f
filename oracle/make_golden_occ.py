#!/usr/bin/env python
"""Generate tests/golden/occ_k*.npz: what the UNMODIFIED reference's ContigDivider::getOccurrenceArray
(kmer_divide.cpp:151-197) computes, printed by its own dumpKmerCoverage through oracle/ref_occ_harness.cpp.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (`make -C oracle ref ref_occ` first).  Per case: the reference
program counts the first reads of a committed input (`assemble -kmer_occ_only -n 1`), the harness looks up every window
of a contig file in that PREFIX_kmer_occ.bin.  Stored: the contig FASTA text, the table as its sorted (key, count) dump
(k-mers of the contigs' source reads), and the expected occurrence per window start.
"""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT
sys.path.insert(1, os.path.join(ROOT, "tests"))

from oracle import oracle as O          # noqa: E402
import golden_cases as G                # noqa: E402

CASES = [(21, "small.fq"), (32, "small.fq"), (33, "small.fq"), (75, "small.fq"), (97, "small.fa"), (129, "small.fa"), (200, "small.fa")]
N_TABLE_READS = 60


def main():
    for k, src in CASES:
        rd = O.Reads()
        rd.add_file(os.path.join(G.INPUTS, src))
        b, o = rd.arrays()
        reads = [bytes(b[int(o[i]):int(o[i + 1])]).decode() for i in range(len(o) - 1)]
        long_reads = [r for r in reads if len(r) >= k + 20]
        with tempfile.TemporaryDirectory() as tmp:
            fa_in = os.path.join(tmp, "table_reads.fa")
            with open(fa_in, "w") as fh:
                for i, r in enumerate(long_reads[:N_TABLE_READS]):
                    fh.write(f">t{i}\n{r.upper()}\n")
            p = subprocess.run([O.REF_BINARY, "assemble", "-kmer_occ_only", "-k", str(k), "-n", "1", "-t", "1", "-m", "1", "-tmp", tmp,
                                "-o", os.path.join(tmp, "ref"), "-f", fa_in], capture_output=True, text=True, cwd=tmp)
            assert p.returncode == 0, p.stderr
            bin_path = os.path.join(tmp, "ref_kmer_occ.bin")
            keys, counts = O.read_bin(bin_path).sorted_dump()
            # contigs: reads that are in the table, reads that are not, an N run, lowercase, a k-1 and a k long sequence
            a, c = long_reads[0], long_reads[N_TABLE_READS + 3]
            contigs = [
                long_reads[1] + long_reads[2][:k + 5],
                a[:k + 7] + "N" + a[k + 8:] + c,
                long_reads[4].lower()[:k + 11] + "nn" + long_reads[5][:2 * k],
                long_reads[6][:k - 1],
                long_reads[7][:k],
                "N" * 3 + long_reads[8][:k + 3] + "N",
            ]
            fa = os.path.join(tmp, "contigs.fa")
            text = ""
            for i, s in enumerate(contigs):
                text += f">seq{i + 1}_len{len(s)}_cov27_read150_maxK{k}\n"
                text += "".join(s[j:j + 70] + "\n" for j in range(0, len(s), 70))
            open(fa, "w").write(text)
            got = O.run_ref_occurrence(bin_path, fa, tmp)
            assert len(got) == len(contigs)
            for (name, arr), s in zip(got, contigs):
                assert len(arr) == max(len(s) - k + 1, 0), (name, len(arr), len(s))
            flat = np.concatenate([arr for _, arr in got]).astype(np.uint16)
            lens = np.array([len(arr) for _, arr in got], np.int64)
        out = os.path.join(G.GOLDEN, f"occ_k{k}.npz")
        np.savez_compressed(out, k=k, contigs_fa=text, keys=keys, counts=counts, occ=flat, occ_lens=lens)
        print(f"{out}: table {len(counts)} k-mers, {len(contigs)} contigs, {int(lens.sum())} windows, "
              f"{int((flat > 0).sum())} found, {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
