#!/usr/bin/env python
"""Generate tests/golden/iter_k*.npz: outputs of the UNMODIFIED reference's Counter::pickupReadMatchedEdgeKmer
(counter.h:870-910) and Counter::makeKmerReadDistributionConsideringPreviousGraph (counter.h:663-750), driven by
oracle/ref_iter_harness.cpp (`make -C oracle ref ref_iter` first).  TEST INFRASTRUCTURE ONLY.

Per case: the table is what the reference program counts from the first reads of a committed input
(`assemble -kmer_occ_only -n 2`, so that some of their k-mers are missing from it and the values differ from the counts
over all reads); the reads are ALL reads of that input (N, lowercase, short reads included), dealt to 3 temp files.
Stored: the table as its sorted dump, the reads, which reads pickup keeps, and the sorted (key, count) records + maxOccurrence
that the seeded counting leaves in kmerFP."""
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

CASES = [(21, "small.fq"), (32, "small.fq"), (40, "cov.fq"), (75, "small.fq"), (97, "small.fa")]
N_TABLE_READS, THREADS = 60, 3


def main():
    for k, src in CASES:
        rd = O.Reads()
        rd.add_file(os.path.join(G.INPUTS, src))
        b, o = rd.arrays()
        reads = [bytes(b[int(o[i]):int(o[i + 1])]).decode() for i in range(len(o) - 1)]
        reads = [r for r in reads if r][:600]
        long_reads = [r for r in reads if len(r) >= k + 10]
        with tempfile.TemporaryDirectory() as tmp:
            fa_in = os.path.join(tmp, "table_reads.fa")
            with open(fa_in, "w") as fh:
                for i, r in enumerate(long_reads[:N_TABLE_READS]):
                    fh.write(f">t{i}\n{r.upper()}\n")
            p = subprocess.run([O.REF_BINARY, "assemble", "-kmer_occ_only", "-k", str(k), "-n", "2" if src == "cov.fq" else "1", "-t", "1",
                                "-m", "1", "-tmp", tmp, "-o", os.path.join(tmp, "ref"), "-f", fa_in], capture_output=True, text=True, cwd=tmp)
            assert p.returncode == 0, p.stderr
            bin_path = os.path.join(tmp, "ref_kmer_occ.bin")
            tkeys, tcounts = O.read_bin(bin_path).sorted_dump()
            kept_seqs = O.run_ref_iter("pickup", bin_path, reads, THREADS, tmp, k)
            keys, counts, max_occ = O.run_ref_iter("count", bin_path, reads, THREADS, tmp, k)
        # survivors come file by file (file j holds reads j, j + T, ...), in order: map them back to a mask over the reads
        kept = np.zeros(len(reads), bool)
        at = 0
        norm = lambda s: "".join(c if c in "ACGT" else "N" for c in s.upper())
        for j in range(THREADS):
            for i in range(j, len(reads), THREADS):
                if at < len(kept_seqs) and norm(reads[i]) == kept_seqs[at]:
                    kept[i] = True
                    at += 1
        assert at == len(kept_seqs), (at, len(kept_seqs))
        out = os.path.join(G.GOLDEN, f"iter_k{k}.npz")
        np.savez_compressed(out, k=k, reads="\n".join(reads), table_keys=tkeys, table_counts=tcounts, kept=kept,
                            keys=keys, counts=counts, max_occ=max_occ)
        print(f"{out}: table {len(tcounts)}, {len(reads)} reads, {int(kept.sum())} kept, {len(counts)} records, max {max_occ}, "
              f"{os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
