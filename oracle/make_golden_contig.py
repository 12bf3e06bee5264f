#!/usr/bin/env python
"""Generate tests/golden/contig_k*.npz: outputs of the UNMODIFIED reference's Counter::makeKmerReadDistributionFromContig
(counter.h:511-593) through oracle/ref_iter_harness.cpp (mode `contig`).  TEST INFRASTRUCTURE ONLY.
Contigs: overlapping pieces of a committed input's reads with different coverages in their headers (so that shared k-mers
must take the larger value), one shorter than k, lower case included; no N (the reference does not skip N windows here); the largest coverage is 65534 (65535 makes the reference index its histogram out of bounds, counter.h:496)."""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT
sys.path.insert(1, os.path.join(ROOT, "tests"))

from oracle import oracle as O          # noqa: E402
import golden_cases as G                # noqa: E402

CASES = [(21, 1), (32, 1), (33, 3), (75, 1), (97, 1), (161, 2)]


def main():
    rd = O.Reads()
    rd.add_file(os.path.join(G.INPUTS, "small.fa"))
    b, o = rd.arrays()
    reads = [bytes(b[int(o[i]):int(o[i + 1])]).decode().upper() for i in range(len(o) - 1)]
    reads = [r for r in reads if len(r) >= 250 and "N" not in r]
    for k, min_occ in CASES:
        a, c, d = reads[0], reads[1], reads[2]
        contigs = [(a, 12), (a[40:] + c[:120], 40), (c[60:].lower(), 7), (d[:k - 1], 99), (d[:k], 0), (c[100:200] + d, 300), (a[:k + 30], 65534)]
        text = ""
        for i, (s, cov) in enumerate(contigs):
            text += f">seq{i + 1}_len{len(s)}_cov{cov}_read150_maxK{k}\n" + "".join(s[j:j + 70] + "\n" for j in range(0, len(s), 70))
        with tempfile.TemporaryDirectory() as tmp:
            fa = os.path.join(tmp, "c.fa")
            open(fa, "w").write(text)
            keys, counts, max_occ = O.run_ref_contig(k, fa, min_occ, tmp)
        out = os.path.join(G.GOLDEN, f"contig_k{k}.npz")
        np.savez_compressed(out, k=k, min_occ=min_occ, contigs_fa=text, coverage=np.array([cv for _, cv in contigs], np.int64),
                            keys=keys, counts=counts, max_occ=max_occ)
        print(f"{out}: {len(counts)} k-mers, values {sorted(set(counts.tolist()))}, max {max_occ}, {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
