#!/usr/bin/env python
"""Generate tests/golden/flags_k*.npz: the eight neighbour-presence probes BruijnGraph::makeInitialBruijnGraph makes per
k-mer (graph.h:337-375), computed with the UNMODIFIED reference's own KMER primitives and Counter::findValue by
oracle/ref_iter_harness.cpp (mode `flags`; `make -C oracle ref ref_iter` first).  TEST INFRASTRUCTURE ONLY.

Per case: the table is what the reference program writes for a committed input (`assemble -kmer_occ_only`, auto cutoff or
-n), i.e. exactly the kept entries loadKmer leaves in memory when the graph builder starts.  Stored: the table's sorted dump
and one byte per key, (leftFlags << 4) | rightFlags -- the value Junction::out receives (graph.h:398)."""
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

CASES = [(21, "cov.fq", 0), (32, "cov.fq", 0), (33, "cov.fq", 2), (64, "cov.fq", 2), (75, "cov.fq", 0), (97, "small.fa", 1), (129, "small.fa", 1)]
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_iter_harness")


def main():
    for k, src, n_opt in CASES:
        with tempfile.TemporaryDirectory() as tmp:
            cmd = [O.REF_BINARY, "assemble", "-kmer_occ_only", "-k", str(k), "-t", "1", "-m", "1", "-tmp", tmp, "-o", os.path.join(tmp, "ref"),
                   "-f", os.path.join(G.INPUTS, src)] + (["-n", str(n_opt)] if n_opt else [])
            p = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
            assert p.returncode == 0, p.stderr
            bin_path = os.path.join(tmp, "ref_kmer_occ.bin")
            keys, counts = O.read_bin(bin_path).sorted_dump()
            out = os.path.join(tmp, "flags.bin")
            p = subprocess.run([HARNESS, "flags", bin_path, "-", "1", out], capture_output=True, text=True, cwd=tmp)
            assert p.returncode == 0, p.stderr
            W = (k + 31) // 32
            raw = np.fromfile(out, dtype=np.uint8).reshape(-1, 8 * W + 1)
            hk = np.ascontiguousarray(raw[:, :8 * W]).view(np.uint64).reshape(-1, W)
            flags = raw[:, 8 * W].copy()
        assert np.array_equal(hk, keys), "the harness walks the keys in the order of the sorted dump"
        path = os.path.join(ROOT, "tests", "golden", f"flags_k{k}.npz")
        np.savez_compressed(path, k=np.int64(k), keys=keys, counts=counts, flags=flags)
        n_j = int(((np.unpackbits(flags[:, None] >> 4, axis=1).sum(1) > 1) | (np.unpackbits(flags[:, None] & 15, axis=1).sum(1) > 1)).sum())
        print(f"{os.path.basename(path)}: {len(flags)} keys, {n_j} with a branching side, {int((flags == 0).sum())} isolated")


if __name__ == "__main__":
    main()
