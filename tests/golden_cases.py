"""Golden-vector cases for the k-mer counting path.

The reference ships no tests or fixtures for this path, so the vectors are
outputs of the unmodified reference binary run in the build container by oracle/make_golden.py.
Inputs are small committed text files under tests/golden/inputs/ (plus one generated, RNG-free
saturation input); expected outputs are tests/golden/<case>.npz.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
INPUTS = os.path.join(GOLDEN, "inputs")


@dataclass(frozen=True)
class Case:
    name: str
    files: tuple
    k: int
    n_opt: int = 1          # -n (0 = auto cutoff)
    repeat: bool = False    # -repeat
    expect_fail: bool = False


def _sweep(prefix, files, ks, **kw):
    return [Case(f"{prefix}_k{k}", files, k, **kw) for k in ks]


CASES = (
    [Case("kat_k4", ("kat.fa",), 4)]
    + _sweep("smallfq", ("small.fq",), (21, 31, 32, 33, 63, 64, 65, 75, 96, 97))
    + _sweep("smallfa", ("small.fa",), (32, 75, 128, 129, 160, 161, 200))
    + [Case("cov_k21_auto", ("cov.fq",), 21, n_opt=0),
       Case("cov_k21_repeat", ("cov.fq",), 21, n_opt=0, repeat=True),
       Case("cov_k32_auto", ("cov.fq",), 32, n_opt=0),
       Case("cov_k40_auto", ("cov.fq",), 40, n_opt=0),
       Case("multi_k32_n2", ("small.fq", "small.fa"), 32, n_opt=2),
       Case("tailhdr_k8", ("tail_header.fa",), 8),
       Case("sat_k32", ("@sat.fa",), 32),
       Case("empty_k8", ("empty.fa",), 8, expect_fail=True)]
)
CASE_BY_NAME = {c.name: c for c in CASES}

SAT_READ = b"ACGGTCATTGACCTAGGATCCAGTTACGATCGGATTCAGCA"      # 41 bp -> 10 windows at k=32
SAT_COPIES = 70000                                           # > 65534: count must saturate


def materialise(case: Case, workdir: str) -> list:
    """Paths of the case's input files ('@name' inputs are generated into workdir)."""
    out = []
    for f in case.files:
        if f.startswith("@"):
            path = os.path.join(workdir, f[1:])
            if f == "@sat.fa":
                with open(path, "wb") as fh:
                    fh.write(b"".join(b">s\n" + SAT_READ + b"\n" for _ in range(SAT_COPIES)))
            else:
                raise KeyError(f)
            out.append(path)
        else:
            out.append(os.path.join(INPUTS, f))
    return out


def golden_path(case: Case) -> str:
    return os.path.join(GOLDEN, case.name + ".npz")
