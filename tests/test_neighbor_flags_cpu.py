"""the eight neighbour-presence probes BruijnGraph::makeInitialBruijnGraph makes per k-mer
(graph.h:337-375).  tests/golden/flags_k*.npz come from the reference's own KMER primitives + Counter::findValue
(oracle/ref_iter_harness.cpp, mode flags; generator oracle/make_golden_flags.py).  Here: the oracle's restatement against them
(k = 21 .. 129: one to five key words)."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = sorted(glob.glob(os.path.join(HERE, "golden", "flags_k*.npz")))


def test_golden_flag_cases_exist():
    assert len(CASES) >= 6


@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_neighbor_flags_match_the_reference(oracle, path):
    g = np.load(path, allow_pickle=False)
    got = oracle.neighbor_flags(int(g["k"]), g["keys"], g["counts"], 1)
    assert np.array_equal(got, g["flags"])
    assert (g["flags"] != 0).mean() > 0.9                      # a real graph: nearly every k-mer has a neighbour


def test_min_count_hides_entries_like_loadkmer_does(oracle):
    """entries below the cutoff are not in the table loadKmer builds (counter.h:600-640): they neither get flags nor count as
    somebody's neighbour"""
    g = np.load(CASES[1], allow_pickle=False)
    k, keys, counts = int(g["k"]), g["keys"], g["counts"].copy()
    cut = int(np.median(counts))
    kept = counts >= cut
    sub = oracle.neighbor_flags(k, keys[kept], counts[kept], 1)
    full = oracle.neighbor_flags(k, keys, counts, cut)
    assert np.array_equal(full[kept], sub) and not full[~kept].any()
