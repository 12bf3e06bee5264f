"""Several devices behind one handle (include/pbk.h, pbk_group_*): the library splits the batches, drives one context per
device and exchanges the keys itself -- k <= 32: pull exchange over peer-mapped memory, k > 32: (key, count) records with
cudaMemcpyPeer.  With one physical GPU the members are logical shards on the same device; `gpurun --gpus 2` runs the same
cases on two real devices (PBK_TEST_GROUP_DEVICES=0,1).  Results must equal the unsharded oracle bit for bit."""
import os

import numpy as np
import pytest

import golden_cases as G
from platanus_b_b200 import KmerGroup, load_library, synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]


def _devices(n):
    env = os.environ.get("PBK_TEST_GROUP_DEVICES")
    if env:
        phys = [int(x) for x in env.split(",")]
        return [phys[i % len(phys)] for i in range(n)]
    return [0] * n


@pytest.mark.parametrize("k,n_members,n_pushes", [(32, 2, 1), (21, 3, 3), (32, 8, 2), (75, 2, 2), (97, 3, 1)])
def test_group_matches_the_unsharded_oracle(oracle, k, n_members, n_pushes):
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 60))
    b, o = rs.flat()
    rd = O.Reads()
    rd.add_array(b, o)
    want = O.count(rd, k)
    n = len(o) - 1
    with KmerGroup(k, _devices(n_members), partition="force") as g:
        for rep in range(2):                                       # second pass: reset, stores and tables reused
            if rep:
                g.reset()
            for p in range(n_pushes):                              # pushes of different sizes: the pull layout is re-planned when one is larger
                lo, hi = n * p // (n_pushes + 1) if p else 0, n * (p + 1) // (n_pushes + 1) if p + 1 < n_pushes else n
                g.push_reads(b[int(o[lo]):int(o[hi])], o[lo:hi + 1] - o[lo])
            g.finalize()
            keys, counts = g.export(1, sorted=True)
            assert g.n_instances == want.n_instances and g.n_distinct == want.n_distinct
            assert np.array_equal(g.occ_hist, want.occ_hist) and np.array_equal(g.len_hist, want.len_hist)
            assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts)
            ku, cu = g.export(3, sorted=False)
            sel = want.counts >= 3
            order = np.lexsort(tuple(ku[:, w] for w in range(ku.shape[1])))
            assert np.array_equal(ku[order], want.keys[sel]) and np.array_equal(cu[order], want.counts[sel])
        # graph hand-off (pbk_group_neighbor_flags): every shard answers for the neighbours it owns, the group ORs them
        cutoff = O.coverage_cutoff(want.occ_hist, want.max_occ)
        kk, _ = g.export(cutoff, sorted=True)
        sel = want.counts >= cutoff
        assert np.array_equal(g.neighbor_flags(kk, cutoff), O.neighbor_flags(k, want.keys, want.counts, cutoff)[sel])
        sizes = [g.member_stats(i)["n_distinct"] for i in range(n_members)]
        assert sum(sizes) == want.n_distinct and max(sizes) < 1.25 * (sum(sizes) / n_members) + 64


def test_group_saturation_and_full_segments(oracle, tmp_path):
    """70 000 copies of one read on four members: the heavily repeated keys overflow their segments (record route inside the
    group) and saturate at 65 534 after summing"""
    O = oracle
    case = G.CASE_BY_NAME["sat_k32"]
    rd = O.Reads()
    for f in G.materialise(case, str(tmp_path)):
        rd.add_file(f)
    want = O.count(rd, case.k)
    b, o = rd.arrays()
    with KmerGroup(case.k, _devices(4), partition="force") as g:
        g.push_reads(b, o)
        g.finalize()
        keys, counts = g.export(1, sorted=True)
    assert int(counts.max()) == 65534
    assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts) and np.array_equal(g.occ_hist, want.occ_hist)


def test_device_count_is_exported():
    assert load_library().pbk_device_count() >= 1
