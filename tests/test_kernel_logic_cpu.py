"""Logic check of the PRODUCT kernels without a GPU: pbk_kernels_impl.cuh compiled for the host by
tests/cpu_emul (one sequential thread per kernel) against the oracle.  The real parity tests run the
same kernels on a B200 (tests/test_gpu_parity.py, -m gpu)."""
import numpy as np
import pytest

import golden_cases as G
from emul_helper import emul_count, emul_insert_records


def _reads(O, case, tmp_path):
    rd = O.Reads()
    for f in G.materialise(case, str(tmp_path)):
        rd.add_file(f)
    return rd


NAMES = ["kat_k4", "smallfq_k21", "smallfq_k31", "smallfq_k32", "smallfq_k33", "smallfq_k63", "smallfq_k64",
         "smallfq_k65", "smallfq_k75", "smallfq_k96", "smallfq_k97", "smallfa_k128", "smallfa_k129", "smallfa_k160",
         "smallfa_k161", "smallfa_k200", "tailhdr_k8", "multi_k32_n2"]


@pytest.mark.parametrize("partition", [False, True], ids=["direct", "partitioned"])
@pytest.mark.parametrize("name", NAMES)
def test_kernels_match_oracle(oracle, name, partition, tmp_path):
    O = oracle
    case = G.CASE_BY_NAME[name]
    rd = _reads(O, case, tmp_path)
    want = O.count(rd, case.k)
    bases, offs = rd.arrays()
    got = emul_count(bases, offs, case.k, partition=partition)
    assert got["err"] == 0
    assert got["n_instances"] == want.n_instances
    assert np.array_equal(got["keys"], want.keys)
    assert np.array_equal(got["counts"], want.counts)
    assert np.array_equal(got["occ_hist"], want.occ_hist)
    assert np.array_equal(got["len_hist"], want.len_hist)


def test_kernels_saturate_at_65534(oracle, tmp_path):
    O = oracle
    case = G.CASE_BY_NAME["sat_k32"]
    rd = _reads(O, case, tmp_path)
    want = O.count(rd, case.k)
    bases, offs = rd.arrays()
    got = emul_count(bases, offs, case.k)
    assert int(got["counts"].max()) == 65534
    assert np.array_equal(got["keys"], want.keys) and np.array_equal(got["counts"], want.counts)
    assert np.array_equal(got["occ_hist"], want.occ_hist)


@pytest.mark.parametrize("k", [21, 32, 75])
def test_platanus_encoding_matches_ascii(oracle, k, tmp_path):
    """PBK_ENC_PLATANUS (SEQ temp-file form: codes 0..3 + N position list, common.h:426-448)."""
    O = oracle
    rd = _reads(O, G.CASE_BY_NAME["smallfq_k32"], tmp_path)
    bases, offs = rd.arrays()
    want = O.count(rd, k)
    codes = np.zeros_like(bases)
    npos, npo = [], [0]
    for r in range(len(offs) - 1):
        s = bases[int(offs[r]):int(offs[r + 1])]
        c = np.array([O.char2bin(int(x)) for x in s], dtype=np.uint8)
        isn = c == 4
        npos.extend(np.nonzero(isn)[0].tolist())
        npo.append(len(npos))
        c[isn] = 3                      # stale byte under an N: must be ignored
        codes[int(offs[r]):int(offs[r + 1])] = c
    got = emul_count(codes, offs, k, encoding=1, n_pos=np.array(npos or [0], np.int32), n_pos_off=np.array(npo, np.uint64))
    assert np.array_equal(got["keys"], want.keys) and np.array_equal(got["counts"], want.counts)


def test_overflow_path_and_rehash(oracle, tmp_path):
    """A table far too small: keys spill to the overflow list, the table is rebuilt, nothing is lost."""
    O = oracle
    rd = _reads(O, G.CASE_BY_NAME["cov_k32_auto"], tmp_path)
    want = O.count(rd, 32)
    bases, offs = rd.arrays()
    got = emul_count(bases, offs, 32, table_slots=want.n_distinct + 7)
    assert np.array_equal(got["keys"], want.keys) and np.array_equal(got["counts"], want.counts)


def test_bad_base_is_flagged(oracle):
    bases = np.frombuffer(b"ACGTRACGTACGTACGTACGTACGTACGTACGTACGTACGT", dtype=np.uint8)
    got = emul_count(bases, np.array([0, len(bases)], np.uint64), 8)
    assert got["err"] & 1


@pytest.mark.parametrize("k,n_shards", [(32, 2), (32, 3), (75, 4)])
def test_hash_range_sharding_is_result_invariant(oracle, k, n_shards, tmp_path):
    """Each rank counts the keys it owns and stages the rest; merging what every rank receives must
    reproduce the unsharded table (result invariance)."""
    O = oracle
    rd = _reads(O, G.CASE_BY_NAME["smallfq_k32"], tmp_path)
    want = O.count(rd, k)
    bases, offs = rd.arrays()
    n_reads = len(offs) - 1
    W = (k + 31) // 32
    cuts = [n_reads * r // n_shards for r in range(n_shards + 1)]
    owned, staged = [], []
    for r in range(n_shards):
        lo, hi = cuts[r], cuts[r + 1]
        b = bases[int(offs[lo]):int(offs[hi])]
        o = offs[lo:hi + 1] - offs[lo]
        got = emul_count(b, o, k, n_shards=n_shards, rank=r)
        owned.append(np.concatenate([got["keys"], got["counts"].astype(np.uint64)[:, None]], axis=1))
        staged.append(got["remote"])
    import ctypes as C
    from platanus_b_b200 import capi
    L = capi.load_library()
    for dest in range(n_shards):
        recs = [owned[dest]]
        for src in range(n_shards):
            if src == dest:
                continue
            part = staged[src]
            own = np.array([L.pbk_shard_of_key(np.ascontiguousarray(row[:W]).ctypes.data_as(C.c_void_p), k, n_shards)
                            for row in part], dtype=np.int64) if len(part) else np.zeros(0, np.int64)
            recs.append(part[own == dest])
        keys, counts = emul_insert_records(np.concatenate(recs), k)
        sel = np.array([L.pbk_shard_of_key(np.ascontiguousarray(row).ctypes.data_as(C.c_void_p), k, n_shards) == dest
                        for row in want.keys], dtype=bool)
        assert np.array_equal(keys, want.keys[sel]) and np.array_equal(counts, want.counts[sel])


@pytest.mark.parametrize("k,n_shards,n_regions,seg_cap", [(32, 2, 4, 4096), (21, 3, 2, 4096), (32, 4, 8, 40), (31, 8, 4, 4096)])
def test_key_exchange_is_result_invariant(oracle, k, n_shards, n_regions, seg_cap, tmp_path):
    """Second form of the sharding (pbk_keyx_*): Pass A writes [dest][region] segments of hashes, the all-to-all hands
    every rank its [source][region] segments, Pass B (gather mode) inserts them region by region.  The merged tables must
    be the unsharded one, every key on the rank pbk_shard_of_key names.  seg_cap = 40: most keys find their segment full
    and take the record route instead."""
    import ctypes as C
    from emul_helper import emul_keyx_insert, emul_keyx_partition
    from platanus_b_b200 import capi
    L = capi.load_library()
    O = oracle
    rd = _reads(O, G.CASE_BY_NAME["smallfq_k32"], tmp_path)
    want = O.count(rd, k)
    bases, offs = rd.arrays()
    n_reads = len(offs) - 1
    cuts = [n_reads * r // n_shards for r in range(n_shards + 1)]
    sends, cursors, spills, inst = [], [], [], 0
    for r in range(n_shards):
        lo, hi = cuts[r], cuts[r + 1]
        s, c, n, sp = emul_keyx_partition(bases[int(offs[lo]):int(offs[hi])], offs[lo:hi + 1] - offs[lo], k, n_shards, n_regions, seg_cap)
        sends.append(s); cursors.append(c); spills.append(sp); inst += n
        for d in range(n_shards):                 # segment (d, j) holds hashes owned by shard d that live in table region j
            for j in range(n_regions):
                h = s[d, j, :min(int(c[d, j]), seg_cap)]
                assert np.all(h >> np.uint64(64 - int(np.log2(n_regions))) == j)
                assert np.all(((h & np.uint64(0xFFFFFFFF)) * np.uint64(n_shards)) >> np.uint64(32) == d)
    assert inst == want.n_instances
    assert (seg_cap == 40) == any(len(sp) for sp in spills)
    owner_of = lambda key: L.pbk_shard_of_key(np.array([key], np.uint64).ctypes.data_as(C.c_void_p), k, n_shards)
    spilled = np.concatenate(spills) if any(len(sp) for sp in spills) else np.zeros((0, 2), np.uint64)
    spill_owner = np.array([owner_of(int(x)) for x in spilled[:, 0]], dtype=np.int64)
    all_keys, all_counts = [], []
    for dest in range(n_shards):
        recv = np.stack([sends[src][dest] for src in range(n_shards)])              # what all_to_all_single delivers
        rcur = np.stack([cursors[src][dest] for src in range(n_shards)])
        keys, counts = emul_keyx_insert(recv, rcur, seg_cap, k, extra=spilled[spill_owner == dest])
        assert all(owner_of(int(x)) == dest for x in keys[:, 0])
        all_keys.append(keys); all_counts.append(counts)
    keys = np.concatenate(all_keys); counts = np.concatenate(all_counts)
    order = np.argsort(keys[:, 0], kind="stable")
    assert np.array_equal(keys[order], want.keys) and np.array_equal(counts[order], want.counts)
