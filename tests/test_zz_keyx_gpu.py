"""Sharding, second form (include/pbk.h, pbk_keyx_*): the k-mers are exchanged BEFORE counting.  G logical shards on one
B200, the all-to-all replaced by device-to-device copies of its equal splits; results must equal the unsharded oracle
(result invariance).  Kept in a file that sorts last: of these cases only (k = 32, G = 2) had been
run on a B200 when round 1 ended (GPU budget spent); the kernels' logic is covered on CPU by
tests/test_kernel_logic_cpu.py::test_key_exchange_is_result_invariant and tests/test_sharded_gloo_cpu.py."""
import numpy as np
import pytest

import golden_cases as G
from platanus_b_b200 import KmerCounter, synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]      # a hung kernel must not hang the box


def _reads(O, case, tmp_path):
    rd = O.Reads()
    for f in G.materialise(case, str(tmp_path)):
        rd.add_file(f)
    return rd


def _oracle_reads_from_set(O, rs):
    rd = O.Reads()
    b, o = rs.flat()
    rd.add_array(b, o)
    return rd


def _keyx_logical_shards(O, b, o, k, n_shards):
    """G logical shards on one GPU through pbk_keyx_*: partition per rank, the all-to-all replaced by device-to-device
    copies of the equal splits, gathered Pass B per rank, then the record route for whatever was staged."""
    from devbuf import DevBuf
    n = len(o) - 1
    ctxs = [KmerCounter(k, n_shards=n_shards, shard_rank=r) for r in range(n_shards)]
    bufs = []
    try:
        parts = [(b[int(o[n * r // n_shards]):int(o[n * (r + 1) // n_shards])],
                  o[n * r // n_shards:n * (r + 1) // n_shards + 1] - o[n * r // n_shards]) for r in range(n_shards)]
        max_w = max(int(po[-1]) - (len(po) - 1) * (k - 1) for _, po in parts)
        lays = [kc.keyx_plan(max_w) for kc in ctxs]
        lay = lays[0]
        assert all((l.n_regions, l.seg_cap, l.bytes_per_dest) == (lay.n_regions, lay.seg_cap, lay.bytes_per_dest) for l in lays)
        assert lay.n_dest == n_shards and lay.bytes_per_dest == lay.n_regions * lay.seg_cap * 8
        bpd, cpd = int(lay.bytes_per_dest), int(lay.cursors_per_dest) * 8
        sends, curs = [], []
        for kc, (pb, po) in zip(ctxs, parts):
            sends.append(DevBuf(n_shards * bpd)); curs.append(DevBuf(n_shards * cpd))
            bufs += [sends[-1], curs[-1]]
            kc.keyx_partition(pb, po, sends[-1].ptr, curs[-1].ptr)
        sent = sum(int(np.minimum(c.to_host(), lay.seg_cap).sum()) for c in curs)
        for dest, kc in enumerate(ctxs):
            recv, rcur = DevBuf(n_shards * bpd), DevBuf(n_shards * cpd)
            bufs += [recv, rcur]
            for src in range(n_shards):
                recv.copy_from(sends[src], src * bpd, dest * bpd, bpd)
                rcur.copy_from(curs[src], src * cpd, dest * cpd, cpd)
            kc.keyx_insert_device(recv.ptr, rcur.ptr)
        # record route: keys that found their segment full
        W = 1
        counts = [kc.shard_send_counts(n_shards) for kc in ctxs]
        staged = int(sum(c.sum() for c in counts))
        packs = []
        for kc, cnt in zip(ctxs, counts):
            packs.append(DevBuf((int(cnt.sum()) + 1) * (W + 1) * 8)); bufs.append(packs[-1])
            kc.shard_pack_device(packs[-1].ptr, int(cnt.sum()) + 1)
        for dest, kc in enumerate(ctxs):
            for src in range(n_shards):
                m = int(counts[src][dest])
                if src == dest or m == 0:
                    continue
                part = DevBuf(m * (W + 1) * 8); bufs.append(part)
                part.copy_from(packs[src], 0, int(counts[src][:dest].sum()) * (W + 1) * 8, m * (W + 1) * 8)
                kc.shard_insert_device(part.ptr, m)
        keys, cts, inst, hist = [], [], 0, np.zeros(65535, np.uint64)
        for kc in ctxs:
            kc.finalize()
            kk, cc = kc.export(1, sorted=True)
            keys.append(kk); cts.append(cc); inst += kc.n_instances; hist += kc.occ_hist
        keys = np.concatenate(keys); cts = np.concatenate(cts)
        order = np.argsort(keys[:, 0], kind="stable")
        return keys[order], cts[order], inst, hist, sent, staged, [int(kc.n_distinct) for kc in ctxs]
    finally:
        for kc in ctxs:
            kc.close()
        for d in bufs:
            d.free()


@pytest.mark.parametrize("k,n_shards", [(32, 2), (21, 3), (32, 8)])
def test_key_exchange_logical_shards_match_unsharded(oracle, k, n_shards):
    """SURVEY.md section 8e, second form of the exchange (include/pbk.h, pbk_keyx_*): the keys travel before counting."""
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 60))
    b, o = rs.flat()
    want = O.count(_oracle_reads_from_set(O, rs), k)
    keys, cts, inst, hist, sent, staged, sizes = _keyx_logical_shards(O, b, o, k, n_shards)
    assert staged == 0 and sent == want.n_instances                  # uniform hashes: no segment overflows
    assert np.array_equal(keys, want.keys) and np.array_equal(cts, want.counts)
    assert inst == want.n_instances and np.array_equal(hist, want.occ_hist)
    assert max(sizes) < 1.2 * (sum(sizes) / n_shards) + 64


def test_key_exchange_full_segments_take_the_record_route(oracle, tmp_path):
    """70 000 copies of one read: ten k-mers with 70 000 instances each overflow their (destination, region) segments;
    the surplus is pre-aggregated in the remote-staging table and reaches its owner as (key, count) records.  Counts
    saturate at 65 534 after summing."""
    O = oracle
    case = G.CASE_BY_NAME["sat_k32"]
    rd = _reads(O, case, tmp_path)
    want = O.count(rd, case.k)
    b, o = rd.arrays()
    keys, cts, inst, hist, sent, staged, _ = _keyx_logical_shards(O, b, o, case.k, 4)
    assert staged > 0 and sent < want.n_instances
    assert np.array_equal(keys, want.keys) and np.array_equal(cts, want.counts)
    assert inst == want.n_instances and np.array_equal(hist, want.occ_hist)


def test_heavily_duplicated_input_cannot_exhaust_the_overflow_list(oracle):
    """Amplicon-like input: 1.2 M copies of one 41 bp read = 12 M windows of ten k-mers.  Each k-mer has 1.2 M instances
    but its bucket segment holds ~0.2 M, so ~10 M windows spill in Pass A -- more than the 4 M records the overflow list
    used to hold (ERR_OVERFLOW_LOST).  The list is now as long as the batch; the counts saturate at 65 534 like the
    reference's (counter.h:468)."""
    O = oracle
    copies = 1_200_000
    read = np.frombuffer(G.SAT_READ, dtype=np.uint8)
    b = np.tile(read, copies)
    o = np.arange(copies + 1, dtype=np.uint64) * np.uint64(len(read))
    rd = O.Reads()
    rd.add_array(b, o)
    want = O.count(rd, 32)
    assert want.n_instances == copies * 10 and int(want.counts.max()) == 65534
    with KmerCounter(32) as kc:
        kc.push_reads(b, o)
        kc.finalize()
        keys, counts = kc.export(1, sorted=True)
        assert kc.n_instances == want.n_instances
        assert kc.stats()["launches_partition"] >= 1            # the bucket pass was taken
    assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts)
    assert np.array_equal(kc.occ_hist, want.occ_hist)


# ---- third form: the pull exchange (pbk_keyx_pull_*) -- Pass B reads the peers' bucket stores in place ---------------------------

def _pull_logical_shards(O, b, o, k, n_shards, n_batches=1):
    """G logical shards on one GPU (contexts of one process, pbk_keyx_pull_connect_local): per batch every shard partitions into
    its own store, then -- all partitions done: the host is the barrier here -- every shard inserts what its peers hold for it."""
    n = len(o) - 1
    ctxs = [KmerCounter(k, n_shards=n_shards, shard_rank=r) for r in range(n_shards)]
    try:
        cuts = [n * i // (n_shards * n_batches) for i in range(n_shards * n_batches + 1)]
        parts = [(b[int(o[cuts[i]]):int(o[cuts[i + 1]])], o[cuts[i]:cuts[i + 1] + 1] - o[cuts[i]]) for i in range(n_shards * n_batches)]
        max_w = max(max(int(po[-1]) - (len(po) - 1) * (k - 1), 0) for _, po in parts)
        lays = [kc.keyx_pull_setup(max_w) for kc in ctxs]
        assert all((l.n_regions, l.seg_cap) == (lays[0].n_regions, lays[0].seg_cap) for l in lays)
        for r, kc in enumerate(ctxs):
            for s, peer in enumerate(ctxs):
                if s != r:
                    kc.keyx_pull_connect_local(s, peer)
        staged = 0
        for bi in range(n_batches):                               # successive batches alternate between the two store parities
            for r, kc in enumerate(ctxs):
                pb, po = parts[bi * n_shards + r]
                kc.keyx_pull_partition(pb, po)
            for kc in ctxs:
                kc.keyx_pull_insert()
            counts = [kc.shard_send_counts(n_shards) for kc in ctxs]      # record route: keys that found their segment full
            staged += int(sum(c.sum() for c in counts))
            from devbuf import DevBuf
            packs, tmp = [], []
            for kc, cnt in zip(ctxs, counts):
                packs.append(DevBuf((int(cnt.sum()) + 1) * 16)); tmp.append(packs[-1])
                kc.shard_pack_device(packs[-1].ptr, int(cnt.sum()) + 1)
            for dest, kc in enumerate(ctxs):
                for src in range(n_shards):
                    m = int(counts[src][dest])
                    if src == dest or m == 0:
                        continue
                    part = DevBuf(m * 16); tmp.append(part)
                    part.copy_from(packs[src], 0, int(counts[src][:dest].sum()) * 16, m * 16)
                    kc.shard_insert_device(part.ptr, m)
            for d in tmp:
                d.free()
        keys, cts, inst, hist = [], [], 0, np.zeros(65535, np.uint64)
        for kc in ctxs:
            kc.finalize()
            kk, cc = kc.export(1, sorted=True)
            keys.append(kk); cts.append(cc); inst += kc.n_instances; hist += kc.occ_hist
        keys = np.concatenate(keys); cts = np.concatenate(cts)
        order = np.argsort(keys[:, 0], kind="stable")
        return keys[order], cts[order], inst, hist, staged, [int(kc.n_distinct) for kc in ctxs]
    finally:
        for kc in ctxs:
            kc.close()


@pytest.mark.parametrize("k,n_shards,n_batches", [(32, 2, 1), (21, 3, 3), (32, 8, 2)])
def test_pull_exchange_logical_shards_match_unsharded(oracle, k, n_shards, n_batches):
    """SURVEY.md section 8e, third form (include/pbk.h, pbk_keyx_pull_*): nobody sends keys, every shard's Pass B reads its peers'
    stores through mapped pointers.  Several batches: the two store parities alternate and the table is kept between them."""
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 60))
    b, o = rs.flat()
    want = O.count(_oracle_reads_from_set(O, rs), k)
    keys, cts, inst, hist, staged, sizes = _pull_logical_shards(O, b, o, k, n_shards, n_batches)
    assert staged == 0
    assert np.array_equal(keys, want.keys) and np.array_equal(cts, want.counts)
    assert inst == want.n_instances and np.array_equal(hist, want.occ_hist)
    assert max(sizes) < 1.2 * (sum(sizes) / n_shards) + 64


def test_pull_exchange_full_segments_take_the_record_route(oracle, tmp_path):
    O = oracle
    case = G.CASE_BY_NAME["sat_k32"]
    rd = _reads(O, case, tmp_path)
    want = O.count(rd, case.k)
    b, o = rd.arrays()
    keys, cts, inst, hist, staged, _ = _pull_logical_shards(O, b, o, case.k, 4)
    assert staged > 0
    assert np.array_equal(keys, want.keys) and np.array_equal(cts, want.counts)
    assert inst == want.n_instances and np.array_equal(hist, want.occ_hist)
