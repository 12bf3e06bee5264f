"""Parity tests proper: the CUDA path through the C ABI (libpbk.so) against the oracle, the committed
reference-binary golden vectors and -- at full BASELINE sizes -- size-independent properties.
Run on a B200 with `pytest -m gpu`.  Bar: bit-exact (integer work)."""
import dataclasses
import os

import numpy as np
import pytest

import golden_cases as G
from platanus_b_b200 import KmerCounter, PbkError, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reads(O, case, tmp_path):
    rd = O.Reads()
    for f in G.materialise(case, str(tmp_path)):
        rd.add_file(f)
    return rd


def _reads_from_arrays(O, b, o):
    rd = O.Reads()
    rd.add_array(b, o)
    return rd


def _oracle_reads_from_set(O, rs):
    rd = O.Reads()
    b, o = rs.flat()
    rd.add_array(b, o)
    return rd


@pytest.mark.parametrize("case", [c for c in G.CASES if not c.expect_fail], ids=lambda c: c.name)
def test_golden_cases_bit_exact(oracle, case, tmp_path):
    """Every committed output of the reference binary: .tsv text, cutoff, header sizes and the sorted
    (key, count) dump of kmer_occ.bin, reproduced by the GPU path end to end (incl. our .bin writer)."""
    O = oracle
    g = np.load(G.golden_path(case), allow_pickle=False)
    rd = _reads(O, case, tmp_path)
    bases, offs = rd.arrays()
    with KmerCounter(case.k) as kc:
        dh = kc.make_kmer_read_distribution(bases, offs, 10 ** 9)
        cutoff = kc.coverage_cutoff(case.n_opt, case.repeat)
        assert cutoff == int(g["cutoff"])
        tsv = str(tmp_path / "g.tsv")
        kc.output_occurrence_distribution(tsv)
        assert open(tsv).read() == str(g["tsv"])
        ave_len = kc.calc_length_distribution_average()
        assert "%g" % ave_len == str(g["ave_read_len"])
        ave_cov = kc.calc_occurrence_distribution_average(cutoff, kc.get_max_occurrence())
        ave_cov = ave_cov * ave_len / (ave_len - case.k + 1.0)
        assert "%g" % (ave_cov * (ave_len - case.k + 1.0) / ave_len) == str(g["kmer_coverage"])
        keys, counts = kc.sorted_key_from_kmer_file(cutoff)
        assert np.array_equal(keys, g["keys"]) and np.array_equal(counts, g["counts"])
        path = str(tmp_path / "g_kmer_occ.bin")
        kc.output_occurrence_table_binary(path, cutoff, dh)
        t = O.read_bin(path)
        assert t.reachable and t.k == case.k and t.index_size == int(g["index_size"])
        k2, c2 = t.sorted_dump()
        assert np.array_equal(k2, g["keys"]) and np.array_equal(c2, g["counts"])
        # and the full table against the oracle (golden only holds entries >= cutoff)
        want = O.count(rd, case.k)
        allk, allc = kc.export(1, sorted=True)
        assert np.array_equal(allk, want.keys) and np.array_equal(allc, want.counts)
        assert kc.n_instances == want.n_instances
        assert np.array_equal(kc.len_hist, want.len_hist)


@pytest.mark.parametrize("name", ["kat_k4", "smallfq_k32", "smallfq_k65", "smallfa_k129", "sat_k32", "multi_k32_n2"])
def test_forced_partition_path_on_golden_cases(oracle, name, tmp_path):
    """The hash-range bucket pass (normally used for batches >= 4M windows) forced on small inputs."""
    O = oracle
    case = G.CASE_BY_NAME[name]
    rd = _reads(O, case, tmp_path)
    bases, offs = rd.arrays()
    want = O.count(rd, case.k)
    with KmerCounter(case.k, partition="force") as kc:
        half = (len(offs) - 1) // 2
        kc.push_reads(bases[:int(offs[half])], offs[:half + 1])               # two pushes: two Pass A/Pass B rounds
        kc.push_reads(bases[int(offs[half]):], offs[half:] - offs[half])
        kc.finalize()
        keys, counts = kc.export(1, sorted=True)
        assert kc.n_instances == want.n_instances
        assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts)
        assert np.array_equal(kc.occ_hist, want.occ_hist) and np.array_equal(kc.len_hist, want.len_hist)


def test_direct_and_partitioned_paths_agree(oracle):
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 40))
    b, o = rs.flat()
    out = []
    for mode in (True, False):
        with KmerCounter(32, partition=mode) as kc:
            kc.push_reads(b, o)
            kc.finalize()
            out.append(kc.export(1, sorted=True) + (kc.occ_hist.copy(), kc.n_instances))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2]) and out[0][3] == out[1][3]


def test_empty_distribution_raises_kmer_dist_error(oracle, tmp_path):
    """reference: KmerDistError from calcDistributionAverage (counter.h:225-237), exit code 6"""
    O = oracle
    case = G.CASE_BY_NAME["empty_k8"]
    rd = _reads(O, case, tmp_path)
    bases, offs = rd.arrays()
    with KmerCounter(case.k) as kc:
        kc.make_kmer_read_distribution(bases, offs, 10 ** 9)
        assert kc.n_distinct == 0 and kc.get_max_occurrence() == 0
        with pytest.raises(PbkError) as e:
            kc.calc_occurrence_distribution_average(kc.coverage_cutoff(), kc.get_max_occurrence())
        assert e.value.status == -7


@pytest.mark.parametrize("k", [32, 42, 52, 62, 72, 75])
def test_c2_k_sweep_scaled_vs_oracle(oracle, k):
    """BASELINE config 2 (k sweep 32..75, 1-3 word keys) on a 1/40-scale genome."""
    O = oracle
    rs = synth.make_reads(synth.config("C2", scale=1 / 40))
    want = O.count(_oracle_reads_from_set(O, rs), k)
    b, o = rs.flat()
    with KmerCounter(k) as kc:
        kc.push_reads(b, o)
        kc.finalize()
        keys, counts = kc.export(1, sorted=True)
        assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts)
        assert np.array_equal(kc.occ_hist, want.occ_hist)
        assert kc.coverage_cutoff() == O.coverage_cutoff(want.occ_hist, want.max_occ)


@pytest.mark.parametrize("name,scale", [("C3", 1 / 100), ("C4", 1 / 200), ("C5", 1 / 20)])
def test_other_configs_scaled_vs_oracle(oracle, name, scale):
    O = oracle
    spec = synth.config(name, scale=scale)
    rs = synth.make_reads(spec)
    want = O.count(_oracle_reads_from_set(O, rs), spec.k)
    b, o = rs.flat()
    with KmerCounter(spec.k) as kc:
        kc.push_reads(b, o)
        kc.finalize()
        keys, counts = kc.export(1, sorted=True)
        assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts)
        assert kc.coverage_cutoff(0, spec.repeat_mode) == O.coverage_cutoff(want.occ_hist, want.max_occ, 0, spec.repeat_mode)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "platanus_b")), reason="reference binary not built")
@pytest.mark.parametrize("k,repeat", [(32, False), (75, False), (32, True)])
def test_against_the_reference_binary_on_the_box(oracle, k, repeat, tmp_path):
    """The unmodified reference (`platanus_b assemble -kmer_occ_only`, OpenMP, all its threads) and the
    GPU path on the same FASTQ files: .tsv bytes, auto cutoff and sorted table dump."""
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 20))
    files = synth.write_fastq(rs, str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq"))
    ref = O.run_reference(files, k, str(tmp_path), threads=min(8, os.cpu_count() or 1), mem_gb=1, repeat=repeat)
    assert ref.returncode == 0, ref.stderr
    b, o = rs.flat()
    with KmerCounter(k) as kc:
        dh = kc.make_kmer_read_distribution(b, o, 10 ** 9)
        cutoff = kc.coverage_cutoff(0, repeat)
        assert cutoff == ref.cutoff
        tsv = str(tmp_path / "g.tsv")
        kc.output_occurrence_distribution(tsv)
        assert open(tsv).read() == ref.tsv
        keys, counts = kc.sorted_key_from_kmer_file(cutoff)
        rk, rc = ref.table.sorted_dump()
        assert np.array_equal(keys, rk) and np.array_equal(counts, rc)
        assert "%g" % kc.calc_length_distribution_average() == ref.ave_read_len
        path = str(tmp_path / "g_kmer_occ.bin")
        kc.output_occurrence_table_binary(path, cutoff, dh)
        t = O.read_bin(path)
        assert t.reachable and t.index_size == ref.table.index_size


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "platanus_b")


def _compare_with_reference_run(O, rs, k, repeat, tmp_path, mem_gb, counter_kwargs=None):
    """the unmodified reference program and the GPU path on the same FASTQ files: .tsv bytes, auto cutoff, averages as printed,
    .bin header and the sorted (key, count) dump of everything >= cutoff"""
    files = synth.write_fastq(rs, str(tmp_path / "r_1.fq"), str(tmp_path / "r_2.fq"))
    ref = O.run_reference(files, k, str(tmp_path), threads=os.cpu_count() or 1, mem_gb=mem_gb, repeat=repeat)
    assert ref.returncode == 0, ref.stderr[-2000:]
    b, o = rs.flat()
    with KmerCounter(k, **(counter_kwargs or {})) as kc:
        dh = kc.make_kmer_read_distribution(b, o, mem_gb * 10 ** 9)
        cutoff = kc.coverage_cutoff(0, repeat)
        assert cutoff == ref.cutoff
        tsv = str(tmp_path / "g.tsv")
        kc.output_occurrence_distribution(tsv)
        assert open(tsv).read() == ref.tsv
        assert "%g" % kc.calc_length_distribution_average() == ref.ave_read_len
        keys, counts = kc.sorted_key_from_kmer_file(cutoff)
        rk, rc = ref.table.sorted_dump()
        assert len(rc) > 0 and np.array_equal(keys, rk) and np.array_equal(counts, rc)
        if keys.shape[1] > 1:                                         # no key twice (a torn slot read would split a key's count)
            assert not np.any(np.all(keys[1:] == keys[:-1], axis=1))
        assert dh == ref.table.index_size + 1 or O.load_size(len(rc)) > dh
        return kc.n_instances, len(rc)


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference binary not built")
@pytest.mark.parametrize("k", [32, 75])
def test_full_size_c1_against_the_reference_binary(oracle, k, tmp_path):
    """BASELINE config 1 at FULL size (4.6 Mb, 2x150 bp, 100x: 364 M / 233 M k-mer instances) against the unmodified reference
    on the box's host cores with its default -m 16 -- the two k of the target (one-word and three-word keys)."""
    rs = synth.make_reads(synth.config("C1"))
    inst, kept = _compare_with_reference_run(oracle, rs, k, False, tmp_path, 16)
    assert inst > 200_000_000 and 0.9 * 4_600_000 < kept < 1.1 * 4_600_000


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="reference binary not built")
@pytest.mark.parametrize("name,scale", [("C3", 0.1), ("C5", 0.1)])
def test_tenth_scale_c3_c5_against_the_reference_binary(oracle, name, scale, tmp_path):
    """BASELINE configs 3 (high GC, 2x250 bp, 300x, 1 % errors: mostly singletons) and 5 (-repeat, 7 operon copies at 1000x: a few
    thousand k-mers with ~5 000 occurrences each -- same-address traffic in Pass B) at a tenth of their genome sizes."""
    spec = synth.config(name, scale=scale)
    rs = synth.make_reads(spec)
    _compare_with_reference_run(oracle, rs, spec.k, spec.repeat_mode, tmp_path, 4)


def test_c1_full_size_properties():
    """BASELINE config 1 at full size (4.6 Mb, 2x150 bp, 100x): too big for the oracle in a test, so
    check what must hold at any size."""
    spec = synth.config("C1")
    rs = synth.make_reads(spec)
    b, o = rs.flat()
    n, L = rs.reads.shape
    with KmerCounter(32, timing=True) as kc:
        kc.push_reads(b, o)
        kc.finalize()
        occ = kc.occ_hist.astype(np.int64)
        # every window without an N is counted exactly once (nothing saturates at 100x)
        is_n = (rs.reads == ord("N"))
        csum = np.concatenate([np.zeros((n, 1), np.int32), np.cumsum(is_n, axis=1, dtype=np.int32)], axis=1)
        clean = int(((csum[:, 32:] - csum[:, :-32]) == 0).sum())
        assert kc.n_instances == clean
        assert int((np.arange(65535) * occ).sum()) == clean
        assert int(occ.sum()) == kc.n_distinct
        assert int(kc.len_hist[L]) == n and int(kc.len_hist.sum()) == n
        cutoff = kc.coverage_cutoff()
        keys, counts = kc.export(cutoff, sorted=True)
        assert len(counts) == int(occ[cutoff:].sum())
        assert np.all(keys[1:, 0] > keys[:-1, 0])                      # strictly ascending, distinct
        assert int(counts.min()) >= cutoff
        # genome-sized solid set: kept k-mers ~ genome length (both strands collapse to one key)
        assert 0.95 * spec.total_genome < len(counts) < 1.10 * spec.total_genome
        first = (keys.copy(), counts.copy(), kc.occ_hist.copy())
        # idempotence: reset + recount in two pushes of different sizes gives the identical table
        kc.reset()
        cut = (n // 3) * L
        kc.push_reads(b[:cut], o[: n // 3 + 1])
        kc.push_reads(b[cut:], o[n // 3:] - o[n // 3])
        kc.finalize()
        k2, c2 = kc.export(cutoff, sorted=True)
        assert np.array_equal(first[0], k2) and np.array_equal(first[1], c2) and np.array_equal(first[2], kc.occ_hist)


def test_device_resident_push_equals_host_push(oracle):
    import torch
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 50))
    b, o = rs.flat()
    want = O.count(_oracle_reads_from_set(O, rs), 32)
    db = torch.from_numpy(b.copy()).cuda()
    do = torch.from_numpy(o.astype(np.int64)).cuda()
    torch.cuda.synchronize()
    with KmerCounter(32) as kc:
        kc.push_reads_device(db.data_ptr(), do.data_ptr(), len(o) - 1, len(b))
        kc.finalize()
        keys, counts = kc.export(1, sorted=True)
    assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts)


def test_unsorted_export_is_a_permutation_of_sorted(oracle):
    rs = synth.make_reads(synth.config("C1", scale=1 / 100))
    b, o = rs.flat()
    with KmerCounter(52) as kc:
        kc.push_reads(b, o)
        kc.finalize()
        ks, cs = kc.export(2, sorted=True)
        ku, cu = kc.export(2, sorted=False)
    order = np.lexsort((ku[:, 0], ku[:, 1]))
    assert np.array_equal(ku[order], ks) and np.array_equal(cu[order], cs)


@pytest.mark.parametrize("k,n_shards", [(32, 2), (75, 3), (32, 8)])
def test_logical_shards_on_one_gpu_match_unsharded(oracle, k, n_shards):
    """with fewer physical GPUs than shards, run G logical shards sequentially on
    one device -- same partition function, the all-to-all replaced by local copies -- and compare
    with G = 1.  Exercises pbk_shard_send_counts / pack_device / insert_device."""
    import torch
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 60))
    b, o = rs.flat()
    want = O.count(_oracle_reads_from_set(O, rs), k)
    n = len(o) - 1
    W = (k + 31) // 32
    ctxs = [KmerCounter(k, n_shards=n_shards, shard_rank=r) for r in range(n_shards)]
    try:
        sends, counts = [], []
        for r, kc in enumerate(ctxs):
            lo, hi = n * r // n_shards, n * (r + 1) // n_shards
            kc.push_reads(b[int(o[lo]):int(o[hi])], o[lo:hi + 1] - o[lo])
            cnt = kc.shard_send_counts(n_shards)
            assert cnt[r] == 0
            buf = torch.zeros((int(cnt.sum()) + 1, W + 1), dtype=torch.int64, device="cuda")
            kc.shard_pack_device(buf.data_ptr(), int(cnt.sum()) + 1)
            sends.append(buf)
            counts.append(cnt)
        torch.cuda.synchronize()
        for dest, kc in enumerate(ctxs):
            for src in range(n_shards):
                if src == dest or counts[src][dest] == 0:
                    continue
                start = int(counts[src][:dest].sum())
                part = sends[src][start:start + int(counts[src][dest])].contiguous()
                kc.shard_insert_device(part.data_ptr(), part.shape[0])
        keys, cts, inst, hist = [], [], 0, np.zeros(65535, np.uint64)
        for kc in ctxs:
            kc.finalize()
            kk, cc = kc.export(1, sorted=True)
            keys.append(kk); cts.append(cc); inst += kc.n_instances; hist += kc.occ_hist
        keys = np.concatenate(keys); cts = np.concatenate(cts)
        order = np.lexsort(tuple(keys[:, w] for w in range(W)))
        assert np.array_equal(keys[order], want.keys) and np.array_equal(cts[order], want.counts)
        assert inst == want.n_instances and np.array_equal(hist, want.occ_hist)
        sizes = [len(c) for c in cts] if False else [int(kc.n_distinct) for kc in ctxs]
        assert max(sizes) < 1.2 * (sum(sizes) / n_shards) + 64            # the mixed hash balances owners
    finally:
        for kc in ctxs:
            kc.close()


@pytest.mark.parametrize("k,partition", [(21, True), (32, "force"), (40, "force"), (75, True), (75, "force")])
def test_hash_range_passes_add_up_to_the_whole_count(oracle, k, partition):
    """pbk_config.n_passes: three contexts, each counting one hash range of the same reads (pushed in two batches), hold disjoint
    key sets whose union, histogram sum and instance sum are the unsharded count -- which is the oracle's."""
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 400))
    want = O.count(_oracle_reads_from_set(O, rs), k)
    b, o = rs.flat()
    n = len(o) - 1
    h = n // 2
    keys, cts, hist, inst = [], [], np.zeros(65535, np.uint64), 0
    for p in range(3):
        with KmerCounter(k, partition=partition, n_passes=3, pass_index=p) as kc:
            kc.push_reads(b[:int(o[h])], o[:h + 1])
            kc.push_reads(b[int(o[h]):], o[h:] - o[h])
            kc.finalize()
            kk, cc = kc.export(1, sorted=True)
            keys.append(kk); cts.append(cc); hist += kc.occ_hist; inst += kc.n_instances
            assert len(kk) > 0
    keys = np.concatenate(keys); cts = np.concatenate(cts)
    order = np.lexsort(tuple(keys[:, w] for w in range(keys.shape[1])))
    assert inst == want.n_instances
    assert np.array_equal(hist, want.occ_hist)
    assert np.array_equal(keys[order], want.keys) and np.array_equal(cts[order], want.counts)


def test_table_growth_from_a_tiny_hint(oracle):
    O = oracle
    rs = synth.make_reads(synth.config("C3", scale=1 / 200))
    b, o = rs.flat()
    want = O.count(_oracle_reads_from_set(O, rs), 40)
    # k = 40: wide slots, any capacity (k <= 32 tables start at 2^27 slots and would not need to grow here)
    with KmerCounter(40, table_slots_hint=1024) as kc:
        kc.push_reads(b, o)
        kc.finalize()
        keys, counts = kc.export(1, sorted=True)
        assert kc.stats()["n_grow"] >= 1
    assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts)


def test_config_layouts_and_pass_arguments():
    """pbk_create takes the earlier, shorter pbk_config (struct_size 40: no n_passes / pass_index) as "no passes", refuses anything
    shorter, and refuses pass arguments that make no sense (index >= count, passes together with shards)."""
    import ctypes as C
    from platanus_b_b200 import capi
    L = capi.load_library()

    class ConfigV1(C.Structure):
        _fields_ = [("struct_size", C.c_uint32), ("k", C.c_uint32), ("device", C.c_int32), ("flags", C.c_uint32),
                    ("n_shards", C.c_uint32), ("shard_rank", C.c_uint32), ("table_slots_hint", C.c_uint64), ("hbm_budget_bytes", C.c_uint64)]

    assert C.sizeof(ConfigV1) == 40 and C.sizeof(capi.PbkConfig) == 48
    create = L.pbk_create
    ctx = C.c_void_p()
    v1 = ConfigV1(40, 21, -1, 0, 0, 0, 0, 0)
    rc = create(C.byref(ctx), C.cast(C.byref(v1), C.POINTER(capi.PbkConfig)))
    assert rc == 0 and ctx.value
    L.pbk_destroy(ctx)
    v1.struct_size = 39
    assert create(C.byref(ctx), C.cast(C.byref(v1), C.POINTER(capi.PbkConfig))) == -1           # PBK_E_ARG
    for kw in (dict(n_passes=3, pass_index=3), dict(n_passes=2, pass_index=0, n_shards=2, shard_rank=1)):
        with pytest.raises(PbkError) as e:
            KmerCounter(21, **kw)
        assert e.value.status == -1
    with KmerCounter(21, n_passes=2, pass_index=1) as kc:                                             # contigs are not available in a pass
        with pytest.raises(PbkError) as e:
            kc.push_contigs(np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGT", np.uint8), np.array([0, 28], np.uint64), np.array([5], np.uint16), 1)
        assert e.value.status == -8


def test_error_behaviour():
    bad = np.frombuffer(b"ACGTACGTACGTACGTRYACGTACGTACGTACGTACGTACGTACGT", dtype=np.uint8)
    with KmerCounter(8) as kc:
        with pytest.raises(PbkError) as e:
            kc.push_reads(bad, np.array([0, len(bad)], np.uint64))
        assert e.value.status == -6                                    # PBK_E_BAD_BASE
    long_read = np.full(500000, ord("A"), np.uint8)
    with KmerCounter(8) as kc:
        with pytest.raises(PbkError) as e:
            kc.push_reads(long_read, np.array([0, 500000], np.uint64))
        assert e.value.status == -5                                    # platanus::ReadError, common.h:465
    ok = np.full(499999, ord("A"), np.uint8)
    with KmerCounter(8) as kc:
        with pytest.raises(PbkError) as e:
            kc.export(1)
        assert e.value.status == -8                                    # export before finalize
        kc.push_reads(ok, np.array([0, 499999], np.uint64))
        kc.finalize()
        keys, counts = kc.export(1)
        assert len(counts) == 1 and int(counts[0]) == 65534 and int(keys[0, 0]) == 0   # saturated poly-A
        with pytest.raises(PbkError):
            kc.push_reads(ok, np.array([0, 10], np.uint64))            # push after finalize


def test_pipelined_pass_overlap_gives_the_identical_table(oracle):
    """k <= 32, >= 4 chunks, new-key ratio known from an earlier batch: Pass B of a sub-batch runs on a second stream
    while Pass A of the next one runs (pbk_api.cu, `Pipe`).  Same reads through the serial and the pipelined path."""
    rs = synth.make_reads(synth.config("C1", scale=0.3))
    b, o = rs.flat()
    with KmerCounter(32, timing=True) as kc:
        kc.push_reads(b, o)                                   # first large batch of the context: serial, with pilot
        kc.finalize()
        assert kc.stats()["n_pipelined_batches"] == 0
        first = kc.export(1, sorted=True) + (kc.occ_hist.copy(), kc.n_instances)
        kc.reset()
        kc.push_reads(b, o)                                   # now pipelined
        kc.finalize()
        st = kc.stats()
        assert st["n_pipelined_batches"] == 1 and st["ms_count_elapsed"] > 0
        second = kc.export(1, sorted=True) + (kc.occ_hist.copy(), kc.n_instances)
    for x, y in zip(first[:3], second[:3]):
        assert np.array_equal(x, y)
    assert first[3] == second[3]
    with KmerCounter(32, pipeline=False) as kc:
        kc.push_reads(b, o)
        kc.reset()
        kc.push_reads(b, o)
        kc.finalize()
        assert kc.stats()["n_pipelined_batches"] == 0
        k3, c3 = kc.export(1, sorted=True)
    assert np.array_equal(k3, first[0]) and np.array_equal(c3, first[1])


def test_large_pushes_are_cut_into_internal_batches(oracle):
    """pbk_push_reads splits a push at read boundaries when it exceeds the internal batch limit (here forced down to
    ~1/5 of the input through PBK_MAX_PUSH_BASES); ASCII and PLATANUS encodings."""
    import subprocess
    import sys
    code = r'''
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from oracle import oracle as O
from platanus_b_b200 import KmerCounter, synth
rs = synth.make_reads(synth.config("C1", scale=1 / 80))
b, o = rs.flat()
rd = O.Reads(); rd.add_array(b, o)
want = O.count(rd, 32)
with KmerCounter(32) as kc:
    kc.push_reads(b, o); kc.finalize()
    keys, counts = kc.export(1, sorted=True)
    assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts) and kc.n_instances == want.n_instances
    assert np.array_equal(kc.len_hist, want.len_hist)
codes = np.array([O.char2bin(int(x)) for x in range(256)], dtype=np.uint8)[b]
isn = codes == 4
npos, npo = [], [0]
L = rs.read_len
for r in range(len(o) - 1):
    w = np.nonzero(isn[r * L:(r + 1) * L])[0]
    npos.extend(w.tolist()); npo.append(len(npos))
codes[isn] = 2
with KmerCounter(32) as kc:
    kc.push_reads(codes, o, encoding=1, n_pos=np.array(npos or [0], np.int32), n_pos_offsets=np.array(npo, np.uint64)); kc.finalize()
    keys, counts = kc.export(1, sorted=True)
    assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts)
print("ok")
'''
    env = dict(os.environ, PBK_MAX_PUSH_BASES=str(1_700_000))
    p = subprocess.run([sys.executable, "-c", code, ROOT], capture_output=True, text=True, env=env)
    assert p.returncode == 0 and "ok" in p.stdout, p.stderr[-2000:]


@pytest.mark.parametrize("k,partition", [(32, True), (32, "force"), (75, "force"), (21, True)])
def test_packed_two_bit_host_input_equals_ascii_input(oracle, k, partition):
    """Opt-in host form PBK_ENC_PACKED2 (pbk_pack_reads + pbk_push_reads_packed): the host packs 2 bits per base as it copies,
    the H2D copy lands in the stream buffer itself.  Lower case, N (also in the last, partial word) and several pushes; the
    result must be the oracle's, like the ASCII form's."""
    from platanus_b_b200 import capi
    O = oracle
    rs = synth.make_reads(dataclasses.replace(synth.config("C1", scale=1 / 100), n_rate=0.002))
    b, o = rs.flat()
    b = b.copy()
    b[::7] |= 0x20                                                     # lower case is the same base (Char2Bin looks at the low nibble)
    b[-5:] = ord("N")
    want = O.count(_oracle_reads_from_set(O, rs), k)
    n = len(o) - 1
    cut = n // 3
    with KmerCounter(k, partition=partition) as kc:
        for lo, hi in ((0, cut), (cut, n)):
            bb = b[int(o[lo]):int(o[hi])]
            words, npos = capi.pack_reads(bb)
            assert len(words) == (len(bb) + 31) // 32
            kc.push_reads_packed(words, o[lo:hi + 1] - o[lo], npos)
        kc.finalize()
        keys, counts = kc.export(1, sorted=True)
        assert kc.stats()["launches_pack"] == 0                        # no pack kernel ran
    want2 = O.count(_reads_from_arrays(O, b, o), k)
    assert np.array_equal(keys, want2.keys) and np.array_equal(counts, want2.counts) and np.array_equal(kc.occ_hist, want2.occ_hist)
    assert kc.n_instances == want2.n_instances and want2.n_instances < want.n_instances      # the extra N removed windows


def test_pack_reads_rejects_characters_without_a_code():
    from platanus_b_b200 import capi
    with pytest.raises(capi.PbkError) as e:
        capi.pack_reads(np.frombuffer(b"ACGTRACGT", dtype=np.uint8))
    assert e.value.status == -6


def test_unknown_characters_can_count_as_n(oracle):
    """PBK_F_UNKNOWN_AS_N: IUPAC ambiguity codes and other characters without a Char2Bin code behave like N (the default stays
    PBK_E_BAD_BASE; the reference silently miscodes them)"""
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 200))
    b, o = rs.flat()
    b = b.copy()
    iupac = np.frombuffer(b"RYKMBHVryk-", dtype=np.uint8)      # (S, W, D have the low nibbles of C, G, T: Char2Bin gives them a code)
    idx = np.arange(17, len(b), 997)
    b[idx] = iupac[np.arange(len(idx)) % len(iupac)]
    as_n = b.copy()
    as_n[idx] = ord("N")
    want = O.count(_reads_from_arrays(O, as_n, o), 32)
    with KmerCounter(32) as kc:
        with pytest.raises(PbkError) as e:
            kc.push_reads(b, o)
        assert e.value.status == -6
    with KmerCounter(32, unknown_as_n=True) as kc:
        kc.push_reads(b, o)
        kc.finalize()
        keys, counts = kc.export(1, sorted=True)
        assert kc.n_instances == want.n_instances
    assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts)
