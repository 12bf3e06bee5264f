"""Occurrence lookup on the B200 through the C ABI (pbk_lookup, pbk_load_entries, pbk_read_kmer_occ_bin; SURVEY.md
section 8f rows 1-2) against the reference's own getOccurrenceArray output (tests/golden/occ_k*.npz) and the oracle.
First seen to pass on a B200 at the end of round 1 (GPUTEST_r01); the kernel's logic is also covered on CPU by
tests/test_lookup_cpu.py through the host emulation."""
import ctypes as C
import dataclasses
import os

import numpy as np
import pytest

import golden_cases as G
from platanus_b_b200 import KmerCounter, synth
from test_lookup_cpu import CASES, contig_reads, expected_per_base

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300, method="thread")]


@pytest.mark.parametrize("path", CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_lookup_matches_the_reference_occurrence_arrays(oracle, path):
    O = oracle
    g = np.load(path, allow_pickle=False)
    k = int(g["k"])
    rd, seqs = contig_reads(O, g)
    bases, offs = rd.arrays()
    with KmerCounter(k) as kc:
        kc.load_entries(g["keys"], g["counts"])
        got = kc.lookup(bases, offs)
        assert np.array_equal(got, expected_per_base(g, seqs, k))
        kc.finalize()                                               # the loaded table is an ordinary table
        keys, counts = kc.export(1, sorted=True)
        assert np.array_equal(keys, g["keys"]) and np.array_equal(counts, g["counts"])
        assert kc.stats()["n_instances"] == 0                       # lookups and loads are not counted reads


@pytest.mark.parametrize("k", [32, 75])
def test_lookup_of_the_counted_reads(oracle, k):
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 100))
    b, o = rs.flat()
    rd = O.Reads()
    rd.add_array(b, o)
    want = O.count(rd, k)
    ref = O.occurrence_array(rd, k, want.keys, want.counts)
    with KmerCounter(k) as kc:
        kc.push_reads(b, o)
        got = kc.lookup(b, o)                                        # no finalize needed: the table is quiescent between calls
        assert np.array_equal(got, ref)
        assert int((got > 0).sum()) == want.n_instances
        kc.finalize()
        assert kc.n_instances == want.n_instances
        assert np.array_equal(kc.len_hist, want.len_hist)           # the lookup's reads did not enter the histograms
        # foreign sequence: nothing found; bad character: flagged like pbk_push_reads
        rng = np.random.default_rng(5)
        junk = np.frombuffer(bytes(rng.choice(list(b"ACGT"), 4000).astype(np.uint8)), dtype=np.uint8)
        assert int(kc.lookup(junk, np.array([0, len(junk)], np.uint64)).sum()) == 0
        bad = junk.copy(); bad[100] = ord("R")
        with pytest.raises(Exception):
            kc.lookup(bad, np.array([0, len(bad)], np.uint64))


def test_read_occurrence_table_binary_then_lookup(oracle, tmp_path):
    """Counter::readOccurrenceTableBinary (counter.h:967-993): a kmer_occ.bin written by our writer (verified against the
    reference's reader elsewhere) loaded into a fresh counter of another k."""
    O = oracle
    from platanus_b_b200 import capi
    L = capi.load_library()
    k = 75
    rd = O.Reads()
    rd.add_file(os.path.join(G.INPUTS, "cov.fq"))
    want = O.count(rd, k)
    keep = want.counts >= 2
    path = str(tmp_path / "t_kmer_occ.bin")
    assert L.pbk_write_kmer_occ_bin(path.encode(), k, np.ascontiguousarray(want.keys[keep]).ctypes.data_as(C.c_void_p),
                                    np.ascontiguousarray(want.counts[keep]).ctypes.data_as(C.c_void_p), int(keep.sum()),
                                    O.double_hash_size(10 ** 9, k), None) == 0
    bases, offs = rd.arrays()
    with KmerCounter(32) as kc:
        assert kc.read_occurrence_table_binary(path) == int(keep.sum())
        assert kc.k == k
        got = kc.lookup(bases, offs)
        assert np.array_equal(got, O.occurrence_array(rd, k, want.keys[keep], want.counts[keep]))


# ---- iterative-k steps (pbk_match_reads, pbk_seed_entries) against the reference's own outputs ---------------------------
from test_lookup_cpu import ITER_CASES, iter_reads          # noqa: E402


@pytest.mark.parametrize("path", ITER_CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_iterative_k_steps_match_the_reference(oracle, path):
    """tests/golden/iter_k*.npz = outputs of the unmodified Counter<KMER>::pickupReadMatchedEdgeKmer and
    makeKmerReadDistributionConsideringPreviousGraph (oracle/ref_iter_harness.cpp)."""
    O = oracle
    g = np.load(path, allow_pickle=False)
    k = int(g["k"])
    rd = iter_reads(O, g)
    bases, offs = rd.arrays()
    with KmerCounter(k) as kc:                                   # the table the previous round left
        kc.load_entries(g["table_keys"], g["table_counts"])
        assert np.array_equal(kc.match_reads(bases, offs), g["kept"])
    for order, partition in (("seed first", True), ("reads first", "force")):
        with KmerCounter(k, partition=partition) as kc:
            if order == "seed first":
                kc.seed_entries(g["table_keys"], g["table_counts"])
                kc.push_reads(bases, offs)
            else:
                kc.push_reads(bases, offs)
                kc.seed_entries(g["table_keys"], g["table_counts"])
            kc.finalize()
            keys, counts = kc.export(1, sorted=True)
            assert np.array_equal(keys, g["keys"]) and np.array_equal(counts, g["counts"])
            assert kc.max_occurrence == int(g["max_occ"])
            want = O.count(rd, k, g["table_keys"], g["table_counts"])
            assert np.array_equal(kc.occ_hist, want.occ_hist)
            kc.reset()                                               # the seeds are forgotten with the counts
            kc.push_reads(bases, offs)
            kc.finalize()
            plain = O.count(rd, k)
            keys, counts = kc.export(1, sorted=True)
            assert np.array_equal(keys, plain.keys) and np.array_equal(counts, plain.counts)


@pytest.mark.parametrize("k,n_shards", [(32, 2), (75, 3)])
def test_lookup_and_seeds_with_hash_range_shards(oracle, k, n_shards):
    """Sharded contexts keep only the entries they own (pbk_load_entries / pbk_seed_entries filter by pbk_shard_of_key), so
    the per-shard lookups of one sequence set add up to the unsharded answer, and the seeded tables partition the seeded count."""
    O = oracle
    g = np.load(ITER_CASES[1] if k == 32 else ITER_CASES[3], allow_pickle=False)
    assert int(g["k"]) == k
    rd = iter_reads(O, g)
    bases, offs = rd.arrays()
    want_occ = O.occurrence_array(rd, k, g["table_keys"], g["table_counts"])
    total = np.zeros(len(bases), np.uint32)
    keys_all, counts_all = [], []
    for r in range(n_shards):
        with KmerCounter(k, n_shards=n_shards, shard_rank=r) as kc:
            kc.load_entries(g["table_keys"], g["table_counts"])
            total += kc.lookup(bases, offs)
        with KmerCounter(k, n_shards=n_shards, shard_rank=r) as kc:
            kc.seed_entries(g["table_keys"], g["table_counts"])
            kc.finalize()                                            # seeds alone: a contig-seeded table without reads
            kk, cc = kc.export(1, sorted=True)
            keys_all.append(kk); counts_all.append(cc)
    assert np.array_equal(total.astype(np.uint16), want_occ)
    keys = np.concatenate(keys_all); counts = np.concatenate(counts_all)
    W = (k + 31) // 32
    order = np.lexsort(tuple(keys[:, w] for w in range(W)))
    assert np.array_equal(keys[order], g["table_keys"]) and np.array_equal(counts[order], g["table_counts"])


from test_lookup_cpu import CONTIG_CASES, contig_seqs          # noqa: E402


@pytest.mark.parametrize("path", CONTIG_CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_contig_table_matches_the_reference(oracle, path):
    """pbk_push_contigs = Counter::makeKmerReadDistributionFromContig (counter.h:511-593) against the reference's own output;
    pushed in two calls and in the other order as well: the maximum does not care."""
    O = oracle
    g = np.load(path, allow_pickle=False)
    k, min_occ = int(g["k"]), int(g["min_occ"])
    rd = contig_seqs(O, g)
    bases, offs = rd.arrays()
    cov = g["coverage"].astype(np.uint16)
    want = O.count_contigs(rd, k, cov, min_occ)
    n = len(offs) - 1
    for order in ("one call", "two calls, reversed"):
        with KmerCounter(k) as kc:
            if order == "one call":
                kc.push_contigs(bases, offs, cov, min_occ)
            else:
                h = n // 2
                kc.push_contigs(bases[int(offs[h]):], offs[h:] - offs[h], cov[h:], min_occ)
                kc.push_contigs(bases[:int(offs[h])], offs[:h + 1], cov[:h], min_occ)
            kc.finalize()
            keys, counts = kc.export(1, sorted=True)
            assert np.array_equal(keys, g["keys"]) and np.array_equal(counts, g["counts"])
            assert kc.max_occurrence == int(g["max_occ"]) and np.array_equal(kc.occ_hist, want.occ_hist)
            assert kc.n_instances == 0


# ---- neighbour-presence flags for the graph builder (pbk_neighbor_flags) -------------------------------
import glob as _glob

FLAG_CASES = sorted(_glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flags_k*.npz")))


@pytest.mark.parametrize("path", FLAG_CASES, ids=lambda p: os.path.basename(p)[:-4])
def test_neighbor_flags_match_the_reference_probes(path):
    """table loaded from the golden (key, count) list; the device answers makeInitialBruijnGraph's eight findValue probes per key
    (graph.h:337-375) exactly like the reference's own primitives did"""
    g = np.load(path, allow_pickle=False)
    k = int(g["k"])
    with KmerCounter(k) as kc:
        kc.load_entries(g["keys"], g["counts"])
        kc.finalize()
        assert np.array_equal(kc.neighbor_flags(g["keys"], 1), g["flags"])


@pytest.mark.parametrize("k", [32, 75])
def test_neighbor_flags_after_counting_match_the_oracle(oracle, k):
    """the real hand-off: count reads, cutoff, sorted export, flags of the kept k-mers against the resident table (which also
    holds the k-mers below the cutoff: they must not count as neighbours)"""
    O = oracle
    rs = synth.make_reads(synth.config("C1", scale=1 / 100))
    b, o = rs.flat()
    rd = O.Reads()
    rd.add_array(b, o)
    want = O.count(rd, k)
    with KmerCounter(k) as kc:
        kc.push_reads(b, o)
        kc.finalize()
        cutoff = kc.coverage_cutoff()
        keys, counts = kc.export(cutoff, sorted=True)
        flags = kc.neighbor_flags(keys, cutoff)
    sel = want.counts >= cutoff
    assert np.array_equal(keys, want.keys[sel])
    assert np.array_equal(flags, O.neighbor_flags(k, want.keys, want.counts, cutoff)[sel])
    # a 100x genome: nearly all kept k-mers are interior nodes of a path (exactly one neighbour on each side)
    assert (flags != 0).mean() > 0.99


def test_lookup_takes_contigs_longer_than_the_read_limit(oracle):
    """ContigDivider::getOccurrenceArray (kmer_divide.cpp:151-197) walks contigs of any length: a 600 kb sequence (reads are
    limited to 499 999 bases, contigs are not)"""
    O = oracle
    spec = synth.config("C1", scale=600_000 / 4_600_000)
    genome = np.frombuffer(b"ACGT", dtype=np.uint8)[synth.make_genome(spec.genome_lengths[0], spec.gc[0], spec.genome_seeds[0])]
    assert len(genome) >= 500_000
    rs = synth.make_reads(dataclasses.replace(spec, coverage=6.0))
    b, o = rs.flat()
    rd = O.Reads()
    rd.add_array(b, o)
    want = O.count(rd, 32)
    with KmerCounter(32) as kc:
        kc.push_reads(b, o)
        kc.finalize()
        got = kc.lookup(genome, np.array([0, len(genome)], np.uint64))
    seqs = O.Reads()
    seqs.add_array(genome, np.array([0, len(genome)], np.uint64))
    assert np.array_equal(got, O.occurrence_array(seqs, 32, want.keys, want.counts))
    assert (got[: len(genome) - 31] > 0).mean() > 0.9
