#!/usr/bin/env python
"""Differential fuzz of the C ABI's host logic + kernels (compiled for the host, tests/cpu_emul) against the oracle:
random read sets (N, lower case, reads shorter than k, duplicates), random k, one or several pushes, direct / forced-partition
/ no-pipeline routes, seeded counting, lookups, read matching, and the two sharded exchanges with 2-5 logical shards.
TEST TOOL (CPU only, never shipped):  python scripts/fuzz_emulated_abi.py [--seconds 300] [--seed 1]"""
import argparse
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ["PBK_TEST_EMULATED_ABI"] = "1"
import emul_helper          # noqa: E402
from platanus_b_b200 import build          # noqa: E402
build.LIB = emul_helper.abi_lib_path()
from devbuf import DevBuf          # noqa: E402
from oracle import oracle as O          # noqa: E402
from platanus_b_b200 import KmerCounter, capi          # noqa: E402


def random_reads(rng, k):
    genome = "".join(rng.choice("ACGT") for _ in range(rng.randint(50, 3000)))
    n = rng.randint(1, 400)
    reads = []
    for _ in range(n):
        L = rng.choice([rng.randint(1, max(2, k + 5)), rng.randint(k, k + 150), rng.randint(1, 300)])
        st = rng.randrange(0, max(1, len(genome) - L + 1))
        r = list(genome[st:st + L]) or ["A"]
        for i in range(len(r)):
            x = rng.random()
            if x < 0.01:
                r[i] = rng.choice("ACGT")
            elif x < 0.013:
                r[i] = "N"
            elif x < 0.016:
                r[i] = r[i].lower()
        reads.append("".join(r))
    if rng.random() < 0.3:
        reads += [reads[0]] * rng.randint(1, 50)
    return reads


def as_arrays(reads):
    b = np.frombuffer("".join(reads).encode(), dtype=np.uint8).copy()
    o = np.zeros(len(reads) + 1, np.uint64)
    o[1:] = np.cumsum([len(r) for r in reads])
    return b, o


def oracle_reads(reads):
    rd = O.Reads()
    for r in reads:
        rd.add(r.encode())
    return rd


def check_table(kc, want, what):
    keys, counts = kc.export(1, sorted=True)
    assert np.array_equal(keys, want.keys) and np.array_equal(counts, want.counts), what
    assert np.array_equal(kc.occ_hist, want.occ_hist), what


def one_case(rng, case_no):
    k = rng.choice([rng.randint(1, 32), rng.randint(33, 96), rng.randint(97, 200), 31, 32, 33, 64, 65])
    W = (k + 31) // 32
    reads = random_reads(rng, k)
    rd = oracle_reads(reads)
    want = O.count(rd, k)
    b, o = as_arrays(reads)
    mode = rng.choice(["plain", "multi_push", "seeded", "lookup", "records", "keys", "contigs", "passes"])
    if mode == "keys" and W > 1:
        mode = "records"
    partition = rng.choice([True, "force", False])
    pipeline = rng.random() < 0.7
    if "PBK_FUZZ_FIXED_FORM" not in os.environ:      # which form of Pass B (k <= 32) the contexts of this case use: read by pbk_create
        os.environ["PBK_PASSB2"] = os.environ["PBK_PASSB2_GATHER"] = rng.choice(["0", "1"])
    desc = f"case {case_no}: k={k} mode={mode} partition={partition} pipeline={pipeline} passb2={os.environ.get('PBK_PASSB2')} reads={len(reads)}"
    if mode in ("plain", "multi_push"):
        with KmerCounter(k, partition=partition, pipeline=pipeline, table_slots_hint=rng.choice([0, 0, 64])) as kc:
            cuts = sorted(set([0, len(reads)] + ([rng.randint(0, len(reads)) for _ in range(rng.randint(1, 3))] if mode == "multi_push" else [])))
            for a, c in zip(cuts, cuts[1:]):
                kc.push_reads(b[int(o[a]):int(o[c])], o[a:c + 1] - o[a])
            kc.finalize()
            assert kc.n_instances == want.n_instances, desc
            check_table(kc, want, desc)
            assert np.array_equal(kc.len_hist, want.len_hist), desc
            kc.reset()                                           # second life of the context, other route
            kc.push_reads(b, o)
            kc.finalize()
            check_table(kc, want, desc + " (after reset)")
    elif mode == "passes":
        # hash-range passes on one GPU (pbk_config.n_passes): disjoint key sets that add up to the whole count; seeds are handed
        # to every pass, the library keeps those of its range
        P = rng.randint(2, 5)
        seeded = rng.random() < 0.4 and len(want.counts) > 0
        ref = want
        if seeded:
            sel = np.array([rng.random() < 0.3 for _ in range(len(want.counts))], bool)
            sk, sc = np.ascontiguousarray(want.keys[sel]), np.array([rng.randint(1, 300) for _ in range(int(sel.sum()))], np.uint16)
            ref = O.count(rd, k, sk, sc)
        keys, cts, hist, inst = [], [], np.zeros(65535, np.uint64), 0
        for p in range(P):
            with KmerCounter(k, partition=partition, pipeline=pipeline, n_passes=P, pass_index=p) as kc:
                h = rng.randint(0, len(reads))
                if seeded:
                    kc.seed_entries(sk, sc)
                kc.push_reads(b[:int(o[h])], o[:h + 1]); kc.push_reads(b[int(o[h]):], o[h:] - o[h])
                kc.finalize()
                kk, cc = kc.export(1, sorted=True)
                keys.append(kk); cts.append(cc); hist += kc.occ_hist; inst += kc.n_instances
        keys = np.concatenate(keys); cts = np.concatenate(cts)
        order = np.lexsort(tuple(keys[:, w] for w in range(W)))
        assert np.array_equal(keys[order], ref.keys) and np.array_equal(cts[order], ref.counts), desc
        assert np.array_equal(hist, ref.occ_hist), desc
        if not seeded:
            assert inst == want.n_instances, desc
    elif mode == "seeded":
        sel = np.array([rng.random() < 0.3 for _ in range(len(want.counts))], bool)
        sk, sc = want.keys[sel], np.array([rng.randint(0, 300) for _ in range(int(sel.sum()))], np.uint16)
        extra = np.array([[rng.getrandbits(64) & ((1 << (2 * ((k - 1) % 32 + 1))) - 1) if w == W - 1 else rng.getrandbits(64) for w in range(W)]
                          for _ in range(rng.randint(0, 5))], np.uint64).reshape(-1, W)
        ek = {tuple(r) for r in want.keys.tolist()}
        extra = np.array(sorted({tuple(r) for r in extra.tolist() if tuple(r) not in ek}), np.uint64).reshape(-1, W)   # a table's keys are distinct
        sk = np.concatenate([sk, extra]); sc = np.concatenate([sc, np.full(len(extra), 7, np.uint16)])
        order = np.lexsort(tuple(sk[:, w] for w in range(W)))
        sk, sc = np.ascontiguousarray(sk[order]), np.ascontiguousarray(sc[order])
        ref = O.count(rd, k, sk, sc)
        with KmerCounter(k, partition=partition, pipeline=pipeline) as kc:
            if rng.random() < 0.5:
                kc.seed_entries(sk, sc); kc.push_reads(b, o)
            else:
                kc.push_reads(b, o); kc.seed_entries(sk, sc)
            kc.finalize()
            check_table(kc, ref, desc)
    elif mode == "contigs":
        seqs = [r.replace("N", "A").replace("n", "a") for r in reads]           # (the reference does not skip N windows of contigs)
        crd = oracle_reads(seqs)
        cb, co = as_arrays(seqs)
        cov = np.array([rng.choice([0, 1, 5, 40, 300, 65534]) for _ in seqs], np.uint16)
        min_occ = rng.choice([0, 1, 3])
        ref = O.count_contigs(crd, k, cov, min_occ)
        with KmerCounter(k) as kc:
            h = rng.randint(0, len(seqs))
            kc.push_contigs(cb[:int(co[h])], co[:h + 1], cov[:h], min_occ)
            kc.push_contigs(cb[int(co[h]):], co[h:] - co[h], cov[h:], min_occ)
            kc.finalize()
            check_table(kc, ref, desc)
    elif mode == "lookup":
        sel = np.array([rng.random() < 0.5 for _ in range(len(want.counts))], bool)
        tk, tc = np.ascontiguousarray(want.keys[sel]), np.ascontiguousarray(want.counts[sel])
        with KmerCounter(k) as kc:
            kc.load_entries(tk, tc)
            assert np.array_equal(kc.lookup(b, o), O.occurrence_array(rd, k, tk, tc)), desc
            assert np.array_equal(kc.match_reads(b, o), O.match_reads(rd, k, tk, tc)), desc
    else:
        G = rng.randint(2, 5)
        n = len(reads)
        parts = [(np.ascontiguousarray(b[int(o[n * r // G]):int(o[n * (r + 1) // G])]), (o[n * r // G:n * (r + 1) // G + 1] - o[n * r // G]).astype(np.uint64)) for r in range(G)]
        ctxs = [KmerCounter(k, n_shards=G, shard_rank=r, partition=partition, pipeline=pipeline) for r in range(G)]
        bufs = []
        try:
            if mode == "keys":
                max_w = max(max(int(po[-1]) - (len(po) - 1) * (k - 1), 0) for _, po in parts)
                lay = [kc.keyx_plan(max_w) for kc in ctxs][0]
                bpd, cpd = int(lay.bytes_per_dest), int(lay.cursors_per_dest) * 8
                for rep in range(2):                             # second step: the queued (no host round trip) insert
                    sends, curs = [], []
                    for kc, (pb, po) in zip(ctxs, parts):
                        if rep:
                            kc.reset()
                        sends.append(DevBuf(G * bpd)); curs.append(DevBuf(G * cpd)); bufs += [sends[-1], curs[-1]]
                        kc.keyx_partition(pb, po, sends[-1].ptr, curs[-1].ptr)
                    for dest, kc in enumerate(ctxs):
                        recv, rcur = DevBuf(G * bpd), DevBuf(G * cpd); bufs += [recv, rcur]
                        for src in range(G):
                            recv.copy_from(sends[src], src * bpd, dest * bpd, bpd)
                            rcur.copy_from(curs[src], src * cpd, dest * cpd, cpd)
                        kc.keyx_insert_device(recv.ptr, rcur.ptr)
            else:
                for kc, (pb, po) in zip(ctxs, parts):
                    kc.push_reads(pb, po)
            counts = [kc.shard_send_counts(G) for kc in ctxs]
            packs = []
            for kc, cnt in zip(ctxs, counts):
                packs.append(DevBuf((int(cnt.sum()) + 1) * (W + 1) * 8)); bufs.append(packs[-1])
                kc.shard_pack_device(packs[-1].ptr, int(cnt.sum()) + 1)
            for dest, kc in enumerate(ctxs):
                for src in range(G):
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
            order = np.lexsort(tuple(keys[:, w] for w in range(W)))
            assert np.array_equal(keys[order], want.keys) and np.array_equal(cts[order], want.counts), desc
            assert inst == want.n_instances and np.array_equal(hist, want.occ_hist), desc
        finally:
            for kc in ctxs:
                kc.close()
            for d in bufs:
                d.free()
    return desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=300)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    O.build()
    rng = random.Random(args.seed)
    t0, n = time.time(), 0
    while time.time() - t0 < args.seconds:
        state = rng.getstate()
        try:
            desc = one_case(rng, n)
        except Exception:
            print(f"FAILED at case {n} (seed {args.seed}); rng state saved to /tmp/fuzz_state_{args.seed}_{n}.pkl", flush=True)
            import pickle
            pickle.dump(state, open(f"/tmp/fuzz_state_{args.seed}_{n}.pkl", "wb"))
            raise
        n += 1
        if n % 20 == 0:
            print(f"{n} cases ok ({time.time() - t0:.0f} s); last: {desc}", flush=True)
    print(f"done: {n} cases ok in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
