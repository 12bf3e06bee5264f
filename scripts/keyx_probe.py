#!/usr/bin/env python
"""Round-2 opener, ONE GPU: what the key exchange (pbk_keyx_*) costs per pass before any NVLink is involved.

For G in {2, 4, 8} logical shards it times, on C1-sized read sets resident in HBM (CUDA events inside libpbk, PBK_F_TIMING):
  * Pass A in the exchange layout (G x R buckets, destination-major) against plain Pass A (64 buckets);
  * Pass B in gather mode over a receive buffer assembled from G independent read samples (each contributing the 1/G of its
    k-mers that shard 0 owns -- the per-GPU work of a weak-scaling step) against plain Pass B;
and checks the merged result against an unsharded count of the same reads.  Prints one JSON line per G.

    python scripts/keyx_probe.py [--scale 1.0] [--shards 2,4,8] [--reps 3]
"""
import argparse
import dataclasses
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from devbuf import DevBuf, _cudart          # noqa: E402  (raw device buffers without torch)
from platanus_b_b200 import KmerCounter, synth          # noqa: E402

K = 32


def upload(a: np.ndarray) -> DevBuf:
    import ctypes as C
    a = np.ascontiguousarray(a)
    d = DevBuf(a.nbytes)
    assert _cudart().cudaMemcpy(d.ptr, a.ctypes.data_as(C.c_void_p), a.nbytes, 1) == 0
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--shards", default="2,4,8")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    spec0 = synth.config("C1", scale=args.scale)
    for G in [int(x) for x in args.shards.split(",")]:
        samples = []
        for s in range(G):
            rs = synth.make_reads(dataclasses.replace(spec0, read_seed=spec0.read_seed + 7919 * s))
            b, o = rs.flat()
            samples.append((upload(b), upload(o.astype(np.uint64)), rs.n_reads, int(b.shape[0])))
        n_reads, n_bases = samples[0][2], samples[0][3]
        max_w = max(nb - nr * (K - 1) for _, _, nr, nb in samples)
        out = {"G": G, "k": K, "reads_per_sample": n_reads, "bases_per_sample": n_bases}
        # plain single-GPU passes on sample 0, for reference
        with KmerCounter(K, timing=True) as kc:
            for _ in range(args.reps + 1):
                kc.reset()
                s0 = kc.stats()
                kc.push_reads_device(samples[0][0].ptr, samples[0][1].ptr, n_reads, n_bases)
                kc.finalize_light()
                s1 = kc.stats()
            out["plain"] = {"ms_partition": s1["ms_partition"] - s0["ms_partition"], "ms_insert": s1["ms_insert"] - s0["ms_insert"],
                            "instances": kc.n_instances, "distinct": kc.n_distinct}
        # exchange layout: every sample partitioned by "its" rank's context; shard 0 inserts what it would receive
        ctxs = [KmerCounter(K, n_shards=G, shard_rank=r, timing=True) for r in range(G)]
        try:
            lay = ctxs[0].keyx_plan(max_w)
            for kc in ctxs[1:]:
                kc.keyx_plan(max_w)
            bpd, cpd = int(lay.bytes_per_dest), int(lay.cursors_per_dest) * 8
            sends = [DevBuf(G * bpd) for _ in range(G)]
            curs = [DevBuf(G * cpd) for _ in range(G)]
            recv, rcur = DevBuf(G * bpd), DevBuf(G * cpd)
            out["layout"] = {"n_regions": int(lay.n_regions), "seg_cap": int(lay.seg_cap), "send_MB": G * bpd / 1e6}
            part_ms, ins_ms = [], []
            for rep in range(args.reps + 1):
                for r, kc in enumerate(ctxs):
                    kc.reset()
                    s0 = kc.stats()
                    kc.keyx_partition_device(samples[r][0].ptr, samples[r][1].ptr, samples[r][2], samples[r][3], sends[r].ptr, curs[r].ptr)
                    if r == 0 and rep > 0:
                        part_ms.append(kc.stats()["ms_partition"] - s0["ms_partition"])
                for src in range(G):
                    recv.copy_from(sends[src], src * bpd, 0, bpd)          # destination 0's split of every source
                    rcur.copy_from(curs[src], src * cpd, 0, cpd)
                s0 = ctxs[0].stats()
                ctxs[0].keyx_insert_device(recv.ptr, rcur.ptr)
                if rep > 0:
                    ins_ms.append(ctxs[0].stats()["ms_insert"] - s0["ms_insert"])
            ctxs[0].finalize_light()
            received = int(np.minimum(rcur.to_host(), lay.seg_cap).sum())
            out["keyx"] = {"ms_partition": float(np.median(part_ms)), "ms_insert_gathered": float(np.median(ins_ms)),
                           "keys_received_by_shard0": received, "distinct_shard0": int(ctxs[0].n_distinct),
                           "staged_records": int(sum(int(kc.shard_send_counts(G).sum()) for kc in ctxs))}
        finally:
            for kc in ctxs:
                kc.close()
            for d in sends + curs + [recv, rcur] + [x for smp in samples for x in smp[:2]]:
                d.free()
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
