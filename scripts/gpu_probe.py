#!/usr/bin/env python
"""Random read-modify-write microbenchmark (R_atomic) -> gpurun_out/atomics.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from platanus_b_b200 import microbench_atomics  # noqa: E402

out = []
MB = 1 << 20
SIZES = [int(x) for x in os.environ.get("PROBE_MB", "16,32,64,128,256,1024,4096,16384").split(",")]
MODES = [int(x) for x in os.environ.get("PROBE_MODES", "0,1,2").split(",")]
for mb in SIZES:
    for mode in MODES:
        slots = (mb * MB) // 16
        n_ops = slots // 2 if mode == 2 else 1 << 28
        try:
            r = microbench_atomics(mb * MB, n_ops, mode)
        except Exception as e:  # noqa: BLE001
            r = None
            print("failed", mb, mode, e, flush=True)
        labels = ["red", "ld+red", "cas-insert", "atom64-ret(8B slot)", "ld64+red64(8B slot)", "atom64-ret skewed 7/8 on 1/28", "stream key + atom64", "stream key + atom64 + 1/8 follow-up", "same, 8 keys per thread batched", "sweep 128 regions, no prefetch", "sweep + prefetch.L2 2 regions ahead", "sweep + ld.cg 2 regions ahead"]
        if mode >= 100:
            label = f"sweep {1 << ((mode - 100) >> 2)} regions (shift/mask), " + ["atom64-ret", "atom64-ret + prefetch.L2 next region", "stream key + atom64-ret", "stream key + atom64-ret + 1/8 red"][(mode - 100) & 3]
        else:
            label = labels[mode]
        rec = {"table_mb": mb, "mode": label, "n_ops": n_ops, "gops": None if r is None else r / 1e9}
        print(json.dumps(rec), flush=True)
        out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.environ.get("PROBE_OUT", "gpurun_out/atomics.json"), "w"), indent=1)
