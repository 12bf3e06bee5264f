#!/usr/bin/env python
"""Pass A / Pass B tuning sweep on the C1 workload (device-resident reads): one JSON line per knob setting.
Knobs are environment variables the launch wrappers read on every call (PBK_REGION_MB, PBK_PF_DIST,
PBK_PASSB_CTAS, PBK_PASSB_HINT).  Output: gpurun_out/tune_passb.jsonl"""
import itertools
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from platanus_b_b200 import KmerCounter, synth  # noqa: E402

spec = synth.config(os.environ.get("TUNE_WORKLOAD", "C1"), scale=float(os.environ.get("TUNE_SCALE", "1")))
rs = synth.make_reads(spec)
b, o = rs.flat()
db = torch.from_numpy(b.copy()).cuda()
do = torch.from_numpy(o.astype(np.int64)).cuda()
torch.cuda.synchronize()
n_reads, n_bases = len(o) - 1, len(b)
grid = json.loads(os.environ.get("TUNE_GRID", '{"PBK_REGION_MB": [4, 8, 16, 32, 64], "PBK_PF_DIST": [0, 1], "PBK_PASSB_HINT": [0, 1], "PBK_PASSB_CTAS": [3], "PBK_PART_SMEM_KB": [100]}'))
names = sorted(grid)
out = open("gpurun_out/tune_passb.jsonl", "a")
K = int(os.environ.get("TUNE_K", "32"))
kc = KmerCounter(K, timing=True)
for combo in itertools.product(*[grid[n] for n in names]):
    for n, v in zip(names, combo):
        os.environ[n] = str(v)
    fin = (lambda: kc.finalize_light()) if int(os.environ.get("PBK_PASSB_HINT", "1")) < 2 else (lambda: None)
    for _ in range(2):
        kc.reset(); kc.push_reads_device(db.data_ptr(), do.data_ptr(), n_reads, n_bases); fin()
    s0 = kc.stats()
    kc.timer_mark(0)
    steps = 3
    for _ in range(steps):
        kc.reset(); kc.push_reads_device(db.data_ptr(), do.data_ptr(), n_reads, n_bases); fin()
    kc.timer_mark(1)
    ms = kc.timer_elapsed_ms(0, 1) / steps
    s1 = kc.stats()
    rec = dict(zip(names, combo))
    rec["k"] = K
    rec.update(ms_step=ms, ms_partition=(s1["ms_partition"] - s0["ms_partition"]) / steps,
               ms_insert=(s1["ms_insert"] - s0["ms_insert"]) / steps, gkmers=kc.stats()["n_instances"] / ms / 1e6)
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n"); out.flush()
