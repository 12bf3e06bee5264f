#!/usr/bin/env python
"""Two steps of the multi-word-key path (k from PROF_K, default 42) on a half-size C1, for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from platanus_b_b200 import KmerCounter, synth
k = int(os.environ.get("PROF_K", "42"))
rs = synth.make_reads(synth.config("C1", scale=0.5))
b, o = rs.flat()
db = torch.from_numpy(b.copy()).cuda(); do = torch.from_numpy(o.astype(np.int64)).cuda()
torch.cuda.synchronize()
kc = KmerCounter(k, timing=True)
for _ in range(2):
    kc.reset(); kc.push_reads_device(db.data_ptr(), do.data_ptr(), len(o) - 1, len(b)); kc.finalize_light()
s = kc.stats()
print({x: s[x] for x in ("ms_partition", "ms_insert", "n_instances", "n_distinct", "table_bytes")})
