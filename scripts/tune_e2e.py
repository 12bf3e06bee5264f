#!/usr/bin/env python
"""e2e (host buffers) step time on C1 for different numbers of chained sub-batches (PBK_N_SB)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from platanus_b_b200 import KmerCounter, synth
rs = synth.make_reads(synth.config("C1"))
b, o = rs.flat()
hb = torch.from_numpy(b.copy()).pin_memory(); ho = torch.from_numpy(o.astype(np.int64)).pin_memory()
kc = KmerCounter(32, timing=True)
for nsb in [int(x) for x in os.environ.get("SB_LIST", "1,2,4,7").split(",")]:
    os.environ["PBK_N_SB"] = str(nsb)
    for _ in range(3):
        kc.reset(); kc.push_reads_ptr(hb.data_ptr(), ho.data_ptr(), len(o) - 1); kc.finalize_light()
    kc.timer_mark(0)
    for _ in range(5):
        kc.reset(); kc.push_reads_ptr(hb.data_ptr(), ho.data_ptr(), len(o) - 1); kc.finalize_light()
    kc.timer_mark(1)
    print(json.dumps({"n_sb": nsb, "ms_step": kc.timer_elapsed_ms(0, 1) / 5}), flush=True)
