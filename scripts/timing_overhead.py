#!/usr/bin/env python
"""How much the per-launch CUDA events of PBK_F_TIMING cost: C1 step time with and without them."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from platanus_b_b200 import KmerCounter, synth
rs = synth.make_reads(synth.config("C1"))
b, o = rs.flat()
hb = torch.from_numpy(b.copy()).pin_memory(); ho = torch.from_numpy(o.astype(np.int64)).pin_memory()
db, do = hb.cuda(), ho.cuda()
torch.cuda.synchronize()
for timing in (True, False):
    kc = KmerCounter(32, timing=timing)
    for mode in ("resident", "host"):
        def step():
            kc.reset()
            if mode == "resident": kc.push_reads_device(db.data_ptr(), do.data_ptr(), len(o) - 1, len(b))
            else: kc.push_reads_ptr(hb.data_ptr(), ho.data_ptr(), len(o) - 1)
            kc.finalize_light()
        for _ in range(3): step()
        kc.timer_mark(0)
        for _ in range(8): step()
        kc.timer_mark(1)
        print(json.dumps({"timing_events": timing, "mode": mode, "ms_step": kc.timer_elapsed_ms(0, 1) / 8}), flush=True)
    kc.close()
