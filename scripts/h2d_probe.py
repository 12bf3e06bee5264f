#!/usr/bin/env python
"""Pinned host -> device copy bandwidth of this box (what bounds bench.py's e2e): one big copy vs 32 MiB chunks."""
import json
import time

import torch

n = 484 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for name, chunk in (("one_copy", n), ("32MiB_chunks", 32 << 20), ("128MiB_chunks", 128 << 20)):
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for o in range(0, n, chunk):
            d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    out[name] = {"ms": best, "GBps": n / best / 1e6}
print(json.dumps(out))
