#!/usr/bin/env python
"""Wall-clock timings of the table CONSUMERS on a C1-sized table, through the C ABI with host buffers:
pbk_export (sorted, >= cutoff), pbk_neighbor_flags, pbk_lookup (the genome as one contig; a quarter of the reads), pbk_match_reads (all reads),
seeded counting (pbk_seed_entries + pbk_push_reads + pbk_finalize).  One JSON line; median of 3 after one warm-up call each.
    gpurun -- 'python scripts/time_consumers.py > gpurun_out/consumers.json'"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from platanus_b_b200 import KmerCounter, synth      # noqa: E402

K = int(os.environ.get("CONS_K", "32"))


def med(f, n=3):
    f()
    t = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        t.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(t))


spec = synth.config("C1")
rs = synth.make_reads(spec)
b, o = rs.flat()
genome = np.frombuffer(b"ACGT", dtype=np.uint8)[synth.make_genome(spec.genome_lengths[0], spec.gc[0], spec.genome_seeds[0])]
out = {"workload": f"C1, k={K}", "n_reads": int(rs.n_reads), "n_bases": int(len(b))}
def kernel_ms(kc, f):
    """device time of the call's own kernels (per-launch CUDA events, PBK_F_TIMING): pack + other classes"""
    s0 = kc.stats()
    f()
    s1 = kc.stats()
    return {"pack": round(s1["ms_pack"] - s0["ms_pack"], 3), "kernels": round(s1["ms_other"] - s0["ms_other"], 3)}


with KmerCounter(K, timing=True) as kc:
    kc.set_timing(False)
    kc.push_reads(b, o)
    kc.finalize()
    cutoff = kc.coverage_cutoff()
    keys, counts = kc.export(cutoff, sorted=True)
    out.update(distinct=int(kc.n_distinct), cutoff=int(cutoff), kept=int(len(counts)))
    out["export_sorted_ms"] = med(lambda: kc.export(cutoff, sorted=True))
    flags = kc.neighbor_flags(keys, cutoff)
    out["neighbor_flags_ms"] = med(lambda: kc.neighbor_flags(keys, cutoff))
    out["neighbor_flags_Mkeys_per_s"] = len(counts) / out["neighbor_flags_ms"] / 1e3
    out["interior_nodes_frac"] = float(((flags >> 4 != 0) & (flags & 15 != 0)).mean())
    go = np.array([0, len(genome)], np.uint64)
    occ = kc.lookup(genome, go)
    out["lookup_genome_ms"] = med(lambda: kc.lookup(genome, go))
    out["genome_windows_found_frac"] = float((occ[: len(genome) - K + 1] >= cutoff).mean())
    q = rs.n_reads // 4
    bq, oq = b[: int(o[q])], o[: q + 1]
    out["lookup_quarter_reads_ms"] = med(lambda: kc.lookup(bq, oq))
    out["lookup_quarter_reads_Gwindows_per_s"] = (len(bq) - q * (K - 1)) / out["lookup_quarter_reads_ms"] / 1e6
    m = kc.match_reads(b, o)
    out["match_all_reads_ms"] = med(lambda: kc.match_reads(b, o))
    out["reads_matched_frac"] = float(m.mean())
    # device time of the kernels alone (the wall-clock figures above are dominated by pageable H2D / D2H copies of the arguments)
    kc.set_timing(True)
    out["kernel_ms"] = {"neighbor_flags": kernel_ms(kc, lambda: kc.neighbor_flags(keys, cutoff)),
                        "lookup_quarter_reads": kernel_ms(kc, lambda: kc.lookup(bq, oq)),
                        "match_all_reads": kernel_ms(kc, lambda: kc.match_reads(b, o)),
                        "export_sorted": kernel_ms(kc, lambda: kc.export(cutoff, sorted=True))}
    out["lookup_kernel_Gwindows_per_s"] = (len(bq) - q * (K - 1)) / max(out["kernel_ms"]["lookup_quarter_reads"]["kernels"], 1e-9) / 1e6
    kc.set_timing(False)
with KmerCounter(K) as kc:
    def seeded():
        kc.reset()
        kc.seed_entries(keys, counts)
        kc.push_reads(b, o)
        kc.finalize()
    out["seeded_count_ms"] = med(seeded, 2)
    out["seeded_distinct"] = int(kc.n_distinct)
print(json.dumps(out))
