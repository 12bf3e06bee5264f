#!/bin/bash
# k=75 / k=42 with compile-time variants of the wide Pass B (keys per thread and round x resident CTAs per SM), built into platanus_b_b200/_lib_<tag>/
mkdir -p gpurun_out
: > gpurun_out/r2j_wide_occupancy.jsonl
for t in "" k1c4 k2c4 k1c5; do for k in 75 42; do
  PBK_LIB_TAG=$t timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-packed --k $k > gpurun_out/_c.json 2> gpurun_out/_c.err || { echo "$t $k FAILED"; tail -2 gpurun_out/_c.err; continue; }
  python - "${t:-default_k2c3}" "$k" <<'PY' | tee -a gpurun_out/r2j_wide_occupancy.jsonl
import json, sys
l = json.loads(open("gpurun_out/_c.json").read().strip().splitlines()[-1])
print(json.dumps({"variant": sys.argv[1], "k": int(sys.argv[2]), "G_kmers_s": round(l["value"] / 1e9, 2), "ms_per_step": round(l["ms_per_step"], 3),
                  "insert_ms": round(l["kernel_ms_per_step"]["count_insert_pass"], 3), "partition_ms": round(l["kernel_ms_per_step"]["count_partition_pass"], 3)}))
PY
done; done
