#!/bin/bash
mkdir -p gpurun_out
for k in 75 42 32; do
  timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --k $k > gpurun_out/_c.json 2> gpurun_out/_c.err || tail -3 gpurun_out/_c.err
  python - "$k" <<'PY' | tee -a gpurun_out/r2i_chained_wide.jsonl
import json, sys
l = json.loads(open("gpurun_out/_c.json").read().strip().splitlines()[-1])
print(json.dumps({"k": int(sys.argv[1]), "G_kmers_s": round(l["value"] / 1e9, 2), "ms_per_step": round(l["ms_per_step"], 3),
                  "kernel_ms": {a: round(b, 3) for a, b in l["kernel_ms_per_step"].items()}, "frac": round(l["roofline"]["frac"], 3), "frac_of_step": round(l["roofline"]["frac_of_step"], 3),
                  "e2e_G": round(l["e2e"]["value"] / 1e9, 2), "e2e_ms": round(l["e2e"]["ms_per_step"], 3), "packed_G": round((l.get("e2e_packed2") or {"value": 0})["value"] / 1e9, 2),
                  "verified": l["verified"]["instances"]}))
PY
  cp gpurun_out/_c.json gpurun_out/r2i_bench_k$k.json
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cli_gpu.py -m gpu -q -p no:cacheprovider -k "full_size_c1 or sweep or golden or pipelined or matches_the_reference_program or large_pushes" > gpurun_out/r2i_gpu_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r2i_gpu_tests.log
tail -4 gpurun_out/r2i_gpu_tests.log
