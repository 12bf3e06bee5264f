#!/bin/bash
# Tuning run on one B200: environment switches and compile-time variants (built by `PBK_LIB_TAG=.. PBK_EXTRA_CFLAGS=.. python -m
# platanus_b_b200.build` into platanus_b_b200/_lib_<tag>/) of the k=32 and k=75 steps; one JSON line per variant.
#   gpurun --timeout 1200 -- 'bash scripts/tune_variants.sh'
mkdir -p gpurun_out
OUT=gpurun_out/r2_tune_variants.jsonl
: > $OUT
run() {  # label, k, env...
  label=$1; k=$2; shift 2
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --k $k > gpurun_out/_tune.json 2> gpurun_out/_tune.err || { echo "$label FAILED"; tail -3 gpurun_out/_tune.err; return; }
  python - "$label" "$k" <<'PY' | tee -a gpurun_out/r2_tune_variants.jsonl
import json, sys
l = json.loads(open("gpurun_out/_tune.json").read().strip().splitlines()[-1])
print(json.dumps({"variant": sys.argv[1], "k": int(sys.argv[2]), "G_kmers_s": round(l["value"] / 1e9, 2), "ms_per_step": round(l["ms_per_step"], 3),
                  "kernel_ms": {a: round(b, 3) for a, b in l["kernel_ms_per_step"].items()}, "e2e_G": round(l["e2e"]["value"] / 1e9, 2),
                  "frac": round(l["roofline"]["frac"], 3)}))
PY
}
run k32_staged 32 PBK_PASSB_STAGED=1
run k32_unstaged 32 PBK_PASSB_STAGED=0
run k75_staged 75 PBK_WIDE_STAGED=1
run k75_unstaged 75 PBK_WIDE_STAGED=0
run k75_load066 75 PBK_MAX_LOAD_WIDE=0.66
run k75_win4 75 PBK_LIB_TAG=win4
run k75_kpt2 75 PBK_LIB_TAG=kpt2
run k75_win4kpt2 75 PBK_LIB_TAG=win4kpt2
run k42_staged 42 PBK_WIDE_STAGED=1
run k42_kpt2 42 PBK_LIB_TAG=kpt2
