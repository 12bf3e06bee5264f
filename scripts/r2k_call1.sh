#!/bin/bash
# single B200: the second form of Pass B (PBK_PASSB2=1: split_kernel + region_build_kernel) -- checked bench line first, A/B against the
# first form and two occupancy variants, ncu launch list + full capture of the two new kernels, then the whole GPU suite through it
mkdir -p gpurun_out
T=r2k
export PBK_PASSB2=1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_k32_split_build.json 2> gpurun_out/${T}_bench_k32_split_build.err; echo "bench split_build rc=$?"; tail -3 gpurun_out/${T}_bench_k32_split_build.err
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-packed > gpurun_out/_v.json 2> gpurun_out/_v.err || { echo "$label FAILED"; tail -3 gpurun_out/_v.err; return; }
  python - "$label" <<'PY' | tee -a gpurun_out/r2k_variants.jsonl
import json, sys
l = json.loads(open("gpurun_out/_v.json").read().strip().splitlines()[-1])
print(json.dumps({"variant": sys.argv[1], "G_kmers_s": round(l["value"] / 1e9, 2), "ms_per_step": round(l["ms_per_step"], 3),
                  "kernel_ms": {a: round(b, 3) for a, b in l["kernel_ms_per_step"].items()}, "e2e_G": round(l["e2e"]["value"] / 1e9, 2), "e2e_ms": round(l["e2e"]["ms_per_step"], 3),
                  "frac": round(l["roofline"]["frac"], 3), "frac_of_step": round(l["roofline"]["frac_of_step"], 3), "verified": l["verified"]["instances"]}))
PY
}
: > gpurun_out/r2k_variants.jsonl
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r2k_bench_k32_split_build.json").read().strip().splitlines()[-1])
    print("split_build", {k: l.get(k) for k in ("value", "ms_per_step", "kernel_ms_per_step", "verified")}, "e2e", l["e2e"]["value"], l["e2e"]["ms_per_step"], "packed", l["e2e_packed2"]["value"], l["roofline"]["frac"], l["roofline"]["frac_of_step"], l["roofline"]["kernel"][:60])
except Exception as e:
    print("no line:", e)
PY
run first_form PBK_PASSB2=0
run split_build PBK_PASSB2=1
run build_2ctas PBK_PASSB2=1 PBK_BUILD_CTAS=2
run split_1cta PBK_PASSB2=1 PBK_SPLIT_CTAS=1
run table_always_read PBK_PASSB2=1 PBK_PASSB2_FRESH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-packed > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'split_kernel|region_build_kernel' -s 4 -c 2 \
    -o gpurun_out/${T}_split_build -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-packed > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 700 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/${T}_gpu_tests_split_build.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/${T}_gpu_tests_split_build.log
tail -16 gpurun_out/${T}_gpu_tests_split_build.log
