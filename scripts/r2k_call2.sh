#!/bin/bash
# single B200, HEAD: Pass B's second form (split_kernel + region_build_kernel with per-warp retry lists) is the default for a context's own
# bucket store.  Bench line first (its result is checked), A/B against the first form, ncu launch list + full capture of the two kernels,
# then every GPU test that takes the partitioned route for one-word keys through the second form, and a subset through the first.
mkdir -p gpurun_out
T=r2l
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${T}_smoke.log
timeout 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench_k32.json 2> gpurun_out/${T}_bench_k32.err; echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench_k32.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r2l_bench_k32.json").read().strip().splitlines()[-1])
    print("default", {k: l.get(k) for k in ("value", "ms_per_step", "kernel_ms_per_step", "verified")}, "e2e", l["e2e"]["value"], l["e2e"]["ms_per_step"], "packed", l["e2e_packed2"]["value"], l["roofline"]["frac"], l["roofline"]["frac_of_step"], l["roofline"]["kernel"][:60])
except Exception as e:
    print("no line:", e)
PY
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-packed > gpurun_out/_v.json 2> gpurun_out/_v.err || { echo "$label FAILED"; tail -3 gpurun_out/_v.err; return; }
  python - "$label" <<'PY' | tee -a gpurun_out/r2l_variants.jsonl
import json, sys
l = json.loads(open("gpurun_out/_v.json").read().strip().splitlines()[-1])
print(json.dumps({"variant": sys.argv[1], "G_kmers_s": round(l["value"] / 1e9, 2), "ms_per_step": round(l["ms_per_step"], 3),
                  "kernel_ms": {a: round(b, 3) for a, b in l["kernel_ms_per_step"].items()}, "e2e_G": round(l["e2e"]["value"] / 1e9, 2), "e2e_ms": round(l["e2e"]["ms_per_step"], 3),
                  "frac": round(l["roofline"]["frac"], 3), "frac_of_step": round(l["roofline"]["frac_of_step"], 3), "verified": l["verified"]["instances"]}))
PY
}
: > gpurun_out/r2l_variants.jsonl
run first_form PBK_PASSB2=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-packed > gpurun_out/${T}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'split_kernel|region_build_kernel' -s 4 -c 2 \
    -o gpurun_out/${T}_split_build -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-packed > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
SEL="forced_partition or direct_and_partitioned or pipelined or full_size_c1 or c1_full_size_properties or large_pushes or device_resident or packed or C5-0.05 or logical_shards or key_exchange or pull or group"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_keyx_gpu.py tests/test_zz_group_gpu.py -m gpu -q -p no:cacheprovider -k "($SEL) and not 75" > gpurun_out/${T}_gpu_tests_second_form.log 2>&1; echo "second form tests rc=$?" | tee -a gpurun_out/${T}_gpu_tests_second_form.log
tail -4 gpurun_out/${T}_gpu_tests_second_form.log
PBK_PASSB2_GATHER=1 timeout 300 python -m pytest tests/test_zz_keyx_gpu.py tests/test_zz_group_gpu.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "(key_exchange or pull or group or logical_shards) and not 75" > gpurun_out/${T}_gpu_tests_gather_second_form.log 2>&1; echo "gather second form tests rc=$?" | tee -a gpurun_out/${T}_gpu_tests_gather_second_form.log
tail -3 gpurun_out/${T}_gpu_tests_gather_second_form.log
PBK_PASSB2=0 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "(forced_partition or direct_and_partitioned or pipelined or c1_full_size_properties or large_pushes) and not 75" > gpurun_out/${T}_gpu_tests_first_form.log 2>&1; echo "first form tests rc=$?" | tee -a gpurun_out/${T}_gpu_tests_first_form.log
tail -3 gpurun_out/${T}_gpu_tests_first_form.log
