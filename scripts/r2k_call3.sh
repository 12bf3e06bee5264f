#!/bin/bash
# single B200, last call of the round: the reworked second form of Pass B (32-bit slot arithmetic + key prefetch in region_build_kernel, flat
# copy-out in split_kernel) against the first form, its targeted tests, then the default bench line with the CPU baseline and the k = 75 line
mkdir -p gpurun_out
T=r2m
run() {  # label, bench args, env...
  label=$1; shift; extra=$1; shift
  env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-packed $extra > gpurun_out/_v.json 2> gpurun_out/_v.err || { echo "$label FAILED"; tail -3 gpurun_out/_v.err; return; }
  cp gpurun_out/_v.json gpurun_out/${T}_bench_$label.json
  python - "$label" <<'PY' | tee -a gpurun_out/r2m_variants.jsonl
import json, sys
l = json.loads(open("gpurun_out/_v.json").read().strip().splitlines()[-1])
print(json.dumps({"variant": sys.argv[1], "G_kmers_s": round(l["value"] / 1e9, 2), "ms_per_step": round(l["ms_per_step"], 3),
                  "kernel_ms": {a: round(b, 3) for a, b in l["kernel_ms_per_step"].items()}, "e2e_G": round(l["e2e"]["value"] / 1e9, 2), "e2e_ms": round(l["e2e"]["ms_per_step"], 3),
                  "frac": round(l["roofline"]["frac"], 3), "frac_of_step": round(l["roofline"]["frac_of_step"], 3), "verified": l["verified"]["instances"]}))
PY
}
: > gpurun_out/r2m_variants.jsonl
run split_build_k32 "" PBK_PASSB2=1
SEL="forced_partition or direct_and_partitioned or pipelined or full_size_c1 or c1_full_size_properties or large_pushes or device_resident or packed or C5-0.05 or logical_shards or key_exchange or pull or group"
PBK_PASSB2=1 PBK_PASSB2_GATHER=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_zz_keyx_gpu.py tests/test_zz_group_gpu.py -m gpu -q -p no:cacheprovider -k "($SEL) and not 75" > gpurun_out/${T}_gpu_tests_second_form.log 2>&1; echo "second form tests rc=$?" | tee -a gpurun_out/${T}_gpu_tests_second_form.log
tail -3 gpurun_out/${T}_gpu_tests_second_form.log
timeout 400 python bench.py --steps 8 --warmup 3 > gpurun_out/${T}_bench_k32.json 2> gpurun_out/${T}_bench_k32.err; echo "default bench rc=$?"; tail -2 gpurun_out/${T}_bench_k32.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r2m_bench_k32.json").read().strip().splitlines()[-1])
    print("default", {k: l.get(k) for k in ("value", "ms_per_step", "kernel_ms_per_step", "verified", "cpu_baseline")}, "e2e", l["e2e"]["value"], l["e2e"]["ms_per_step"], "packed", l["e2e_packed2"]["value"], l["roofline"]["frac"], l["roofline"]["frac_of_step"], l["roofline"]["kernel"][:60])
except Exception as e:
    print("no line:", e)
PY
run split_build_k32_256thr "" PBK_PASSB2=1 PBK_SPLIT_THREADS=256
run first_form_k75 "--k 75" PBK_PASSB2=0
