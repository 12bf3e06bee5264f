#!/bin/bash
# single B200: whole GPU suite at HEAD, smoke, default bench (k=32, with the cpu_baseline leg) and k=75
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider --durations=8 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/${TAG}_gpu_tests.log
tail -14 gpurun_out/${TAG}_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --k 75 > gpurun_out/${TAG}_bench_k75.json 2> gpurun_out/${TAG}_bench_k75.err; echo "bench k75 rc=$?"; tail -2 gpurun_out/${TAG}_bench_k75.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json", "gpurun_out/${TAG}_bench_k75.json"):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: l.get(k) for k in ("value", "ms_per_step", "kernel_ms_per_step", "cpu_baseline")}, "e2e", l["e2e"]["value"], l["e2e"]["ms_per_step"], "packed", l.get("e2e_packed2"), l["roofline"]["frac"], l["roofline"]["frac_of_step"])
    except Exception as e:
        print(f, "no line:", e)
PY
