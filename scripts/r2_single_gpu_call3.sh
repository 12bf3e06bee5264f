#!/bin/bash
# single B200: whole GPU suite at HEAD + full-size C3 / C5 bench lines (BASELINE configs 3 and 5)
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=10 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/${TAG}_gpu_tests.log
tail -18 gpurun_out/${TAG}_gpu_tests.log
for wl in C3 C5; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu-baseline --no-packed > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/${TAG}_bench_$wl.err
done
python - <<PY
import json
for wl in ("C3", "C5"):
    try:
        l = json.loads(open("gpurun_out/${TAG}_bench_%s.json" % wl).read().strip().splitlines()[-1])
        print(wl, {k: l.get(k) for k in ("value", "ms_per_step", "kernel_ms_per_step", "verified", "table")}, "e2e", l["e2e"]["value"], l["config"]["instances_per_step"])
    except Exception as e:
        print(wl, "no line:", e)
PY
