#!/bin/bash
# single B200: A/B of the side-stream overlap, consumer timings, fresh ncu captures of the round-2 kernels
mkdir -p gpurun_out
for v in overlap nooverlap; do
  if [ $v = nooverlap ]; then export PBK_NO_OVERLAP=1; else unset PBK_NO_OVERLAP; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_$v.json 2> gpurun_out/r2g_bench_$v.err; echo "bench $v rc=$?"; tail -2 gpurun_out/r2g_bench_$v.err
done
unset PBK_NO_OVERLAP
python - <<PY
import json
for v in ("overlap", "nooverlap"):
    try:
        l = json.loads(open("gpurun_out/r2g_bench_%s.json" % v).read().strip().splitlines()[-1])
        print(v, {k: l.get(k) for k in ("value", "ms_per_step", "kernel_ms_per_step", "verified")}, "e2e", l["e2e"]["value"], "packed", l["e2e_packed2"]["value"], l["roofline"]["frac"], l["roofline"]["frac_of_step"])
    except Exception as e:
        print(v, "no line:", e)
PY
timeout 600 python scripts/time_consumers.py > gpurun_out/r2g_consumers_k32.json 2> gpurun_out/r2g_consumers.err; echo "consumers rc=$?"; tail -2 gpurun_out/r2g_consumers.err; cat gpurun_out/r2g_consumers_k32.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-packed > gpurun_out/r2g_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bucket_insert_compact_kernel|histogram_kernel|export_kernel' -s 1 -c 3 \
    -o gpurun_out/r2g_k32 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-packed > gpurun_out/r2g_ncu_k32.log 2>&1; echo "ncu k32 rc=$?"
