#!/bin/bash
mkdir -p gpurun_out
for v in nooverlap overlap nooverlap overlap; do
  if [ $v = nooverlap ]; then export PBK_NO_OVERLAP=1; else unset PBK_NO_OVERLAP; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-packed > gpurun_out/_ab.json 2> gpurun_out/_ab.err || tail -3 gpurun_out/_ab.err
  python - "$v" <<'PY'
import json, sys
l = json.loads(open("gpurun_out/_ab.json").read().strip().splitlines()[-1])
print(sys.argv[1], "value ms", round(l["ms_per_step"], 3), "e2e ms", round(l["e2e"]["ms_per_step"], 3))
PY
done
unset PBK_NO_OVERLAP
timeout 600 python scripts/time_consumers.py > gpurun_out/r2g_consumers_k32.json 2> gpurun_out/r2g_consumers.err; echo "consumers rc=$?"; tail -2 gpurun_out/r2g_consumers.err; cat gpurun_out/r2g_consumers_k32.json
