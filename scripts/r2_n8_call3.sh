#!/bin/bash
# N-GPU timing of the pull exchange after the leaner step (one collective + no counter read-back between the passes)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {   # name, env assignments..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 $TR --master-port $((29520 + RANDOM % 200)) bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline "$@" \
      > gpurun_out/r2f_bench_n${N}_${name}.json 2> gpurun_out/r2f_bench_n${N}_${name}.err
  echo "bench N=$N $name rc=$?"; grep -v "ProcessGroupNCCL\|OMP_NUM\|^\*\*\*" gpurun_out/r2f_bench_n${N}_${name}.err | tail -3
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2f_bench_n${N}_${name}.json").read().strip().splitlines()[-1])
    print({k: l.get(k) for k in ("value", "ms_per_step", "verified", "kernel_ms_per_step", "exchange_note")}, "e2e", l["e2e"]["value"], l["e2e"]["ms_per_step"])
except Exception as e:
    print("no line:", e)
PY
}
run C1_pull PBK_X=0 -- --workload C1
run C1_pull_r16 PBK_KEYX_REGIONS=16 -- --workload C1
run C4full_pull PBK_X=0 -- --workload C4full --no-verify-recount
