#!/bin/bash
# Round-2 multi-GPU gate + measurements, one gpurun call:  gpurun --gpus N --timeout 1500 -- 'bash scripts/r2_multi_gpu_call.sh N'
#   1. tests/sharded_check.py over NCCL on N real GPUs: every exchange form against the unsharded oracle
#   2. bench.py --gpus N for each exchange form (each run ends with the checked step + single-GPU recount, bench.py verify_result)
N=${1:-2}
FORMS=${2:-"records keys keys_async"}
WORKLOADS=${3:-"C1"}
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/sharded_check.py > gpurun_out/r2_sharded_check_n$N.log 2>&1; echo "sharded_check rc=$?" | tee -a gpurun_out/r2_sharded_check_n$N.log
grep -E "^ok|rc=" gpurun_out/r2_sharded_check_n$N.log
port=29520
for wl in $WORKLOADS; do
for form in $FORMS; do
  case $form in
    records)    extra="--exchange records" ;;
    keys)       extra="--exchange keys --keyx-chunks 4" ;;
    keys_async) extra="--exchange keys --keyx-async --keyx-chunks 4" ;;
    keys_async2) extra="--exchange keys --keyx-async --keyx-chunks 2" ;;
    keys_async8) extra="--exchange keys --keyx-async --keyx-chunks 8" ;;
    pull)       extra="--exchange pull" ;;
  esac
  port=$((port+1))
  timeout 600 $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-cpu-baseline $extra \
      > gpurun_out/r2_bench_n${N}_${wl}_${form}.json 2> gpurun_out/r2_bench_n${N}_${wl}_${form}.err
  echo "bench N=$N $wl $form rc=$?"; tail -c 600 gpurun_out/r2_bench_n${N}_${wl}_${form}.err | tail -3
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2_bench_n${N}_${wl}_${form}.json").read().strip().splitlines()[-1])
    print({k: l.get(k) for k in ("value", "ms_per_step", "exchange_bytes_sent_per_gpu_per_step", "verified", "kernel_ms_per_step")}, l["e2e"])
except Exception as e:
    print("no line:", e)
PY
done
done
