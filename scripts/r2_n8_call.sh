#!/bin/bash
# N-GPU gate + scaling measurements of round 2 (gpurun --gpus 8): sharded_check over NCCL on all ranks (every exchange form against the
# unsharded oracle), then bench.py: default exchange (pull) and records on the C1 replica workload, pull on C4 slices and on the whole C4.
N=${1:-8}
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29511 tests/sharded_check.py > gpurun_out/r2_sharded_check_n$N.log 2>&1; echo "sharded_check rc=$?" | tee -a gpurun_out/r2_sharded_check_n$N.log
grep -E "^ok|rc=" gpurun_out/r2_sharded_check_n$N.log
run() {   # name, args...
  name=$1; shift
  timeout 600 $TR --master-port $((29520 + RANDOM % 200)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline "$@" \
      > gpurun_out/r2_bench_n${N}_${name}.json 2> gpurun_out/r2_bench_n${N}_${name}.err
  echo "bench N=$N $name rc=$?"; grep -v ProcessGroupNCCL gpurun_out/r2_bench_n${N}_${name}.err | tail -3
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/r2_bench_n${N}_${name}.json").read().strip().splitlines()[-1])
    print({k: l.get(k) for k in ("value", "ms_per_step", "exchange_bytes_sent_per_gpu_per_step", "verified", "kernel_ms_per_step", "exchange_note")}, l["e2e"])
except Exception as e:
    print("no line:", e)
PY
}
run C1_pull --workload C1
run C1_records --workload C1 --exchange records
run C4_pull --workload C4
[ "$N" -ge 4 ] && run C4full_pull --workload C4full --no-verify-recount
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
