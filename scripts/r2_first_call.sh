#!/bin/bash
# Round-2 opener: everything that was built after round 1's GPU budget ran out, on ONE B200, in one gpurun call (~6 min):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/r2_first_call.sh'
# Then, on 2 GPUs (gpurun --gpus 2), the correctness gate of every exchange form over NCCL and the first timings:
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/sharded_check.py
#   ... bench.py --gpus 2 --exchange records | --exchange keys | --exchange keys --keyx-async   (then the same at --gpus 8)
set -u
mkdir -p gpurun_out
# 1. the GPU cases that have only been seen to pass against the host-emulated ABI (xfail marks ignored: failures are real here)
timeout 600 python -m pytest tests/test_zz_keyx_gpu.py tests/test_zz_lookup_gpu.py -m gpu -q --runxfail -p no:cacheprovider \
    > gpurun_out/r2_zz_tests.log 2>&1; echo "zz tests rc=$?" | tee -a gpurun_out/r2_zz_tests.log
# 2. the whole suite as the driver runs it
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_gpu_tests.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/r2_gpu_tests.log
# 3. single-GPU bench (unchanged hot path: must reproduce r1c_bench.json) and the per-pass cost of the key-exchange layout
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 600 python scripts/keyx_probe.py --shards 2,4,8 > gpurun_out/r2_keyx_probe.jsonl 2> gpurun_out/r2_keyx_probe.err; echo "probe rc=$?"
tail -3 gpurun_out/r2_zz_tests.log gpurun_out/r2_gpu_tests.log; cat gpurun_out/r2_keyx_probe.jsonl
