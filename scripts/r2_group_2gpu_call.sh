#!/bin/bash
# two real B200s in ONE process: the library's own multi-GPU path (pbk_group_*: peer access, P2P loads, cudaMemcpyPeerAsync) and
# pbk_assemble on it, against the oracle / the reference program
mkdir -p gpurun_out
nvidia-smi -L
PBK_TEST_GROUP_DEVICES=0,1 timeout 900 python -m pytest tests/test_zz_group_gpu.py tests/test_cli_gpu.py -m gpu -q -p no:cacheprovider -k "group or several_gpus or device_count" \
    > gpurun_out/r2_group_2gpu_tests.log 2>&1; echo "group tests rc=$?" | tee -a gpurun_out/r2_group_2gpu_tests.log
tail -5 gpurun_out/r2_group_2gpu_tests.log
# process against process on the whole C1: the reference (all host cores), pbk_assemble on one GPU, pbk_assemble on two
python scripts/cli_e2e.py --gpus 1,2 > gpurun_out/r2_cli_e2e_2gpu.json 2> gpurun_out/r2_cli_e2e_2gpu.err; echo "cli_e2e rc=$?"; tail -3 gpurun_out/r2_cli_e2e_2gpu.err; cat gpurun_out/r2_cli_e2e_2gpu.json | tail -3
