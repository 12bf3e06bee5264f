#!/bin/bash
# single B200, the last seconds: table consumers (seeded entries / loaded entries pass through the new range filter) and the CLI's two seams at HEAD
mkdir -p gpurun_out
timeout 55 python -m pytest tests/test_zz_lookup_gpu.py tests/test_cli_gpu.py -m gpu -q -p no:cacheprovider -k "test_zz_lookup_gpu or 32-False-True or 32-False-False" > gpurun_out/r2n_gpu_tests_head_subset2.log 2>&1; echo "subset2 rc=$?" | tee -a gpurun_out/r2n_gpu_tests_head_subset2.log
tail -3 gpurun_out/r2n_gpu_tests_head_subset2.log
