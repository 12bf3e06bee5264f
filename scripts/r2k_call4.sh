#!/bin/bash
# single B200, the last 2 GPU-minutes of the round: the hash-range-pass tests on real hardware
mkdir -p gpurun_out
timeout 110 python -m pytest "tests/test_gpu_parity.py::test_hash_range_passes_add_up_to_the_whole_count" "tests/test_cli_gpu.py::test_pbk_assemble_counts_in_hash_range_passes_when_the_table_does_not_fit" -m gpu -q -p no:cacheprovider > gpurun_out/r2n_gpu_tests_passes.log 2>&1; echo "passes tests rc=$?" | tee -a gpurun_out/r2n_gpu_tests_passes.log
tail -5 gpurun_out/r2n_gpu_tests_passes.log
