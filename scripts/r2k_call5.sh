#!/bin/bash
# single B200, what is left of the round's GPU time: smoke() and the golden cases at HEAD (count_kernel's direct route changed last)
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2n_smoke.log
timeout 70 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "golden_cases_bit_exact or error_behaviour or table_growth or direct_and_partitioned" > gpurun_out/r2n_gpu_tests_head_subset.log 2>&1; echo "subset rc=$?" | tee -a gpurun_out/r2n_gpu_tests_head_subset.log
tail -3 gpurun_out/r2n_gpu_tests_head_subset.log
