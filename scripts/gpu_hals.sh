#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_fd.py -x -q -k "hals or HALS" > gpurun_out/pytest_gpu_hals.log 2>&1
echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_hals.log
timeout 300 python scripts/hals_scale.py --N 512 --T 4194304 --K 128 --L 32 --iters 2 > gpurun_out/hals_c5_T4M.log 2>&1; tail -2 gpurun_out/hals_c5_T4M.log
timeout 300 python scripts/hals_scale.py --N 512 --T 1048576 --K 64 --L 32 --iters 2 > gpurun_out/hals_K64_T1M.log 2>&1; tail -2 gpurun_out/hals_K64_T1M.log
