#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_fd.py -x -q -k "hals or HALS" > gpurun_out/pytest_gpu_hals_ovl.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_hals_ovl.log
for f in 0 1; do CMF_HALS_OVERLAP=$f timeout 60 python scripts/hals_scale.py --N 512 --T 1048576 --K 128 --L 32 --iters 2 2>&1 | tail -1; done
