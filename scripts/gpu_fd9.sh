#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_fd9.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_fd9.log
tail -3 gpurun_out/pytest_gpu_fd9.log
timeout 300 python scripts/hals_scale.py --N 512 --T 4194304 --K 128 --L 32 --iters 2 > gpurun_out/fd9_hals_c5_T4M.log 2>&1; tail -3 gpurun_out/fd9_hals_c5_T4M.log
timeout 300 python scripts/hals_scale.py --N 512 --T 1048576 --K 64 --L 32 --iters 2 > gpurun_out/fd9_hals_K64_T1M.log 2>&1; tail -3 gpurun_out/fd9_hals_K64_T1M.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/fd9_c4.json 2> gpurun_out/fd9_c4.err
python -c "import json; d=json.load(open('gpurun_out/fd9_c4.json')); print(d['value'], d['ms_per_step'], d['loss'], d['roofline']['kernel_ms'])"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_fd9.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_fd9.log
