#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_fd8.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_fd8.log
tail -3 gpurun_out/pytest_gpu_fd8.log
timeout 300 python scripts/hals_scale.py --N 512 --T 4194304 --K 128 --L 32 --iters 2 > gpurun_out/fd8_hals_c5_T4M.log 2>&1; tail -3 gpurun_out/fd8_hals_c5_T4M.log
timeout 300 python scripts/hals_scale.py --N 512 --T 1048576 --K 64 --L 32 --iters 2 > gpurun_out/fd8_hals_K64_T1M.log 2>&1; tail -3 gpurun_out/fd8_hals_K64_T1M.log
timeout 900 python bench.py > gpurun_out/bench_c4_fd8.json 2> gpurun_out/bench_c4_fd8.err; echo "bench exit $?"; tail -2 gpurun_out/bench_c4_fd8.err
python -c "import json; d=json.load(open('gpurun_out/bench_c4_fd8.json')); print(d['value'], d['ms_per_step'], d['value_direct_loss'], d['e2e'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], d['cpu_baseline'])"
