#!/bin/bash
# round 2, pass N: W-side transforms with 8 complex columns per CTA at B = 1024 -- parity (default and forced B = 1024) and the c4 iteration
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_fd.py tests/test_gpu_scale.py tests/test_gpu_bench_shape.py -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2n_pytest.log
CMF_FD_B=1024 timeout 80 python -m pytest tests/test_gpu_fd.py -x -q > gpurun_out/r2n_pytest_b1024.log 2>&1; echo "pytest B=1024 rc=$?"; tail -2 gpurun_out/r2n_pytest_b1024.log
timeout 120 python bench.py --steps 6 --warmup 2 --no-e2e --no-cpu --no-calibrated --no-direct > gpurun_out/r2n_c4.json 2> gpurun_out/r2n_c4.err; echo "c4 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2n_c4.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['loss'])"
