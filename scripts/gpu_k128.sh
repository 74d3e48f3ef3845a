#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_fd.py -x -q -s > gpurun_out/pytest_gpu_k128.log 2>&1
echo "pytest fd exit $?"; grep -E "K=80|passed|failed|Error|error|assert" gpurun_out/pytest_gpu_k128.log | tail -12
timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/k128_c4.json 2> gpurun_out/k128_c4.err
python -c "import json; d=json.load(open('gpurun_out/k128_c4.json')); print(d['value'], d['ms_per_step'], d['value_direct_loss'], d['loss'])"
timeout 120 python scripts/hals_scale.py --N 512 --T 1048576 --K 128 --L 32 --iters 2 2>&1 | tail -3
