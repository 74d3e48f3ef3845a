#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fd.py -x -q > gpurun_out/pytest_gpu_fd10.log 2>&1
echo "pytest fd exit $?"; tail -15 gpurun_out/pytest_gpu_fd10.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/fd10_c4.json 2> gpurun_out/fd10_c4.err; tail -2 gpurun_out/fd10_c4.err
python -c "import json; d=json.load(open('gpurun_out/fd10_c4.json')); print(d['value'], d['ms_per_step'], d['value_direct_loss'], d['loss'], d['roofline']['kernel_ms'])"
