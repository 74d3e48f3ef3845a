#!/bin/bash
# round 2, pass C: radix-8 FFT passes + memory planning: full GPU suite, c4 bench + live trace, config 5 on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_pytest.log
CMF_TRACE=1 CMF_TRACE_SKIP=200 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-calibrated > gpurun_out/r2c_trace.json 2> gpurun_out/r2c_trace.err; echo "trace rc=$?"
grep CMF_TRACE gpurun_out/r2c_trace.err | sort | head -60
python -c "import json; d=json.load(open('gpurun_out/r2c_trace.json')); print(d['value'], d['ms_per_step'], d['value_direct_loss'], d['roofline']['kernel_ms'], d['loss'])"
timeout 900 python bench.py --config c5 --alg hals --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2c_c5_1gpu.json 2> gpurun_out/r2c_c5_1gpu.err; echo "c5 rc=$?"; tail -3 gpurun_out/r2c_c5_1gpu.err
python -c "import json; d=json.load(open('gpurun_out/r2c_c5_1gpu.json')); print(d['value'], d['ms_per_step'], d['critical_path'], d['loss'], d['roofline']['moved']['note'])"
