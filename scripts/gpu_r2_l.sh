#!/bin/bash
# round 2, pass L: HALS parity after the tail-table / prepare changes, config 5 on one GPU with the live trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_fd.py -q -k "hals or HALS or exact" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2l_pytest.log
CMF_TRACE=1 CMF_TRACE_SKIP=40 timeout 600 python bench.py --config c5 --alg hals --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2l_c5.json 2> gpurun_out/r2l_c5.err; echo "c5 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2l_c5.json')); print('c5', d['value'], d['ms_per_step'], d['critical_path']['ms_per_sweep'], d['loss'])"
grep "CMF_TRACE " gpurun_out/r2l_c5.err | awk '{print $2, $3, $4, $6, $9, $11}' | sort -k4 -n -r | head -16
