#!/bin/bash
# round 2, pass B: live launch trace of the c4 iteration, ncu --set full of the H-side FFT kernels, numW accuracy at full c4,
# config 5 (HALS, T = 2^24) on one GPU
mkdir -p gpurun_out
CMF_TRACE=1 CMF_TRACE_SKIP=200 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-calibrated > gpurun_out/r2b_trace.json 2> gpurun_out/r2b_trace.err; echo "trace rc=$?"
grep CMF_TRACE gpurun_out/r2b_trace.err | sort -t'l' -k3 | head -60
timeout 900 python scripts/numw_accuracy.py > gpurun_out/r2b_numw_accuracy.log 2>&1; echo "numw rc=$?"; tail -5 gpurun_out/r2b_numw_accuracy.log
FCMD="python bench.py --T 1048576 --steps 1 --warmup 1 --no-e2e --no-cpu --no-calibrated"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fft_h_kernel|ifft_numH_kernel|mu_update_kernel|fft_x_kernel' -s 6 -c 8 -o gpurun_out/r2_fft_kernels $FCMD > gpurun_out/r2b_ncu_fft.log 2>&1; echo "ncu fft rc=$?"; tail -2 gpurun_out/r2b_ncu_fft.log
timeout 900 python bench.py --config c5 --alg hals --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2b_c5_1gpu.json 2> gpurun_out/r2b_c5_1gpu.err; echo "c5 rc=$?"; tail -3 gpurun_out/r2b_c5_1gpu.err
python -c "import json; d=json.load(open('gpurun_out/r2b_c5_1gpu.json')); print(d['value'], d['ms_per_step'], d['critical_path'], d['loss'])"
