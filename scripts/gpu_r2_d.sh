#!/bin/bash
# round 2, pass D: A/B of the FFT kernels (radix-8 at 3 CTAs/SM, radix-8 at 2 CTAs/SM, round-1 radix-2^2) live inside the c4
# iteration and isolated under ncu; parity of the fused <numH,H'> / vectorised MU update
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fd.py tests/test_gpu_bench_shape.py tests/test_gpu_scale.py -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2d_pytest.log
for v in default minb2 oldfft default; do
  LIB=""; [ $v != default ] && LIB=$PWD/scripts/ab/libcmf_$v.so
  CMF_SM100_LIB=$LIB CMF_TRACE=1 CMF_TRACE_SKIP=150 timeout 600 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu --no-calibrated --no-direct > gpurun_out/r2d_$v.json 2> gpurun_out/r2d_$v.err; echo "$v rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/r2d_$v.json')); print('$v', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['loss']['final'])"
  grep CMF_TRACE gpurun_out/r2d_$v.err | grep -E "fd_spectrum_H|fd_gram  *:6[0-9][0-9] |fd_denomH|fd_transconv|fd_corr|launch_mu|expansion" | awk '{print "   ", $2, $3, $4, $8, $9}'
done
FCMD="python bench.py --T 1048576 --steps 1 --warmup 1 --no-e2e --no-cpu --no-calibrated --no-direct"
for v in default oldfft minb2; do
  LIB=""; [ $v != default ] && LIB=$PWD/scripts/ab/libcmf_$v.so
  CMF_SM100_LIB=$LIB timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,launch__registers_per_thread --clock-control none -k regex:'fft_h_kernel|ifft_numH_kernel|mu_update' -s 6 -c 8 --csv --log-file gpurun_out/r2d_ncu_$v.csv $FCMD > /dev/null 2>&1; echo "ncu $v rc=$?"
done
