#!/bin/bash
# round 2, pass E: FFT kernels with batched global loads (default = 3 CTAs/SM, minb2 = 2 CTAs/SM, oldfft = round-1 kernels), live and under ncu;
# parity; live launch trace of config 1 (is it launch-bound?)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fd.py tests/test_gpu_bench_shape.py tests/test_gpu_scale.py -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2e_pytest.log
for v in default minb2 oldfft; do
  LIB=""; [ $v != default ] && LIB=$PWD/scripts/ab/libcmf_$v.so
  EXTRA="--no-direct"; [ $v == default ] && EXTRA=""
  CMF_SM100_LIB=$LIB CMF_TRACE=1 CMF_TRACE_SKIP=150 timeout 600 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu --no-calibrated $EXTRA > gpurun_out/r2e_$v.json 2> gpurun_out/r2e_$v.err; echo "$v rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/r2e_$v.json')); print('$v', d['value'], d['ms_per_step'], d['value_direct_loss'], d['roofline']['kernel_ms'], d['loss']['final'])"
  grep CMF_TRACE gpurun_out/r2e_$v.err | grep -E "fd_spectrum_H|fd_gram|fd_denomH|fd_transconv|fd_corr|launch_mu|expansion|fd_conv_loss" | awk '{print "   ", $2, $3, $4, $8, $9}'
done
FCMD="python bench.py --T 1048576 --steps 1 --warmup 1 --no-e2e --no-cpu --no-calibrated --no-direct"
for v in default minb2; do
  LIB=""; [ $v != default ] && LIB=$PWD/scripts/ab/libcmf_$v.so
  CMF_SM100_LIB=$LIB timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread --clock-control none -k regex:'fft_h_kernel|ifft_numH_kernel|mu_update' -s 6 -c 8 --csv --log-file gpurun_out/r2e_ncu_$v.csv $FCMD > /dev/null 2>&1; echo "ncu $v rc=$?"
done
CMF_TRACE=1 timeout 600 python scripts/configs12.py > gpurun_out/r2e_configs12.json 2> gpurun_out/r2e_configs12.err; echo "configs12 rc=$?"
grep "CMF_TRACE:" gpurun_out/r2e_configs12.err | head; python -c "import json; d=json.load(open('gpurun_out/r2e_configs12.json')); print({k:(v['gpu_f64_s_per_100it'], v['gpu_f32_s_per_100it']) for k,v in d.items()})"
