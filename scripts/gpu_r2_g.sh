#!/bin/bash
# round 2, pass G: scheduling order of the transforms' 1-D grids (CMF_FD_ORDER=0 tile fastest, 1 block fastest), live and under ncu
mkdir -p gpurun_out
for o in 0 1 0 1; do
  CMF_FD_ORDER=$o CMF_TRACE=1 CMF_TRACE_SKIP=150 timeout 600 python bench.py --steps 6 --warmup 2 --no-e2e --no-cpu --no-calibrated > gpurun_out/r2g_o$o.json 2> gpurun_out/r2g_o$o.err; echo "order $o rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/r2g_o$o.json')); print('order $o', d['value'], d['ms_per_step'], d['value_direct_loss'], d['roofline']['kernel_ms'])"
  grep CMF_TRACE gpurun_out/r2g_o$o.err | grep -E "fd_spectrum_H|fd_gram  *:62|fd_denomH|fd_transconv  *:66|fd_conv_loss|fd_build" | awk '{print "   ", $2, $3, $4, $8, $9}'
done
FCMD="python bench.py --T 1048576 --steps 1 --warmup 1 --no-e2e --no-cpu --no-calibrated"
for o in 0 1; do
  CMF_FD_ORDER=$o timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct --clock-control none -k regex:'fft_h_kernel|ifft_numH_kernel|ifft_resid|fft_x_kernel' -c 14 --csv --log-file gpurun_out/r2g_ncu_o$o.csv $FCMD > /dev/null 2>&1; echo "ncu order $o rc=$?"
done
