#!/bin/bash
# round 2, pass I: evidence of the final build -- ncu launch list of the default command, ncu --set full of the transforms and the two
# products at T = 2^20, block length 1024 A/B, live trace of config 5 (what is outside the H sweep)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-calibrated"
timeout 600 $CMD > gpurun_out/r2i_plain_c4.json 2> gpurun_out/r2i_plain_c4.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c4_final.csv $CMD > gpurun_out/r2i_ncu_launch.log 2>&1; echo "launch list rc=$?"
FCMD="python bench.py --T 1048576 --steps 1 --warmup 1 --no-e2e --no-cpu --no-calibrated --no-direct"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fft_h_kernel|ifft_numH_kernel|mu_update_vec4|tc_kernel' -s 8 -c 14 -o gpurun_out/r2_final_kernels $FCMD > gpurun_out/r2i_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/r2i_ncu_full.log
CMF_FD_B=1024 timeout 600 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu --no-calibrated --no-direct > gpurun_out/r2i_c4_b1024.json 2> gpurun_out/r2i_c4_b1024.err; echo "b1024 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2i_c4_b1024.json')); print('B=1024', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['moved']['note']); d=json.load(open('gpurun_out/r2i_plain_c4.json')); print('B=512', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
CMF_TRACE=1 CMF_TRACE_SKIP=40 timeout 900 python bench.py --config c5 --alg hals --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2i_c5_trace.json 2> gpurun_out/r2i_c5_trace.err; echo "c5 rc=$?"
grep "CMF_TRACE " gpurun_out/r2i_c5_trace.err | awk '{print $2, $3, $4, $6, $9, $11}' | sort -k4 -n -r | head -24
