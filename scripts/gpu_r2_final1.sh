#!/bin/bash
# round 2, final 1-GPU pass: smoke, the default bench command (e2e + CPU baseline), reference arm
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke_final.log
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c4_final.json 2> gpurun_out/r2_bench_c4_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_c4_final.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_c4_final.json')); print(d['value'], d['ms_per_step'], d['value_direct_loss'], d['value_calibrated_loss'], d['e2e'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'], d['cpu_baseline'], d['roofline_hbm'], d['clocks'])"
