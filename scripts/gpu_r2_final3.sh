#!/bin/bash
# round 2, closing pass on the final build: full GPU suite, default bench command, ncu launch list of the default command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c4_final.json 2> gpurun_out/r2_bench_c4_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_c4_final.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_c4_final.json')); c=d['value_calibrated_loss']; print(d['value'], d['ms_per_step'], d['value_direct_loss'], c and c['value'], d['e2e']['value'], d['e2e']['phases_s'], d['roofline']['frac'], d['roofline']['moved']['note'], d['roofline']['kernel_ms'], d['roofline_hbm']['frac'], d['clocks'])"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-calibrated"
timeout 200 $CMD > gpurun_out/r2_plain_c4_final.json 2> gpurun_out/r2_plain_c4_final.err; echo "plain rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c4_final.csv $CMD > gpurun_out/r2_ncu_launch_final.log 2>&1; echo "launch list rc=$?"
