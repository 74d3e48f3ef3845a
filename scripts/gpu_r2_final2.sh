#!/bin/bash
# round 2, last pass: full GPU suite and the default bench command on the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_c4_final.json 2> gpurun_out/r2_bench_c4_final.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_bench_c4_final.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_c4_final.json')); c=d['value_calibrated_loss']; print(d['value'], d['ms_per_step'], d['value_direct_loss'], c and c['value'], d['e2e'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['moved'], d['roofline']['kernel_ms'], d['roofline_hbm']['frac'], d['clocks'], d['cpu_baseline'])"
