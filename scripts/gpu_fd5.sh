#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fd.py tests/test_gpu_scale.py -x -q -s > gpurun_out/pytest_gpu_fd5.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_fd5.log
grep -E "loss_mode|passed|failed|rror" gpurun_out/pytest_gpu_fd5.log | tail -12
timeout 900 python bench.py > gpurun_out/bench_c4_fd5.json 2> gpurun_out/bench_c4_fd5.err; echo "bench exit $?"; tail -2 gpurun_out/bench_c4_fd5.err
python -c "import json; d=json.load(open('gpurun_out/bench_c4_fd5.json')); print(d['value'], d['ms_per_step'], d['value_direct_loss'], d['e2e'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['gpu_launches'])"
CMD="python bench.py --config c4 --T 524288 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_fd5_T512k.json 2> gpurun_out/plain_fd5_T512k.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fd5_T512k.csv $CMD > gpurun_out/ncu_launch_fd5.log 2>&1
echo "launch list exit $?"
CMD2="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD2 > gpurun_out/plain_fd5_c4full.json 2> gpurun_out/plain_fd5_c4full.err && ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 8 -c 3 -o gpurun_out/prof_fd5_c4full_r1 $CMD2 > gpurun_out/ncu_fd5_c4full.log 2>&1
echo "ncu full exit $?"; grep -c "Profiling" gpurun_out/ncu_fd5_c4full.log
