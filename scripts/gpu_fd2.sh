#!/bin/bash
# full GPU suite + smoke + default bench + ncu launch list + ncu --set full of the two frequency-domain products
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_fd.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_fd.log
tail -5 gpurun_out/pytest_gpu_fd.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_fd.log 2>&1; echo "smoke exit $?"; tail -12 gpurun_out/smoke_fd.log
timeout 900 python bench.py > gpurun_out/bench_c4_fd.json 2> gpurun_out/bench_c4_fd.err; echo "bench exit $?"; tail -2 gpurun_out/bench_c4_fd.err
python -c "import json; d=json.load(open('gpurun_out/bench_c4_fd.json')); print(d['value'], d['ms_per_step'], d['value_direct_loss'], d['e2e'], d['roofline']['achieved'], d['roofline']['frac'], d['cpu_baseline'])"
CMD="python bench.py --config c4 --T 524288 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_fd_T512k.json 2> gpurun_out/plain_fd_T512k.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fd_T512k.csv $CMD > gpurun_out/ncu_launch_fd.log 2>&1
echo "launch list exit $?"
CMD2="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD2 > gpurun_out/plain_fd_c4full.json 2> gpurun_out/plain_fd_c4full.err && ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 7 -c 3 -o gpurun_out/prof_fd_c4full_r1 $CMD2 > gpurun_out/ncu_fd_c4full.log 2>&1
echo "ncu full exit $?"; tail -5 gpurun_out/ncu_fd_c4full.log
