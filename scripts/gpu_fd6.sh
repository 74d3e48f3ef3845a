#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_fd6.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_fd6.log
tail -3 gpurun_out/pytest_gpu_fd6.log
for B in 512 1024; do
  CMF_FD_B=$B timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/fd6_c4_B$B.json 2> gpurun_out/fd6_c4_B$B.err
  echo "B=$B exit $?"; tail -2 gpurun_out/fd6_c4_B$B.err
  python -c "import json; d=json.load(open('gpurun_out/fd6_c4_B$B.json')); print(d['value'], d['ms_per_step'], d['loss'], d['roofline']['kernel_ms'], d['clocks'])"
done
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_fd6_c4.json 2> gpurun_out/plain_fd6_c4.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fd6_c4.csv $CMD > gpurun_out/ncu_launch_fd6.log 2>&1
echo "launch list exit $?"
CMD2="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu"
$CMD2 > gpurun_out/plain_fd6_c4full.json 2> gpurun_out/plain_fd6_c4full.err && ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 11 -c 1 -o gpurun_out/prof_fd6_fqt_c4full_r1 $CMD2 > gpurun_out/ncu_fd6_c4full.log 2>&1
echo "ncu full exit $?"; grep -c "Profiling" gpurun_out/ncu_fd6_c4full.log
