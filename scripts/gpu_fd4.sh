#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_fd4.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_fd4.log
tail -5 gpurun_out/pytest_gpu_fd4.log
for B in 512 1024; do
  CMF_FD_B=$B timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/fd4_c4_B$B.json 2> gpurun_out/fd4_c4_B$B.err
  echo "B=$B exit $?"; tail -2 gpurun_out/fd4_c4_B$B.err
  python -c "import json; d=json.load(open('gpurun_out/fd4_c4_B$B.json')); print(d['value'], d['ms_per_step'], d['loss'], d['roofline']['kernel_ms'], d['clocks'])"
done
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/fd4_c3.json 2> gpurun_out/fd4_c3.err
python -c "import json; d=json.load(open('gpurun_out/fd4_c3.json')); print('c3', d['value'], d['ms_per_step'], d['loss'], d['roofline']['kernel_ms'])"
CMD="python bench.py --config c4 --T 524288 --steps 2 --warmup 3 --no-e2e --no-cpu --loss-mode 1"
$CMD > gpurun_out/plain_fd4_T512k.json 2> gpurun_out/plain_fd4_T512k.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fd4_T512k.csv $CMD > gpurun_out/ncu_launch_fd4.log 2>&1
echo "launch list exit $?"
