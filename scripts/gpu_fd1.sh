#!/bin/bash
# first GPU check of the frequency-domain engine: parity tests, then A/B against the time-domain engine
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fd.py -q -s > gpurun_out/fd_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/fd_tests.log
tail -30 gpurun_out/fd_tests.log
for e in 1 2; do
  timeout 600 python bench.py --config c4 --T 524288 --steps 3 --warmup 3 --no-e2e --no-cpu --engine $e > gpurun_out/fd_ab_T512k_e$e.json 2> gpurun_out/fd_ab_T512k_e$e.err
  echo "engine $e exit $?"; cat gpurun_out/fd_ab_T512k_e$e.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['loss'], d['roofline']['kernel_ms'])"
done
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --no-e2e --no-cpu --engine 2 > gpurun_out/fd_c4_e2.json 2> gpurun_out/fd_c4_e2.err
echo "c4 engine 2 exit $?"; tail -3 gpurun_out/fd_c4_e2.err; cat gpurun_out/fd_c4_e2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value_direct_loss'], d['loss'], d['roofline']['kernel_ms'])"
