#!/bin/bash
# round 2, multi-GPU pass: usage gpu_r2_multi.sh <ngpu>
N=$1
mkdir -p gpurun_out
if [ "$N" == "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r2_pytest_multi.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_c4_${N}gpu_final.json 2> gpurun_out/r2_bench_c4_${N}gpu_final.err; echo "c4 x$N rc=$?"; tail -2 gpurun_out/r2_bench_c4_${N}gpu_final.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_c4_${N}gpu_final.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['value_direct_loss'], d['value_calibrated_loss'] and d['value_calibrated_loss']['value'], d['e2e'] and d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['contraction_share_of_step'])"
if [ "$N" == "8" ]; then
  timeout 900 $TR bench.py --gpus $N --config c5 --alg hals --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/r2_bench_c5_${N}gpu_final.json 2> gpurun_out/r2_bench_c5_${N}gpu_final.err; echo "c5 x$N rc=$?"; tail -2 gpurun_out/r2_bench_c5_${N}gpu_final.err
  python -c "import json; d=json.load(open('gpurun_out/r2_bench_c5_${N}gpu_final.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['critical_path']['ms_per_sweep'], d['loss'])"
fi
