#!/bin/bash
# final round-1 pass: full GPU suite, smoke, default bench (e2e + cpu), ncu launch list of the default command at full size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_final.log
timeout 900 python bench.py > gpurun_out/bench_c4_final.json 2> gpurun_out/bench_c4_final.err; echo "bench exit $?"; tail -2 gpurun_out/bench_c4_final.err
python -c "import json; d=json.load(open('gpurun_out/bench_c4_final.json')); print(d['value'], d['ms_per_step'], d['value_direct_loss'], d['e2e'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['moved'], d['roofline']['kernel_ms'], d['gpu_launches'], d['cpu_baseline'], d['roofline_hbm'], d['clocks'])"
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3_final.json 2> gpurun_out/bench_c3_final.err
python -c "import json; d=json.load(open('gpurun_out/bench_c3_final.json')); print('c3', d['value'], d['ms_per_step'], d['value_direct_loss'], d['loss'])"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_final_c4.json 2> gpurun_out/plain_final_c4.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final_c4.csv $CMD > gpurun_out/ncu_launch_final.log 2>&1
echo "launch list exit $?"
