#!/bin/bash
# round 2, pass A: calibrated-expansion parity at the benchmark (K, L), c4 bench with the three loss regimes, ncu launch list of
# the default command, ncu --set full of the HALS round kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bench_shape.py -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2a_pytest.log
grep "max rel err" gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2a_c4.json 2> gpurun_out/r2a_c4.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2a_c4.json')); print(d['value'], d['ms_per_step'], d['value_direct_loss'], d['value_calibrated_loss'], d['roofline']['kernel_ms'], d['clocks'])"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-calibrated"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_c4.csv $CMD > gpurun_out/r2a_ncu_launch.log 2>&1; echo "launch list rc=$?"
HCMD="python scripts/hals_scale.py --N 512 --T 1048576 --K 128 --L 32 --iters 1"
timeout 300 $HCMD > gpurun_out/r2a_hals_plain.log 2>&1; echo "hals plain rc=$?"; tail -2 gpurun_out/r2a_hals_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hals2_sweep -c 1 -o gpurun_out/r2_hals2_sweep $HCMD > gpurun_out/r2a_ncu_hals.log 2>&1; echo "ncu hals rc=$?"; tail -3 gpurun_out/r2a_ncu_hals.log
