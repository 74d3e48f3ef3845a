#!/bin/bash
# round 2, pass K: new default block length (pow2 >= 8L, <= 1024): full GPU suite, c4 / c3 / c5 bench, ncu --set full of the two products at full c4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2k_c4.json 2> gpurun_out/r2k_c4.err; echo "c4 rc=$?"; tail -2 gpurun_out/r2k_c4.err
python -c "import json; d=json.load(open('gpurun_out/r2k_c4.json')); c=d['value_calibrated_loss']; print(d['value'], d['ms_per_step'], d['value_direct_loss'], c and c['value'], d['e2e'] and d['e2e']['value'], d['e2e'] and d['e2e']['phases_s'], d['roofline']['frac'], d['roofline']['moved'], d['roofline']['kernel_ms'], d['roofline_hbm']['frac'])"
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2k_c3.json 2> gpurun_out/r2k_c3.err; echo "c3 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2k_c3.json')); c=d['value_calibrated_loss']; print('c3', d['value'], d['ms_per_step'], d['value_direct_loss'], c and c['value'], d['roofline']['moved']['note'])"
timeout 900 python bench.py --config c5 --alg hals --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2k_c5.json 2> gpurun_out/r2k_c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/r2k_c5.err
python -c "import json; d=json.load(open('gpurun_out/r2k_c5.json')); print('c5', d['value'], d['ms_per_step'], d['critical_path']['ms_per_sweep'], d['loss'], d['roofline']['moved']['note'])"
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-calibrated --no-direct"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_kernel' -s 6 -c 8 -o gpurun_out/r2_products_c4_full $CMD > gpurun_out/r2k_ncu_products.log 2>&1; echo "ncu products rc=$?"; tail -2 gpurun_out/r2k_ncu_products.log
