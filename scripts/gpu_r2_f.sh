#!/bin/bash
# round 2, pass F: full GPU suite on the build with batched loads / 1-D tile-fastest grids / N-split transconv; c4 bench with trace;
# configs 1-2 timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2f_pytest.log
CMF_TRACE=1 CMF_TRACE_SKIP=150 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2f_c4.json 2> gpurun_out/r2f_c4.err; echo "c4 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2f_c4.json')); print(d['value'], d['ms_per_step'], d['value_direct_loss'], d['value_calibrated_loss'] and d['value_calibrated_loss']['value'], d['roofline']['kernel_ms'], d['loss']['final'])"
grep CMF_TRACE gpurun_out/r2f_c4.err | grep -E "fd_spectrum_H|fd_gram|fd_denomH|fd_transconv|fd_corr|launch_mu|expansion|fd_conv_loss|fd_build" | awk '{print "   ", $2, $3, $4, $8, $9}'
timeout 600 python scripts/configs12.py > gpurun_out/r2f_configs12.json 2> gpurun_out/r2f_configs12.err; echo "configs12 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2f_configs12.json')); print({k:(v['gpu_f64_s_per_100it'], v['gpu_f32_s_per_100it'], v['cpu_oracle_s_per_100it'], v['max_rel_loss_err_f64'], v['max_rel_loss_err_f32']) for k,v in d.items()})"
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2f_c3.json 2> gpurun_out/r2f_c3.err; echo "c3 rc=$?"
python -c "import json; d=json.load(open('gpurun_out/r2f_c3.json')); print('c3', d['value'], d['ms_per_step'], d['value_direct_loss'], d['value_calibrated_loss'] and d['value_calibrated_loss']['value'], d['loss'])"
