#!/bin/bash
# round 2, pass J: block length A/B with the round-2 transforms (c4: 512 vs 1024, H-side tiles of 16 or 8 column pairs; c3: 256 vs 512)
mkdir -p gpurun_out
run() { # name config env...
  name=$1; cfg=$2; shift 2
  env "$@" timeout 600 python bench.py --config $cfg --steps 5 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2j_$name.json 2> gpurun_out/r2j_$name.err; echo "$name rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/r2j_$name.json')); c=d['value_calibrated_loss']; print('$name', round(d['ms_per_step'],2), 'it/s', round(d['value'],2), 'direct', round(d['value_direct_loss'],2), 'calibrated', c and round(c['value'],2), {k:round(v['total_ms']/max(v['launches'],1),2) for k,v in d['roofline']['kernel_ms'].items()}, d['loss']['final'])"
}
run c4_b512 c4 CMF_X=0
run c4_b1024 c4 CMF_FD_B=1024
run c4_b1024_c8 c4 CMF_FD_B=1024 CMF_FD_COLS=8
run c4_b512_again c4 CMF_X=0
run c3_b256 c3 CMF_X=0
run c3_b512 c3 CMF_FD_B=512
run c3_b1024 c3 CMF_FD_B=1024
