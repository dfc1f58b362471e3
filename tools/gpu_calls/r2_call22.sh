#!/bin/bash
set -u
OUT=gpurun_out/r2c22; mkdir -p $OUT
for m in 62500 500000; do
  timeout 90 python tools/timeline.py libnadm_b200_tl.so $m > $OUT/timeline_dec_$m.txt 2>&1; tail -1 $OUT/timeline_dec_$m.txt
done
for m in 4096 16384 31250 62500; do
  timeout 120 python tools/step_breakdown.py --snps $m --out $OUT/breakdown_$m.json > /dev/null 2> $OUT/breakdown_$m.err
  python -c "
import json; d=json.load(open('$OUT/breakdown_$m.json')); print('M',$m,'step_us',round(d['step_ms']*1e3,1), {k:round(v,1) for k,v in d['calls_us'].items()})"
done
timeout 300 ncu --kernel-name 'regex:(enc_|dec_|mlp_|reduce_|step_)' --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_62500.csv python tools/step_breakdown.py --snps 62500 --steps 5 > /dev/null 2>&1
python tools/launch_summary.py $OUT/launches_62500.csv 2>&1 | head -20
