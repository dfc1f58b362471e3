#!/bin/bash
set -u
OUT=gpurun_out/r2c24; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for m in 62500 500000; do
  timeout 90 python tools/timeline_enc.py libnadm_b200_tl.so $m > $OUT/timeline_enc_$m.txt 2>&1; tail -2 $OUT/timeline_enc_$m.txt
  timeout 90 python tools/timeline.py libnadm_b200_tl.so $m > $OUT/timeline_dec_$m.txt 2>&1; tail -2 $OUT/timeline_dec_$m.txt
done
for gb in 0 1; do for m in 62500 500000; do
  NADM_NO_GRIDBAR=$gb timeout 120 python tools/step_breakdown.py --snps $m --out $OUT/breakdown_nogb${gb}_$m.json > /dev/null 2> $OUT/breakdown_nogb${gb}_$m.err
  python -c "
import json; d=json.load(open('$OUT/breakdown_nogb${gb}_$m.json')); print('no_gridbar',$gb,'M',$m,'step_us',round(d['step_ms']*1e3,1), {k:round(v,1) for k,v in d['calls_us'].items()})"
done; done
