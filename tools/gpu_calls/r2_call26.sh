#!/bin/bash
set -u
OUT=gpurun_out/r2c26; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
for m in 62500 500000; do
  timeout 120 python tools/step_breakdown.py --snps $m --out $OUT/breakdown_$m.json > /dev/null 2> $OUT/breakdown_$m.err
  python -c "
import json; d=json.load(open('$OUT/breakdown_$m.json')); print('M',$m,'step_us',round(d['step_ms']*1e3,1), {k:round(v,1) for k,v in d['calls_us'].items()})"
done
timeout 300 python bench.py --rows 20000 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err; head -c 400 $OUT/bench.json; echo
