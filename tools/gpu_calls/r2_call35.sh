#!/bin/bash
set -u
OUT=gpurun_out/r2c35; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
for nd in 0 1 0 1; do for m in 62500 500000; do
  NADM_NO_DEFER=$nd timeout 120 python tools/step_breakdown.py --snps $m --out $OUT/bd_nd${nd}_$m.json > /dev/null 2> $OUT/bd_nd${nd}_$m.err
  python -c "
import json; d=json.load(open('$OUT/bd_nd${nd}_$m.json')); print('no_defer',$nd,'M',$m,'step_us',round(d['step_ms']*1e3,1), {k:round(v,1) for k,v in d['calls_us'].items()})"
done; done
