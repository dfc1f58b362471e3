#!/bin/bash
set -u
OUT=gpurun_out/r2c36; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deferred or training_fixtures or step_fixture or graph" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for nd in 0 1 0 1; do for m in 62500 500000; do
  NADM_NO_DEFER=$nd timeout 120 python tools/step_breakdown.py --snps $m --out $OUT/bd_nd${nd}_$m.json > /dev/null 2> $OUT/bd_nd${nd}_$m.err
  python -c "
import json; d=json.load(open('$OUT/bd_nd${nd}_$m.json')); print('no_defer',$nd,'M',$m,'step_us',round(d['step_ms']*1e3,1), {k:round(v,1) for k,v in d['calls_us'].items()})"
done; done
