#!/bin/bash
set -u
OUT=gpurun_out/r2c33; mkdir -p $OUT
NADM_LIB=libnadm_bwd512.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mlp or deferred or training_fixtures or step_fixture" > $OUT/pytest_512.log 2>&1; echo "pytest512 rc=$?"; tail -2 $OUT/pytest_512.log
for lib in libnadm_b200.so libnadm_bwd512.so libnadm_b200.so libnadm_bwd512.so; do for m in 62500 500000; do
  NADM_LIB=$lib timeout 120 python tools/step_breakdown.py --snps $m --out $OUT/bd_${lib}_$m.json > /dev/null 2> $OUT/bd_${lib}_$m.err
  python -c "
import json; d=json.load(open('$OUT/bd_${lib}_$m.json')); print('$lib','M',$m,'step_us',round(d['step_ms']*1e3,1), {k:round(v,1) for k,v in d['calls_us'].items()})"
done; done
