#!/bin/bash
set -u
OUT=gpurun_out/r2c27; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for nd in 0 1; do for m in 62500 500000; do
  NADM_NO_DEFER=$nd timeout 120 python tools/step_breakdown.py --snps $m --out $OUT/breakdown_nd${nd}_$m.json > /dev/null 2> $OUT/breakdown_nd${nd}_$m.err
  python -c "
import json; d=json.load(open('$OUT/breakdown_nd${nd}_$m.json')); print('no_defer',$nd,'M',$m,'step_us',round(d['step_ms']*1e3,1), {k:round(v,1) for k,v in d['calls_us'].items()})"
done; done
timeout 300 ncu --kernel-name 'regex:(enc_|dec_|mlp_|reduce_|step_)' --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_62500.csv python tools/step_breakdown.py --snps 62500 --steps 5 > /dev/null 2>&1
python tools/launch_summary.py $OUT/launches_62500.csv 2>&1 | head -14
