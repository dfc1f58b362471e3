#!/bin/bash
set -u
OUT=gpurun_out/r2c21; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
timeout 300 python bench.py --workload cfg4 --rows 8000 --steps 30 --warmup 3 --no-cpu > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err; echo "cfg4 rc=$?"; tail -3 $OUT/bench_cfg4.err; head -c 600 $OUT/bench_cfg4.json; echo
for m in 62500 500000; do
  timeout 90 python tools/timeline_enc.py libnadm_b200_tl.so $m > $OUT/timeline_enc_$m.txt 2>&1; tail -1 $OUT/timeline_enc_$m.txt
  timeout 90 python tools/timeline.py libnadm_b200_tl.so $m > $OUT/timeline_dec_$m.txt 2>&1; tail -3 $OUT/timeline_dec_$m.txt
done
for pdl in 0 1; do for m in 62500 125000 500000; do
  NADM_PDL=$pdl timeout 120 python tools/step_breakdown.py --snps $m --out $OUT/breakdown_pdl${pdl}_$m.json > /dev/null 2> $OUT/breakdown_pdl${pdl}_$m.err
  python -c "
import json; d=json.load(open('$OUT/breakdown_pdl${pdl}_$m.json')); print('pdl',$pdl,'M',$m,'step_us',round(d['step_ms']*1e3,1), {k:round(v,1) for k,v in d['calls_us'].items()})"
done; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_62500.csv python tools/step_breakdown.py --snps 62500 --steps 5 > /dev/null 2>&1
python tools/launch_summary.py $OUT/launches_62500.csv 2>&1 | head -20
