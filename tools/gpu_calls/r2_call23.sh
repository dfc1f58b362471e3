#!/bin/bash
set -u
OUT=gpurun_out/r2c23; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
for gb in 0 1; do for m in 62500 500000; do
  NADM_NO_GRIDBAR=$gb timeout 120 python tools/step_breakdown.py --snps $m --out $OUT/breakdown_nogb${gb}_$m.json > /dev/null 2> $OUT/breakdown_nogb${gb}_$m.err
  python -c "
import json; d=json.load(open('$OUT/breakdown_nogb${gb}_$m.json')); print('no_gridbar',$gb,'M',$m,'step_us',round(d['step_ms']*1e3,1), {k:round(v,1) for k,v in d['calls_us'].items()})"
done; done
timeout 300 python bench.py --rows 20000 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -2 $OUT/bench.err; head -c 400 $OUT/bench.json; echo
timeout 300 python bench.py --workload cfg4 --rows 8000 --steps 30 --warmup 3 --no-cpu --no-e2e > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err; echo "cfg4 rc=$?"; head -c 400 $OUT/bench_cfg4.json; echo
timeout 300 ncu --kernel-name 'regex:(enc_|dec_|mlp_|reduce_|step_)' --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_62500.csv python tools/step_breakdown.py --snps 62500 --steps 5 > /dev/null 2>&1
python tools/launch_summary.py $OUT/launches_62500.csv 2>&1 | head -14
