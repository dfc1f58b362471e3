#!/bin/bash
set -u
OUT=gpurun_out/r2c17; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:nadm:: -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/launches_bench.log 2>&1
python tools/launch_summary.py $OUT/launches.csv | tee $OUT/launch_summary.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:dec_tc_kernel --launch-skip 4 -c 1 \
   -o $OUT/dec_full -f python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_dec.log 2>&1; echo "ncu_dec rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:enc_.*_tc_kernel --launch-skip 8 -c 2 \
   -o $OUT/enc_full -f python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_enc.log 2>&1; echo "ncu_enc rc=$?"
timeout 300 python bench.py --rows 20000 --steps 200 --warmup 5 --no-cpu > $OUT/bench.json 2> $OUT/bench.err
python -c "import json;d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]);print('ms/step',round(d['ms_per_step'],4),'grad_only',round(d['grad_only']['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'late',round(d['late_training']['ms_per_launch'],4),'infer',round(d['infer']['value']),'e2e',round(d['e2e']['value']))" || tail -5 $OUT/bench.err
timeout 150 python tools/step_breakdown.py --out $OUT/breakdown.json 2> $OUT/breakdown.err | tail -1
ls -la $OUT
