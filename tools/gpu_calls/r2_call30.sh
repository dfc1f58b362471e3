#!/bin/bash
# final single-GPU validation of round 2: tests, smoke, launch list, ncu captures, the bench lines of every workload / arm
set -u
OUT=gpurun_out/r2c30; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $OUT/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.txt
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:nadm:: -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/launches_bench.log 2>&1
python tools/launch_summary.py $OUT/launches.csv | tee $OUT/launch_summary.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:dec_tc_kernel --launch-skip 4 -c 1 \
   -o $OUT/dec_full -f python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_dec.log 2>&1; echo "ncu_dec rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:enc_.*_tc_kernel --launch-skip 8 -c 2 \
   -o $OUT/enc_full -f python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_enc.log 2>&1; echo "ncu_enc rc=$?"
timeout 240 ncu --set full --clock-control none -k regex:mlp_ --launch-skip 12 -c 3 \
   -o $OUT/mlp_full -f python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_mlp.log 2>&1; echo "ncu_mlp rc=$?"
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default rc=$?"; tail -2 $OUT/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --impl reference-cuda --steps 10 --warmup 3 > $OUT/bench_reference_cuda.json 2> $OUT/bench_reference_cuda.err; echo "refcuda rc=$?"
timeout 300 python bench.py --workload cfg2 --no-cpu > $OUT/bench_cfg2.json 2> $OUT/bench_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --workload cfg4 --rows 8000 --steps 50 --warmup 3 --no-cpu --no-e2e > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err; echo "cfg4 rc=$?"
timeout 300 python bench.py --workload cfg5 --no-cpu > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err; echo "cfg5 rc=$?"
timeout 150 python tools/step_breakdown.py --out $OUT/breakdown.json 2> $OUT/breakdown.err | tail -1
python - <<'PY'
import json
for f in ['default','reference','reference_cuda','cfg2','cfg4','cfg5']:
    try:
        d=json.loads(open(f'gpurun_out/r2c30/bench_{f}.json').read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'launches', d.get('gpu_launches'))
    except Exception as e: print(f, 'ERR', e)
PY
ls $OUT
