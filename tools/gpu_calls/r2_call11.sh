#!/bin/bash
set -u
OUT=gpurun_out/r2c11; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for v1 in 1 0 1 0; do
NADM_ENC_FWD_V1=$v1 timeout 200 python bench.py --rows 20000 --steps 200 --warmup 5 --no-cpu --no-e2e > $OUT/bench_v1_$v1.json 2> $OUT/bench_v1_$v1.err
python -c "import json;d=json.loads(open('$OUT/bench_v1_$v1.json').read().strip().splitlines()[-1]);print('[fwd V1=$v1] ms/step',round(d['ms_per_step'],4),'grad_only',round(d['grad_only']['ms_per_step'],4),'infer',round(d['infer']['value']))" || tail -5 $OUT/bench_v1_$v1.err
done
for m in 250000 125000 62500; do timeout 150 python tools/step_breakdown.py --snps $m --out $OUT/breakdown_$m.json 2> $OUT/breakdown_$m.err | tail -1; done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 120 -k "host_fed or graph" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 300 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu > $OUT/bench_e2e.json 2> $OUT/bench_e2e.err
python -c "import json;d=json.loads(open('$OUT/bench_e2e.json').read().strip().splitlines()[-1]);print('e2e',d['e2e'])" || tail -5 $OUT/bench_e2e.err
