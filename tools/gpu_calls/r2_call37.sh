#!/bin/bash
set -u
OUT=gpurun_out/r2c37; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "8-peer" > $OUT/pytest_sharded.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_sharded.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 300 --warmup 5 --rows 20000 --no-e2e --breakdown > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "n8 rc=$?"; tail -2 $OUT/bench_n8.err
timeout 300 python bench.py --gpus 1 --steps 300 --warmup 5 --rows 20000 --no-cpu --no-e2e > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "n1 rc=$?"
python - <<'PY'
import json
for t in ('n8','n1'):
    d=json.loads(open(f'gpurun_out/r2c37/bench_{t}.json').read().strip().splitlines()[-1])
    print(t, 'ms', round(d['ms_per_step'],4), 'grad_only', round(d['grad_only']['ms_per_step'],4), d.get('calls_us'), d['loss'])
PY
