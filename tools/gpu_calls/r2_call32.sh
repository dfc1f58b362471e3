#!/bin/bash
# 8 GPUs: the driver's own command (default rows, e2e leg included) + the per-call breakdown
set -u
OUT=gpurun_out/r2c32; mkdir -p $OUT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 200 --warmup 5 --breakdown > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "n8 rc=$?"; tail -3 $OUT/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c32/bench_n8.json').read().strip().splitlines()[-1])
print('n8 ms', round(d['ms_per_step'],4), 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'calls', d['calls_us']['rank0'], d['calls_us']['max_over_ranks'], 'infer', d['infer']['value'], 'grad_only', d['grad_only']['ms_per_step'])
PY
