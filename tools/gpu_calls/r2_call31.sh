#!/bin/bash
# 2 GPUs: the driver's own command (default rows, e2e leg included) + the per-call breakdown
set -u
OUT=gpurun_out/r2c31; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 200 --warmup 5 --breakdown > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "n2 rc=$?"; tail -3 $OUT/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c31/bench_n2.json').read().strip().splitlines()[-1])
print('n2 ms', round(d['ms_per_step'],4), 'e2e', d['e2e'], 'calls', d['calls_us'], 'infer', d['infer']['value'])
PY
