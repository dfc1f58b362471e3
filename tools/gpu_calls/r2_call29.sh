#!/bin/bash
# 8-GPU box, final build: sharded parity at 4 and 8 ranks (2 ranks were checked on a 2-GPU box), scaling lines N = 1, 2, 4, 8, cfg5 at 8
set -u
OUT=gpurun_out/r2c29; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "4-peer or 8-peer" > $OUT/pytest_sharded.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_sharded.log
run() { n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then env "$@" timeout 300 python bench.py --gpus 1 --rows 20000 --no-cpu --no-e2e > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  else env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --rows 20000 --no-cpu --no-e2e > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err; fi
  echo "$tag rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('$OUT/bench_$tag.json').read().strip().splitlines()[-1]); print('$tag', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'value', round(d['value']), (d.get('exchange') or '')[:30], d.get('loss'))
except Exception as e: print('$tag parse error', e)
PY
}
run 8 n8 A=1
run 4 n4 A=1
run 1 n1 A=1
