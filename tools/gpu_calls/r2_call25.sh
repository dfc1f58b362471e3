#!/bin/bash
# 8-GPU box: sharded parity at 2/4/8 ranks, then the scaling bench lines the driver will run (N = 1, 2, 4, 8), both exchanges at 8
set -u
OUT=gpurun_out/r2c25; mkdir -p $OUT
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c > $OUT/gpu.txt
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $OUT/pytest_sharded.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_sharded.log
run() { # n tag extra-env...
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then env "$@" timeout 400 python bench.py --gpus 1 --rows 20000 --no-cpu --no-e2e > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  else env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --rows 20000 --no-cpu --no-e2e > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err; fi
  echo "$tag rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('$OUT/bench_$tag.json').read().strip().splitlines()[-1]); print('$tag', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'value', round(d['value']), d.get('exchange','')[:40], d.get('loss'))
except Exception as e: print('$tag parse error', e)
PY
}
run 1 n1 A=1
run 2 n2 A=1
run 4 n4 A=1
run 8 n8 A=1
run 8 n8_nccl NADM_XCHG=nccl
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --workload cfg5 --no-cpu > $OUT/bench_cfg5_n8.json 2> $OUT/bench_cfg5_n8.err; echo "cfg5 n8 rc=$?"; head -c 500 $OUT/bench_cfg5_n8.json; echo
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 4 --workload cfg4 --rows 8000 --no-cpu --no-e2e > $OUT/bench_cfg4_n4.json 2> $OUT/bench_cfg4_n4.err; echo "cfg4 n4 rc=$?"; head -c 500 $OUT/bench_cfg4_n4.json; echo
