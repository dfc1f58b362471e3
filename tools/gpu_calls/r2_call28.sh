#!/bin/bash
set -u
OUT=gpurun_out/r2c28; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > $OUT/pytest_sharded.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_sharded.log
for x in peer nccl; do
NADM_XCHG=$x timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --rows 20000 --no-cpu --no-e2e > $OUT/bench_n2_$x.json 2> $OUT/bench_n2_$x.err; echo "n2 $x rc=$?"
python - <<PY
import json
d=json.loads(open('$OUT/bench_n2_$x.json').read().strip().splitlines()[-1]); print('$x', d['n_gpus'], 'ms', round(d['ms_per_step'],4), d.get('exchange','')[:50], d.get('loss'))
PY
done
timeout 300 python bench.py --gpus 1 --rows 20000 --no-cpu --no-e2e > $OUT/bench_n1.json 2> $OUT/bench_n1.err; python -c "
import json; d=json.loads(open('$OUT/bench_n1.json').read().strip().splitlines()[-1]); print('n1 ms', round(d['ms_per_step'],4), d.get('loss'))"
