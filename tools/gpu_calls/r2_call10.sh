#!/bin/bash
# 2 GPUs: sharded parity with the fused peer exchange and with NCCL, bench at N=2 both ways
set -u
OUT=gpurun_out/r2c10; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q --timeout 280 > $OUT/pytest_sharded.log 2>&1; echo "pytest sharded rc=$?"; tail -25 $OUT/pytest_sharded.log
for x in peer nccl peer nccl; do
NADM_XCHG=$x timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --rows 20000 --steps 200 --warmup 5 --no-e2e > $OUT/bench2_$x.json 2> $OUT/bench2_$x.err; echo "bench2 $x rc=$?"
python -c "import json;d=json.loads(open('$OUT/bench2_$x.json').read().strip().splitlines()[-1]);print('[$x] 2gpu ms/step',round(d['ms_per_step'],4),'grad_only',round(d['grad_only']['ms_per_step'],4),'infer',round(d['infer']['value']),d['exchange'][:30],d['loss'])" || tail -5 $OUT/bench2_$x.err
done
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 -k "not sharded" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 200 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu --no-e2e > $OUT/bench1.json 2> $OUT/bench1.err
python -c "import json;d=json.loads(open('$OUT/bench1.json').read().strip().splitlines()[-1]);print('1gpu ms/step',round(d['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'infer',round(d['infer']['value']), d['loss'])" || tail -5 $OUT/bench1.err
timeout 150 python tools/step_breakdown.py --out $OUT/breakdown.json 2> $OUT/breakdown.err | tail -1
