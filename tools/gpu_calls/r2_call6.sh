#!/bin/bash
set -u
OUT=gpurun_out/r2c6; mkdir -p $OUT
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
for v1 in 1 0 1 0; do
  NADM_ENC_FWD_V1=$v1 timeout 60 python tools/enc_probe.py fwd 500000 20000 2>&1 | tail -1 | sed "s/^/[V1=$v1] /"
  NADM_ENC_FWD_V1=$v1 timeout 60 python tools/enc_probe.py bwd 500000 20000 2>&1 | tail -1 | sed "s/^/[V1=$v1] /"
done
} 2>&1 | tee $OUT/enc_slab.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest.log
for v1 in 1 0; do
NADM_ENC_FWD_V1=$v1 timeout 200 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu --no-e2e > $OUT/bench_v1_$v1.json 2> $OUT/bench_v1_$v1.err
python -c "import json;d=json.loads(open('$OUT/bench_v1_$v1.json').read().strip().splitlines()[-1]);print('[V1=$v1] ms/step',round(d['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'late',d['late_training'],'infer',round(d['infer']['value']), 'loss', d['loss'])" || tail -5 $OUT/bench_v1_$v1.err
done
