#!/bin/bash
set -u
OUT=gpurun_out/r2c4; mkdir -p $OUT
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for wg in 3 4 3 4; do
  for loss in 1 0; do NADM_DEC_WGS=$wg timeout 60 python tools/dec_probe.py 500000 20000 8 800 $loss 2>&1 | tail -1 | sed "s/^/[WGS=$wg] /"; done
done
for wg in 3 4; do for B in 768 1024; do NADM_DEC_WGS=$wg timeout 60 python tools/dec_probe.py 500000 20000 8 $B 1 2>&1 | tail -1 | sed "s/^/[WGS=$wg] /"; done; done
NADM_DEC_WGS=4 NADM_DEC_SLOTS=3 timeout 60 python tools/dec_probe.py 500000 20000 8 800 1 2>&1 | tail -1 | sed "s/^/[WGS=4 SLOTS=3] /"
NADM_DEC_WGS=4 timeout 60 python tools/dec_probe.py 500000 20000 5 800 1 2>&1 | tail -1 | sed "s/^/[WGS=4 k=5] /"
} 2>&1 | tee $OUT/dec_wgs.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest.log
for wg in 3 4; do
NADM_DEC_WGS=$wg timeout 150 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu --no-e2e > $OUT/bench_wg$wg.json 2> $OUT/bench_wg$wg.err
python -c "import json;d=json.loads(open('$OUT/bench_wg$wg.json').read().strip().splitlines()[-1]);print('[WGS=$wg] ms/step',round(d['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'grad_only',round(d['grad_only']['ms_per_step'],4))" || tail -3 $OUT/bench_wg$wg.err
done
