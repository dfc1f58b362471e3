#!/bin/bash
set -u
OUT=gpurun_out/r2c8; mkdir -p $OUT
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
for w in fwd bwd bwd_adam; do timeout 60 python tools/enc_probe.py $w 500000 20000 2>&1 | tail -1; done
NADM_ENC_BWD_SLAB=1 timeout 60 python tools/enc_probe.py bwd_adam 500000 20000 2>&1 | tail -1 | sed "s/^/[bwd slab] /"
for loss in 1 0; do timeout 60 python tools/dec_probe.py 500000 20000 8 800 $loss 2>&1 | tail -1; done
} 2>&1 | tee $OUT/probes.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout 200 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu --no-e2e > $OUT/bench.json 2> $OUT/bench.err
python -c "import json;d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]);print('ms/step',round(d['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'infer',round(d['infer']['value']))" || tail -5 $OUT/bench.err
timeout 150 python tools/step_breakdown.py --out $OUT/breakdown.json 2> $OUT/breakdown.err | tail -1
