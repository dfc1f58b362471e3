#!/bin/bash
set -u
OUT=gpurun_out/r2c14; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for w in fwd bwd bwd_adam; do timeout 60 python tools/enc_probe.py $w 500000 20000 2>&1 | tail -1; done
for loss in 1 0; do timeout 60 python tools/dec_probe.py 500000 20000 8 800 $loss 2>&1 | tail -1; done
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
NADM_ENC_FWD_SLAB=1 NADM_ENC_BWD_SLAB=1 timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 -k "encoder or training or fullsize or cfg2 or rsvd or geno or smoke" > $OUT/pytest_slab.log 2>&1; echo "pytest slab rc=$?"; tail -3 $OUT/pytest_slab.log
timeout 300 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu > $OUT/bench.json 2> $OUT/bench.err
python -c "import json;d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]);print('ms/step',round(d['ms_per_step'],4),'grad_only',round(d['grad_only']['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'late',{k:v for k,v in d['late_training'].items() if k!='what'},'infer',round(d['infer']['value']),'e2e',round(d['e2e']['value']))" || tail -5 $OUT/bench.err
