#!/bin/bash
set -u
OUT=gpurun_out/r2c12; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 tests/cuda/gather_probe.bin 20000 16 2>&1 | tee $OUT/gather_probe.txt
echo "--- 900-row matrix: one row set, L2-resident ---" | tee -a $OUT/gather_probe.txt
timeout 300 tests/cuda/gather_probe.bin 900 16 2>&1 | tee -a $OUT/gather_probe.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
NADM_ENC_FWD_SLAB=1 NADM_ENC_BWD_SLAB=1 timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 -k "encoder or training or fullsize or cfg2 or rsvd or geno or smoke" > $OUT/pytest_slab.log 2>&1; echo "pytest slab rc=$?"; tail -3 $OUT/pytest_slab.log
timeout 300 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu > $OUT/bench.json 2> $OUT/bench.err
python -c "import json;d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]);print('ms/step',round(d['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'late',{k:v for k,v in d['late_training'].items() if k!='what'},'infer',round(d['infer']['value']),'e2e',round(d['e2e']['value']))" || tail -5 $OUT/bench.err
