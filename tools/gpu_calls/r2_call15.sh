#!/bin/bash
set -u
OUT=gpurun_out/r2c15; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for lib in libnadm_decold.so libnadm_b200.so libnadm_decold.so libnadm_b200.so; do
  for loss in 1 0; do NADM_LIB=$lib timeout 60 python tools/dec_probe.py 500000 20000 8 800 $loss 2>&1 | tail -1 | sed "s/^/[$lib] /"; done
done
for epw in 4 8 4 8; do NADM_ENC_BWD_EPW=$epw timeout 60 python tools/enc_probe.py bwd_adam 500000 20000 2>&1 | tail -1 | sed "s/^/[EPW=$epw] /"; done
NADM_ENC_BWD_EPW=8 timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 -k "encoder_bwd or training or adam or geno or rsvd" > $OUT/pytest_epw8.log 2>&1; echo "pytest epw8 rc=$?"; tail -2 $OUT/pytest_epw8.log
for cfg in "libnadm_decold.so 4" "libnadm_b200.so 4" "libnadm_b200.so 8"; do set -- $cfg
NADM_LIB=$1 NADM_ENC_BWD_EPW=$2 timeout 200 python bench.py --rows 20000 --steps 200 --warmup 5 --no-cpu --no-e2e > $OUT/bench_$1_$2.json 2> $OUT/bench_$1_$2.err
python -c "import json;d=json.loads(open('$OUT/bench_$1_$2.json').read().strip().splitlines()[-1]);print('[$1 EPW=$2] ms/step',round(d['ms_per_step'],4),'grad_only',round(d['grad_only']['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'late',round(d['late_training']['ms_per_launch'],4))" || tail -5 $OUT/bench_$1_$2.err
done
