#!/bin/bash
set -u
OUT=gpurun_out/r2c16; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for cfg in "libnadm_decold.so 4" "libnadm_decold.so 3" "libnadm_b200.so 4" "libnadm_b200.so 3" "libnadm_decold.so 4" "libnadm_b200.so 3"; do set -- $cfg
  for loss in 1 0; do NADM_LIB=$1 NADM_DEC_WGS=$2 timeout 60 python tools/dec_probe.py 500000 20000 8 800 $loss 2>&1 | tail -1 | sed "s/^/[$1 WGS=$2] /"; done
done
for cfg in "libnadm_decold.so 4" "libnadm_b200.so 4" "libnadm_b200.so 3"; do set -- $cfg
NADM_LIB=$1 NADM_DEC_WGS=$2 timeout 200 python bench.py --rows 20000 --steps 200 --warmup 5 --no-cpu --no-e2e > $OUT/bench_$1_$2.json 2> $OUT/bench_$1_$2.err
python -c "import json;d=json.loads(open('$OUT/bench_$1_$2.json').read().strip().splitlines()[-1]);print('[$1 WGS=$2] ms/step',round(d['ms_per_step'],4),'grad_only',round(d['grad_only']['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'late',round(d['late_training']['ms_per_launch'],4))" || tail -5 $OUT/bench_$1_$2.err
done
