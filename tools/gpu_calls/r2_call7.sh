#!/bin/bash
set -u
OUT=gpurun_out/r2c7; mkdir -p $OUT
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for lib in libnadm_b200.so libnadm_s4d1.so libnadm_s2d2_kol.so libnadm_s2d2_kow.so; do
  NADM_LIB=$lib timeout 60 python tools/enc_probe.py fwd 500000 20000 2>&1 | tail -1 | sed "s/^/[$lib] /"
  NADM_LIB=$lib timeout 60 python tools/enc_probe.py bwd 500000 20000 2>&1 | tail -1 | sed "s/^/[$lib] /"
done
NADM_ENC_FWD_V1=1 timeout 60 python tools/enc_probe.py fwd 500000 20000 2>&1 | tail -1 | sed "s/^/[v1] /"
NADM_ENC_FWD_V1=1 timeout 60 python tools/enc_probe.py bwd 500000 20000 2>&1 | tail -1 | sed "s/^/[v1] /"
NADM_ENC_ISSUERS=1 timeout 60 python tools/enc_probe.py fwd 500000 20000 2>&1 | tail -1 | sed "s/^/[1 issuer] /"
NADM_ENC_ISSUERS=1 timeout 60 python tools/enc_probe.py bwd 500000 20000 2>&1 | tail -1 | sed "s/^/[1 issuer] /"
} 2>&1 | tee $OUT/enc_slab.txt
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
timeout 200 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu --no-e2e > $OUT/bench.json 2> $OUT/bench.err
python -c "import json;d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]);print('ms/step',round(d['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'infer',round(d['infer']['value']))" || tail -5 $OUT/bench.err
