#!/bin/bash
# round 2, GPU visit 1: baseline parity on this box, the prepared encoder variants (TS, TS + half-stage hand-over),
# decoder pipeline floor (decode skipped), encoder timelines.  Outputs under gpurun_out/r2c1/.
set -u
OUT=gpurun_out/r2c1; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q --timeout 60 > $OUT/pytest_base.log 2>&1; echo "pytest base rc=$?"; tail -3 $OUT/pytest_base.log
# half-stage variant: smoke, then the suite
NADM_LIB=libnadm_b200_half.so NADM_ENC_TS=1 timeout 90 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/half_smoke.txt 2>&1; rc=$?; tail -2 $OUT/half_smoke.txt; echo "half smoke rc=$rc"
if [ $rc -eq 0 ]; then
  NADM_LIB=libnadm_b200_half.so NADM_ENC_TS=1 timeout 300 python -m pytest tests -m gpu -x -q --timeout 60 > $OUT/pytest_half.log 2>&1; echo "pytest half rc=$?"; tail -3 $OUT/pytest_half.log
fi
for v in "libnadm_b200.so 0" "libnadm_b200.so 1" "libnadm_b200_half.so 1" "libnadm_b200.so 0" "libnadm_b200_half.so 1"; do
  set -- $v
  for w in fwd; do NADM_LIB=$1 NADM_ENC_TS=$2 timeout 60 python tools/enc_probe.py $w 500000 20000 2>&1 | tail -1 | sed "s/^/[$1 TS=$2] /"; done
  NADM_LIB=$1 NADM_ENC_TS=$2 timeout 120 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu --no-e2e > $OUT/ab_$1_ts$2.json 2> $OUT/ab_$1_ts$2.err
  python -c "import json;d=json.loads(open('$OUT/ab_$1_ts$2.json').read().strip().splitlines()[-1]);print('[$1 TS=$2] ms/step',round(d['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4),'infer',round(d['infer']['value']))" || echo "[$1 TS=$2] bench failed"
done
NADM_LIB=libnadm_b200.so timeout 60 python tools/enc_probe.py bwd 500000 20000 2>&1 | tail -1
# decoder with the CUDA-core decode skipped: floor of everything else in that kernel (results are garbage by design)
NADM_LIB=libnadm_b200_skip.so timeout 120 python bench.py --rows 20000 --steps 60 --warmup 5 --no-cpu --no-e2e > $OUT/ab_skip.json 2> $OUT/ab_skip.err
python -c "import json;d=json.loads(open('$OUT/ab_skip.json').read().strip().splitlines()[-1]);print('[skipdecode] ms/step',round(d['ms_per_step'],4),'dec',round(d['roofline']['ms_per_launch'],4))" || echo "skip bench failed"
NADM_ENC_TS=0 timeout 90 python tools/timeline_enc.py libnadm_b200_tl.so > $OUT/timeline_enc_ss.txt 2>&1; tail -3 $OUT/timeline_enc_ss.txt
NADM_ENC_TS=1 timeout 90 python tools/timeline_enc.py libnadm_b200_tl.so > $OUT/timeline_enc_ts.txt 2>&1; tail -3 $OUT/timeline_enc_ts.txt
NADM_ENC_TS=1 timeout 90 python tools/timeline_enc.py libnadm_b200_tl_half.so > $OUT/timeline_enc_ts_half.txt 2>&1; tail -3 $OUT/timeline_enc_ts_half.txt
timeout 150 python tools/step_breakdown.py --out $OUT/breakdown.json 2> $OUT/breakdown.err | tail -1
ls $OUT
