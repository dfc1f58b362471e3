#!/bin/bash
set -u
OUT=gpurun_out/r2c18; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest.log
for k in 8 9 12 16; do timeout 60 python tools/dec_probe.py 500000 20000 $k 800 1 2>&1 | tail -1; done
NADM_GENERIC=1 timeout 60 python tools/dec_probe.py 500000 20000 12 800 1 2>&1 | tail -1 | sed "s/^/[generic] /"
timeout 400 python bench.py --workload cfg4 --rows 8000 --steps 30 --warmup 3 --no-cpu --no-e2e > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err
python -c "import json;d=json.loads(open('$OUT/bench_cfg4.json').read().strip().splitlines()[-1]);print('cfg4 1gpu ms/step',round(d['ms_per_step'],4),'value',round(d['value']),'generic',d['generic_kernel_launches'],'step_frac',round(d['roofline']['step_frac'],3))" || tail -5 $OUT/bench_cfg4.err
timeout 300 python bench.py --workload cfg2 --steps 200 --warmup 5 --no-cpu > $OUT/bench_cfg2.json 2> $OUT/bench_cfg2.err
python -c "import json;d=json.loads(open('$OUT/bench_cfg2.json').read().strip().splitlines()[-1]);print('cfg2 ms/step',round(d['ms_per_step'],4),'value',round(d['value']),'e2e',round(d['e2e']['value']),'step_frac',round(d['roofline']['step_frac'],3))" || tail -5 $OUT/bench_cfg2.err
