#!/bin/bash
set -u
OUT=gpurun_out/r2c38; mkdir -p $OUT
timeout 100 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
timeout 60 python bench.py --rows 20000 --steps 300 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]); print('ms', round(d['ms_per_step'],4), 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], 'grad_only', d['grad_only']['ms_per_step'])"
