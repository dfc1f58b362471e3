#!/bin/bash
set -u
OUT=gpurun_out/r2c19; mkdir -p $OUT
for k in 4 5 7 9 10 11 12; do timeout 60 python tools/dec_probe.py 300000 8000 $k 800 1 2>&1 | tail -1; done
NADM_NO_GRAPH=1 CUDA_LAUNCH_BLOCKING=1 timeout 300 python bench.py --workload cfg4 --rows 8000 --steps 3 --warmup 1 --no-cpu --no-e2e > $OUT/bench_cfg4_dbg.json 2> $OUT/bench_cfg4_dbg.err; echo "rc=$?"; tail -25 $OUT/bench_cfg4_dbg.err
