#!/bin/bash
set -u
OUT=gpurun_out/r2c20; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $OUT/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
timeout 300 python bench.py --workload cfg4 --rows 8000 --steps 30 --warmup 3 --no-cpu > $OUT/bench_cfg4.json 2> $OUT/bench_cfg4.err; echo "cfg4 rc=$?"; tail -3 $OUT/bench_cfg4.err
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default rc=$?"; tail -3 $OUT/bench_default.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; tail -3 $OUT/bench_reference.err
timeout 300 python bench.py --impl reference-cuda --steps 10 --warmup 3 > $OUT/bench_reference_cuda.json 2> $OUT/bench_reference_cuda.err; echo "refcuda rc=$?"; tail -3 $OUT/bench_reference_cuda.err
timeout 300 python bench.py --workload cfg2 --no-cpu > $OUT/bench_cfg2.json 2> $OUT/bench_cfg2.err; echo "cfg2 rc=$?"; tail -3 $OUT/bench_cfg2.err
timeout 300 python bench.py --workload cfg5 --no-cpu > $OUT/bench_cfg5.json 2> $OUT/bench_cfg5.err; echo "cfg5 rc=$?"; tail -3 $OUT/bench_cfg5.err
for f in cfg4 default reference reference_cuda cfg2 cfg5; do echo "== $f"; head -c 900 $OUT/bench_$f.json; echo; done
