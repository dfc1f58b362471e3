#!/bin/bash
set -u
OUT=gpurun_out/r2c39; mkdir -p $OUT
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:nadm:: -c 300 --csv --log-file $OUT/launches.csv \
   python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/launches_bench.log 2>&1
python tools/launch_summary.py $OUT/launches.csv | tee $OUT/launch_summary.txt
