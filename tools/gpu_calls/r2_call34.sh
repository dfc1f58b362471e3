#!/bin/bash
set -u
OUT=gpurun_out/r2c34; mkdir -p $OUT
for late in 0 1; do for loss in 1 0; do timeout 90 python tools/dec_probe.py 500000 20000 8 800 $loss $late 2>&1 | tail -1; done; done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dec_tc_kernel --launch-skip 5 -c 1 -o $OUT/dec_late -f python tools/dec_probe.py 500000 20000 8 800 1 1 > $OUT/ncu_late.log 2>&1; echo "ncu late rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dec_tc_kernel --launch-skip 5 -c 1 -o $OUT/dec_noloss -f python tools/dec_probe.py 500000 20000 8 800 0 0 > $OUT/ncu_noloss.log 2>&1; echo "ncu noloss rc=$?"
