#!/bin/bash
set -u
OUT=gpurun_out/r2c3; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $OUT/gpu.txt
timeout 300 tests/cuda/gather_probe.bin 20000 2>&1 | tee $OUT/gather_probe.txt
echo "--- 900-row matrix (112 MB: the gathered rows mostly stay in the 126 MB L2) ---" | tee -a $OUT/gather_probe.txt
timeout 300 tests/cuda/gather_probe.bin 900 2>&1 | tee -a $OUT/gather_probe.txt
