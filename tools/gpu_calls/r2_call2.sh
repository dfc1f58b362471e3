#!/bin/bash
# round 2, GPU visit 2: knock-out builds — where does the time of the three tensor-core kernels go?
set -u
OUT=gpurun_out/r2c2; mkdir -p $OUT
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== smoke of the default build (sanity of this round's edits)"
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
for lib in libnadm_b200.so libnadm_ko_widen.so libnadm_ko_load.so libnadm_ko_mma.so libnadm_ko_wl.so libnadm_ko_wlm.so; do
  for ts in 0 1; do
    NADM_LIB=$lib NADM_ENC_TS=$ts timeout 60 python tools/enc_probe.py fwd 500000 20000 2>&1 | tail -1 | sed "s/^/[$lib TS=$ts] /"
  done
  NADM_LIB=$lib timeout 60 python tools/enc_probe.py bwd 500000 20000 2>&1 | tail -1 | sed "s/^/[$lib] /"
done
for lib in libnadm_b200.so libnadm_b200_skip.so libnadm_ko_mma.so; do
  for loss in 1 0; do
    NADM_LIB=$lib timeout 60 python tools/dec_probe.py 500000 20000 8 800 $loss 2>&1 | tail -1 | sed "s/^/[$lib] /"
  done
done
for B in 768 896 1024; do NADM_LIB=libnadm_b200.so timeout 60 python tools/dec_probe.py 500000 20000 8 $B 1 2>&1 | tail -1; done
} 2>&1 | tee $OUT/knockouts.txt
