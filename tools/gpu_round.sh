#!/bin/bash
# One GPU-box visit: parity tests, A/B bench legs, ncu launch list + one full capture.  Outputs under gpurun_out/.
# usage: tools/gpu_round.sh <tag> [steps...]   steps: test ab bench launches ncu_dec ncu_enc ncu_mlp
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
for step in "$@"; do
case $step in
test)
  timeout 240 python -m pytest tests -m gpu -x -q --timeout 60 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log; tail -5 $OUT/pytest.log;;
ab)
  for wg in 3 4; do export NADM_DEC_SLOTS=$wg;
    timeout 150 python bench.py --rows 20000 --steps 60 --warmup 5 --no-cpu --no-e2e > $OUT/ab_wg$wg.json 2> $OUT/ab_wg$wg.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/ab_wg$wg.json").read().strip().splitlines()[-1])
    print("WG=$wg ms/step", round(d["ms_per_step"],4), "dec ms", round(d["roofline"]["ms_per_launch"],4), "grad_only ms", round(d["grad_only"]["ms_per_step"],4))
except Exception as e:
    print("WG=$wg failed", e)
PY
  done; unset NADM_DEC_SLOTS;;
abpf)
  for lib in ${NADM_AB_LIBS:-libnadm_b200.so libnadm_b200_pf.so}; do
    NADM_LIB=$lib timeout 150 python bench.py --rows 20000 --steps 60 --warmup 5 --no-cpu --no-e2e > $OUT/ab_$lib.json 2> $OUT/ab_$lib.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/ab_$lib.json").read().strip().splitlines()[-1])
    print("$lib ms/step", round(d["ms_per_step"],4), "dec ms", round(d["roofline"]["ms_per_launch"],4), "grad_only ms", round(d["grad_only"]["ms_per_step"],4))
except Exception as e:
    print("$lib failed", e)
PY
  done;;
abpdl)
  for pdl in 0 1 0 1; do
    NADM_PDL=$pdl timeout 150 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu --no-e2e > $OUT/ab_pdl$pdl.json 2> $OUT/ab_pdl$pdl.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/ab_pdl$pdl.json").read().strip().splitlines()[-1])
    print("NADM_PDL=$pdl ms/step", round(d["ms_per_step"],4), "dec ms", round(d["roofline"]["ms_per_launch"],4), "grad_only ms", round(d["grad_only"]["ms_per_step"],4), "infer", round(d["infer"]["value"]))
except Exception as e:
    print("NADM_PDL=$pdl failed", e)
PY
  done;;
encts)
  NADM_ENC_TS=1 timeout 60 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/encts_smoke.txt 2>&1; rc=$?; tail -2 $OUT/encts_smoke.txt; echo "encts smoke rc=$rc"
  if [ $rc -eq 0 ]; then
    for ts in 0 1 0 1; do
      NADM_ENC_TS=$ts timeout 100 python bench.py --rows 20000 --steps 100 --warmup 5 --no-cpu --no-e2e > $OUT/ab_encts$ts.json 2> $OUT/ab_encts$ts.err
      python -c "import json;d=json.loads(open('$OUT/ab_encts$ts.json').read().strip().splitlines()[-1]);print('NADM_ENC_TS=$ts ms/step',round(d['ms_per_step'],4),'infer',round(d['infer']['value']))" || echo "NADM_ENC_TS=$ts failed"
    done
    NADM_ENC_TS=1 timeout 200 python -m pytest tests -m gpu -x -q --timeout 40 > $OUT/encts_pytest.log 2>&1; echo "encts pytest rc=$?"; tail -3 $OUT/encts_pytest.log
  fi;;
bench)
  timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 3000 $OUT/bench.json;;
launches)
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:nadm:: -c 400 --csv --log-file $OUT/launches.csv \
     python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/launches_bench.log 2>&1
  python tools/launch_summary.py $OUT/launches.csv | tee $OUT/launch_summary.txt;;
ncu_dec)
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:dec_tc_kernel --launch-skip 4 -c 1 \
     -o $OUT/dec_full -f python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_dec.log 2>&1; echo "ncu_dec rc=$?";;
ncu_enc)
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:enc_.*_tc_kernel --launch-skip 8 -c 2 \
     -o $OUT/enc_full -f python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_enc.log 2>&1; echo "ncu_enc rc=$?";;
ncu_mlp)
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:mlp_ --launch-skip 20 -c 5 \
     -o $OUT/mlp_full -f python bench.py --rows 20000 --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_mlp.log 2>&1; echo "ncu_mlp rc=$?";;
sharded)
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > $OUT/sharded.txt 2>&1; echo "sharded rc=$?"; tail -9 $OUT/sharded.txt
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 --no-e2e > $OUT/bench2.json 2> $OUT/bench2.err; echo "bench2 rc=$?"; tail -c 1500 $OUT/bench2.json; tail -3 $OUT/bench2.err
  NADM_NO_GRAPH=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 5 --no-e2e > $OUT/bench2_eager.json 2> $OUT/bench2_eager.err; echo "bench2 eager rc=$?"; python -c "import json;d=json.loads(open('$OUT/bench2_eager.json').read().strip().splitlines()[-1]);print('eager 2gpu ms/step',d['ms_per_step'])";;
shard2)
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > $OUT/sharded.txt 2>&1; echo "sharded rc=$?"; tail -9 $OUT/sharded.txt;;
exit2)
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 100 --warmup 5 --rows 20000 --no-e2e > $OUT/bench2q.json 2> $OUT/bench2q.err; echo "bench2 quick rc=$?"; python -c "import json;d=json.loads(open('$OUT/bench2q.json').read().strip().splitlines()[-1]);print('2gpu ms/step',d['ms_per_step'], d['step_launch'])";;
smoke)
  timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3;;
refarm)
  timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -c 900 $OUT/bench_ref.json;;
l2probe)
  for w in fwd bwd; do for n in 200 20000; do timeout 60 python tools/enc_probe.py $w 500000 $n 2>&1 | tail -1; done; done;;
sanitize)
  timeout 280 compute-sanitizer --tool memcheck --print-limit 20 python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -6 $OUT/memcheck.txt;;
exit4)
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --steps 100 --warmup 5 --rows 20000 --no-e2e > $OUT/bench4q.json 2> $OUT/bench4q.err; echo "bench4 quick rc=$?"; python -c "import json;d=json.loads(open('$OUT/bench4q.json').read().strip().splitlines()[-1]);print('4gpu ms/step',d['ms_per_step'], d['step_launch'], d['infer']['value'])";;
encprobe)
  for w in fwd bwd; do for m in 20000 100000; do NADM_ENC_ISSUERS=2 timeout 40 python tools/enc_probe.py $w $m 2>&1 | tail -1; echo "probe $w $m rc=$?"; done; done;;
timeline_enc)
  timeout 90 python tools/timeline_enc.py > $OUT/timeline_enc.txt 2>&1; tail -30 $OUT/timeline_enc.txt;;
timeline)
  timeout 90 python tools/timeline.py libnadm_b200_tl.so > $OUT/timeline.txt 2>&1; tail -8 $OUT/timeline.txt;;
breakdown)
  timeout 150 python tools/step_breakdown.py --out $OUT/breakdown.json 2> $OUT/breakdown.err | tail -1;;
esac
done
ls -la $OUT
