#!/usr/bin/env python
"""Warm, in-pipeline duration of every library call of the training step (CUDA events around each call, all on the
step's stream), plus the total step time and the host time per step.  ncu's per-launch times are cold-cache and
serialised; this is the in-situ view.   python tools/step_breakdown.py [--rows N] [--steps K] [--workload cfg3]"""
import argparse
import json
import sys
import time
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from neural_admixture_b200 import ops  # noqa: E402
from neural_admixture_b200.model.neural_admixture import NeuralAdmixture  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=20000)
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--workload", default="cfg3")
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--out", default=None)
ap.add_argument("--snps", type=int, default=None, help="override M (e.g. 62500 = one rank's shard of cfg3 at 8 GPUs)")
a = ap.parse_args()
N, M, ks, B = bench.WORKLOADS[a.workload]
N = a.rows or N
M = a.snps or M
B = a.batch or B
dev = torch.device("cuda:0")
pg = bench.synth_packed(ops, N, M, 0, M, bench.SEED, dev)
V, P = bench.synth_init(M, 0, M, ks, bench.SEED, dev)
torch.manual_seed(bench.SEED)
k = ks[0] if len(ks) == 1 else None
na = NeuralAdmixture(k, 1, B, bench.LR, dev, bench.SEED, 1, True, "nadm_b200", None if k else min(ks), None if k else max(ks))
na.prepare(P, pg, bench.HIDDEN, bench.NCOMP, V, M, N)
order = torch.cat([na.epoch_order(N) for _ in range((a.steps + 20) * B // N + 2)]).to(dev)
losses = torch.zeros(a.steps + 20, device=dev)

records = defaultdict(list)
names = ["encoder_fwd", "mlp_fwd", "decoder_step", "mlp_bwd", "encoder_bwd"]
orig = {n: getattr(ops, n) for n in names}
timing = {"on": False}


def wrap(n):
    f = orig[n]

    def g(*args, **kw):
        if not timing["on"]:
            return f(*args, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f(*args, **kw)
        e1.record()
        records[n].append((e0, e1))
    return g


for n in names:
    setattr(ops, n, wrap(n))

na.train_steps(order, 10, True, first=0)
torch.cuda.synchronize()
# plain timing (no per-call events)
t0 = time.perf_counter()
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record()
na.train_steps(order, a.steps, True, first=10)
a1.record()
host_ms = (time.perf_counter() - t0) * 1e3 / a.steps
torch.cuda.synchronize()
step_ms = a0.elapsed_time(a1) / a.steps
timing["on"] = True
s0 = 10 + a.steps
for s in range(s0, s0 + 10):
    na._train_step(order[(s % (a.steps + 10)) * B:(s % (a.steps + 10) + 1) * B], None, losses[0:1])
torch.cuda.synchronize()
out = {"workload": a.workload, "rows": N, "B": B, "step_ms": step_ms, "host_ms_per_step": host_ms, "calls_us": {}}
tot = 0.0
for n in names:
    ts = [e0.elapsed_time(e1) * 1e3 for e0, e1 in records[n]]
    per_step = sum(ts) / 10
    out["calls_us"][n] = per_step
    tot += per_step
out["calls_sum_us"] = tot
print(json.dumps(out))
if a.out:
    Path(a.out).write_text(json.dumps(out, indent=1))
