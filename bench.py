#!/usr/bin/env python
"""bench.py — training samples/sec of the Neural ADMIXTURE hot path on synthetic N x M genotype matrices.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg3|cfg2|...]

One "step" = one minibatch (B = 800 samples, the reference default) through the whole fused path: encoder X.V ->
RMSNorm/MLP/softmax -> fused decoder + BCE loss + backward + Adam + clamp on P -> MLP backward + Adam -> encoder
backward + Adam on V.  Workload at every N: BASELINE.json configs[2] = 100k samples x 500k SNPs, K = 8 (it fits one
B200: 12.5 GB packed); for N > 1 the SNP axis is sharded across the ranks (strong scaling: total work fixed) and the
B x C partial projection and the B x K dQ (+ loss) are all-reduced over NCCL every step.

The JSON line carries: value (HBM-resident whole-job samples/s), e2e (same metric with every step's minibatch
streamed from pinned HOST memory and its loss read back), roofline (dominant kernel: the fused decoder step),
cpu_baseline (the reference's CPU path — its torch op sequence, oracle/nadm_torch_port.py — on this box's cores, on
a bounded sample), clocks, gpu_launches.  `--impl reference` prints the CPU arm alone.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (N samples, M SNPs, ks, batch)   — BASELINE.json configs[1..4]
    "cfg2": (10_000, 100_000, [8], 800),
    "cfg3": (100_000, 500_000, [8], 800),
    "cfg4": (50_000, 300_000, list(range(4, 13)), 800),
    "cfg5": (1_000_000, 500_000, [8], 2048),      # projective inference (use --rows to bound N on one GPU)
}
METRICS = {"cfg2": "training samples/sec at N=10k x M=100k SNPs K=8",
           "cfg3": "training samples/sec at N=100k x M=500k SNPs K=8",
           "cfg4": "training samples/sec at N=50k x M=300k SNPs, heads K=4..12",
           "cfg5": "inference samples/sec at N=1M x M=500k SNPs K=8"}
METRIC = METRICS["cfg3"]
UNIT = "samples/s"
HIDDEN, NCOMP, LR, SEED, MISSING = 1024, 8, 2e-3, 42, 0.005


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------------
# synthetic admixture-model genotypes
# ------------------------------------------------------------------------------------------------------------------
GEN_BLOCKS = 64   # the SNP axis is generated in 64 fixed column blocks, each from its own seeded stream: a rank that owns
                  # a slice of the axis generates exactly the columns a single GPU would (world sizes that divide 64)


def gen_bounds(M):
    blocks = (M + 63) // 64
    return [min(M, (blocks * j) // GEN_BLOCKS * 64) for j in range(GEN_BLOCKS + 1)]


def synth_params(N, M, K_true, seed, device):
    """Sample ancestries Q_true (N x K) and allele frequencies P_true (K x M, full SNP axis): the same on every rank."""
    import torch
    gq = torch.Generator(device="cpu").manual_seed(seed)
    alpha = torch.full((K_true,), 0.2)
    gam = torch._standard_gamma(alpha.expand(N, K_true).contiguous(), generator=gq)
    Qt = (gam / gam.sum(1, keepdim=True).clamp_min(1e-30)).to(device)
    gp = torch.Generator(device=device).manual_seed(seed * 1000 + 17)
    anc = torch.rand((1, M), device=device, generator=gp) * 0.45 + 0.03      # minor-allele oriented
    Pt = (anc + 0.12 * torch.randn((K_true, M), device=device, generator=gp)).clamp_(0.01, 0.99)
    return Qt, Pt


def synth_rows(Qt_rows, Pt, gen, missing=MISSING):
    """uint8 codes {0,1,2,3}: g ~ Binomial(2, Q_true P_true), missing rate `missing` coded 3."""
    import torch
    prob = Qt_rows @ Pt
    g = (torch.rand(prob.shape, device=prob.device, generator=gen) < prob).to(torch.uint8)
    g += (torch.rand(prob.shape, device=prob.device, generator=gen) < prob).to(torch.uint8)
    g[torch.rand(prob.shape, device=prob.device, generator=gen) < missing] = 3
    return g


def synth_packed(ops, N, M, c0, c1, seed, device, chunk=2048):
    """2-bit packed N x (c1 - c0) slice of the synthetic N x M matrix.  Column block j of GEN_BLOCKS draws from its own
    generator, so the genotypes of a column do not depend on how the SNP axis is sharded."""
    import torch
    Qt, Pt = synth_params(N, M, 8, seed, device)
    gb = gen_bounds(M)
    mine = [j for j in range(GEN_BLOCKS) if gb[j] < gb[j + 1] and gb[j] >= c0 and gb[j + 1] <= c1]
    assert sum(gb[j + 1] - gb[j] for j in mine) == c1 - c0, "SNP slice is not a union of generation blocks"
    gens = {j: torch.Generator(device=device).manual_seed(seed * 100003 + j) for j in mine}
    pg = ops.PackedGenotypes.empty(N, c1 - c0, device)
    codes = torch.empty((min(chunk, N), c1 - c0), dtype=torch.uint8, device=device)
    for r0 in range(0, N, chunk):
        r1 = min(N, r0 + chunk)
        for j in mine:
            codes[: r1 - r0, gb[j] - c0:gb[j + 1] - c0] = synth_rows(Qt[r0:r1], Pt[:, gb[j]:gb[j + 1]], gens[j])
        ops.pack2bit(codes[: r1 - r0], pg.storage[r0:r1], c1 - c0)
    return pg


def synth_init(M, c0, c1, ks, seed, device):
    """V: orthonormalised N(0,1) columns (M x C); P_init ~ U(0.05, 0.95) (sum K x M) — RSVD + GMM initialisation is
    bypassed at this size.  Generated for the full SNP axis, then sliced to [c0, c1)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed * 7 + 3)
    V = torch.linalg.qr(torch.randn((M, NCOMP), device=device, generator=g))[0][c0:c1].contiguous()
    P = (torch.rand((sum(ks), M), device=device, generator=g) * 0.9 + 0.05)[:, c0:c1].contiguous()
    return V, P


def snp_slice(M, rank, world, align=64):
    blocks = (M + align - 1) // align
    return min(M, (blocks * rank) // world * align), min(M, (blocks * (rank + 1)) // world * align)


# ------------------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(index)], stdout=subprocess.PIPE, text=True,
                                         stderr=subprocess.DEVNULL)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        self.t_start = self.t_stop = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_start(self):
        self.t_start = time.time()

    def mark_stop(self):
        self.t_stop = time.time()

    def summary(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if self.t_start - 0.05 <= t <= self.t_stop + 0.15 and len(r) >= 7]
        if not rows:
            rows = [r for _, r in self.rows[-3:] if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return float("nan")
        return {"sm_mhz": statistics.median(num(r[0]) for r in rows), "sm_max_mhz": num(rows[0][1]),
                "power_w_max": max(num(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------------------------
# reference arms: the reference's torch op sequence (oracle/nadm_torch_port.py) on the host cores, or on cuda:0
# ------------------------------------------------------------------------------------------------------------------
def _port_problem(ks, Ms, rows, gen_dev, data_dev):
    """A rows x Ms uint8 problem + TorchPort with the bench's synthetic generator (its own bounded sample)."""
    import torch
    sys.path.insert(0, str(ROOT / "oracle"))
    from nadm_torch_port import TorchPort
    Qt, Pt = synth_params(rows, Ms, 8, SEED, gen_dev)
    gen = torch.Generator(device=gen_dev).manual_seed(SEED * 100003)
    data = torch.empty((rows, Ms), dtype=torch.uint8, device=data_dev)
    for r0 in range(0, rows, 256):
        data[r0:r0 + 256] = synth_rows(Qt[r0:r0 + 256], Pt, gen).to(data_dev)
    V, P = synth_init(Ms, 0, Ms, ks, SEED, gen_dev)
    Ps, off = [], 0
    for k in ks:
        Ps.append(P[off:off + k].T.contiguous().to(data_dev))
        off += k
    return data, TorchPort(V.to(data_dev), Ps, HIDDEN, lr=LR, seed=SEED, as_shipped=True, device=data_dev)


def cpu_reference(wl, ks, M, B, steps, warmup, budget_s, as_line=False, n_gpus=1):
    """EXACTLY `steps` timed steps after `warmup` untimed ones; what is bounded is the SNP width of each step (a
    subsample of the workload's columns, step time scaled by M / Ms: the op sequence is linear in M)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen_dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    cpu = torch.device("cpu")
    rows = 4 * B                                           # step throughput does not depend on N

    def run(data, port, n):
        g = torch.Generator().manual_seed(SEED)
        ts = []
        for _ in range(n):
            idx = torch.randperm(rows, generator=g)[:B]
            t0 = time.perf_counter()
            port.step(port.gather(data, idx))
            ts.append(time.perf_counter() - t0)
        return ts

    # probe at M/20 to size the sample: cost is linear in M
    probe_M = max(1024, M // 20 // 64 * 64)
    data, port = _port_problem(ks, probe_M, rows, gen_dev, cpu)
    run(data, port, 1)
    t_probe = min(run(data, port, 2))
    est_full = t_probe * M / probe_M
    frac = min(1.0, budget_s / (est_full * (steps + warmup)))
    Ms = M if frac >= 1.0 else max(1024, int(M * frac) // 64 * 64)
    if Ms != probe_M:
        del data, port
        data, port = _port_problem(ks, Ms, rows, gen_dev, cpu)
    run(data, port, warmup)
    ts = run(data, port, steps)
    t_step_full = (sum(ts) / len(ts)) * (M / Ms)
    value = B / t_step_full
    sample = (f"{steps} timed steps (+{warmup} warm-up) of B={B} on a {rows} x {Ms} host uint8 matrix"
              + ("" if Ms == M else f" (SNP subsample {Ms}/{M}; step time scaled by M/Ms, cost is linear in M)")
              + "; reference op sequence as shipped (matmul precision 'medium', fused Adam, per-row gather)")
    cb = {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    if not as_line:
        return cb
    return {"impl": "reference", "metric": METRICS[wl], "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": t_step_full * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl}: {WORKLOADS[wl][0]} samples x {M} SNPs, K={ks}, B={B}, CPU",
                       "hidden": HIDDEN, "n_components": NCOMP},
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def cuda_reference(wl, ks, M, B, steps, warmup):
    """Second stated baseline (`--impl reference-cuda`): the reference's eager op sequence on cuda:0 — per step the
    minibatch's 2-bit rows are unpacked to a B x M uint8 tensor (reference :404-406; torch bit ops stand in for its
    pack2bit.cu kernel, which cannot travel to the GPU box), then Q_P forward, BCELoss(sum), backward, fused Adam,
    restrict_P, loss.item() — precision as shipped ('medium').  N is truncated to 4 B rows (step cost does not depend
    on N); SNP width bounded by the memory the eager path needs (about 40 bytes per (row, SNP) element)."""
    import torch
    dev = torch.device("cuda:0")
    rows = 4 * B
    free = torch.cuda.mem_get_info(dev)[0]
    Ms = min(M, int(free * 0.5 / (B * 44)) // 64 * 64)
    data, port = _port_problem(ks, Ms, rows, dev, dev)
    shifts = torch.tensor([0, 2, 4, 6], dtype=torch.uint8, device=dev)
    packed = torch.zeros((rows, (Ms + 3) // 4), dtype=torch.uint8, device=dev)
    pad = (-Ms) % 4
    d4 = torch.nn.functional.pad(data, (0, pad)).view(rows, -1, 4)
    packed = (d4[:, :, 0] | (d4[:, :, 1] << 2) | (d4[:, :, 2] << 4) | (d4[:, :, 3] << 6)).contiguous()
    del data, d4

    def one(idx):
        pb = packed[idx]
        g = ((pb.unsqueeze(-1) >> shifts) & 3).view(B, -1)[:, :Ms]
        return port.step(g)

    gcpu = torch.Generator().manual_seed(SEED)
    for _ in range(warmup):
        one(torch.randperm(rows, generator=gcpu)[:B].to(dev))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one(torch.randperm(rows, generator=gcpu)[:B].to(dev))
    e1.record()
    torch.cuda.synchronize()
    t_step = e0.elapsed_time(e1) * 1e-3 / steps * (M / Ms)
    value = B / t_step
    sample = (f"{steps} timed steps (+{warmup} warm-up) of B={B} on a {rows} x {Ms} packed device matrix"
              + ("" if Ms == M else f" (SNP subsample {Ms}/{M}: the eager path needs ~44 B of device memory per (row, SNP); "
                                    "step time scaled by M/Ms)")
              + "; reference eager op sequence on cuda:0, matmul precision 'medium' as shipped, loss.item() every step")
    return {"impl": "reference-cuda", "metric": METRICS[wl], "value": value, "unit": UNIT, "n_gpus": 1, "steps": steps,
            "warmup": warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (tf32/bf16 matmuls: 'medium')", "data": "synthetic",
            "config": {"workload": f"{wl}: {WORKLOADS[wl][0]} samples x {M} SNPs, K={ks}, B={B}, eager torch on cuda:0",
                       "hidden": HIDDEN, "n_components": NCOMP},
            "cpu_baseline": None, "baseline": {"kind": "port-on-cuda", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ------------------------------------------------------------------------------------------------------------------
# committed ncu digests: only used when they were captured for THIS configuration
# ------------------------------------------------------------------------------------------------------------------
def ncu_digest(M_loc, B, ks):
    """(dram bytes per launch, warp instructions per launch) of dec_tc_kernel from profiles/*ncu_summary.json whose
    recorded config equals this run's (M_loc, B, heads); (None, None) otherwise — never a number from another shape."""
    for name in ("r2_final_ncu_summary.json", "r2_ncu_summary.json"):
        f = ROOT / "profiles" / name
        if not f.exists():
            continue
        try:
            summ = json.loads(f.read_text())
            cfg = summ.get("_config", {})
            if cfg.get("M_loc") == M_loc and cfg.get("B") == B and cfg.get("ks") == list(ks):
                v = next(v for k, v in summ.items() if k.startswith("dec_tc_kernel"))
                return v.get("dram_bytes"), v.get("warp_instructions")
        except Exception:
            pass
    return None, None


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--rows", type=int, default=None, help="override N (debug)")
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--breakdown", action="store_true",
                    help="add `calls_us`: in-pipeline CUDA-event time of every library call of a step (eager launches "
                         "after the timed region), on rank 0 and as the max over ranks: how a rank's step is apportioned")
    args = ap.parse_args()
    assert args.warmup >= 0 and args.steps >= 1

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload
    N, M, ks, B = WORKLOADS[wl]
    B = args.batch or B
    N = args.rows or N

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(cpu_reference(wl, ks, M, B, args.steps, args.warmup, 150.0, as_line=True, n_gpus=args.gpus)),
                  flush=True)
        return
    if args.impl == "reference-cuda":
        if rank == 0:
            print(json.dumps(cuda_reference(wl, ks, M, B, min(args.steps, 50), max(args.warmup, 3))), flush=True)
        return

    import torch
    import torch.distributed as dist
    from neural_admixture_b200 import ops
    from neural_admixture_b200.model.neural_admixture import NeuralAdmixture

    assert torch.cuda.is_available(), "bench.py (impl b200) needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    assert GEN_BLOCKS % world == 0, "world size must divide 64 (generation blocks of the synthetic SNP axis)"

    if wl == "cfg5":
        return infer_bench(args, ops, dist, rank, world, dev, N, M, ks)

    c0, c1 = snp_slice(M, rank, world)
    M_loc = c1 - c0
    pg = synth_packed(ops, N, M, c0, c1, SEED, dev)
    V, P = synth_init(M, c0, c1, ks, SEED, dev)
    torch.manual_seed(SEED)
    k = ks[0] if len(ks) == 1 else None
    na = NeuralAdmixture(k, 1, B, LR, dev, SEED, world, rank == 0, "nadm_b200", None if k else min(ks),
                         None if k else max(ks))
    na.prepare(P, pg, HIDDEN, NCOMP, V, M_loc, N)
    del P
    sumK = sum(ks)

    total = args.warmup + args.steps
    orders = []
    while sum(o.numel() // B for o in orders) < total + 8:
        o = na.epoch_order(N)
        orders.append(o[: (o.numel() // B) * B])               # full batches only inside the timed region
    order = torch.cat(orders).to(dev)
    losses = torch.zeros(total + 8, dtype=torch.float32, device=dev)

    def run_steps(s0, n):
        # the public step loop: CUDA-graph replayed steps (NADM_NO_GRAPH=1: eager launches), loss evaluated every step
        losses[s0:s0 + n].copy_(na.train_steps(order, n, True, first=s0))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- HBM-resident timing ------------------------------------------------------------------------------------
    run_steps(0, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    sync_all()
    l0 = ops.launch_count() + na.graph_kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.mark_start()
    ev0.record()
    run_steps(args.warmup, args.steps)
    ev1.record()
    sync_all()
    if sampler:
        sampler.mark_stop()
    launches = ops.launch_count() + na.graph_kernel_launches - l0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    # same loop without evaluating the reconstruction loss (what NeuralAdmixture does on epochs whose loss the reference
    # does not print: 4 of 5 epochs) — reported next to the headline, which evaluates the loss on every step
    n_go = min(args.steps, 40)
    na.train_steps(order, 1, False, first=args.warmup)           # (captures the loss-free step graph outside the timing)
    sync_all()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    na.train_steps(order, n_go, False, first=args.warmup)
    g1.record()
    sync_all()
    ms_go = max_over_ranks(g0.elapsed_time(g1))
    ms_step = ms_total / args.steps
    value = B * args.steps / (ms_total * 1e-3)
    loss_first = float(losses[args.warmup].item())
    loss_last = float(losses[args.warmup + args.steps - 1].item())
    clocks = sampler.summary() if sampler else None

    # ---- dominant kernel (fused decoder step) timed alone on its stream, same inputs ------------------------------
    peak, peak_src = peaks()
    hyper = ops.adam_hyper(LR, 10_000)                              # late-training Adam coefficients
    fb = na.raw_model._fwd_buffers(B)
    sb = na._step_buffers(B)
    s_extra = total

    def time_decoder(Qbuf, Pl, Pm, Pv, hyp, iters=6):
        out = []
        for it in range(iters):
            idx = order[(s_extra + it) * B:(s_extra + it + 1) * B]
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            off = 0
            for i, kk in enumerate(ks):
                ops.decoder_step(pg, Qbuf, sb["dQ"], off, kk, Pl[i], Pm[i], Pv[i], hyp, sb["loss"], fb["ws"], row_idx=idx)
                off += kk
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                out.append(a.elapsed_time(b))
        return sum(out) / len(out)

    dec_t = time_decoder(fb["Q"], [d.weight.data for d in na.raw_model.decoders.decoders], na.optimizer.m["P"],
                         na.optimizer.v["P"], hyper)
    pitch_bytes = (M_loc + 3) // 4
    dec_bytes = len(ks) * B * pitch_bytes + 24 * M_loc * sumK       # genotype pass per head + {P,m,v} read+write
    step_bytes = (2 + len(ks)) * B * pitch_bytes + 24 * M_loc * (NCOMP + sumK)
    traffic, warp_instr = ncu_digest(M_loc, B, ks)
    sm_hz = (clocks or {}).get("sm_mhz") or 0
    issue_floor_ms = (warp_instr / (148 * 4) / (sm_hz * 1e6) * 1e3) if (warp_instr and sm_hz) else None
    roofline = {"bound": "hbm", "kernel": "dec_tc_kernel (tcgen05 fused decoder: Q.P^T + BCE + backward + dQ/dP + Adam + clamp)",
                "achieved": dec_bytes / (dec_t * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": dec_bytes / (dec_t * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dec_bytes, "ms_per_launch": dec_t,
                "traffic_source": ("dram__bytes_read.sum + dram__bytes_write.sum of one launch at this configuration, "
                                   "profiles/r2_final_ncu_summary.json") if traffic else
                                  "null: no ncu --set full capture committed for this (M_loc, B, heads)",
                "issue_floor_ms": issue_floor_ms,
                "issue_floor_what": "warp instructions of one launch (ncu capture, this configuration) / (148 SMs x 4 "
                                    "schedulers) / SM clock sampled in this run: the kernel is bound by instruction "
                                    "issue on the SM, not by HBM (DESIGN.md section 4)" if issue_floor_ms else None,
                "step_frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak,
                "step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9,
                         "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak, "ms": ms_step}}

    # ---- the decoder on LATE-TRAINING inputs: exact 0 / 1 allele frequencies and concentrated Q send 16-SNP groups to
    # the general (clamp / floor / one log per element) path instead of the fast one the early steps above take -----
    gl = torch.Generator(device=dev).manual_seed(SEED + 99)
    P_late, zeros = [], []
    for kk in ks:
        Pl = torch.rand((M_loc, kk), device=dev, generator=gl) * 0.9 + 0.05
        u = torch.rand((M_loc, kk), device=dev, generator=gl)
        Pl[u < 0.2] = 0.0
        Pl[u > 0.8] = 1.0
        P_late.append(Pl.contiguous())
        zeros.append(torch.zeros_like(Pl))
    gq = torch.Generator(device="cpu").manual_seed(SEED + 98)
    gam = torch._standard_gamma(torch.full((B, sumK), 0.05), generator=gq).clamp_min(1e-30)
    Q_late = torch.empty((B, sumK), device=dev)
    off = 0
    for kk in ks:
        blk = gam[:, off:off + kk]
        Q_late[:, off:off + kk] = (blk / blk.sum(1, keepdim=True)).to(dev)
        off += kk
    late_t = time_decoder(Q_late, P_late, zeros, [z.clone() for z in zeros], ops.adam_hyper(0.0, 10_000))
    ms_ = min(M_loc, 65536) // 16 * 16
    R = Q_late[:, :ks[0]] @ P_late[0][:ms_].T                        # head 0, a 65536-SNP sample, torch fp32
    prod = R * (1 - R)
    general = ((prod < 2.0 ** -15) | (R > 1)).view(B, ms_ // 16, 16).any(dim=2).float().mean().item()
    late = {"ms_per_launch": late_t, "vs_early": late_t / dec_t, "general_path_fraction": general,
            "what": "same kernel, same shapes; P: 20 % exact 0, 20 % exact 1, rest U(.05,.95); Q ~ Dirichlet(0.05); "
                    "lr = 0 so that P stays put.  general_path_fraction: share of (row, 16-SNP group)s with "
                    "min R(1-R) < 2^-15, estimated with torch on head 0 and a 65536-SNP sample"}
    del P_late, zeros, Q_late, R, prod

    # ---- forward-only half of the path (post-training Q pass / `infer`, BASELINE configs[4]): Q for consecutive rows ----
    n_inf = min(N, 16384)
    inf_pg = ops.PackedGenotypes(pg.storage[:n_inf], n_inf, M_loc)
    na.raw_model.infer_packed(inf_pg, 2048, **na.comm())
    sync_all()
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for _ in range(3):
        na.raw_model.infer_packed(inf_pg, 2048, **na.comm())
    i1.record()
    sync_all()
    t_inf = max_over_ranks(i0.elapsed_time(i1))
    infer = {"value": 3 * n_inf / (t_inf * 1e-3), "unit": UNIT, "rows": n_inf, "batch": 2048,
             "algorithmic_gbs": 3 * n_inf * pitch_bytes / (t_inf * 1e-3) / 1e9,
             "what": "Q_P.infer_packed on this run's SNP-sharded training matrix (post-training Q pass: one all-reduce of "
                     "Z per batch when sharded); `--workload cfg5` is the inference benchmark proper (sample-sharded)"}

    # ---- end to end: minibatches streamed from pinned host memory, loss read back every step ---------------------
    e2e, host = None, None
    if not args.no_e2e:
        nb = 8
        host = [pg.storage[order[(total - 1 - j) * B:(total - j) * B]].cpu().pin_memory() for j in range(nb)]
        n_e2e = max(8, min(args.steps, 40))
        na.train_from_host([host[j % nb] for j in range(3)])
        sync_all()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hl = na.train_from_host([host[j % nb] for j in range(n_e2e)])
        e1.record()
        sync_all()
        t = max_over_ranks(e0.elapsed_time(e1))
        e2e = {"value": B * n_e2e / (t * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(host[0].numel()),
               "d2h_bytes_per_step": 4, "steps": n_e2e, "ms_per_step": t / n_e2e,
               "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / n_e2e,
               "what": "NeuralAdmixture.train_from_host: per step the minibatch's 2-bit packed rows are copied from "
                       "pinned host memory (double-buffered side stream), the fused step runs, loss read back"}
        assert all(math.isfinite(x) for x in hl)

    # ---- optional: how the step is apportioned (eager launches with an event pair around every library call) --------
    calls_us = None
    if args.breakdown:
        from collections import defaultdict
        names = ["encoder_fwd", "mlp_fwd", "decoder_step", "mlp_bwd", "encoder_bwd"]
        orig = {n_: getattr(ops, n_) for n_ in names}
        rec = defaultdict(list)

        def wrap(n_):
            f = orig[n_]

            def g(*a, **kw):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                f(*a, **kw)
                ev1.record()
                rec[n_].append((ev0, ev1))
            return g

        for n_ in names:
            setattr(ops, n_, wrap(n_))
        lb = torch.zeros(1, device=dev)
        for s_ in range(24):
            na._train_step(order[s_ * B:(s_ + 1) * B], None, lb)
        sync_all()
        for n_ in names:
            setattr(ops, n_, orig[n_])
        mine = {n_: sum(a.elapsed_time(b) for a, b in v[4:]) * 1e3 / max(1, len(v) - 4) for n_, v in rec.items()}
        calls_us = {"rank0": {k_: round(mine[k_], 1) for k_ in names},
                    "max_over_ranks": {k_: round(max_over_ranks(mine[k_]), 1) for k_ in names},
                    "what": "eager launches, CUDA events around each library call on the step's stream (20 steps); "
                            "mlp_fwd / mlp_bwd include the wait for the peers inside the fused exchange; the graph-"
                            "replayed step (ms_per_step) has none of the host launch gaps these calls see"}

    step_launch = "cuda-graph replay (one graph launch per step)" if na.use_graph else "eager launches"
    exchange = None if world == 1 else (
        "fused into the MLP kernels: 8-byte {value, exchange number} stores over NVLink into peer-mapped exchange areas, "
        "no fence, no flag, no NCCL kernel in the step" if na.exchange is not None else "2 NCCL all-reduces per step (inside the replayed graph)")
    generic = int(na.generic_kernel_launches)
    fallback = na.graph_fallback
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        host = None
        del pg, na
        torch.cuda.empty_cache()
        cpu = cpu_reference(wl, ks, M, B, 3, 1, args.cpu_budget)

    if rank == 0:
        line = {"metric": METRICS[wl], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{wl}: {N} samples x {M} SNPs, heads K={ks}, batch B={B}, "
                                       f"SNP axis sharded over {world} GPU(s) ({M_loc} SNPs on rank 0)",
                           "hidden": HIDDEN, "n_components": NCOMP, "lr": LR, "missing_rate": MISSING,
                           "l2": f"no flush: every step gathers a fresh random minibatch ({B * pitch_bytes / 1e6:.0f} MB "
                                 f"of rows out of {N * pitch_bytes / 1e9:.2f} GB) and streams "
                                 f"{24 * M_loc * (NCOMP + sumK) / 1e6:.0f} MB of parameters + Adam state "
                                 "(vs 126 MB L2)",
                           "generator": "column blocks with their own seeds: identical genotypes, V and P_init at "
                                        "every world size (loss.first/last_timed_step agree across N)"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "loss": {"first_timed_step": loss_first, "last_timed_step": loss_last,
                                           "schedule": "evaluated on every timed step, as the reference does"},
                "step_launch": step_launch, "graph_fallback": fallback, "generic_kernel_launches": generic,
                "exchange": exchange, "calls_us": calls_us,
                "infer": infer, "late_training": late,
                "grad_only": {"value": B * n_go / (ms_go * 1e-3), "unit": UNIT, "steps": n_go,
                              "ms_per_step": ms_go / n_go,
                              "what": "same step with the loss value not evaluated (loss pointer NULL)"}}
        print(json.dumps(line), flush=True)
    teardown(dist, world, locals().get("na"))


def teardown(dist, world, na):
    """Release order that lets a sharded run exit cleanly: the captured step graphs hold NCCL kernels, so they go first
    (NeuralAdmixture.release_graphs), then every rank drains its device, meets at a barrier and destroys the process
    group.  A failsafe timer only guards against a wedged teardown (it says so on stderr if it ever fires)."""
    if world <= 1:
        return
    import gc
    import torch

    def failsafe():
        sys.stderr.write("bench.py: teardown did not finish within 60 s; forcing exit (the JSON line is already out)\n")
        sys.stderr.flush()
        os._exit(0)
    wd = threading.Timer(60.0, failsafe)
    wd.daemon = True
    wd.start()
    if na is not None:
        na.close_exchange()
        na.release_graphs()
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    wd.cancel()


def infer_bench(args, ops, dist, rank, world, dev, N, M, ks):
    """`--workload cfg5` (BASELINE configs[4]): projective inference, N samples x M SNPs through a trained encoder.
    Samples are independent, so the path shards on the SAMPLE axis with no data-path collective at all: every rank
    holds the full V / MLP (16 MB) and its own rows of the packed matrix; value = rows of all ranks / max time."""
    import torch
    from neural_admixture_b200.model.neural_admixture import Q_P
    n_loc = (N + world - 1) // world
    r0 = min(N, rank * n_loc)
    n_loc = min(n_loc, N - r0)
    # this rank's rows of the synthetic matrix (rows are i.i.d. given the generator: a per-rank seed is enough here)
    pg = synth_packed(ops, n_loc, M, 0, M, SEED + 1000 * rank, dev)
    V, _ = synth_init(M, 0, M, [ks[0]], SEED, dev)
    torch.manual_seed(SEED)
    model = Q_P(HIDDEN, NCOMP, ks_list=ks, V=V, is_train=False).to(dev)
    model.bind()
    batch = 2048

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    warm = ops.PackedGenotypes(pg.storage[: min(n_loc, 4 * batch)], min(n_loc, 4 * batch), M)
    for _ in range(max(1, min(args.warmup, 3))):
        model.infer_packed(warm, batch)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) if rank == 0 else None
    sync_all()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.mark_start()
    e0.record()
    passes = max(1, min(args.steps, 3))
    for _ in range(passes):
        Qs = model.infer_packed(pg, batch)
    e1.record()
    sync_all()
    if sampler:
        sampler.mark_stop()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / passes
    peak, peak_src = peaks()
    pitch_bytes = (M + 3) // 4
    bytes_rank = n_loc * pitch_bytes + 4 * M * NCOMP * ((n_loc + batch - 1) // batch)
    # end to end: rows arrive from pinned host memory in 2048-row batches, Q goes back to the host
    nb = min(n_loc, 8 * batch)
    host_rows = pg.storage[:nb].cpu().pin_memory()
    q_host = torch.empty((nb, ks[0]), dtype=torch.float32).pin_memory()
    stage = ops.PackedGenotypes(torch.empty((batch, pg.pitch), dtype=torch.uint8, device=dev), batch, M)
    sync_all()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for b0 in range(0, nb - batch + 1, batch):
        stage.storage.copy_(host_rows[b0:b0 + batch], non_blocking=True)
        q = model.infer_packed(stage, batch)[0]
        q_host[b0:b0 + batch].copy_(q, non_blocking=True)
    h1.record()
    sync_all()
    th = torch.tensor([h0.elapsed_time(h1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(th, op=dist.ReduceOp.MAX)
    n_e2e = (nb // batch) * batch
    if rank == 0:
        line = {"metric": METRICS["cfg5"], "value": N / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": passes,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"cfg5: projective inference of {N} samples x {M} SNPs, heads K={ks}, "
                                       f"sample axis sharded over {world} GPU(s) ({n_loc} rows on rank 0), batch {batch}",
                           "hidden": HIDDEN, "n_components": NCOMP,
                           "l2": f"inputs larger than L2: {n_loc * pitch_bytes / 1e9:.2f} GB of packed rows per rank"},
                "roofline": {"bound": "hbm", "kernel": "enc_fwd_tc_kernel (+ reduce, MLP forward)",
                             "achieved": bytes_rank / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": bytes_rank / (ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": batch * pitch_bytes + 4 * M * NCOMP},
                "cpu_baseline": None,
                "e2e": {"value": world * n_e2e / (float(th.item()) * 1e-3), "unit": UNIT,
                        "h2d_bytes_per_step": batch * pg.pitch, "d2h_bytes_per_step": batch * ks[0] * 4,
                        "what": "2048-row batches copied from pinned host memory, Q copied back"},
                "gpu_launches": int(ops.launch_count() - l0), "clocks": sampler.summary() if sampler else None}
        print(json.dumps(line), flush=True)
    teardown(dist, world, None)


if __name__ == "__main__":
    main()
