"""Bounded probe of the fused decoder kernel alone at the bench shapes (A/B and knock-out builds: run under `timeout`).
usage: NADM_LIB=<variant> python tools/dec_probe.py [M] [rows] [k] [B] [loss:0|1] [late:0|1]
late=1: late-training inputs as in bench.py's `late_training` leg (P: 20 % exact 0, 20 % exact 1; Q ~ Dirichlet(0.05); lr 0)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from neural_admixture_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 8
B = int(sys.argv[4]) if len(sys.argv) > 4 else 800
want_loss = (int(sys.argv[5]) if len(sys.argv) > 5 else 1) != 0
late = (int(sys.argv[6]) if len(sys.argv) > 6 else 0) != 0
dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(1)
pg = ops.PackedGenotypes.empty(N, M, dev)
for r0 in range(0, N, 100):
    nr = min(100, N - r0)
    ops.pack2bit(torch.randint(0, 3, (nr, M), dtype=torch.uint8, device=dev, generator=gen), pg.storage[r0:r0 + nr], M)
idx = torch.randint(0, N, (B,), device=dev, generator=gen).contiguous()
Q = torch.softmax(torch.randn((B, k), device=dev, generator=gen), dim=1).contiguous()
dQ = torch.zeros_like(Q)
P = (torch.rand((M, k), device=dev, generator=gen) * 0.9 + 0.05).contiguous()
Pm, Pv = torch.zeros_like(P), torch.zeros_like(P)
loss = torch.zeros(1, device=dev)
ws = torch.empty(ops.workspace_bytes(B, M, 8, 1024, k), dtype=torch.uint8, device=dev)
hyper = ops.adam_hyper(1e-6, 10_000)
if late:
    u = torch.rand((M, k), device=dev, generator=gen)
    P = torch.where(u < 0.2, torch.zeros_like(P), torch.where(u < 0.4, torch.ones_like(P), P)).contiguous()
    Q = torch.distributions.Dirichlet(torch.full((k,), 0.05, device=dev)).sample((B,)).float().contiguous()
    hyper = ops.adam_hyper(0.0, 10_000)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(13):
    if it == 3:
        e0.record()
    ops.decoder_step(pg, Q, dQ, 0, k, P, Pm, Pv, hyper, loss if want_loss else None, ws, row_idx=idx)
e1.record()
torch.cuda.synchronize()
print('dec', M, 'k', k, 'B', B, 'loss', int(want_loss), 'late', int(late), 'ok', round(e0.elapsed_time(e1) * 100, 1), 'us per call', flush=True)
