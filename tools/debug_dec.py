import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import nadm_oracle as orc
from neural_admixture_b200 import ops
dev = torch.device('cuda:0')
def t(a, dtype=torch.float32): return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=dev).contiguous()
def relF(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
cases = [(40, 1024, 3, 40, True), (40, 1024, 3, 40, False), (90, 515, 8, 77, False), (20, 9, 2, 20, True), (64, 203, 5, 48, True)]
for (N, M, k, B, edge) in cases:
    rng = np.random.default_rng(M + k)
    G = rng.integers(0, 3, size=(N, M), dtype=np.uint8); G[rng.random((N, M)) < 0.02] = 3
    P = rng.uniform(0.02, 0.98, size=(M, k)).astype(np.float32)
    if edge:
        P[5 % M, :] = 0.0; P[min(17, M - 1), 0] = 0.0; P[min(23, M - 1), k - 1] = 1.0
    Q = rng.dirichlet(0.3 * np.ones(k), size=B).astype(np.float32)
    idx = rng.permutation(N)[:B]
    pg = ops.PackedGenotypes.from_unpacked_host(torch.as_tensor(G), dev)
    dQ = torch.zeros((B, k), device=dev); dP = torch.empty((M, k), device=dev); loss = torch.zeros(1, device=dev)
    ws = torch.empty(ops.workspace_bytes(B, M, 8, 64, k), dtype=torch.uint8, device=dev)
    ops.decoder_step(pg, t(Q), dQ, 0, k, t(P), None, None, None, loss, ws, row_idx=t(idx, torch.int64), dP_out=dP)
    x = orc.genotype_to_x(G[idx])
    l_ref, dQ_ref, dP_ref = orc.decoder_loss_grads(x, Q.astype(np.float64), P.astype(np.float64))
    dq, dp = dQ.cpu().numpy().astype(np.float64), dP.cpu().numpy().astype(np.float64)
    print(f"case {N,M,k,B,edge}: loss rel {abs(loss.item()-l_ref)/abs(l_ref):.2e}  dQ relF {relF(dq,dQ_ref):.2e}  dP relF {relF(dp,dP_ref):.2e}")
    err = np.abs(dq - dQ_ref); b, kk = np.unravel_index(err.argmax(), err.shape)
    print(f"   worst dQ at b={b},k={kk}: got {dq[b,kk]:.6e} ref {dQ_ref[b,kk]:.6e}  Q[b]={Q[b]}")
    # per-element contributions of the worst row
    raw = Q[b].astype(np.float64) @ P.astype(np.float64).T
    Gm = (np.clip(raw,0,1) - x[b]) / np.maximum(raw*(1-raw), 1e-12) * ((raw>=0)&(raw<=1))
    terms = Gm * P[:, kk]
    top = np.argsort(-np.abs(terms))[:4]
    print("   largest terms:", [(int(m), float(terms[m]), float(raw[m]), float(x[b, m])) for m in top], " sum|terms|=%.3e" % np.abs(terms).sum())
    errp = np.abs(dp - dP_ref); m, kk = np.unravel_index(errp.argmax(), errp.shape)
    print(f"   worst dP at m={m},k={kk}: got {dp[m,kk]:.6e} ref {dP_ref[m,kk]:.6e}")
