import sys, os, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
import neural_admixture_b200._lib as L
from pathlib import Path
L.LIB_PATH = Path('/root/repo/neural_admixture_b200/csrc/' + (sys.argv[1] if len(sys.argv) > 1 else 'libnadm_b200_tl.so'))
import torch
from neural_admixture_b200 import ops
dev = torch.device('cuda:0')
N, M, k, B = 4000, (int(sys.argv[2]) if len(sys.argv) > 2 else 500000), 8, 800
gen = torch.Generator(device=dev).manual_seed(1)
pg = ops.PackedGenotypes.empty(N, M, dev)
for r0 in range(0, N, 500):
    codes = torch.randint(0, 3, (500, M), dtype=torch.uint8, device=dev, generator=gen)
    ops.pack2bit(codes, pg.storage[r0:r0 + 500], M)
idx = torch.randperm(N, device=dev, generator=gen)[:B].contiguous()
Q = torch.softmax(torch.randn((B, k), device=dev, generator=gen), 1).contiguous()
P = (torch.rand((M, k), device=dev, generator=gen) * 0.9 + 0.05).contiguous()
dQ = torch.zeros((B, k), device=dev); loss = torch.zeros(1, device=dev)
Pm = torch.zeros_like(P); Pv = torch.zeros_like(P)
ws = torch.empty(ops.workspace_bytes(B, M, 8, 1024, k), dtype=torch.uint8, device=dev)
for it in range(3):
    ops.decoder_step(pg, Q, dQ, 0, k, P, Pm, Pv, ops.adam_hyper(2e-3, it + 1), loss, ws, row_idx=idx)
torch.cuda.synchronize()
out = np.zeros((8, 512), dtype=np.int64)
lib = L.load()
lib.nadm_debug_timeline.argtypes = [ctypes.c_void_p]
rc = lib.nadm_debug_timeline(out.ctypes.data)
t0 = out[0][0]
names = ["wg_wait_start", "wg_raw_seen", "wg_G_written", "issA_wait", "issA_G_seen", "issA_done", "wg_gt_free", "issB_done"]
print("unit " + " ".join(f"{n:>14s}" for n in names))
for u in range(60, 84):
    print(f"{u:4d} " + " ".join(f"{int(out[r][u] - t0):14d}" for r in range(8)))
span = out[5][300] - out[5][100]
print(f"cycles per unit (units 100..300, issuer A done): {span / 200:.1f}")
dec = (out[2][100:300] - out[1][100:300]).mean()
wait = (out[1][100:300] - out[0][100:300]).mean()
print(f"compute warpgroup: decode {dec:.0f} cycles/unit, wait for raw {wait:.0f} cycles/unit")
print(f"issuer A: wait for G {(out[4][100:300] - out[3][100:300]).mean():.0f}, issue {(out[5][100:300] - out[4][100:300]).mean():.0f} cycles/unit")
ph = out[0][500:506]
print("phases (cycles): setup", int(ph[1] - ph[0]), " units of warpgroup 0", int(ph[2] - ph[1]), " wait for the last MMA2", int(ph[3] - ph[2]),
      " dQ partial", int(ph[4] - ph[3]), " tail (last dP epilogue, sync)", int(ph[5] - ph[4]), " total", int(ph[5] - ph[0]))
ph2 = out[0][505:508]
print("fused reduction (cycles): grid barrier", int(ph2[1] - ph2[0]), " reduce share", int(ph2[2] - ph2[1]))
