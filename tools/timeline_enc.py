import sys, ctypes, numpy as np
sys.path.insert(0, '/root/repo')
import neural_admixture_b200._lib as L
from pathlib import Path
L.LIB_PATH = Path('/root/repo/neural_admixture_b200/csrc') / (sys.argv[1] if len(sys.argv) > 1 else 'libnadm_b200_tl.so')
import torch
from neural_admixture_b200 import ops
dev = torch.device('cuda:0')
N, M, C, B = 4000, (int(sys.argv[2]) if len(sys.argv) > 2 else 500000), 8, 800
gen = torch.Generator(device=dev).manual_seed(1)
pg = ops.PackedGenotypes.empty(N, M, dev)
for r0 in range(0, N, 500):
    codes = torch.randint(0, 3, (500, M), dtype=torch.uint8, device=dev, generator=gen)
    ops.pack2bit(codes, pg.storage[r0:r0 + 500], M)
idx = torch.randperm(N, device=dev, generator=gen)[:B].contiguous()
V = (torch.randn((M, C), device=dev, generator=gen) / M ** 0.5).contiguous()
Z = torch.empty((B, C), device=dev)
ws = torch.empty(ops.workspace_bytes(B, M, 8, 1024, 8), dtype=torch.uint8, device=dev)
for it in range(3):
    ops.encoder_fwd(pg, V, Z, ws, row_idx=idx)
torch.cuda.synchronize()
out = np.zeros((8, 512), dtype=np.int64)
lib = L.load()
lib.nadm_debug_enc_timeline.argtypes = [ctypes.c_void_p]
lib.nadm_debug_enc_timeline(out.ctypes.data)
t0 = out[0][0]
names = ["p_start", "p_emptyA", "p_cpwait", "p_widened", "p_arrived", "p_issued", "m_fullA", "m_done"]
print("tile " + " ".join(f"{n:>10s}" for n in names))
for u in range(40, 62):
    print(f"{u:4d} " + " ".join(f"{int(out[r][u] - t0):10d}" for r in range(8)))

t = out[2]
print("phases (cycles): setup", int(t[1] - t[0]), " first tile start after setup", int(out[0][0] - t[1]),
      " producers' loop", int(t[2] - out[0][0]), " drain MMAs", int(t[3] - t[2]), " epilogue", int(t[4] - t[3]),
      " total", int(t[4] - t[0]))
print("fused reduction (cycles): grid barrier", int(t[5] - t[4]), " reduce share", int(t[6] - t[5]))
