#!/usr/bin/env python
"""Multi-GPU parity of the SNP-sharded engine (NCCL) against the single-GPU engine and the CPU oracle, and the
sharded train -> save -> load -> infer round trip.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py

Prints one line ``SHARDED PARITY (n GPUs): PASS|FAIL`` (tests/test_gpu_sharded.py runs this under torchrun)."""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))
from neural_admixture_b200 import ops  # noqa: E402
from neural_admixture_b200.model.neural_admixture import NeuralAdmixture  # noqa: E402
from neural_admixture_b200.model.train import snp_slice  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(5)
    N, M, ks, C, H, B, epochs = 1500, 40_003, [5, 6, 7, 8], 8, 128, 512, 2
    Pt = rng.uniform(0.05, 0.95, size=(6, M))
    Qt = rng.dirichlet(0.3 * np.ones(6), size=N)
    G = rng.binomial(2, Qt @ Pt).astype(np.uint8)
    G[rng.random((N, M)) < 0.01] = 3
    V = np.linalg.qr(rng.standard_normal((M, C)))[0].astype(np.float32)
    P0 = rng.uniform(0.05, 0.95, size=(sum(ks), M)).astype(np.float32)
    init = {}
    exch = ["?"]

    def run(sharded):
        c0, c1 = snp_slice(M, rank, world) if sharded else (0, M)
        torch.manual_seed(0)
        na = NeuralAdmixture(None, epochs, B, 2e-3, dev, 0, world if sharded else 0, rank == 0, "nadm_b200", min(ks), max(ks))
        na.keep_loss_history = True
        orig = na.initialize_model

        def init_and_capture(*a):
            orig(*a)
            if not sharded:
                init.update({n: t.detach().cpu().numpy().astype(np.float64) for n, t in na.raw_model.state_dict().items()})

        na.initialize_model = init_and_capture
        packed = ops.PackedGenotypes.from_unpacked_host(torch.as_tensor(G), dev, c0, c1)
        orig_open = na.open_exchange

        def open_and_note(*a, **kw):
            orig_open(*a, **kw)
            exch[0] = "fused peer exchange" if na.exchange is not None else "NCCL all-reduce"

        na.open_exchange = open_and_note
        Qs, Ps, raw = na.launch_training(torch.as_tensor(P0[:, c0:c1].copy(), device=dev), packed, H, C,
                                         torch.as_tensor(V[c0:c1].copy(), device=dev), c1 - c0, N)
        if rank == 0:
            print(f"  {'sharded' if sharded else 'single '} run: cuda-graph steps = {na.use_graph}, "
                  f"graph-replayed kernels = {na.graph_kernel_launches}, generic kernels = {na.generic_kernel_launches}, "
                  f"exchange = {exch[0] if sharded else 'none'}")
        return Qs, Ps, na.loss_history, raw

    Qs_s, Ps_s, loss_s, raw_s = run(True)
    dist.barrier()
    ok = True
    if rank == 0:
        rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
        # ---- the checkpoint a sharded run writes (reference main.py:41-43) is the single-GPU checkpoint ----
        from neural_admixture_b200.src import inference
        with tempfile.TemporaryDirectory() as d:
            sd = {n: t.detach().cpu() for n, t in raw_s.state_dict().items() if not n.startswith("decoders")}
            assert tuple(sd["V"].shape) == (M, C), f"sharded run returned a truncated V {tuple(sd['V'].shape)}"
            torch.save(sd, f"{d}/run.pt")
            raw_s.save_config("run", d)
            model = inference.load_model(d, "run", dev)
            full = ops.PackedGenotypes.from_unpacked_host(torch.as_tensor(G), dev)
            Qi = [q.cpu().numpy() for q in model.infer_packed(full, 1024)]
        e_rt = max(rel(Qi[i], Qs_s[i]) for i in range(len(ks)))
        print(f"sharded train -> save -> load_model -> infer: max relF(Q) vs the training run's Q = {e_rt:.2e}")
        ok &= e_rt < 1e-5
        # ---- sharded vs single GPU ----
        Qs_1, Ps_1, loss_1, _ = run(False)
        worst = 0.0
        for i in range(len(ks)):
            eq, ep = rel(Qs_s[i], Qs_1[i]), rel(Ps_s[i], Ps_1[i])
            worst = max(worst, eq, ep)
            print(f"head K={ks[i]}: relF(Q sharded vs single) = {eq:.2e}   relF(P) = {ep:.2e}")
        el = max(abs(a - b) / abs(b) for a, b in zip(loss_s, loss_1))
        print(f"epoch losses sharded {loss_s} single {loss_1} (max rel diff {el:.2e})")
        ok &= worst < 1e-4 and el < 1e-5
        # ---- sharded vs the fp64 oracle (same initial parameters, same sampler stream) ----
        import nadm_oracle as orc
        from helpers import state_from_sd
        st = state_from_sd(init, ks)
        gen = torch.Generator().manual_seed(0)
        orders = [np.array(list(torch.utils.data.RandomSampler(range(N), generator=gen))) for _ in range(epochs)]
        lo, Qo = orc.train(st, G, orders, B, 2e-3)
        wo = max(max(rel(Qs_s[i], Qo[i]), rel(Ps_s[i], st.P[i])) for i in range(len(ks)))
        elo = max(abs(a - b) / abs(b) for a, b in zip(loss_s, lo))
        print(f"sharded vs fp64 oracle: max relF(Q, P) = {wo:.2e}, max rel epoch-loss diff {elo:.2e}")
        ok &= wo < 1e-4 and elo < 5e-5
        print(f"SHARDED PARITY ({world} GPUs): {'PASS' if ok else 'FAIL'}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
