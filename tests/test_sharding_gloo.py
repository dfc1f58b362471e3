"""SNP-axis sharding checked with two CPU processes over gloo.

Each rank owns a contiguous, 64-SNP aligned slice of the genotype columns, of V and of P; the only exchanges per step
are the all-reduce of the partial projection Z (B x C) and of [dQ | loss] (B x sumK + 1).  The ranks run the numpy
oracle on their slice (the CUDA kernels need a GPU; the slicing / exchange / replication logic is what is tested
here) and must reproduce the single-process oracle step exactly up to summation order.  The same orchestration
(`NeuralAdmixture._train_step` with `sharded=True`) runs over NCCL on the GPU box (tools/check_sharded.py)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import nadm_oracle as orc  # noqa: E402
from neural_admixture_b200.model.train import snp_slice  # noqa: E402


def _problem(seed=3, N=96, M=1000, ks=(3, 4), C=8, H=16):
    rng = np.random.default_rng(seed)
    G = rng.integers(0, 3, size=(N, M), dtype=np.uint8)
    G[rng.random((N, M)) < 0.02] = 3
    st = orc.OracleState(
        V=rng.standard_normal((M, C)) / np.sqrt(M), w_rms=1 + 0.1 * rng.standard_normal(C),
        W1=rng.standard_normal((H, C)) / np.sqrt(C), b1=0.1 * rng.standard_normal(H),
        W2=[rng.standard_normal((k, H)) / np.sqrt(H) for k in ks], b2=[0.1 * rng.standard_normal(k) for k in ks],
        P=[rng.uniform(0.02, 0.98, size=(M, k)) for k in ks], ks=list(ks))
    batches = [rng.permutation(N)[:64] for _ in range(3)]
    return G, st, batches


def _all_reduce(a: np.ndarray) -> np.ndarray:
    t = torch.from_numpy(np.ascontiguousarray(a))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


def _sharded_step(st_loc: orc.OracleState, g_loc: np.ndarray, lr: float):
    """One training step on this rank's SNP slice: the order of operations of NeuralAdmixture._train_step."""
    x = orc.genotype_to_x(g_loc)
    Z = _all_reduce(orc.encoder_fwd(x, st_loc.V))                                  # exchange 1: B x C
    Zn, rinv, Hh, Qs = orc.mlp_fwd(Z, st_loc.w_rms, st_loc.W1, st_loc.b1, st_loc.W2, st_loc.b2)
    loss, dQs, grads = 0.0, [], {}
    for i, (Q, P) in enumerate(zip(Qs, st_loc.P)):
        l, dQ, dP = orc.decoder_loss_grads(x, Q, P)
        loss += l
        dQs.append(dQ)
        grads[f"P.{i}"] = dP
    packed = _all_reduce(np.concatenate([np.concatenate(dQs, axis=1).ravel(), [loss]]))   # exchange 2: B x sumK + 1
    loss = float(packed[-1])
    dQ_all = packed[:-1].reshape(len(x), -1)
    off, dQs = 0, []
    for k in st_loc.ks:
        dQs.append(dQ_all[:, off:off + k])
        off += k
    dZ, dw, dW1, db1, dW2, db2 = orc.mlp_bwd(dQs, Qs, Hh, Zn, rinv, Z, st_loc.w_rms, st_loc.W1, st_loc.W2)
    grads.update({"V": orc.encoder_bwd(x, dZ), "w_rms": dw, "W1": dW1, "b1": db1})
    for i in range(len(st_loc.ks)):
        grads[f"W2.{i}"], grads[f"b2.{i}"] = dW2[i], db2[i]
    st_loc.t += 1
    for name, p in st_loc.params().items():
        if name not in st_loc.m:
            st_loc.m[name], st_loc.v[name] = np.zeros_like(p), np.zeros_like(p)
        orc.adam_update(p, grads[name], st_loc.m[name], st_loc.v[name], st_loc.t, lr)
    for P in st_loc.P:
        np.clip(P, 0.0, 1.0, out=P)
    return loss


def _worker(rank: int, world: int, port: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    G, st, batches = _problem()
    c0, c1 = snp_slice(G.shape[1], rank, world)
    loc = st.copy()
    loc.V = st.V[c0:c1].copy()
    loc.P = [P[c0:c1].copy() for P in st.P]
    losses = [_sharded_step(loc, G[idx][:, c0:c1], 2e-3) for idx in batches]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), c0=c0, c1=c1, V=loc.V, losses=np.array(losses), W1=loc.W1,
             w_rms=loc.w_rms, **{f"P{i}": P for i, P in enumerate(loc.P)}, **{f"W2_{i}": w for i, w in enumerate(loc.W2)})
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_snp_slices_partition_the_axis():
    for M, world in [(1000, 2), (500_000, 8), (300_000, 4), (63, 2), (64, 3), (8451, 8)]:
        edges = [snp_slice(M, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == M
        for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
            assert a1 == b0 and a0 <= a1
        assert all(a0 % 64 == 0 for a0, _ in edges)          # slices start on 16 packed bytes


@pytest.mark.timeout(300)
def test_two_rank_sharded_step_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    G, st, batches = _problem()
    ref_losses = [orc.train_step(st, G[idx], 2e-3)[0] for idx in batches]
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    V = np.concatenate([p["V"] for p in parts], axis=0)
    assert [int(p["c0"]) for p in parts] == [0, int(parts[0]["c1"])]
    np.testing.assert_allclose(V, st.V, rtol=1e-9, atol=1e-12)
    for i in range(len(st.ks)):
        np.testing.assert_allclose(np.concatenate([p[f"P{i}"] for p in parts], axis=0), st.P[i], rtol=1e-9, atol=1e-12)
    for p in parts:                                            # replicated parameters stay identical on every rank
        np.testing.assert_allclose(p["losses"], ref_losses, rtol=1e-10)
        np.testing.assert_allclose(p["W1"], st.W1, rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(p["w_rms"], st.w_rms, rtol=1e-8, atol=1e-12)
        for i in range(len(st.ks)):
            np.testing.assert_allclose(p[f"W2_{i}"], st.W2[i], rtol=1e-8, atol=1e-12)
    assert np.array_equal(parts[0]["W1"], parts[1]["W1"])
