"""Parity of the CUDA path (through the C ABI, via neural_admixture_b200.ops) with the CPU oracle and with the golden
fixtures generated from the reference.  Needs a B200: run with ``-m gpu``.

Tolerances (floating point; fp32 kernels vs the fp64 oracle):
  KERNEL_TOL = 2e-5   relative Frobenius error of one kernel's output (fp32 rounding of M-long sums)
  STEP_TOL   = 5e-5   parameters after one/two full optimisation steps (Adam normalises gradients: |dp| = lr)
  QP_TOL     = 1e-4   north-star bar: ||Q - Q_ref||_F / ||Q_ref||_F (and P) after short trainings
Integer work (pack / unpack) is bit-exact."""
from pathlib import Path

import numpy as np
import pytest
import torch

import nadm_oracle as orc
from helpers import load_golden, sub, state_from_sd, sd_name, relF

pytestmark = pytest.mark.gpu

KERNEL_TOL = 2e-5
STEP_TOL = 5e-5
QP_TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "the -m gpu tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from neural_admixture_b200 import ops as _ops
    return _ops


def t(a, dev, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=dev).contiguous()


def rand_genotypes(rng, N, M, miss=0.02):
    G = rng.integers(0, 3, size=(N, M), dtype=np.uint8)
    G[rng.random((N, M)) < miss] = 3
    return G


def packed_from(ops, G, dev):
    return ops.PackedGenotypes.from_unpacked_host(torch.as_tensor(G), dev)


def decoder_grads_with_fp32_raw(x, Q, P):
    """orc.decoder_loss_grads with raw = Q P^T rounded once to fp32 (everything else fp64): the error floor of ANY
    fp32 implementation of the decoder, the reference's own included."""
    raw = (Q @ P.T).astype(np.float32).astype(np.float64)
    R = np.clip(raw, 0.0, 1.0)
    Gm = (R - x) / np.maximum((1.0 - R) * R, 1e-12)
    Gm = np.where((raw >= 0.0) & (raw <= 1.0), Gm, 0.0)
    return None, Gm @ P, Gm.T @ Q


def ws_for(ops, B, M, C, H, sumK, dev):
    return torch.empty(ops.workspace_bytes(B, M, C, H, sumK), dtype=torch.uint8, device=dev)


# ---------------------------------------------------------------------------------------------------------------
# integer work: bit-exact
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,M", [(1, 1), (3, 4), (5, 7), (16, 203), (2, 1025), (33, 4096), (7, 65537)])
def test_pack_unpack_bit_exact(ops, dev, N, M):
    rng = np.random.default_rng(N * 1000 + M)
    G = rng.integers(0, 256, size=(N, M), dtype=np.uint8)          # high bits must be dropped (pack2bit.cu:29)
    src = torch.as_tensor(G, device=dev)
    pc = (M + 3) // 4
    dst = torch.full((N, pc), 0xAA, dtype=torch.uint8, device=dev)
    ops.pack2bit(src, dst)
    assert np.array_equal(dst.cpu().numpy(), orc.pack2bit(G))
    back = torch.full((N, M), 9, dtype=torch.uint8, device=dev)
    ops.unpack2bit(dst, back)
    assert np.array_equal(back.cpu().numpy(), G & 3)
    # padded-pitch container: tail bytes are zero
    pg = packed_from(ops, G & 3, dev)
    st = pg.storage.cpu().numpy()
    assert np.array_equal(st[:, :pc], orc.pack2bit(G)) and not st[:, pc:].any()


def _reference_pack2bit_module():
    """oracle/_ref/pack2bit_ref.so: the reference's own pack2bit.cu compiled unmodified by oracle/build_ref.py (in the
    build container, where /root/reference exists; it travels to the GPU box with the snapshot)."""
    import importlib.util
    so = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "pack2bit_ref.so"
    if not so.exists():
        pytest.skip("oracle/_ref/pack2bit_ref.so was not built (python oracle/build_ref.py needs /root/reference)")
    try:
        spec = importlib.util.spec_from_file_location("pack2bit_ref", so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception as e:        # e.g. a torch build on the box other than the one it was compiled against
        pytest.skip(f"the compiled reference module does not load here: {type(e).__name__}: {e}")
    return mod


@pytest.mark.parametrize("N,M", [(3, 4), (5, 7), (16, 203), (1100, 1025), (33, 4099)])
def test_pack_unpack_against_the_compiled_reference_kernels(ops, dev, N, M):
    """The 2-bit layout pinned to the reference ITSELF: its pack2bit_cpu_to_gpu / unpack2bit_gpu_to_gpu (pack2bit.cu:10-63,
    compiled from /root/reference by oracle/build_ref.py) against nadm_pack2bit / nadm_unpack2bit, bit for bit, both
    directions and crosswise (each side unpacks what the other packed)."""
    ref = _reference_pack2bit_module()
    rng = np.random.default_rng(N * 131 + M)
    G = rng.integers(0, 256, size=(N, M), dtype=np.uint8)          # high bits must be dropped by both
    pc = (M + 3) // 4
    theirs = torch.full((N, pc), 0x55, dtype=torch.uint8, device=dev)
    ref.pack2bit_cpu_to_gpu(torch.as_tensor(G), theirs)
    mine = torch.full((N, pc), 0xAA, dtype=torch.uint8, device=dev)
    ops.pack2bit(torch.as_tensor(G, device=dev), mine)
    torch.cuda.synchronize()
    assert torch.equal(mine, theirs)
    assert np.array_equal(theirs.cpu().numpy(), orc.pack2bit(G))    # and the oracle's restatement of it
    a = torch.full((N, M), 9, dtype=torch.uint8, device=dev)
    b = torch.full((N, M), 7, dtype=torch.uint8, device=dev)
    ref.unpack2bit_gpu_to_gpu(mine, a)                              # the reference unpacks what this library packed
    ops.unpack2bit(theirs, b)                                       # and the other way round
    torch.cuda.synchronize()
    assert torch.equal(a, b) and np.array_equal(a.cpu().numpy(), G & 3)


def test_reference_pack2bit_module_surface(ops, dev):
    """Same two functions, argument meaning and error behaviour as the reference's pybind module
    (pack2bit.cu:65-76,120-130,144-147)."""
    from neural_admixture_b200.src import pack2bit
    from neural_admixture_b200._lib import NadmError
    rng = np.random.default_rng(5)
    G = rand_genotypes(rng, 1500, 777)                              # > 1024 rows: two staging chunks
    out = torch.empty((1500, (777 + 3) // 4), dtype=torch.uint8, device=dev)
    pack2bit.pack2bit_cpu_to_gpu(torch.as_tensor(G), out)
    assert np.array_equal(out.cpu().numpy(), orc.pack2bit(G))
    un = torch.empty((800, 777), dtype=torch.uint8, device=dev)
    pack2bit.unpack2bit_gpu_to_gpu(out[:800].contiguous(), un)
    assert np.array_equal(un.cpu().numpy(), G[:800])
    with pytest.raises(NadmError):
        pack2bit.pack2bit_cpu_to_gpu(torch.as_tensor(G).to(dev), out)
    with pytest.raises(NadmError):
        pack2bit.pack2bit_cpu_to_gpu(torch.as_tensor(G), out[:, :-1])
    with pytest.raises(NadmError):
        pack2bit.unpack2bit_gpu_to_gpu(out.cpu(), un)


# ---------------------------------------------------------------------------------------------------------------
# kernels vs oracle
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,M,C,B,gather", [(64, 203, 8, 48, True), (300, 4099, 8, 300, True), (40, 1024, 8, 40, False),
                                            (1000, 20000, 8, 800, True), (17, 5, 3, 17, True), (520, 777, 16, 513, True)])
def test_encoder_fwd(ops, dev, N, M, C, B, gather):
    rng = np.random.default_rng(M)
    G = rand_genotypes(rng, N, M)
    V = (rng.standard_normal((M, C)) / np.sqrt(M)).astype(np.float32)
    idx = rng.permutation(N)[:B] if gather else np.arange(3 if N > B + 3 else 0, (3 if N > B + 3 else 0) + B)
    pg = packed_from(ops, G, dev)
    Z = torch.empty((B, C), dtype=torch.float32, device=dev)
    ws = ws_for(ops, B, M, C, 64, 8, dev)
    if gather:
        ops.encoder_fwd(pg, t(V, dev), Z, ws, row_idx=t(idx, dev, torch.int64))
    else:
        ops.encoder_fwd(pg, t(V, dev), Z, ws, row0=int(idx[0]), B=B)
    ref = orc.encoder_fwd(orc.genotype_to_x(G[idx]), V.astype(np.float64))
    assert relF(Z.cpu().numpy(), ref) < KERNEL_TOL


@pytest.mark.parametrize("ks,H,B", [([5], 64, 48), ([8], 1024, 800), ([3, 4, 5], 32, 150), (list(range(4, 13)), 1024, 77)])
def test_mlp_fwd_bwd(ops, dev, ks, H, B):
    from neural_admixture_b200._lib import MlpParams
    rng = np.random.default_rng(B)
    C, sumK = 8, sum(ks)
    Z = rng.standard_normal((B, C)) * 0.3
    w_rms = 1 + 0.1 * rng.standard_normal(C)
    W1 = rng.standard_normal((H, C)) / np.sqrt(C)
    b1 = 0.1 * rng.standard_normal(H)
    W2 = [rng.standard_normal((k, H)) / np.sqrt(H) for k in ks]
    b2 = [0.1 * rng.standard_normal(k) for k in ks]
    Zn, rinv, Hh, Qs = orc.mlp_fwd(Z, w_rms, W1, b1, W2, b2)
    d = {"Z": t(Z, dev), "w_rms": t(w_rms, dev), "W1": t(W1, dev), "b1": t(b1, dev),
         "W2": t(np.concatenate(W2, 0), dev), "b2": t(np.concatenate(b2, 0), dev)}
    rinv_d = torch.empty(B, device=dev)
    Hh_d = torch.empty((B, H), device=dev)
    Q_d = torch.empty((B, sumK), device=dev)
    ops.mlp_fwd(d["Z"], d["w_rms"], d["W1"], d["b1"], d["W2"], d["b2"], ks, rinv_d, Hh_d, Q_d)
    assert relF(Q_d.cpu().numpy(), np.concatenate(Qs, 1)) < KERNEL_TOL
    assert relF(Hh_d.cpu().numpy(), Hh) < KERNEL_TOL
    # backward with raw gradients out (no Adam), supervised term on when single head
    dQs = [rng.standard_normal(q.shape) * 10 for q in Qs]
    y = rng.integers(0, ks[0], size=B) if len(ks) == 1 else None
    dQs_ref = [g.copy() for g in dQs]
    sup_loss = 0.0
    if y is not None:
        sup_loss, dsup = orc.supervised_loss_grads(Qs[0], y, 100.0)
        dQs_ref[0] = dQs_ref[0] + dsup
    dZ, dw, dW1, db1, dW2, db2 = orc.mlp_bwd(dQs_ref, Qs, Hh, Zn, rinv, Z, w_rms, W1, W2)
    g = {n: torch.zeros_like(d[n]) for n in ["w_rms", "W1", "b1", "W2", "b2"]}
    p = MlpParams()
    for n in g:
        setattr(p, n, d[n].data_ptr())
        setattr(p, "g_" + n, g[n].data_ptr())
    dZ_d = torch.empty((B, C), device=dev)
    loss = torch.zeros(1, device=dev)
    ws = ws_for(ops, B, 1000, C, H, sumK, dev)
    ops.mlp_bwd(t(np.concatenate(dQs, 1), dev), Q_d, Hh_d, d["Z"], rinv_d, ks, p, None, dZ_d, loss, ws,
                labels=None if y is None else t(y, dev, torch.int64), sup_weight=100.0 if y is not None else 0.0)
    tol = 5e-5
    assert relF(dZ_d.cpu().numpy(), dZ) < tol
    assert relF(g["W1"].cpu().numpy(), dW1) < tol
    assert relF(g["b1"].cpu().numpy(), db1) < tol
    assert relF(g["W2"].cpu().numpy(), np.concatenate(dW2, 0)) < tol
    assert relF(g["b2"].cpu().numpy(), np.concatenate(db2, 0)) < tol
    assert relF(g["w_rms"].cpu().numpy(), dw) < tol
    if y is not None:
        assert abs(loss.item() - sup_loss) < 1e-5 * abs(sup_loss)


@pytest.mark.parametrize("ks,B,M", [([8], 800, 40000), ([5], 130, 3000), ([3, 12, 6], 300, 9001), ([8], 2048, 20000)])
def test_deferred_reductions_match_separate_kernels(ops, dev, ks, B, M):
    """encoder_fwd(deferred) + mlp_fwd and decoder_step(deferred) + mlp_bwd (the sums over the producers' CTAs are
    completed inside the consuming MLP kernels) against the same calls with their separate reduction kernels: the
    forward sums exact integers in double (same value up to the final rounding), dQ is a re-ordered fp32 sum."""
    from neural_admixture_b200._lib import MlpParams
    rng = np.random.default_rng(B + M)
    N, C, H, sumK = B + 50, 8, 256, sum(ks)
    G = rand_genotypes(rng, N, M)
    pg = packed_from(ops, G, dev)
    idx = t(rng.permutation(N)[:B], dev, torch.int64)
    V = t((rng.standard_normal((M, C)) / np.sqrt(M)), dev)
    prm = {"w_rms": t(1 + 0.1 * rng.standard_normal(C), dev), "W1": t(rng.standard_normal((H, C)) / np.sqrt(C), dev),
           "b1": t(0.1 * rng.standard_normal(H), dev), "W2": t(rng.standard_normal((sumK, H)) / np.sqrt(H), dev),
           "b2": t(0.1 * rng.standard_normal(sumK), dev)}
    Ps = [t(rng.uniform(0.02, 0.98, size=(M, k)), dev) for k in ks]
    ws = ws_for(ops, B, M, C, H, sumK, dev)
    outs = []
    for deferred in (False, True):
        Z = torch.empty((B, C), device=dev)
        rinv, Hh, Q = torch.empty(B, device=dev), torch.empty((B, H), device=dev), torch.empty((B, sumK), device=dev)
        ops.encoder_fwd(pg, V, Z, ws, row_idx=idx, deferred=deferred)
        ops.mlp_fwd(Z, prm["w_rms"], prm["W1"], prm["b1"], prm["W2"], prm["b2"], ks, rinv, Hh, Q)
        dQ = torch.zeros((B, sumK), device=dev)
        loss = torch.zeros(1, device=dev)
        off = 0
        for i, k in enumerate(ks):       # every head deferred: a pending reduction is completed by the next head's call
            ops.decoder_step(pg, Q, dQ, off, k, Ps[i], None, None, None, loss, ws, row_idx=idx, deferred=deferred)
            off += k
        g = {n: torch.zeros_like(prm[n]) for n in prm}
        p = MlpParams()
        for n in prm:
            setattr(p, n, prm[n].data_ptr())
            setattr(p, "g_" + n, g[n].data_ptr())
        dZ = torch.empty((B, C), device=dev)
        ops.mlp_bwd(dQ, Q, Hh, Z, rinv, ks, p, None, dZ, loss, ws)
        outs.append({"Z": Z, "Q": Q, "Hh": Hh, "dQ": dQ, "dZ": dZ, "loss": loss, **{"g_" + n: g[n] for n in g}})
    a, b = outs
    assert torch.equal(a["Z"], b["Z"]) or relF(b["Z"].cpu().numpy(), a["Z"].cpu().numpy()) < 1e-7
    for n in ("Q", "Hh"):
        assert relF(b[n].cpu().numpy(), a[n].cpu().numpy()) < 1e-6, n
    for n in ("dQ", "dZ", "g_W1", "g_W2", "g_b1", "g_b2", "g_w_rms"):
        assert relF(b[n].cpu().numpy(), a[n].cpu().numpy()) < 2e-6, n
    assert abs(a["loss"].item() - b["loss"].item()) <= 2e-7 * abs(a["loss"].item())


@pytest.mark.parametrize("ks,B,M,adam", [([8], 800, 30000, True), ([4, 12], 130, 2000, False)])
def test_deferred_parameter_update_is_bit_identical(ops, dev, ks, B, M, adam):
    """mlp_bwd(deferred_apply) + encoder_bwd (the network's parameter update runs on the encoder backward's epilogue
    warps) against mlp_bwd with its own update kernel: same association of the slab sums -> bit-identical parameters,
    moments, raw gradients and loss; and a pending update that no encoder_bwd picks up is run by the next call that
    needs the parameters (here mlp_fwd)."""
    from neural_admixture_b200._lib import MlpParams
    rng = np.random.default_rng(B)
    N, C, H, sumK = B + 20, 8, 1024, sum(ks)
    G = rand_genotypes(rng, N, M)
    pg = packed_from(ops, G, dev)
    idx = t(rng.permutation(N)[:B], dev, torch.int64)
    Z = t(rng.standard_normal((B, C)) * 0.3, dev)
    dQ0 = t(rng.standard_normal((B, sumK)) * 10, dev)
    labels = t(rng.integers(0, ks[0], size=B), dev, torch.int64)
    init = {"w_rms": 1 + 0.1 * rng.standard_normal(C), "W1": rng.standard_normal((H, C)) / np.sqrt(C),
            "b1": 0.1 * rng.standard_normal(H), "W2": rng.standard_normal((sumK, H)) / np.sqrt(H),
            "b2": 0.1 * rng.standard_normal(sumK)}
    V0 = rng.standard_normal((M, C)) / np.sqrt(M)
    ws = ws_for(ops, B, M, C, H, sumK, dev)
    results = []
    for mode in ("separate", "rides_along", "flushed"):
        prm = {n: t(v, dev) for n, v in init.items()}
        mom = {n: (torch.zeros_like(prm[n]), torch.zeros_like(prm[n])) for n in prm}
        grd = {n: torch.zeros_like(prm[n]) for n in prm}
        p = MlpParams()
        for n in prm:
            setattr(p, n, prm[n].data_ptr())
            setattr(p, "m_" + n, mom[n][0].data_ptr())
            setattr(p, "v_" + n, mom[n][1].data_ptr())
            setattr(p, "g_" + n, grd[n].data_ptr())
        rinv, Hh, Q = torch.empty(B, device=dev), torch.empty((B, H), device=dev), torch.empty((B, sumK), device=dev)
        Zc = Z.clone()
        ops.mlp_fwd(Zc, prm["w_rms"], prm["W1"], prm["b1"], prm["W2"], prm["b2"], ks, rinv, Hh, Q)
        dQ, dZ, loss = dQ0.clone(), torch.empty((B, C), device=dev), torch.zeros(1, device=dev)
        V, Vm, Vv = t(V0, dev), torch.zeros((M, C), device=dev), torch.zeros((M, C), device=dev)
        hyper = ops.adam_hyper(2e-3, 3) if adam else None
        ops.mlp_bwd(dQ, Q, Hh, Zc, rinv, ks, p, hyper, dZ, loss, ws, labels=labels, sup_weight=100.0,
                    deferred_apply=(mode != "separate"))
        if mode == "flushed":          # nobody picks the update up: the next reader of the parameters runs it
            ops.mlp_fwd(Zc, prm["w_rms"], prm["W1"], prm["b1"], prm["W2"], prm["b2"], ks, rinv, Hh, Q)
        ops.encoder_bwd(pg, dZ, V, Vm, Vv, ops.adam_hyper(2e-3, 3), ws, row_idx=idx)
        results.append({**{n: prm[n].clone() for n in prm}, **{"m_" + n: mom[n][0].clone() for n in prm},
                        **{"g_" + n: grd[n].clone() for n in prm}, "loss": loss.clone(), "V": V.clone(), "dZ": dZ.clone()})
    for other in results[1:]:
        for n, a in results[0].items():
            assert torch.equal(a, other[n]), n
    assert results[0]["loss"].item() > 0 and float(results[0]["g_W1"].abs().max()) > 0


@pytest.mark.parametrize("N,M,k,B,edge", [(64, 203, 5, 48, True), (300, 4099, 8, 300, False), (1000, 20000, 8, 800, True),
                                          (40, 1024, 3, 40, True), (90, 515, 12, 77, False), (20, 9, 2, 20, True),
                                          (900, 9001, 16, 800, True), (300, 2050, 9, 300, True)])
def test_decoder_step_grads(ops, dev, N, M, k, B, edge):
    rng = np.random.default_rng(M + k)
    G = rand_genotypes(rng, N, M)
    P = rng.uniform(0.02, 0.98, size=(M, k)).astype(np.float32)
    if edge:  # exact 0 / 1 entries: the 1e-12 floor (+-1e12 gradients) and the inclusive clamp mask
        P[5 % M, :] = 0.0
        P[min(17, M - 1), 0] = 0.0
        P[min(23, M - 1), k - 1] = 1.0
    Q = rng.dirichlet(0.3 * np.ones(k), size=B).astype(np.float32)
    idx = rng.permutation(N)[:B]
    pg = packed_from(ops, G, dev)
    sumK, q_off = k + 3, 2                                       # the head sits inside a wider Q / dQ
    Qw = np.zeros((B, sumK), dtype=np.float32)
    Qw[:, q_off:q_off + k] = Q
    dQ = torch.full((B, sumK), 7.0, device=dev)
    dP = torch.empty((M, k), device=dev)
    loss = torch.zeros(1, device=dev)
    P_d = t(P, dev)
    ws = ws_for(ops, B, M, 8, 64, sumK, dev)
    ops.decoder_step(pg, t(Qw, dev), dQ, q_off, k, P_d, None, None, None, loss, ws, row_idx=t(idx, dev, torch.int64),
                     dP_out=dP)
    x = orc.genotype_to_x(G[idx])
    l_ref, dQ_ref, dP_ref = orc.decoder_loss_grads(x, Q.astype(np.float64), P.astype(np.float64))
    # conditioning: G = (R - x) / (R (1 - R)) amplifies the rounding of raw = Q P^T when R is within ~1e-4 of 0 or 1.
    # The bar is "raw good to 4 fp32 ulps": tolerance = KERNEL_TOL + 8 x the effect of rounding raw to fp32 once.
    _, dQ_r32, dP_r32 = decoder_grads_with_fp32_raw(x, Q.astype(np.float64), P.astype(np.float64))
    assert abs(loss.item() - l_ref) < 1e-5 * abs(l_ref)
    assert relF(dQ[:, q_off:q_off + k].cpu().numpy(), dQ_ref) < KERNEL_TOL + 8 * relF(dQ_r32, dQ_ref)
    assert relF(dP.cpu().numpy(), dP_ref) < KERNEL_TOL + 8 * relF(dP_r32, dP_ref)
    assert torch.all(dQ[:, :q_off] == 7.0) and torch.all(dQ[:, q_off + k:] == 7.0)   # other heads' columns untouched
    assert np.array_equal(P_d.cpu().numpy(), P)                   # no Adam requested: P unchanged
    if edge:
        assert np.abs(dP_ref).max() > 1e10


@pytest.mark.parametrize("with_loss", [True, False])
def test_decoder_step_late_training_ranges(ops, dev, with_loss):
    """Late-training inputs: concentrated Q, allele frequencies at exactly 0 and very close to 0, so that many
    reconstructions have R (1 - R) between the 1e-12 floor of BCELoss' backward and 2^-15 (the kernel's mid path: fast
    gradient formula, one log per element), some below the floor (general path), the rest in the fast range.
    Gradients and loss against the fp64 oracle.  The gradient-only case also pins entries at exactly 1: reconstructions
    within 1e-7 of 1 are resolved by NO fp32 implementation (the loss of the reference's own fp32 path moves by 2.6 %
    when raw = Q P^T is rounded to fp32 once), so the loss is checked on the well-conditioned inputs and the gradients,
    with their conditioning term, on both."""
    rng = np.random.default_rng(77)
    N, M, k, B = 900, 12_007, 8, 800
    G = rand_genotypes(rng, N, M)
    P = rng.uniform(0.05, 0.95, size=(M, k))
    u = rng.random((M, k))
    P[u < 0.15] = 0.0
    if not with_loss:
        P[(u >= 0.15) & (u < 0.30)] = 1.0
    tiny = (u >= 0.30) & (u < 0.50)
    P[tiny] = 10.0 ** rng.uniform(-9, -4, size=int(tiny.sum()))
    P = P.astype(np.float32)
    Q = rng.dirichlet(0.05 * np.ones(k), size=B).astype(np.float32)
    idx = rng.permutation(N)[:B]
    pg = packed_from(ops, G, dev)
    dQ = torch.zeros((B, k), device=dev)
    dP = torch.empty((M, k), device=dev)
    loss = torch.zeros(1, device=dev)
    ws = ws_for(ops, B, M, 8, 64, k, dev)
    ops.decoder_step(pg, t(Q, dev), dQ, 0, k, t(P, dev), None, None, None, loss if with_loss else None, ws,
                     row_idx=t(idx, dev, torch.int64), dP_out=dP)
    x = orc.genotype_to_x(G[idx])
    Q64, P64 = Q.astype(np.float64), P.astype(np.float64)
    l_ref, dQ_ref, dP_ref = orc.decoder_loss_grads(x, Q64, P64)
    R = np.clip(Q64 @ P64.T, 0, 1)
    prod = R * (1 - R)
    mid = ((prod >= 1e-12) & (prod < 2.0 ** -15)).mean()
    assert mid > 0.02 and (prod < 1e-12).mean() > 1e-4 and (prod >= 2.0 ** -15).mean() > 0.2    # all three paths are hit
    _, dQ_r32, dP_r32 = decoder_grads_with_fp32_raw(x, Q64, P64)
    if with_loss:
        assert abs(loss.item() - l_ref) < 1e-5 * abs(l_ref)
    assert relF(dQ.cpu().numpy(), dQ_ref) < KERNEL_TOL + 8 * relF(dQ_r32, dQ_ref)
    assert relF(dP.cpu().numpy(), dP_ref) < KERNEL_TOL + 8 * relF(dP_r32, dP_ref)


@pytest.mark.parametrize("k,B", [(8, 800), (8, 1024), (12, 800), (5, 130)])
def test_decoder_step_without_loss_is_the_same_update(ops, dev, k, B):
    """loss = NULL (gradients only: the schedule used on epochs whose loss the reference does not print) must give
    bit-identical dQ and P updates.  The two cases run different instantiations of the kernel (3 compute warpgroups with
    the loss, 4 without; 4 raw slots for k <= 8 and B <= 896, 3 otherwise), so this also pins every warpgroup / slot
    combination the launcher can pick."""
    rng = np.random.default_rng(11)
    N, M = 1100, 30011
    G = rand_genotypes(rng, N, M)
    pg = packed_from(ops, G, dev)
    P0 = rng.uniform(0.0, 1.0, size=(M, k)).astype(np.float32)
    P0[7] = 0.0
    P0[9] = 1.0
    Q = t(rng.dirichlet(0.3 * np.ones(k), size=B).astype(np.float32), dev)
    idx = t(rng.permutation(N)[:B], dev, torch.int64)
    ws = ws_for(ops, B, M, 8, 64, k, dev)
    outs = []
    for with_loss in (True, False):
        P, Pm, Pv = t(P0, dev), torch.zeros((M, k), device=dev), torch.zeros((M, k), device=dev)
        dQ = torch.zeros((B, k), device=dev)
        loss = torch.zeros(1, device=dev)
        ops.decoder_step(pg, Q, dQ, 0, k, P, Pm, Pv, ops.adam_hyper(2e-3, 1), loss if with_loss else None, ws, row_idx=idx)
        outs.append((dQ.clone(), P.clone(), Pm.clone(), Pv.clone(), loss.item()))
    for a, b in zip(outs[0][:4], outs[1][:4]):
        assert torch.equal(a, b)
    assert outs[0][4] > 0 and outs[1][4] == 0.0


@pytest.mark.parametrize("N,M,C,B", [(64, 203, 8, 48), (300, 4099, 8, 300), (1000, 20000, 8, 800), (30, 6, 5, 30)])
def test_encoder_bwd(ops, dev, N, M, C, B):
    rng = np.random.default_rng(M + 1)
    G = rand_genotypes(rng, N, M)
    dZ = rng.standard_normal((B, C)).astype(np.float32)
    idx = rng.permutation(N)[:B]
    pg = packed_from(ops, G, dev)
    V = torch.zeros((M, C), device=dev)
    dV = torch.empty((M, C), device=dev)
    ops.encoder_bwd(pg, t(dZ, dev), V, None, None, None, ws_for(ops, B, M, C, 64, 8, dev),
                    row_idx=t(idx, dev, torch.int64), dV_out=dV)
    ref = orc.encoder_bwd(orc.genotype_to_x(G[idx]), dZ.astype(np.float64))
    assert relF(dV.cpu().numpy(), ref) < KERNEL_TOL


def test_adam_fused_matches_oracle(ops, dev):
    """Adam + restrict_P fused into the decoder / encoder-backward kernels, three consecutive steps."""
    rng = np.random.default_rng(3)
    N, M, k, C, B = 80, 1029, 6, 8, 64
    G = rand_genotypes(rng, N, M)
    pg = packed_from(ops, G, dev)
    P = rng.uniform(0.0, 1.0, size=(M, k))
    V = rng.standard_normal((M, C)) / np.sqrt(M)
    P_d, Pm, Pv = t(P, dev), torch.zeros((M, k), device=dev), torch.zeros((M, k), device=dev)
    V_d, Vm, Vv = t(V, dev), torch.zeros((M, C), device=dev), torch.zeros((M, C), device=dev)
    P_o, V_o = P_d.cpu().numpy().astype(np.float64), V_d.cpu().numpy().astype(np.float64)
    mo = {n: np.zeros_like(a) for n, a in (("P", P_o), ("V", V_o))}
    vo = {n: np.zeros_like(a) for n, a in (("P", P_o), ("V", V_o))}
    ws = ws_for(ops, B, M, C, 64, k, dev)
    for step in range(1, 4):
        idx = rng.permutation(N)[:B]
        Q = rng.dirichlet(0.3 * np.ones(k), size=B).astype(np.float32)
        dZ = rng.standard_normal((B, C)).astype(np.float32)
        x = orc.genotype_to_x(G[idx])
        _, _, dP_ref = orc.decoder_loss_grads(x, Q.astype(np.float64), P_o)
        orc.adam_update(P_o, dP_ref, mo["P"], vo["P"], step, 2e-3)
        np.clip(P_o, 0, 1, out=P_o)
        orc.adam_update(V_o, orc.encoder_bwd(x, dZ.astype(np.float64)), mo["V"], vo["V"], step, 2e-3)
        hyper = ops.adam_hyper(2e-3, step)
        dQ = torch.zeros((B, k), device=dev)
        loss = torch.zeros(1, device=dev)
        ops.decoder_step(pg, t(Q, dev), dQ, 0, k, P_d, Pm, Pv, hyper, loss, ws, row_idx=t(idx, dev, torch.int64))
        ops.encoder_bwd(pg, t(dZ, dev), V_d, Vm, Vv, hyper, ws, row_idx=t(idx, dev, torch.int64))
        assert relF(P_d.cpu().numpy(), P_o) < STEP_TOL, step
        assert relF(V_d.cpu().numpy(), V_o) < STEP_TOL, step
        assert P_d.min().item() >= 0.0 and P_d.max().item() <= 1.0


def test_loglikelihood(ops, dev):
    rng = np.random.default_rng(9)
    N, M, k = 210, 1333, 5
    G = rand_genotypes(rng, N, M, miss=0.05)
    P = rng.uniform(0, 1, size=(M, k)).astype(np.float32)
    Q = rng.dirichlet(np.ones(k), size=N).astype(np.float32)
    pg = packed_from(ops, G, dev)
    ll = ops.loglikelihood(pg, t(Q, dev), t(P, dev), ws_for(ops, 64, M, 8, 64, k, dev))
    ref = orc.loglikelihood(G, P, Q)
    assert abs(ll - ref) < 1e-9 * abs(ref)


def test_loglikelihood_reference_golden(ops, dev):
    """nadm_loglikelihood against the value the reference's Cython routine returned (tests/golden/loglik.npz)."""
    g = load_golden("loglik.npz")
    pg = packed_from(ops, g["G"], dev)
    for K in (3, 8):
        ll = ops.loglikelihood(pg, t(g[f"Q{K}"], dev), t(g[f"P{K}"], dev), ws_for(ops, 64, g["G"].shape[1], 8, 64, K, dev))
        assert abs(ll - float(g[f"ll{K}"])) < 1e-9 * abs(float(g[f"ll{K}"])), K


def test_error_behaviour(ops, dev):
    from neural_admixture_b200._lib import NadmError
    pg = ops.PackedGenotypes.empty(8, 64, dev)
    Z = torch.empty((8, 8), device=dev)
    with pytest.raises(NadmError, match="unsupported"):
        ops.encoder_fwd(pg, torch.zeros((64, 40), device=dev), torch.empty((8, 40), device=dev),
                        torch.empty(1 << 20, dtype=torch.uint8, device=dev), row0=0, B=8)
    with pytest.raises(NadmError, match="workspace"):
        ops.encoder_fwd(pg, torch.zeros((64, 8), device=dev), Z, torch.empty(16, dtype=torch.uint8, device=dev),
                        row0=0, B=8)
    with pytest.raises(NadmError, match="CUDA tensors only"):
        ops.encoder_fwd(pg, torch.zeros((64, 8)), Z, torch.empty(1 << 20, dtype=torch.uint8, device=dev), row0=0, B=8)


# ---------------------------------------------------------------------------------------------------------------
# golden fixtures produced by the reference itself
# ---------------------------------------------------------------------------------------------------------------
def _engine_from_fixture(g, dev, epochs=None, G=None):
    """NeuralAdmixture on the fixture's data, starting from the fixture's initial parameters (the reference's MLP
    initialisation depends on torch's global RNG state at construction, so parameters are loaded, not re-drawn)."""
    from neural_admixture_b200 import ops as _ops
    from neural_admixture_b200.model.neural_admixture import NeuralAdmixture
    ks = [int(k) for k in g["ks"]]
    init = sub(g, "init/")
    G = g["G"] if G is None else G
    N, M = G.shape
    k = ks[0] if len(ks) == 1 else None
    na = NeuralAdmixture(k, int(g["epochs"]) if epochs is None else epochs, int(g["batch"]), float(g["lr"]), dev,
                         int(g["seed"]), 0, True, "nadm_b200", None if k else min(ks), None if k else max(ks))
    na.keep_loss_history = True
    orig = na.initialize_model

    def init_and_load(P, hidden_size, num_features, V, ks_list):
        orig(P, hidden_size, num_features, V, ks_list)
        na.raw_model.load_state_dict({n: torch.as_tensor(a) for n, a in init.items()})
        na.raw_model.bind()

    na.initialize_model = init_and_load
    P_init = np.concatenate([init[f"decoders.decoders.{i}.weight"].T for i in range(len(ks))], axis=0)
    packed = _ops.PackedGenotypes.from_unpacked_host(torch.as_tensor(G), dev)
    H = init["common_encoder.0.weight"].shape[0]
    pops = torch.as_tensor(g["pops"], dtype=torch.int64) if "pops" in g else None
    Qs, Ps, raw = na.launch_training(t(P_init, dev), packed, H, init["V"].shape[1], t(init["V"], dev), M, N, pops)
    return na, Qs, Ps, raw


def test_step_fixture_two_steps(ops, dev):
    """tests/golden/step_k5.npz: two steps of the reference's Q_P + fused Adam + restrict_P on one batch, with missing
    codes and exact 0/1 entries of P."""
    from neural_admixture_b200.model.neural_admixture import NeuralAdmixture
    g = load_golden("step_k5.npz")
    G, init = g["G"], sub(g, "init/")
    B, M = G.shape
    na = NeuralAdmixture(5, 0, B, float(g["lr"]), dev, 0, 0, True, "nadm_b200", None, None)
    na.M, na.N = M, B
    na.packed = packed_from(ops, G, dev)
    na._train_bufs = {}
    P_init = init["decoders.decoders.0.weight"].T
    na.initialize_model(t(P_init, dev), init["common_encoder.0.weight"].shape[0], 8, t(init["V"], dev), [5])
    na.raw_model.load_state_dict({n: torch.as_tensor(a) for n, a in init.items()})
    na.raw_model.bind()
    na.optimizer = na.raw_model.create_custom_adam(device=dev, lr=float(g["lr"]))
    idx = torch.arange(B, device=dev)
    loss = torch.zeros(1, device=dev)
    for step, key in enumerate(["after1/", "after2/"]):
        na._train_step(idx, None, loss)
        assert abs(loss.item() - g["losses"][step]) < 2e-5 * abs(g["losses"][step])
        ref = sub(g, key)
        sd = {n: a.detach().cpu().numpy() for n, a in na.raw_model.state_dict().items()}
        for name, a in ref.items():
            assert relF(sd[name], a) < STEP_TOL, (key, name)


@pytest.mark.parametrize("fixture", ["train_k3.npz", "train_k3to5.npz", "train_sup_k3.npz", "train_k4to12.npz"])
def test_training_fixtures(dev, fixture):
    g = load_golden(fixture)
    na, Qs, Ps, raw = _engine_from_fixture(g, dev)
    np.testing.assert_allclose(na.loss_history, g["epoch_losses"], rtol=5e-5)
    for i in range(len(g["ks"])):
        assert relF(Qs[i], g[f"Q/{i}"]) < QP_TOL, i
        assert relF(Ps[i], g[f"P/{i}"]) < QP_TOL, i
    final = sub(g, "final/")
    assert relF(raw.V.detach().cpu().numpy(), final["V"]) < QP_TOL
    # the sampler stream is the reference's (loaders.py:29-30)
    from neural_admixture_b200.model.neural_admixture import NeuralAdmixture
    assert na.generic_kernel_launches == 0              # every head (K up to 12 here) stays on the tensor-core kernels
    fresh = NeuralAdmixture(3, 1, 8, 1e-3, dev, int(g["seed"]), 0, True, None, 3, 5)
    for e in range(g["orders"].shape[0]):
        assert np.array_equal(fresh.epoch_order(g["G"].shape[0]).numpy(), g["orders"][e])


def test_demo_fixture(dev):
    """The reference's only integration test (demo/run_demo.sh + run_diagnostics.py): K=7, 5 epochs, seed 42 on the
    shipped BED (105 x 8451), from the reference's own RSVD/GMM initialisation."""
    g = load_golden("demo_k7.npz")
    G = orc.unpack2bit(g["G_packed"], int(g["M"]))
    na, Qs, Ps, raw = _engine_from_fixture(g, dev, G=G)
    np.testing.assert_allclose(na.loss_history, g["epoch_losses"], rtol=5e-5)
    assert relF(Qs[0], g["Q/0"]) < QP_TOL
    assert relF(Ps[0], g["P/0"]) < QP_TOL
    assert relF(Qs[0], g["expected_Q"]) < 5e-3      # the shipped .expected file, at the level the reference itself reaches
    assert (Ps[0] == 0.0).mean() > 0.2               # the clamp / 1e-12-floor branch is hit constantly here


def test_infer_forward_surface(ops, dev):
    """Q_P in inference mode: reference signature forward(uint8 B x M) -> (probs_list, X) (inference.py:56-58,75)."""
    from neural_admixture_b200.model.neural_admixture import Q_P
    g = load_golden("train_k3to5.npz")
    ks = [int(k) for k in g["ks"]]
    final = {n: a for n, a in sub(g, "final/").items() if not n.startswith("decoders")}   # what main.py:41-42 saves
    H, C = final["common_encoder.0.weight"].shape
    model = Q_P(H, C, ks_list=ks, V=torch.as_tensor(final["V"]), is_train=False)
    model.load_state_dict({n: torch.as_tensor(a) for n, a in final.items()})
    model.to(dev)
    X = torch.as_tensor(g["G"][:100], device=dev)
    probs, Xo = model(X)
    assert Xo is X and len(probs) == len(ks)
    for i in range(len(ks)):
        assert relF(probs[i].cpu().numpy(), g[f"Q/{i}"][:100]) < QP_TOL


# ---------------------------------------------------------------------------------------------------------------
# PLINK .bed -> packed device layout (next row f4): bit-exact vs the oracle and vs the reference reader's output
# ---------------------------------------------------------------------------------------------------------------
def test_bed_to_packed_golden(ops, dev, tmp_path):
    from neural_admixture_b200.src import snp_reader
    g = load_golden("bed_demo_slices.npz")
    N, M = int(g["N"]), int(g["M"])
    for tag in "ab":
        base = tmp_path / f"case_{tag}"
        with open(str(base) + ".bed", "wb") as f:
            f.write(bytes([0x6C, 0x1B, 0x01]))
            f.write(g[f"bed_{tag}"].tobytes())
        with open(str(base) + ".fam", "w") as f:
            f.write("".join(f"f{i} i{i} 0 0 0 -9\n" for i in range(N)))
        pg = snp_reader.read_bed_packed(str(base) + ".bed", dev, chunk_snps=512)     # 3 chunks, ragged last one
        assert pg.N == N and pg.M == M
        want = orc.pack2bit(g[f"G_{tag}"])                     # the reference's matrix through its pack layout
        got = pg.storage.cpu().numpy()
        assert np.array_equal(got[:, :want.shape[1]], want)
        assert not got[:, want.shape[1]:].any()                # zero row tails
        # a rank's SNP slice (sharded runs) holds the same codes as the slice of the full matrix, when unflipped
        if tag == "a":
            sl = snp_reader.read_bed_packed(str(base) + ".bed", dev, col0=256, col1=1300)
            assert np.array_equal(sl.storage.cpu().numpy()[:, :(1300 - 256 + 3) // 4], orc.pack2bit(g["G_a"][:, 256:1300]))


@pytest.mark.parametrize("N,M", [(1, 1), (7, 130), (256, 128), (1000, 1000), (1027, 333)])
def test_bed_to_packed_random(ops, dev, N, M):
    rng = np.random.default_rng(N * 7 + M)
    nb = (N + 3) // 4
    bed = rng.integers(0, 256, size=(M, nb), dtype=np.uint8)
    raw = orc.read_bed(bed, N)
    pg = ops.PackedGenotypes.empty(N, M, dev)
    pg.storage.fill_(0xAB)                                     # the kernel must overwrite every byte it owns
    counts = torch.zeros(4, dtype=torch.int64, device=dev)
    ops.bed_to_packed(t(bed, dev, torch.uint8), N, pg, counts=counts)
    want = orc.pack2bit(raw)
    got = pg.storage.cpu().numpy()
    assert np.array_equal(got[:, :want.shape[1]], want)
    pc128 = ((M + 127) // 128) * 32                            # tiles are 128 SNPs wide: zero up to the tile edge
    assert not got[:, want.shape[1]:min(pc128, got.shape[1])].any()
    c = counts.cpu().numpy()
    assert [int(c[1]), int(c[2]), int(c[3])] == [int((raw == v).sum()) for v in (1, 2, 3)]
    # flip: 2 - G in uint8, packed (missing stays 3); in-place flip of the packed matrix gives the same
    pg2 = ops.PackedGenotypes.empty(N, M, dev)
    ops.bed_to_packed(t(bed, dev, torch.uint8), N, pg2, flip=True)
    want_f = orc.pack2bit((2 - raw).astype(np.uint8))
    assert np.array_equal(pg2.storage.cpu().numpy()[:, :want.shape[1]], want_f)
    pg.storage[:, pc128:].zero_() if pc128 < pg.pitch else None
    ops.flip_packed(pg)
    assert np.array_equal(pg.storage.cpu().numpy()[:, :want.shape[1]], want_f)
    assert not pg.storage.cpu().numpy()[:, want.shape[1]:].any()


# ---------------------------------------------------------------------------------------------------------------
# randomized SVD on the packed matrix (next row f2): products and full RSVD vs the reference's own outputs
# ---------------------------------------------------------------------------------------------------------------
def test_rsvd_products_and_rsvd_golden(ops, dev):
    from neural_admixture_b200.src import svd
    g, b = load_golden("rsvd_demo_slices.npz"), load_golden("bed_demo_slices.npz")
    for tag, mv in (("a", 3), ("b", 255)):
        A = b[f"G_{tag}"]                                          # reference matrix: missing = 3 / 255 (flipped)
        pg = packed_from(ops, A, dev)                              # packing keeps the two low bits: 255 -> 3
        ws = ws_for(ops, 1024, A.shape[1], 8, 8, 8, dev)
        Y = ops.geno_matmul(pg, t(g[f"Omega_{tag}"], dev), ws, mv).cpu().numpy()
        Bm = ops.geno_matmul_t(pg, t(g[f"QT_{tag}"], dev), ws, mv).cpu().numpy()
        assert relF(Y, orc.multiply_A_omega(A, g[f"Omega_{tag}"])) < 2e-7          # vs fp64 oracle
        assert relF(Bm, orc.multiply_QT_A(g[f"QT_{tag}"], A)) < 2e-7
        assert relF(Y, g[f"Y_{tag}"]) < 1e-6 and relF(Bm, g[f"B_{tag}"]) < 1e-6     # vs the reference's fp32 loops
        Vt = svd.RSVD(pg, A.shape[0], A.shape[1], 8, 42, missing_value=mv)
        assert Vt.shape == (8, A.shape[1])
        assert relF(Vt, g[f"Vt_{tag}"]) < 1e-4
        assert relF(Vt, orc.rsvd(A, 8, 42)) < 1e-4


def test_geno_matmul_row_batches(ops, dev):
    """More rows than one launch takes (1024): forward batches are independent, transposed batches accumulate."""
    rng = np.random.default_rng(8)
    N, M, K = 2500, 1111, 11
    A = rand_genotypes(rng, N, M)
    pg = packed_from(ops, A, dev)
    ws = ws_for(ops, 1024, M, 8, 8, 8, dev)
    Om = rng.standard_normal((M, K)).astype(np.float32)
    QT = rng.standard_normal((K, N)).astype(np.float32)
    assert relF(ops.geno_matmul(pg, t(Om, dev), ws, 3).cpu().numpy(), orc.multiply_A_omega(A, Om)) < 2e-7
    assert relF(ops.geno_matmul_t(pg, t(QT, dev), ws, 3).cpu().numpy(), orc.multiply_QT_A(QT, A)) < 3e-7


def test_train_packed_pipeline_matches_host_pipeline(dev, tmp_path):
    """.bed -> packed -> device RSVD -> device PCA projection + GMM init -> training, never materialising N x M bytes,
    gives the same initial P and the same fit as the pipeline fed with the reference reader's uint8 matrix."""
    from neural_admixture_b200.model import train as tr
    from neural_admixture_b200.src import snp_reader, svd
    b = load_golden("bed_demo_slices.npz")
    N, M = int(b["N"]), int(b["M"])
    base = tmp_path / "case_a"
    with open(str(base) + ".bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        f.write(b["bed_a"].tobytes())
    with open(str(base) + ".fam", "w") as f:
        f.write("".join(f"f{i} i{i} 0 0 0 -9\n" for i in range(N)))
    pg = snp_reader.read_bed_packed(str(base) + ".bed", dev)
    V = svd.RSVD(pg, N, M, 8, 42)                                   # C x M
    P_dev = tr.gmm_initial_P_packed(pg, V, [4], 42)
    P_host = tr.gmm_initial_P(b["G_a"], V, [4], 8, 42)
    assert relF(P_dev, P_host) < 1e-5
    torch.manual_seed(1)
    Ps1, Qs1, _ = tr.train_packed(3, 64, 2e-3, 4, 42, pg, 64, V)
    torch.manual_seed(1)
    Ps2, Qs2, _ = tr.train(3, 64, 2e-3, 4, 42, torch.as_tensor(b["G_a"]), dev, 0, 64, True, V, None, n_components=8)
    assert relF(Qs1[0], Qs2[0]) < 1e-4 and relF(Ps1[0], Ps2[0]) < 1e-4


# ---------------------------------------------------------------------------------------------------------------
# CUDA-graph replayed steps == eager steps (same kernels; Adam coefficients computed on the device)
# ---------------------------------------------------------------------------------------------------------------
def test_graph_replayed_steps_match_eager(dev, monkeypatch):
    from neural_admixture_b200 import ops
    from neural_admixture_b200.model.neural_admixture import NeuralAdmixture
    rng = np.random.default_rng(3)
    N, M, K, C, H, B = 333, 2051, 4, 8, 64, 100                # 4 steps per epoch, ragged last batch (33 rows)
    G = rand_genotypes(rng, N, M)
    V = np.linalg.qr(rng.standard_normal((M, C)))[0].astype(np.float32)
    P0 = rng.uniform(0.05, 0.95, size=(K, M)).astype(np.float32)
    y = torch.as_tensor(rng.integers(0, K, size=N))
    outs = []
    for graph in (True, False):
        monkeypatch.setattr(NeuralAdmixture, "use_graph", graph)
        torch.manual_seed(5)
        na = NeuralAdmixture(K, 3, B, 2e-3, dev, 7, 0, True, "nadm_b200", None, None)
        na.keep_loss_history = True
        pg = packed_from(ops, G, dev)
        Qs, Ps, _ = na.launch_training(torch.as_tensor(P0, device=dev), pg, H, C, torch.as_tensor(V, device=dev), M, N, y)
        assert na.use_graph == graph                              # no silent fall-back to eager
        assert (na.graph_kernel_launches > 0) == graph
        outs.append((Qs[0], Ps[0], np.array(na.loss_history), na.optimizer.step_count))
    assert outs[0][3] == outs[1][3] == 12
    assert relF(outs[0][0], outs[1][0]) < 1e-6 and relF(outs[0][1], outs[1][1]) < 1e-6
    np.testing.assert_allclose(outs[0][2], outs[1][2], rtol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
# host-fed steps (train_from_host: two streams, staged batches; what bench.py's e2e leg times) == resident-matrix steps
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("graph", [False, True])
def test_host_fed_steps_match_resident_steps(dev, monkeypatch, graph):
    from neural_admixture_b200 import ops
    from neural_admixture_b200.model.neural_admixture import NeuralAdmixture
    rng = np.random.default_rng(11)
    N, M, K, C, H, B, S = 300, 2051, 4, 8, 64, 100, 3
    G = rand_genotypes(rng, N, M)
    V = np.linalg.qr(rng.standard_normal((M, C)))[0].astype(np.float32)
    P0 = rng.uniform(0.05, 0.95, size=(K, M)).astype(np.float32)
    order = torch.as_tensor(rng.permutation(N)[: S * B].astype(np.int64))
    # graph = False: eager steps on both sides (the same five calls per step); True: replayed CUDA graphs on both sides
    # (the host-fed loop then never waits for the device between steps: copies, steps and loss read-backs are ordered
    # by events only)
    monkeypatch.setattr(NeuralAdmixture, "use_graph", graph)
    res = []
    for host_fed in (False, True):
        torch.manual_seed(5)
        na = NeuralAdmixture(K, 1, B, 2e-3, dev, 7, 0, True, "nadm_b200", None, None)
        pg = packed_from(ops, G, dev)
        na.prepare(torch.as_tensor(P0, device=dev), pg, H, C, torch.as_tensor(V, device=dev), M, N)
        if host_fed:
            rows = pg.storage.cpu()                                  # N x pitch, the PackedGenotypes row layout
            batches = [rows[order[s * B:(s + 1) * B]].contiguous().pin_memory() for s in range(S)]
            losses = np.array(na.train_from_host(batches))
        else:
            losses = na.train_steps(order.to(dev), S, True).cpu().numpy()
        torch.cuda.synchronize()
        assert na.use_graph == graph and (na.graph_kernel_launches > 0) == graph
        res.append((na.raw_model.V.detach().cpu().numpy().copy(),
                    na.raw_model.decoders.decoders[0].weight.detach().cpu().numpy().copy(), losses))
    assert relF(res[1][0], res[0][0]) < 1e-6 and relF(res[1][1], res[0][1]) < 1e-6
    np.testing.assert_allclose(res[1][2], res[0][2], rtol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
# full-size properties (config 2: 10k x 100k, K = 8; B = 800) — no oracle needed
# ---------------------------------------------------------------------------------------------------------------
def test_fullsize_properties(ops, dev):
    N, M, C, k, B = 10000, 100000, 8, 8, 800
    gen = torch.Generator(device=dev).manual_seed(42)
    pg = ops.PackedGenotypes.empty(N, M, dev)
    for r0 in range(0, N, 1000):
        codes = torch.randint(0, 4, (1000, M), dtype=torch.uint8, device=dev, generator=gen)
        ops.pack2bit(codes, pg.storage[r0:r0 + 1000], M)
    idx = torch.randperm(N, device=dev, generator=gen)[:B].contiguous()
    ws = ws_for(ops, B, M, C, 1024, k, dev)
    V1 = torch.randn((M, C), device=dev, generator=gen) / M ** 0.5
    V2 = torch.randn((M, C), device=dev, generator=gen) / M ** 0.5
    Z1, Z2, Z12 = (torch.empty((B, C), device=dev) for _ in range(3))
    ops.encoder_fwd(pg, V1, Z1, ws, row_idx=idx)
    ops.encoder_fwd(pg, V2, Z2, ws, row_idx=idx)
    ops.encoder_fwd(pg, (V1 + V2).contiguous(), Z12, ws, row_idx=idx)
    assert relF((Z1 + Z2).cpu().numpy(), Z12.cpu().numpy()) < KERNEL_TOL          # linearity in V
    # V = 1: Z[b, c] = sum_m x[b, m]  -> exact dosage count / 2 per row (integers < 2^24: exact in fp32 partials)
    ones = torch.ones((M, C), device=dev)
    ops.encoder_fwd(pg, ones, Z1, ws, row_idx=idx)
    un = torch.empty((B, M), dtype=torch.uint8, device=dev)
    ops.unpack2bit(pg.storage[idx].contiguous()[:, :(M + 3) // 4].contiguous(), un)
    dosage = torch.where(un == 3, torch.zeros_like(un), un).sum(dim=1, dtype=torch.int64).double() / 2
    assert torch.equal(Z1[:, 0].double(), dosage) and torch.equal(Z1[:, 0], Z1[:, C - 1])
    # adjoint identity: <dZ, X V> == <X^T dZ, V>
    dZ = torch.randn((B, C), device=dev, generator=gen)
    dV = torch.empty((M, C), device=dev)
    ops.encoder_bwd(pg, dZ, torch.zeros((M, C), device=dev), None, None, None, ws, row_idx=idx, dV_out=dV)
    ops.encoder_fwd(pg, V1, Z1, ws, row_idx=idx)
    lhs = (dZ.double() * Z1.double()).sum().item()
    rhs = (dV.double() * V1.double()).sum().item()
    assert abs(lhs - rhs) < 1e-4 * max(abs(lhs), (dZ.double().norm() * Z1.double().norm()).item() * 1e-2)
    # decoder: with Q = one-hot(0) for every row, R[b, m] = P[m, 0]: dP[:, 1:] == 0 and dQ[b, j] = sum_m G[b,m] P[m,j];
    # also <dQ, Q> == <dP, P> (both equal sum G * R).
    Q = torch.zeros((B, k), device=dev)
    Q[:, 0] = 1.0
    P = torch.rand((M, k), device=dev, generator=gen) * 0.9 + 0.05
    dQ = torch.zeros((B, k), device=dev)
    dP = torch.empty((M, k), device=dev)
    loss = torch.zeros(1, device=dev)
    ops.decoder_step(pg, Q, dQ, 0, k, P, None, None, None, loss, ws, row_idx=idx, dP_out=dP)
    assert torch.count_nonzero(dP[:, 1:]).item() == 0
    a = (dQ.double() * Q.double()).sum().item()
    b = (dP.double() * P.double()).sum().item()
    assert abs(a - b) < 1e-4 * abs(a)
    # closed form of the loss for this Q: counts of each code per SNP column
    codes = torch.where(un == 3, torch.zeros_like(un), un)
    p0 = P[:, 0].double()
    n1 = (codes == 1).sum(0).double()
    n2 = (codes == 2).sum(0).double()
    n0 = B - n1 - n2
    closed = -(n2 * p0.log() + n0 * (1 - p0).log() + 0.5 * n1 * (p0.log() + (1 - p0).log())).sum().item()
    assert abs(loss.item() - closed) < 2e-5 * abs(closed)


# ---------------------------------------------------------------------------------------------------------------
# config 2 at FULL size (BASELINE.json configs[1]: 10k x 100k, K = 8, B = 800): three consecutive optimisation steps of
# the product loop against the fp64 oracle on the same rows ("cfg2 full-size ... per-step")
# ---------------------------------------------------------------------------------------------------------------
def test_cfg2_fullsize_steps_vs_oracle(ops, dev):
    from neural_admixture_b200.model.neural_admixture import NeuralAdmixture
    N, M, C, K, B, H, S = 10_000, 100_000, 8, 8, 800, 1024, 3
    gen = torch.Generator(device=dev).manual_seed(2)
    # admixture-model genotypes generated on the device, 0.5 % missing
    Qt = torch.distributions.Dirichlet(torch.full((K,), 0.2)).sample((N,)).to(dev)
    Pt = torch.rand((K, M), device=dev, generator=gen) * 0.96 + 0.02
    pg = ops.PackedGenotypes.empty(N, M, dev)
    for r0 in range(0, N, 1000):
        pr = Qt[r0:r0 + 1000] @ Pt
        g_ = (torch.rand(pr.shape, device=dev, generator=gen) < pr).to(torch.uint8)
        g_ += (torch.rand(pr.shape, device=dev, generator=gen) < pr).to(torch.uint8)
        g_[torch.rand(pr.shape, device=dev, generator=gen) < 0.005] = 3
        ops.pack2bit(g_, pg.storage[r0:r0 + 1000], M)
    V = torch.linalg.qr(torch.randn((M, C), device=dev, generator=gen))[0].contiguous()
    P0 = (torch.rand((K, M), device=dev, generator=gen) * 0.9 + 0.05).contiguous()
    torch.manual_seed(3)
    na = NeuralAdmixture(K, 1, B, 2e-3, dev, 3, 0, True, "nadm_b200", None, None)
    na.prepare(P0, pg, H, C, V, M, N)
    sd0 = {n: a.detach().cpu().numpy().astype(np.float64) for n, a in na.raw_model.state_dict().items()}
    st = state_from_sd(sd0, [K])
    order = torch.randperm(N, generator=torch.Generator().manual_seed(4))[: S * B]
    losses = na.train_steps(order.to(dev), S, True).cpu().numpy()
    assert na.generic_kernel_launches == 0 and na.graph_fallback is None
    for s_ in range(S):
        idx = order[s_ * B:(s_ + 1) * B]
        un = torch.empty((B, M), dtype=torch.uint8, device=dev)
        ops.unpack2bit(pg.storage[idx.to(dev)].contiguous(), un)
        l_ref, _ = orc.train_step(st, un.cpu().numpy(), 2e-3)
        assert abs(losses[s_] - l_ref) < 2e-5 * abs(l_ref), s_
    sd = {n: a.detach().cpu().numpy() for n, a in na.raw_model.state_dict().items()}
    for name, p_ in st.params().items():
        assert relF(sd[sd_name(name)], p_) < STEP_TOL, name
    # Q of the first 2048 samples through the post-training pass, north-star bar
    Qd = na.raw_model.infer_packed(ops.PackedGenotypes(pg.storage[:2048], 2048, M), 1024)[0].cpu().numpy()
    un = torch.empty((2048, M), dtype=torch.uint8, device=dev)
    ops.unpack2bit(pg.storage[:2048].contiguous(), un)
    assert relF(Qd, orc.infer_Q(st, un.cpu().numpy())[0]) < QP_TOL


# ---------------------------------------------------------------------------------------------------------------
# `infer` entry point end to end (reference src/inference.py:16-99): checkpoint + config written the way training
# writes them (main.py:41-43), a PLINK .bed on disk, `{out_name}.{K}.Q` read back
# ---------------------------------------------------------------------------------------------------------------
def test_inference_main_end_to_end(ops, dev, tmp_path):
    import argparse
    import time
    from neural_admixture_b200.model.neural_admixture import Q_P
    from neural_admixture_b200.src import inference
    g = load_golden("train_k3to5.npz")
    ks = [int(k) for k in g["ks"]]
    final = sub(g, "final/")
    saved = {n: torch.as_tensor(a) for n, a in final.items() if not n.startswith("decoders")}      # main.py:41
    torch.save(saved, tmp_path / "run.pt")
    H, C = final["common_encoder.0.weight"].shape
    Q_P(H, C, ks_list=ks, V=torch.as_tensor(final["V"]), is_train=False).save_config("run", str(tmp_path))
    G = g["G"]                                                    # N x M codes the fixture's Q belongs to
    N, M = G.shape
    # the reader keeps the matrix when its mean (missing counted as 3) is < 1 and maps g -> 2 - g otherwise
    # (snp_reader.py:110): store the orientation that reads back as G
    Gfile = G if G.mean() < 1 else np.where(G == 3, 3, 2 - G).astype(np.uint8)
    assert (Gfile.mean() < 1) == (Gfile is G)
    # write it as a .bed: SNP-major, 4 samples per byte, fields 00 -> 2, 01 -> missing, 10 -> 1, 11 -> 0 (utils.pyx:43-68)
    field = np.array([3, 2, 0, 1], dtype=np.uint8)[Gfile.T]       # code -> bed field
    pad = np.zeros((M, (-N) % 4), dtype=np.uint8)
    f4 = np.concatenate([field, pad], axis=1).reshape(M, -1, 4)
    payload = (f4[:, :, 0] | (f4[:, :, 1] << 2) | (f4[:, :, 2] << 4) | (f4[:, :, 3] << 6)).astype(np.uint8)
    with open(tmp_path / "data.bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        f.write(payload.tobytes())
    (tmp_path / "data.fam").write_text("".join(f"f{i} i{i} 0 0 0 -9\n" for i in range(N)))
    args = argparse.Namespace(data_path=str(tmp_path / "data.bed"), out_name="proj", save_dir=str(tmp_path), name="run",
                              seed=42, batch_size=64, num_gpus=1)
    assert inference.main(args, time.time()) == 0
    for i, k in enumerate(ks):
        Q = np.loadtxt(tmp_path / f"proj.{k}.Q")
        assert Q.shape == (N, k)
        assert relF(Q, g[f"Q/{i}"]) < QP_TOL, k
