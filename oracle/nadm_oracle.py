"""CPU oracle for the Neural ADMIXTURE per-minibatch hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy (float64 by default), the algorithm that the reference executes through
PyTorch eager ops.  It is the checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under ``neural_admixture_b200/``
imports it, and the product path raises if its CUDA library is missing.

Where the arithmetic really lives: the reference delegates every hot-path operation to PyTorch (third-party,
not vendored under /root/reference; pinned ``torch<=2.4.0,>2.0.0`` in setup.cfg:26, container has 2.11.0).
Each function below cites the reference call site it follows and restates the published semantics of the torch
op behind it (``BCELoss`` log clamp at -100, its backward's ``max((1-x)*x, 1e-12)`` floor, ``clamp_`` backward's
inclusive mask, ``RMSNorm``, fused ``Adam``).

Pinning: the reference's test-suite holds no vectors for this path (tests/test_placeholder.py is ``assert True``).
The oracle is therefore pinned against outputs of the reference itself, imported in the build container from
/root/reference by ``tests/golden/make_golden.py`` (fixtures committed under ``tests/golden/``; checked by
``tests/test_oracle_golden.py``), and against the demo's shipped ``demo_run.7.{Q,P}.expected`` files at the
~1e-3 level the reference itself reproduces them to.  The 2-bit pack/unpack layout (pack2bit.cu) is CUDA-only in
the reference and cannot be executed in the build container, but it compiles there: ``oracle/build_ref.py`` builds the
reference's file unmodified into ``oracle/_ref/pack2bit_ref.so``, and on the GPU box
``tests/test_gpu_parity.py::test_pack_unpack_against_the_compiled_reference_kernels`` pins this file's ``pack2bit`` /
``unpack2bit`` and the library's kernels to the reference's own, bit for bit.

All citations are ``path:line`` under /root/reference/neural_admixture/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------------------------------
# 2-bit genotype packing  (src/utils_c/pack2bit.cu:10-36 pack, :38-62 unpack)
# --------------------------------------------------------------------------------------------------------------


def packed_cols(M: int) -> int:
    """ceil(M/4) bytes per sample row (pack2bit.cu:72, train.py:121)."""
    return (M + 3) // 4


def pack2bit(G: np.ndarray) -> np.ndarray:
    """uint8 N x M genotype codes -> N x ceil(M/4) bytes; SNP 4c+i sits in bits 2i..2i+1 of byte c,
    only the two low bits of each code are kept, the tail of the last byte is zero (pack2bit.cu:26-35)."""
    G = np.ascontiguousarray(G, dtype=np.uint8)
    N, M = G.shape
    pc = packed_cols(M)
    pad = np.zeros((N, pc * 4), dtype=np.uint8)
    pad[:, :M] = G & 3
    q = pad.reshape(N, pc, 4)
    return (q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)).astype(np.uint8)


def unpack2bit(packed: np.ndarray, M: int) -> np.ndarray:
    """Inverse of :func:`pack2bit` (pack2bit.cu:53-60): byte c -> codes of SNPs 4c..4c+3, truncated to M."""
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    N, pc = packed.shape
    assert pc == packed_cols(M)
    out = np.empty((N, pc, 4), dtype=np.uint8)
    for i in range(4):
        out[:, :, i] = (packed >> (2 * i)) & 3
    return np.ascontiguousarray(out.reshape(N, pc * 4)[:, :M])


# --------------------------------------------------------------------------------------------------------------
# model state
# --------------------------------------------------------------------------------------------------------------


# --------------------------------------------------------------------------------------------------------------
# PLINK .bed decoding  (src/utils_c/utils.pyx:43-68 read_bed; src/snp_reader.py:16-45 _read_bed, :109-110 read_data)
# --------------------------------------------------------------------------------------------------------------


def read_bed(bed_bytes: np.ndarray, N: int) -> np.ndarray:
    """bed_bytes: M x ceil(N/4) uint8, the .bed payload after its 3 magic bytes, one row per SNP (snp_reader.py:33-38).
    Sample 4b+i sits in bits 2i..2i+1 of byte b; field -> genotype code through the table [2, 3, 1, 0]
    (utils.pyx:51,58-67).  Returns the N x M uint8 matrix ``read_bed`` fills (codes 0, 1, 2, 3 = missing)."""
    lut = np.array([2, 3, 1, 0], dtype=np.uint8)
    M = bed_bytes.shape[0]
    fields = np.stack([(bed_bytes >> (2 * i)) & 3 for i in range(4)], axis=2).reshape(M, -1)[:, :N]
    return np.ascontiguousarray(lut[fields].T)


def orient_alleles(G: np.ndarray) -> np.ndarray:
    """``G if G.mean() < 1 else 2 - G`` in uint8 arithmetic (snp_reader.py:110): missing 3 becomes 255, whose two low
    bits — all that pack2bit keeps (pack2bit.cu:29) — are 3 again."""
    assert int(G.min()) == 0 and int(G.max()) in (2, 3)          # snp_reader.py:109
    return G if G.mean() < 1 else (2 - G).astype(np.uint8)


# --------------------------------------------------------------------------------------------------------------
# randomized SVD  (src/svd.py:39-83; products src/utils_c/rsvd.pyx:16-50)
# --------------------------------------------------------------------------------------------------------------


def multiply_A_omega(A: np.ndarray, Omega: np.ndarray) -> np.ndarray:
    """Y[i, j] = sum_l float(A[i, l]) * Omega[l, j] over the raw uint8 values (rsvd.pyx:16-32), in float64."""
    return A.astype(np.float64) @ Omega.astype(np.float64)


def multiply_QT_A(QT: np.ndarray, A: np.ndarray) -> np.ndarray:
    """B[i, j] = sum_l QT[i, l] * float(A[l, j]) (rsvd.pyx:34-50), in float64."""
    return QT.astype(np.float64) @ A.astype(np.float64)


def svd_flip(V: np.ndarray, U: np.ndarray) -> np.ndarray:
    """src/svd.py:16-37."""
    idx = np.argmax(np.abs(U), axis=0)
    return V * np.sign(U[idx, np.arange(U.shape[1])])[:, None]


def rsvd(A: np.ndarray, k: int = 8, seed: int = 42, oversampling: int = 10, power_iterations: int = 2) -> np.ndarray:
    """src/svd.py:39-83 with float64 products (the reference's are float32 running sums)."""
    rng = np.random.default_rng(seed)
    k_prime = max(k + oversampling, 20)
    Omega = rng.standard_normal(size=(A.shape[1], k_prime), dtype=np.float32)
    Y = multiply_A_omega(A, Omega)
    for _ in range(power_iterations):
        Qy, _ = np.linalg.qr(Y, mode="reduced")
        Y = multiply_A_omega(A, multiply_QT_A(Qy.T, A).T)
    Q, _ = np.linalg.qr(Y, mode="reduced")
    B = multiply_QT_A(Q.T, A)
    Ut, St, Vt = np.linalg.svd(B, full_matrices=False)
    return svd_flip(Vt, Ut)[:k]


@dataclass
class OracleState:
    """Parameters of Q_P in the reference's own layouts (model/neural_admixture.py:126-144).

    V      M x C   (``Q_P.V``, :129-130)
    w_rms  C       (``batch_norm.weight``, :135)
    W1,b1  H x C,H (``common_encoder.0``, :138-140)
    W2,b2  per head k x H, k (``multihead_encoder.heads.i``, :29)
    P      per head M x k  (``decoders.decoders.i.weight``, :73-74)
    Adam moments mirror every parameter; ``t`` is the shared step count (:197-204).
    """

    V: np.ndarray
    w_rms: np.ndarray
    W1: np.ndarray
    b1: np.ndarray
    W2: List[np.ndarray]
    b2: List[np.ndarray]
    P: List[np.ndarray]
    ks: List[int]
    t: int = 0
    m: dict = field(default_factory=dict)
    v: dict = field(default_factory=dict)

    def params(self):
        out = {"V": self.V, "w_rms": self.w_rms, "W1": self.W1, "b1": self.b1}
        for i in range(len(self.ks)):
            out[f"W2.{i}"] = self.W2[i]
            out[f"b2.{i}"] = self.b2[i]
            out[f"P.{i}"] = self.P[i]
        return out

    def copy(self) -> "OracleState":
        return OracleState(
            self.V.copy(), self.w_rms.copy(), self.W1.copy(), self.b1.copy(),
            [a.copy() for a in self.W2], [a.copy() for a in self.b2], [a.copy() for a in self.P],
            list(self.ks), self.t, {k: a.copy() for k, a in self.m.items()}, {k: a.copy() for k, a in self.v.items()})


# --------------------------------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------------------------------


def genotype_to_x(g: np.ndarray, dtype=np.float64) -> np.ndarray:
    """``X = g.float()/2 ; X = where(X == 1.5, 0, X)`` (model/neural_admixture.py:169-170): codes 0,1,2 -> 0,.5,1;
    the missing code 3 is trained as 0."""
    x = g.astype(dtype) / 2
    x[g == 3] = 0
    return x


def encoder_fwd(x: np.ndarray, V: np.ndarray) -> np.ndarray:
    """``X_pca = X @ self.V`` (:172)."""
    return x @ V


def mlp_fwd(Z, w_rms, W1, b1, W2, b2, rms_eps=1e-8):
    """RMSNorm(C, eps=1e-8) -> Linear(C,H)+ReLU -> per-head Linear(H,k) -> softmax(dim=1)
    (:135-144 construction, :173-176 forward).  Returns (Zn, rinv, Hh, [Q_k])."""
    rinv = 1.0 / np.sqrt((Z * Z).mean(axis=1, keepdims=True) + rms_eps)  # torch.nn.RMSNorm semantics
    Zn = Z * rinv * w_rms
    Hh = np.maximum(Zn @ W1.T + b1, 0)
    Qs = []
    for W2k, b2k in zip(W2, b2):
        L = Hh @ W2k.T + b2k
        L = L - L.max(axis=1, keepdims=True)
        e = np.exp(L)
        Qs.append(e / e.sum(axis=1, keepdims=True))
    return Zn, rinv, Hh, Qs


def decoder_loss_grads(x: np.ndarray, Q: np.ndarray, P: np.ndarray):
    """One head of the decoder + loss + their backward.

    forward  : ``R = clamp_(Q @ P.T, 0, 1)`` (NeuralDecoder.forward :94-97; P is the Linear weight, M x k, :73-74)
    loss     : ``BCELoss(reduction='sum')(R, X)`` (:288, :431) = sum -[X*max(log R,-100) + (1-X)*max(log1p(-R),-100)]
    backward : BCELoss: (R - X) / max((1-R)*R, 1e-12); clamp_: gradient passes where 0 <= raw <= 1 (inclusive);
               Linear: dP = G.T @ Q, dQ = G @ P  (autograd through :96, invoked by loss.backward() :410).
    Returns (loss, dQ, dP)."""
    raw = Q @ P.T
    R = np.clip(raw, 0.0, 1.0)
    with np.errstate(divide="ignore"):
        logR = np.maximum(np.log(R), -100.0)
        log1mR = np.maximum(np.log1p(-R), -100.0)
    loss = -(x * logR + (1.0 - x) * log1mR).sum()
    G = (R - x) / np.maximum((1.0 - R) * R, 1e-12)
    G = np.where((raw >= 0.0) & (raw <= 1.0), G, 0.0)
    return float(loss), G @ P, G.T @ Q


def supervised_loss_grads(Q: np.ndarray, y: np.ndarray, weight: float):
    """``supervised_loss_weight * CrossEntropyLoss(reduction='sum')(Q, y)`` (:293, :473): the reference feeds the
    already-softmaxed Q as logits (double softmax).  Returns (loss, dQ)."""
    Ls = Q - Q.max(axis=1, keepdims=True)
    lse = np.log(np.exp(Ls).sum(axis=1, keepdims=True))
    logp = Ls - lse
    n = Q.shape[0]
    loss = -logp[np.arange(n), y].sum() * weight
    d = np.exp(logp)
    d[np.arange(n), y] -= 1.0
    return float(loss), d * weight


def mlp_bwd(dQs, Qs, Hh, Zn, rinv, Z, w_rms, W1, W2):
    """Backward of :func:`mlp_fwd` (what autograd derives for :173-176).  Returns
    (dZ, dw_rms, dW1, db1, [dW2_k], [db2_k])."""
    dH = np.zeros_like(Hh)
    dW2, db2 = [], []
    for dQ, Q, W2k in zip(dQs, Qs, W2):
        dL = Q * (dQ - (dQ * Q).sum(axis=1, keepdims=True))
        dW2.append(dL.T @ Hh)
        db2.append(dL.sum(axis=0))
        dH += dL @ W2k
    dHpre = dH * (Hh > 0)
    dW1 = dHpre.T @ Zn
    db1 = dHpre.sum(axis=0)
    dZn = dHpre @ W1
    # RMSNorm backward: y = z * rinv * w
    C = Z.shape[1]
    dw_rms = (dZn * Z * rinv).sum(axis=0)
    gw = dZn * w_rms
    dZ = rinv * gw - Z * (rinv ** 3) * (gw * Z).sum(axis=1, keepdims=True) / C
    return dZ, dw_rms, dW1, db1, dW2, db2


def encoder_bwd(x: np.ndarray, dZ: np.ndarray) -> np.ndarray:
    """dV = X.T @ dZ (autograd through ``X @ self.V`` :172)."""
    return x.T @ dZ


def adam_update(p, g, m, v, t, lr, beta1=0.9, beta2=0.95, eps=1e-8):
    """torch.optim.Adam, no weight decay, no amsgrad, betas (0.9, 0.95) (:197-204).  ``t`` is the 1-based step.
    In-place on p, m, v."""
    m *= beta1
    m += (1 - beta1) * g
    v *= beta2
    v += (1 - beta2) * g * g
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p -= (lr / bc1) * m / denom


# --------------------------------------------------------------------------------------------------------------
# one full training step and the loops around it
# --------------------------------------------------------------------------------------------------------------


def step_grads(st: OracleState, g_batch: np.ndarray, y: Optional[np.ndarray] = None, sup_weight: float = 100.0):
    """Forward + backward of one minibatch (``_run_step`` :419-432 then ``loss.backward()`` :410; supervised
    variant :460-474).  Returns (loss, grads dict keyed like ``OracleState.params()``, aux dict)."""
    x = genotype_to_x(g_batch, dtype=st.V.dtype)
    Z = encoder_fwd(x, st.V)
    Zn, rinv, Hh, Qs = mlp_fwd(Z, st.w_rms, st.W1, st.b1, st.W2, st.b2)
    loss = 0.0
    dQs, grads = [], {}
    for i, (Q, P) in enumerate(zip(Qs, st.P)):
        l, dQ, dP = decoder_loss_grads(x, Q, P)
        loss += l
        if y is not None and i == 0:
            ls, dqs = supervised_loss_grads(Q, y, sup_weight)
            loss += ls
            dQ = dQ + dqs
        dQs.append(dQ)
        grads[f"P.{i}"] = dP
    dZ, dw, dW1, db1, dW2, db2 = mlp_bwd(dQs, Qs, Hh, Zn, rinv, Z, st.w_rms, st.W1, st.W2)
    grads.update({"V": encoder_bwd(x, dZ), "w_rms": dw, "W1": dW1, "b1": db1})
    for i in range(len(st.ks)):
        grads[f"W2.{i}"] = dW2[i]
        grads[f"b2.{i}"] = db2[i]
    aux = {"Z": Z, "Qs": Qs, "dQs": dQs, "dZ": dZ, "Hh": Hh}
    return loss, grads, aux


def train_step(st: OracleState, g_batch: np.ndarray, lr: float, y=None, sup_weight: float = 100.0):
    """zero_grad -> forward -> backward -> Adam step -> restrict_P (``_run_epoch`` body :403-414).
    Mutates ``st``; returns (loss, aux)."""
    loss, grads, aux = step_grads(st, g_batch, y, sup_weight)
    st.t += 1
    for name, p in st.params().items():
        if name not in st.m:
            st.m[name] = np.zeros_like(p)
            st.v[name] = np.zeros_like(p)
        adam_update(p, grads[name], st.m[name], st.v[name], st.t, lr)
    for P in st.P:  # restrict_P (:179-185)
        np.clip(P, 0.0, 1.0, out=P)
    aux["grads"] = grads
    return loss, aux


def infer_Q(st: OracleState, G: np.ndarray, batch: int = 1024) -> List[np.ndarray]:
    """Forward-only Q for every row, sequential batches (post-train pass :369-383; inference.py:71-77)."""
    outs = [[] for _ in st.ks]
    for s in range(0, G.shape[0], batch):
        x = genotype_to_x(G[s:s + batch], dtype=st.V.dtype)
        _, _, _, Qs = mlp_fwd(encoder_fwd(x, st.V), st.w_rms, st.W1, st.b1, st.W2, st.b2)
        for o, q in zip(outs, Qs):
            o.append(q)
    return [np.concatenate(o, axis=0) for o in outs]


def train(st: OracleState, G: np.ndarray, epoch_orders: Sequence[np.ndarray], batch_size: int, lr: float,
          y: Optional[np.ndarray] = None, sup_weight: float = 100.0) -> Tuple[List[float], List[np.ndarray]]:
    """The epoch loop (``launch_training`` :365-366, ``_run_epoch`` :394-417).  ``epoch_orders[e]`` is the row order
    the reference's ``RandomSampler`` yields in epoch e (one ``torch.randperm(N, generator)`` per epoch,
    loaders.py:30 + :283); batches are consecutive slices of it, last one ragged (DataLoader default
    ``drop_last=False``).  Returns (per-epoch loss sums, final Qs)."""
    losses = []
    for order in epoch_orders:
        acc = 0.0
        for s in range(0, len(order), batch_size):
            idx = np.asarray(order[s:s + batch_size])
            l, _ = train_step(st, G[idx], lr, None if y is None else y[idx], sup_weight)
            acc += l
        losses.append(acc)
    return losses, infer_Q(st, G, min(G.shape[0], 1024))


def loglikelihood(G: np.ndarray, P: np.ndarray, Q: np.ndarray, eps: float = 1e-6) -> float:
    """fp64 binomial log-likelihood, missing (3) skipped, rec and g clamped by eps (src/utils_c/utils.pyx:17-40)."""
    rec = np.clip(Q.astype(np.float64) @ P.astype(np.float64).T, eps, 1 - eps)
    g = np.clip(G.astype(np.float64), eps, 2 - eps)
    ll = g * np.log(rec) + (2 - g) * np.log1p(-rec)
    return float(ll[G != 3].sum())


def hudsons_fst(p1: np.ndarray, p2: np.ndarray) -> float:
    """mean((p1-p2)^2) / (mean(p1(1-p2)+p2(1-p1)) + 1e-7)  (model/neural_admixture.py:532-550)."""
    return float(((p1 - p2) ** 2).mean() / ((p1 * (1 - p2) + p2 * (1 - p1)).mean() + 1e-7))
