"""Randomized SVD on the device-resident packed matrix ("next" row f2): mirror of the reference's ``src/svd.py``.

Same algorithm, same random stream and same return value as the reference's ``RSVD`` (:39-83) — Omega ~ N(0,1) from
``np.random.default_rng(seed)``, k' = max(k + oversampling, 20), ``power_iterations`` rounds of QR + two products, QR,
B = Q^T A, SVD of the k' x M matrix B, ``svd_flip``, first k rows of Vt — but the six passes over the N x M matrix
(``rsvd.multiply_A_omega`` / ``multiply_QT_A``: naive OpenMP triple loops over an N x M uint8 host array, rsvd.pyx:16-50)
run on the 2-bit packed matrix through ``nadm_geno_matmul`` / ``nadm_geno_matmul_t`` (exact-integer tensor-core
contraction).  The thin QR of the N x k' factor and the SVD of the k' x M factor stay on the host in numpy, as in the
reference."""
from __future__ import annotations

import logging
import sys
import time

import numpy as np
import torch

from .. import ops

logging.basicConfig(stream=sys.stdout, level=logging.INFO, format="%(message)s")
log = logging.getLogger(__name__)


def svd_flip(V: np.ndarray, U: np.ndarray) -> np.ndarray:
    """Sign convention of the reference (:16-37): row j of V is multiplied by the sign of the entry of column j of U
    that is largest in magnitude, which makes the decomposition deterministic."""
    pivot = np.abs(U).argmax(axis=0)
    signs = np.sign(U[pivot, np.arange(U.shape[1])])
    return V * signs[:, None]


def RSVD(pg: ops.PackedGenotypes, N: int, M: int, k: int = 8, seed: int = 42, oversampling: int = 10,
         power_iterations: int = 2, missing_value: int = 3) -> np.ndarray:
    """Vt[:k] (k x M, float32) of the genotype matrix held in ``pg``.  ``missing_value``: the uint8 value the
    reference's matrix holds for a missing genotype — 3, or 255 when its reader flipped the alleles
    (snp_reader.py:110)."""
    assert pg.N == N and pg.M == M
    dev = pg.storage.device
    rng = np.random.default_rng(seed)
    k_prime = max(k + oversampling, 20)
    ws = torch.empty(ops.workspace_bytes(1024, M, 8, 8, 8), dtype=torch.uint8, device=dev)
    t0 = time.time()

    def a_omega(om: np.ndarray) -> np.ndarray:          # (M, k') -> (N, k')
        return ops.geno_matmul(pg, torch.as_tensor(om, dtype=torch.float32, device=dev), ws, missing_value).cpu().numpy()

    def qt_a(qt: np.ndarray) -> np.ndarray:             # (k', N) -> (k', M)
        return ops.geno_matmul_t(pg, torch.as_tensor(qt, dtype=torch.float32, device=dev), ws, missing_value).cpu().numpy()

    log.info("    RSVD 1/4: random test matrix and Y = A @ Omega")
    Omega = rng.standard_normal(size=(M, k_prime), dtype=np.float32)
    Y = a_omega(Omega)
    for _ in range(power_iterations):
        basis, _ = np.linalg.qr(Y, mode="reduced")
        Y = a_omega(np.ascontiguousarray(qt_a(np.ascontiguousarray(basis.T)).T))
    log.info("    RSVD 2/4: thin QR of Y")
    Q, _ = np.linalg.qr(Y, mode="reduced")
    log.info("    RSVD 3/4: B = Q^T @ A")
    B = qt_a(np.ascontiguousarray(Q.T))
    log.info("    RSVD 4/4: SVD of the k' x M factor")
    Ut, St, Vt = np.linalg.svd(B, full_matrices=False)
    Vt = svd_flip(Vt, Ut)
    log.info(f"    RSVD done in {time.time() - t0:.2f} s")
    return Vt[:k, :]
