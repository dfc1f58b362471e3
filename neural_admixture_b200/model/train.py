"""Mirror of the reference's ``neural_admixture/model/train.py``: initialise P (GMM in the PCA subspace, or per-label
means in supervised mode), put the 2-bit packed genotypes on the device, and run ``NeuralAdmixture``.

Only the call site (reference :115-132) and everything after it is the B200 engine; the one-off initialisation
(reference :47-83: a float PCA projection in 1024-row chunks and scikit-learn's GaussianMixture on N x 8 points) is
host-side plumbing kept as the reference has it.  In sharded mode (``num_gpus > 1`` with torch.distributed
initialised) every rank keeps only its contiguous SNP slice of the genotypes, of P and of V."""
from __future__ import annotations

import logging
import sys
from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .. import ops
from .neural_admixture import NeuralAdmixture

logging.basicConfig(stream=sys.stdout, level=logging.INFO, format="%(message)s")
log = logging.getLogger(__name__)


def snp_slice(M: int, rank: int, world: int, align: int = 64) -> Tuple[int, int]:
    """Contiguous SNP range of ``rank``: boundaries are multiples of ``align`` SNPs (16 packed bytes)."""
    blocks = (M + align - 1) // align
    b0 = (blocks * rank) // world
    b1 = (blocks * (rank + 1)) // world
    return min(M, b0 * align), min(M, b1 * align)


def gmm_initial_P(data: np.ndarray, V: np.ndarray, ks, n_components: int, seed: int) -> np.ndarray:
    """(sum K) x M initial allele frequencies from a Gaussian mixture fitted in the PCA subspace (reference :47-69).
    Note the projection uses genotype/2 WITHOUT the missing->0 mapping, exactly as the reference's init does."""
    from sklearn.mixture import GaussianMixture
    N, M = data.shape
    X_pca = np.zeros((N, n_components), dtype=np.float32)
    for i in range(0, N, 1024):
        j = min(i + 1024, N)
        X_pca[i:j] = (data[i:j].astype(np.float32) / 2) @ V.T
    X_pca = X_pca.astype("float64")
    log.info("")
    log.info("    Running Gaussian Mixture in PCA subspace...")
    log.info("")
    Ps = []
    for k in ks:
        gmm = GaussianMixture(n_components=k, n_init=5, init_params="k-means++", tol=1e-4, covariance_type="full",
                              max_iter=100, random_state=seed).fit(X_pca)
        Ps.append(np.clip(gmm.means_ @ V, 5e-6, 1 - 5e-6))
    return np.concatenate(Ps, axis=0)


def train(epochs: int, batch_size: int, learning_rate: float, K: int, seed: int, data: torch.Tensor,
          device: torch.device, num_gpus: int, hidden_size: int, master: bool, V: np.ndarray, pops,
          min_k: int = None, max_k: int = None, n_components: int = None):
    """Same signature and return value ``(Ps, Qs, model)`` as the reference's ``train`` (:19-149).
    ``data``: host uint8 N x M genotype codes; ``V``: C x M right singular vectors (as RSVD returns them)."""
    N, M = data.shape
    data_np = data.numpy()
    sharded = num_gpus > 1 and dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if sharded else 0
    world = dist.get_world_size() if sharded else 1
    sumK = K if K is not None else sum(range(min_k, max_k + 1))
    y_num = None

    if master:
        if pops is None:
            P = gmm_initial_P(data_np, V, [K] if K is not None else list(range(min_k, max_k + 1)), n_components, seed)
        else:
            log.info("")
            log.info("    Running Supervised Mode...")
            log.info("")
            anc = {a: i for i, a in enumerate(sorted(np.unique([a for a in pops])))}
            assert len(anc) == K, (f"Number of ancestries in training ground truth ({len(anc)}) is not equal to the "
                                   f"value of K ({K})")
            y_num = np.array([anc[a] for a in pops], dtype=np.int64)
            # per-label mean genotype, not halved (reference :82)
            P = np.vstack([data_np[y_num == i].astype(np.float32).mean(axis=0) for i in range(K)])
    if sharded:
        dist.barrier()
        P_full = torch.as_tensor(P, dtype=torch.float32, device=device).contiguous() if master else \
            torch.empty((sumK, M), dtype=torch.float32, device=device)
        V_full = torch.as_tensor(np.ascontiguousarray(V.T), dtype=torch.float32, device=device) if master else \
            torch.empty((M, n_components), dtype=torch.float32, device=device)
        if master:
            log.info("    Broadcasting to all GPUs...")
        dist.broadcast(P_full, src=0)
        dist.broadcast(V_full, src=0)
        if pops is not None:
            y = torch.as_tensor(y_num, dtype=torch.int64, device=device) if master else \
                torch.empty(len(pops), dtype=torch.int64, device=device)
            dist.broadcast(y, src=0)
        else:
            y = None
        dist.barrier()
        c0, c1 = snp_slice(M, rank, world)
        P_init = P_full[:, c0:c1].contiguous()
        V_dev = V_full[c0:c1].contiguous()
        del P_full, V_full
    else:
        c0, c1 = 0, M
        P_init = torch.as_tensor(P, dtype=torch.float32, device=device).contiguous()
        V_dev = torch.as_tensor(np.ascontiguousarray(V.T), dtype=torch.float32, device=device)
        y = torch.as_tensor(y_num, dtype=torch.int64, device=device) if pops is not None else None

    packed = ops.PackedGenotypes.from_unpacked_host(data, device, c0, c1)
    model = NeuralAdmixture(K, epochs, batch_size, learning_rate, device, seed, num_gpus, master, "nadm_b200", min_k,
                            max_k)
    Qs, Ps, raw = model.launch_training(P_init, packed, hidden_size, V_dev.shape[1], V_dev, c1 - c0, N, y)

    # log-likelihood (reference :134-146, utils.pyx:17-40) on the resident packed data: every rank evaluates its own SNP
    # slice against the full Q and its slice of P, the fp64 partial sums are all-reduced (the reference evaluates it on
    # the host matrix; nothing is re-uploaded here)
    ws = torch.empty(ops.workspace_bytes(min(N, 1024), c1 - c0, V_dev.shape[1], hidden_size, sumK), dtype=torch.uint8,
                     device=device)
    ks = [K] if K is not None else list(range(min_k, max_k + 1))
    for i, k in enumerate(ks):
        ll = torch.tensor([ops.loglikelihood(packed, model.last_Q[i].contiguous(),
                                             raw.decoders.decoders[i].weight.data.contiguous(), ws)],
                          dtype=torch.float64, device=device)
        if sharded:
            dist.all_reduce(ll, op=dist.ReduceOp.SUM)
        if master:
            logl = float(ll.item())
            log.info(f"    Log-likelihood: {logl:2f}." if K is not None else f"    Log-likelihood for K={k}: {logl:2f}.")
    return Ps, Qs, raw


# ------------------------------------------------------------------------------------------------------------------
# The same pipeline with the genotypes ONLY in the device-resident packed layout (single GPU): nothing below touches an
# N x M one-byte-per-genotype array (50 GB at 100k x 500k), so real data of the benchmark's size can be fitted.
# ------------------------------------------------------------------------------------------------------------------
def gmm_initial_P_packed(pg: ops.PackedGenotypes, V: np.ndarray, ks, seed: int, missing_value: int = 3) -> np.ndarray:
    """``gmm_initial_P`` with the PCA projection on the device: ``(data / 2) @ V.T`` over the raw uint8 values
    (reference :49-53: no missing -> 0 there, a missing entry counts 3 / 2, or 255 / 2 after the reader's flip) is
    ``0.5 * nadm_geno_matmul``; the Gaussian mixture on the N x C projection stays scikit-learn's (:61-67)."""
    from sklearn.mixture import GaussianMixture
    dev = pg.storage.device
    ws = torch.empty(ops.workspace_bytes(1024, pg.M, 8, 8, 8), dtype=torch.uint8, device=dev)
    Vt = torch.as_tensor(np.ascontiguousarray(V.T), dtype=torch.float32, device=dev)          # M x C
    X_pca = (0.5 * ops.geno_matmul(pg, Vt, ws, missing_value)).cpu().numpy().astype("float64")
    log.info("")
    log.info("    Running Gaussian Mixture in PCA subspace...")
    log.info("")
    Ps = []
    for k in ks:
        gmm = GaussianMixture(n_components=k, n_init=5, init_params="k-means++", tol=1e-4, covariance_type="full",
                              max_iter=100, random_state=seed).fit(X_pca)
        Ps.append(np.clip(gmm.means_ @ V, 5e-6, 1 - 5e-6))
    return np.concatenate(Ps, axis=0)


def train_packed(epochs: int, batch_size: int, learning_rate: float, K: int, seed: int, pg: ops.PackedGenotypes,
                 hidden_size: int, V: np.ndarray, pops=None, min_k: int = None, max_k: int = None,
                 missing_value: int = 3):
    """``train`` (reference model/train.py:19-149) for genotypes that exist only as a ``PackedGenotypes`` on the device
    (e.g. from ``src.snp_reader.read_bed_packed``; V from ``src.svd.RSVD`` on the same object).  Returns
    ``(Ps, Qs, model)``.  ``missing_value``: see ``gmm_initial_P_packed``."""
    device, N, M = pg.storage.device, pg.N, pg.M
    ks = [K] if K is not None else list(range(min_k, max_k + 1))
    y = None
    if pops is None:
        P = gmm_initial_P_packed(pg, V, ks, seed, missing_value)
    else:
        log.info("")
        log.info("    Running Supervised Mode...")
        log.info("")
        anc = {a: i for i, a in enumerate(sorted(np.unique([a for a in pops])))}
        assert len(anc) == K, (f"Number of ancestries in training ground truth ({len(anc)}) is not equal to the "
                               f"value of K ({K})")
        y_num = np.array([anc[a] for a in pops], dtype=np.int64)
        # per-label mean of the raw uint8 values, not halved (reference :78-82): one-hot labels^T @ A / counts
        onehot = torch.zeros((K, N), dtype=torch.float32, device=device)
        onehot[torch.as_tensor(y_num, device=device), torch.arange(N, device=device)] = 1.0
        ws = torch.empty(ops.workspace_bytes(1024, M, 8, 8, 8), dtype=torch.uint8, device=device)
        sums = ops.geno_matmul_t(pg, onehot, ws, missing_value)
        P = (sums / onehot.sum(dim=1, keepdim=True)).cpu().numpy()
        y = torch.as_tensor(y_num, dtype=torch.int64, device=device)
    P_init = torch.as_tensor(P, dtype=torch.float32, device=device).contiguous()
    V_dev = torch.as_tensor(np.ascontiguousarray(V.T), dtype=torch.float32, device=device)
    model = NeuralAdmixture(K, epochs, batch_size, learning_rate, device, seed, 0, True, "nadm_b200", min_k, max_k)
    Qs, Ps, raw = model.launch_training(P_init, pg, hidden_size, V_dev.shape[1], V_dev, M, N, y)
    ws = torch.empty(ops.workspace_bytes(min(N, 1024), M, V_dev.shape[1], hidden_size, sum(ks)), dtype=torch.uint8,
                     device=device)
    for i, k in enumerate(ks):
        logl = ops.loglikelihood(pg, torch.as_tensor(Qs[i], device=device).contiguous(),
                                 torch.as_tensor(Ps[i], device=device).contiguous(), ws)
        log.info(f"    Log-likelihood: {logl:2f}." if K is not None else f"    Log-likelihood for K={k}: {logl:2f}.")
    return Ps, Qs, raw
