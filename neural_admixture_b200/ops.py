"""Tensor-level wrappers over the C ABI (include/nadm_b200.h).  torch is used for device memory and streams only;
every computation below is a call into libnadm_b200.so on the current CUDA stream."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import AdamHyper, MlpParams, NadmError, Xchg, check

PITCH_ALIGN = 128  # bytes; rows of the packed matrix start on 128-byte lines


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t: Optional[torch.Tensor] = None) -> C.c_void_p:
    """The current stream of the device that holds ``t`` (not of whatever device happens to be current)."""
    return C.c_void_p(torch.cuda.current_stream(None if t is None else t.device).cuda_stream)


class _on:
    """Make the tensor's device current for the duration of a library call: the library launches on the current device
    (cudaFuncSetAttribute, SM count and the launch itself), the pointers live on the tensor's device."""

    def __init__(self, t: torch.Tensor):
        self.idx = t.device.index if t.is_cuda else None
        self.prev = None

    def __enter__(self):
        if self.idx is not None and self.idx != torch.cuda.current_device():
            self.prev = torch.cuda.current_device()
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise NadmError("libnadm_b200 operates on CUDA tensors only (there is no CPU path)")


def padded_pitch(M: int) -> int:
    pc = (M + 3) // 4
    return ((pc + PITCH_ALIGN - 1) // PITCH_ALIGN) * PITCH_ALIGN


def launch_count() -> int:
    return int(_lib.load().nadm_launch_count())


def generic_launch_count() -> int:
    """Launches of the slow first-generation kernels so far (shapes outside the tensor-core path)."""
    return int(_lib.load().nadm_generic_launch_count())


def pack2bit(src: torch.Tensor, dst: torch.Tensor, M: Optional[int] = None) -> None:
    """src: rows x M uint8 codes (cuda), dst: rows x >=ceil(M/4) uint8 (cuda); row strides are honoured."""
    _need_cuda(src, dst)
    assert src.dtype == torch.uint8 and dst.dtype == torch.uint8 and src.dim() == 2 and dst.dim() == 2
    assert src.stride(1) == 1 and dst.stride(1) == 1
    rows = src.shape[0]
    M = src.shape[1] if M is None else M
    assert dst.shape[0] == rows
    src_pitch = src.stride(0) if rows > 1 else max(src.shape[1], 1)
    if rows > 1 and dst.stride(0) != dst.shape[1]:
        raise NadmError("pack2bit: destination rows must be contiguous (the kernel zero-fills each row's tail)")
    with _on(src): check(_lib.load().nadm_pack2bit(_ptr(src), rows, M, src_pitch, _ptr(dst), dst.shape[1], _stream(src)))


def unpack2bit(src: torch.Tensor, dst: torch.Tensor) -> None:
    """src: rows x >=ceil(M/4) packed (cuda), dst: rows x M uint8 (cuda)."""
    _need_cuda(src, dst)
    assert src.dtype == torch.uint8 and dst.dtype == torch.uint8 and src.dim() == 2 and dst.dim() == 2
    assert src.stride(1) == 1 and dst.stride(1) == 1
    rows, M = dst.shape
    assert src.shape[0] == rows
    src_pitch = src.stride(0) if rows > 1 else src.shape[1]
    dst_pitch = dst.stride(0) if rows > 1 else max(M, 1)
    with _on(src): check(_lib.load().nadm_unpack2bit(_ptr(src), rows, M, src_pitch, _ptr(dst), dst_pitch, _stream(src)))


def workspace_bytes(B: int, M: int, C_: int, H: int, sumK: int) -> int:
    return int(_lib.load().nadm_workspace_bytes(B, M, C_, H, sumK))


def _ks_array(ks: Sequence[int]):
    return (C.c_int32 * len(ks))(*[int(k) for k in ks])


def adam_hyper(lr: float, step: int, beta1: float = 0.9, beta2: float = 0.95, eps: float = 1e-8,
               device_coef: Optional[torch.Tensor] = None) -> AdamHyper:
    """nadm_adam_t.  ``device_coef`` (8 x float32 CUDA tensor written by ``step_begin``): the kernels read the step's
    coefficients from the device, so the step's launch arguments are the same every step (CUDA-graph replay)."""
    return AdamHyper(lr, beta1, beta2, eps, step, 0, None if device_coef is None else device_coef.data_ptr())


class PackedGenotypes:
    """Sample-major 2-bit genotype matrix on the device, rows padded to PITCH_ALIGN bytes with zero tails."""

    def __init__(self, storage: torch.Tensor, N: int, M: int):
        assert storage.dtype == torch.uint8 and storage.dim() == 2 and storage.is_cuda and storage.is_contiguous()
        assert storage.shape[0] == N and storage.shape[1] % 16 == 0 and storage.shape[1] >= (M + 3) // 4
        if N >= 2 ** 32:
            raise NadmError("at most 2^32 - 1 sample rows per device (the kernels keep row numbers in 32 bits)")
        self.storage, self.N, self.M = storage, N, M

    @property
    def pitch(self) -> int:
        return self.storage.shape[1]

    @classmethod
    def empty(cls, N: int, M: int, device) -> "PackedGenotypes":
        return cls(torch.zeros((N, padded_pitch(M)), dtype=torch.uint8, device=device), N, M)

    @classmethod
    def from_reference_layout(cls, packed: torch.Tensor, M: int) -> "PackedGenotypes":
        """Adopt an N x ceil(M/4) tensor in the reference's layout (model/train.py:121).  Zero-copy when the row
        pitch already satisfies the kernels' 16-byte rule, otherwise one device-side re-pitch copy."""
        N, pc = packed.shape
        assert pc >= (M + 3) // 4
        if packed.is_contiguous() and pc % 16 == 0 and packed.data_ptr() % 16 == 0:
            return cls(packed, N, M)
        out = cls.empty(N, M, packed.device)
        out.storage[:, :pc].copy_(packed)
        return out

    @classmethod
    def from_unpacked_host(cls, G: torch.Tensor, device, col0: int = 0, col1: Optional[int] = None,
                           chunk_rows: int = 1024) -> "PackedGenotypes":
        """Host uint8 N x M codes -> device packed, columns [col0, col1) only (a rank's SNP slice).  Staged through
        pinned memory in chunks of rows, packed on the device (role of pack2bit_cpu_to_gpu, pack2bit.cu:65-117)."""
        assert G.dtype == torch.uint8 and G.dim() == 2 and not G.is_cuda
        N = G.shape[0]
        col1 = G.shape[1] if col1 is None else col1
        Mloc = col1 - col0
        out = cls.empty(N, Mloc, device)
        stage = torch.empty((min(chunk_rows, max(N, 1)), Mloc), dtype=torch.uint8, device=device)
        for r0 in range(0, N, chunk_rows):
            r1 = min(N, r0 + chunk_rows)
            stage[: r1 - r0].copy_(G[r0:r1, col0:col1], non_blocking=False)
            pack2bit(stage[: r1 - r0], out.storage[r0:r1], Mloc)
        return out


_DEFER = os.environ.get("NADM_NO_DEFER", "0") != "1"   # A/B switch: keep the separate reduction kernels


def encoder_fwd(pg: PackedGenotypes, V: torch.Tensor, Z: torch.Tensor, ws: torch.Tensor, *,
                row_idx: Optional[torch.Tensor] = None, row0: int = 0, B: Optional[int] = None,
                deferred: bool = False) -> None:
    """``deferred``: the sum over the kernel's CTAs may be left to the ``mlp_fwd`` call that follows ON THE SAME Z (and the
    same ``ws``, untouched in between): nothing may read Z before that call."""
    _need_cuda(V, Z, ws, row_idx)
    B = (row_idx.numel() if row_idx is not None else B)
    assert V.dtype == torch.float32 and V.is_contiguous() and V.shape[0] == pg.M
    assert Z.dtype == torch.float32 and Z.is_contiguous() and Z.shape == (B, V.shape[1])
    assert row_idx is None or (row_idx.dtype == torch.int64 and row_idx.is_contiguous())
    fn = _lib.load().nadm_encoder_fwd_deferred if (deferred and _DEFER) else _lib.load().nadm_encoder_fwd
    with _on(pg.storage): check(fn(_ptr(pg.storage), pg.pitch, _ptr(row_idx), row0, B, pg.M, _ptr(V), V.shape[1],
                                   _ptr(Z), _ptr(ws), ws.numel() * ws.element_size(), _stream(pg.storage)))


def mlp_fwd(Z, w_rms, W1, b1, W2, b2, ks, rinv, Hh, Q, xchg: Optional[Xchg] = None) -> None:
    _need_cuda(Z, w_rms, W1, b1, W2, b2, rinv, Hh, Q)
    B, C_ = Z.shape
    H = W1.shape[0]
    for t in (Z, w_rms, W1, b1, W2, b2, rinv, Hh, Q):
        assert t.dtype == torch.float32 and t.is_contiguous()
    assert W2.shape == (sum(ks), H) and Q.shape == (B, sum(ks)) and Hh.shape == (B, H)
    with _on(Z): check(_lib.load().nadm_mlp_fwd(_ptr(Z), B, C_, H, _ptr(w_rms), _ptr(W1), _ptr(b1), _ptr(W2), _ptr(b2),
                                   _ks_array(ks), len(ks), _ptr(rinv), _ptr(Hh), _ptr(Q),
                                   None if xchg is None else C.byref(xchg), _stream(Z)))


def decoder_step(pg: PackedGenotypes, Q, dQ, q_off: int, k: int, P, Pm, Pv, adam: Optional[AdamHyper], loss, ws, *,
                 row_idx=None, row0: int = 0, dP_out=None, deferred: bool = False) -> None:
    """``deferred``: the sum over the kernel's CTAs (this head's columns of dQ, the head's loss) may be left to the
    ``mlp_bwd`` call that follows ON THE SAME dQ and ``ws``; a later ``decoder_step`` on the same dQ completes it too."""
    _need_cuda(Q, dQ, P, Pm, Pv, loss, ws, row_idx, dP_out)   # loss may be None: gradients only
    B, q_ld = Q.shape
    assert P.shape == (pg.M, k) and P.is_contiguous() and P.dtype == torch.float32
    assert Q.is_contiguous() and dQ.is_contiguous() and dQ.shape == Q.shape
    assert row_idx is None or (row_idx.dtype == torch.int64 and row_idx.is_contiguous() and row_idx.numel() == B)
    fn = _lib.load().nadm_decoder_step_deferred if (deferred and _DEFER) else _lib.load().nadm_decoder_step
    with _on(pg.storage): check(fn(_ptr(pg.storage), pg.pitch, _ptr(row_idx), row0, B, pg.M, _ptr(Q), _ptr(dQ),
                                   q_ld, q_off, k, _ptr(P), _ptr(Pm), _ptr(Pv),
                                   None if adam is None else C.byref(adam), _ptr(dP_out), _ptr(loss), _ptr(ws),
                                   ws.numel() * ws.element_size(), _stream(pg.storage)))


def mlp_bwd(dQ, Q, Hh, Z, rinv, ks, params: MlpParams, adam: Optional[AdamHyper], dZ, loss, ws, *, labels=None,
            sup_weight: float = 0.0, xchg: Optional[Xchg] = None, deferred_apply: bool = False) -> None:
    """``deferred_apply``: the parameter update is left to the ``encoder_bwd`` call that follows ON THE SAME dZ (it runs on
    that kernel's epilogue warps); any other library call that needs the parameters runs it first."""
    _need_cuda(dQ, Q, Hh, Z, rinv, dZ, loss, ws, labels)
    B, C_ = Z.shape
    H = Hh.shape[1]
    assert labels is None or (labels.dtype == torch.int64 and labels.is_contiguous() and labels.numel() == B)
    fn = _lib.load().nadm_mlp_bwd_deferred if (deferred_apply and _DEFER) else _lib.load().nadm_mlp_bwd
    with _on(dQ): check(fn(_ptr(dQ), _ptr(Q), _ptr(Hh), _ptr(Z), _ptr(rinv), B, C_, H, _ks_array(ks), len(ks),
                                   _ptr(labels), float(sup_weight), C.byref(params),
                                   None if adam is None else C.byref(adam), _ptr(dZ), _ptr(loss), _ptr(ws),
                                   ws.numel() * ws.element_size(), None if xchg is None else C.byref(xchg), _stream(dQ)))


def encoder_bwd(pg: PackedGenotypes, dZ, V, Vm, Vv, adam: Optional[AdamHyper], ws, *, row_idx=None, row0: int = 0,
                dV_out=None) -> None:
    _need_cuda(dZ, V, Vm, Vv, ws, row_idx, dV_out)
    B, C_ = dZ.shape
    assert V.shape == (pg.M, C_) and V.is_contiguous() and dZ.is_contiguous()
    with _on(pg.storage): check(_lib.load().nadm_encoder_bwd(_ptr(pg.storage), pg.pitch, _ptr(row_idx), row0, B, pg.M, _ptr(dZ), C_, _ptr(V),
                                       _ptr(Vm), _ptr(Vv), None if adam is None else C.byref(adam), _ptr(dV_out),
                                       _ptr(ws), ws.numel() * ws.element_size(), _stream(pg.storage)))


def loglikelihood(pg: PackedGenotypes, Q: torch.Tensor, P: torch.Tensor, ws: torch.Tensor, eps: float = 1e-6) -> float:
    _need_cuda(Q, P, ws)
    k = P.shape[1]
    assert Q.shape == (pg.N, k) and P.shape == (pg.M, k) and Q.is_contiguous() and P.is_contiguous()
    out = torch.zeros(1, dtype=torch.float64, device=Q.device)
    with _on(pg.storage): check(_lib.load().nadm_loglikelihood(_ptr(pg.storage), pg.pitch, pg.N, pg.M, _ptr(Q), _ptr(P), k, eps, _ptr(out),
                                         _ptr(ws), ws.numel() * ws.element_size(), _stream(pg.storage)))
    return float(out.item())


def bed_to_packed(bed: torch.Tensor, N: int, dst: PackedGenotypes, snp0: int = 0, flip: bool = False,
                  counts: Optional[torch.Tensor] = None) -> None:
    """bed: (SNPs of this chunk) x ceil(N/4) uint8 on the device — rows of the .bed payload; writes columns
    [snp0, snp0 + bed.shape[0]) of ``dst``.  counts: optional 4 x int64 device tensor (codes 1, 2, 3 accumulate)."""
    _need_cuda(bed, counts)
    assert bed.dtype == torch.uint8 and bed.dim() == 2 and bed.stride(1) == 1
    assert counts is None or (counts.dtype == torch.int64 and counts.numel() == 4 and counts.is_contiguous())
    Mc = bed.shape[0]
    with _on(bed): check(_lib.load().nadm_bed_to_packed(_ptr(bed), bed.stride(0) if Mc > 1 else bed.shape[1], N, Mc, snp0, int(flip),
                                         _ptr(dst.storage), dst.pitch, _ptr(counts), _stream(bed)))


def flip_packed(pg: PackedGenotypes) -> None:
    """In place g -> 2 - g (missing unchanged): the reference's minor-allele orientation (snp_reader.py:110)."""
    with _on(pg.storage): check(_lib.load().nadm_flip_packed(_ptr(pg.storage), pg.pitch, pg.N, pg.M, _stream(pg.storage)))


def step_begin(order: torch.Tensor, counters: torch.Tensor, stride: int, B: int, row_idx_out: torch.Tensor,
               hyper: AdamHyper, coef_out: torch.Tensor, loss_accum: Optional[torch.Tensor] = None) -> None:
    """Device-side start of a step: the minibatch's rows out of the device-resident permutation and the Adam
    coefficients of the step, both indexed by ``counters`` (2 x int64 on the device)."""
    _need_cuda(order, counters, row_idx_out, coef_out)
    assert order.dtype == torch.int64 and counters.dtype == torch.int64 and counters.numel() >= 2
    assert row_idx_out.dtype == torch.int64 and row_idx_out.numel() >= B and coef_out.numel() * coef_out.element_size() >= 32
    with _on(order): check(_lib.load().nadm_step_begin(_ptr(order), order.numel(), _ptr(counters), stride, B, _ptr(row_idx_out),
                                      C.byref(hyper), _ptr(coef_out), _ptr(loss_accum), _stream(order)))


def step_end(counters: torch.Tensor, loss: Optional[torch.Tensor], losses_out: Optional[torch.Tensor]) -> None:
    with _on(counters): check(_lib.load().nadm_step_end(_ptr(counters), _ptr(loss), _ptr(losses_out), _stream(counters)))


def step_next(order: torch.Tensor, counters: torch.Tensor, stride: int, B: int, row_idx_out: torch.Tensor,
              hyper: AdamHyper, coef_out: torch.Tensor, loss_accum: Optional[torch.Tensor], record_loss: bool,
              losses_out: Optional[torch.Tensor]) -> None:
    """``step_end`` of the pending step (if any) + ``step_begin`` of the next in ONE kernel; ``counters``: 4 x int64
    (see nadm_step_next).  The last step of a run is finished by ``step_flush``."""
    _need_cuda(order, counters, row_idx_out, coef_out)
    assert order.dtype == torch.int64 and counters.dtype == torch.int64 and counters.numel() == 4
    assert row_idx_out.dtype == torch.int64 and row_idx_out.numel() >= B and coef_out.numel() * coef_out.element_size() >= 32
    with _on(order): check(_lib.load().nadm_step_next(_ptr(order), order.numel(), _ptr(counters), stride, B, _ptr(row_idx_out),
                                     C.byref(hyper), _ptr(coef_out), _ptr(loss_accum), 1 if record_loss else 0,
                                     _ptr(losses_out), _stream(order)))


def step_flush(counters: torch.Tensor, losses_out: Optional[torch.Tensor]) -> None:
    assert counters.dtype == torch.int64 and counters.numel() == 4
    with _on(counters): check(_lib.load().nadm_step_flush(_ptr(counters), _ptr(losses_out), _stream(counters)))


def geno_matmul(pg: PackedGenotypes, Omega: torch.Tensor, ws: torch.Tensor, missing_value: int = 3) -> torch.Tensor:
    """Y = A @ Omega (N x K) with A the uint8 genotype VALUES (code 3 -> ``missing_value``); Omega: M x K float32 on the
    device, any K (processed in column chunks of 8).  Role of rsvd.multiply_A_omega (rsvd.pyx:56-71)."""
    _need_cuda(Omega, ws)
    assert Omega.dtype == torch.float32 and Omega.dim() == 2 and Omega.shape[0] == pg.M
    K = Omega.shape[1]
    Y = torch.empty((pg.N, K), dtype=torch.float32, device=Omega.device)
    for c0 in range(0, K, 8):
        c1 = min(K, c0 + 8)
        om = Omega[:, c0:c1].contiguous()
        y = torch.empty((pg.N, c1 - c0), dtype=torch.float32, device=Omega.device)
        with _on(pg.storage): check(_lib.load().nadm_geno_matmul(_ptr(pg.storage), pg.pitch, pg.N, pg.M, _ptr(om), c1 - c0, missing_value,
                                           _ptr(y), _ptr(ws), ws.numel() * ws.element_size(), _stream(pg.storage)))
        Y[:, c0:c1] = y
    return Y


def geno_matmul_t(pg: PackedGenotypes, QT: torch.Tensor, ws: torch.Tensor, missing_value: int = 3) -> torch.Tensor:
    """B = QT @ A (K x M) with A the uint8 genotype VALUES; QT: K x N float32 on the device, any K.  Role of
    rsvd.multiply_QT_A (rsvd.pyx:74-93)."""
    _need_cuda(QT, ws)
    assert QT.dtype == torch.float32 and QT.dim() == 2 and QT.shape[1] == pg.N
    K = QT.shape[0]
    Bm = torch.empty((K, pg.M), dtype=torch.float32, device=QT.device)
    for c0 in range(0, K, 8):
        c1 = min(K, c0 + 8)
        q = QT[c0:c1].T.contiguous()                                  # N x k
        bt = torch.empty((pg.M, c1 - c0), dtype=torch.float32, device=QT.device)
        with _on(pg.storage): check(_lib.load().nadm_geno_matmul_t(_ptr(pg.storage), pg.pitch, pg.N, pg.M, _ptr(q), c1 - c0, missing_value,
                                             _ptr(bt), _ptr(ws), ws.numel() * ws.element_size(), _stream(pg.storage)))
        Bm[c0:c1] = bt.T
    return Bm


class PeerExchange:
    """The peer-mapped exchange areas of an SNP-sharded run (nadm_xchg_t): one area per rank (CUDA IPC, one process per
    GPU on one node), opened by every other rank.  ``xchg`` is what ``mlp_fwd`` / ``mlp_bwd`` take.  Collective: every
    rank of ``group`` must construct it at the same point."""

    def __init__(self, slot_floats: int, device: torch.device, group=None):
        import torch.distributed as dist
        lib = _lib.load()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise NadmError("the fused peer exchange supports at most 8 ranks")
        self.device, self.lib, self.peers = device, lib, []
        nbytes = int(lib.nadm_xchg_area_bytes(slot_floats))
        own, handle = C.c_void_p(), (C.c_ubyte * 64)()
        with torch.cuda.device(device):
            check(lib.nadm_ipc_alloc(nbytes, C.byref(own), handle))
        self.own = own
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.xchg = Xchg()
        self.xchg.world, self.xchg.rank, self.xchg.slot_floats = self.world, self.rank, int(slot_floats)
        ok = True
        for r in range(self.world):
            if r == self.rank:
                self.xchg.area[r] = own.value
                continue
            p = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
            with torch.cuda.device(device):
                rc = lib.nadm_ipc_open(buf, C.byref(p))
            if rc != 0:
                ok = False
                self.err = lib.nadm_last_error().decode(errors="replace")
                break
            self.peers.append(p)
            self.xchg.area[r] = p.value
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.seq = torch.zeros(2 * 1024, dtype=torch.int32, device=device)
        self.xchg.seq = self.seq.data_ptr()
        self.ok = bool(flag.item())
        dist.barrier(group=group)

    def close(self) -> None:
        """Collective: unmap the peers' areas, then (after a barrier, when nobody can touch it any more) free one's own."""
        import torch.distributed as dist
        if self.own is None:
            return
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for p in self.peers:
                self.lib.nadm_ipc_close(p)
            self.peers = []
            if dist.is_initialized():
                dist.barrier()
            self.lib.nadm_ipc_free(self.own)
        self.own = None
