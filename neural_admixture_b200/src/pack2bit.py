"""Drop-in for the reference's JIT-built ``pack2bit`` pybind11 module (src/utils_c/pack2bit.cu:144-147): the same
two functions with the same argument meaning, on top of ``nadm_pack2bit`` / ``nadm_unpack2bit`` (C ABI).

Differences: work is issued on torch's current stream (the reference uses the legacy default stream and a device
synchronise after every launch, pack2bit.cu:115,141); errors surface as ``NadmError`` (a RuntimeError, like
TORCH_CHECK's)."""
from __future__ import annotations

import torch

from .. import ops
from .._lib import NadmError

MAX_ROWS_PER_BATCH = 1024  # rows staged per host->device copy (pack2bit.cu:8)


def pack2bit_cpu_to_gpu(input_cpu: torch.Tensor, output_gpu: torch.Tensor) -> None:
    """Host uint8 N x M codes -> device N x ceil(M/4) packed bytes (pack2bit.cu:65-117)."""
    if input_cpu.is_cuda:
        raise NadmError("Input tensor must be on CPU")
    if not output_gpu.is_cuda:
        raise NadmError("Output tensor must be on CUDA device")
    N, M = input_cpu.shape
    pc = (M + 3) // 4
    if output_gpu.shape[0] != N:
        raise NadmError("Output tensor row dimension mismatch")
    if output_gpu.shape[1] != pc:
        raise NadmError("Output tensor column dimension mismatch")
    rows = min(MAX_ROWS_PER_BATCH, max(N, 1))
    stage = torch.empty((rows, M), dtype=torch.uint8, device=output_gpu.device)
    src = input_cpu.contiguous()
    for r0 in range(0, N, rows):
        r1 = min(N, r0 + rows)
        stage[: r1 - r0].copy_(src[r0:r1])
        ops.pack2bit(stage[: r1 - r0], output_gpu[r0:r1], M)


def unpack2bit_gpu_to_gpu(input_gpu: torch.Tensor, output_gpu: torch.Tensor) -> None:
    """Device N x ceil(M/4) packed bytes -> device N x M uint8 codes (pack2bit.cu:120-142)."""
    if not input_gpu.is_cuda:
        raise NadmError("Input tensor must be on CUDA device")
    if not output_gpu.is_cuda:
        raise NadmError("Output tensor must be on CUDA device")
    if input_gpu.device != output_gpu.device:
        raise NadmError("Input and Output tensors must be on the same CUDA device")
    N, M = output_gpu.shape
    if input_gpu.shape[0] != N:
        raise NadmError("Input tensor row dimension mismatch")
    if input_gpu.shape[1] != (M + 3) // 4:
        raise NadmError("Input tensor column dimension mismatch based on output shape")
    ops.unpack2bit(input_gpu, output_gpu)
