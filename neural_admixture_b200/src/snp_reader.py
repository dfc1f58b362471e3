"""PLINK .bed straight into the device-resident 2-bit packed layout ("next" row f4).

Mirror of the reference's ``SNPReader._read_bed`` + ``read_data`` (src/snp_reader.py:16-45,89-110) for the BED
format, with the same results — LUT [2,3,1,0] (utils_c/utils.pyx:43-68), biallelic check, and the global allele flip
``G if G.mean() < 1 else 2 - G`` — but the N x M one-byte-per-genotype host array (50 GB at 100k x 500k) and the
re-packing pass (pack2bit.cu:65-117) never exist: chunks of SNP rows go host -> device as they are in the file and a
2-bit transpose kernel (``nadm_bed_to_packed``) writes the sample-major packed matrix the training kernels read."""
from __future__ import annotations

import logging
import sys
from math import ceil
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import torch

from .. import ops
from .._lib import NadmError

logging.basicConfig(stream=sys.stdout, level=logging.INFO, format="%(message)s")
log = logging.getLogger(__name__)

CHUNK_SNPS = 16384  # SNP rows per host->device copy (multiple of 128)


def bed_shape(file: str) -> Tuple[int, int, Path]:
    """(N samples, M SNPs, path of the .bed) from the .fam line count and the .bed size (snp_reader.py:27-38)."""
    file_path = Path(file)
    fam_file, bed_file = file_path.with_suffix(".fam"), file_path.with_suffix(".bed")
    with open(fam_file, "r") as fam:
        N = sum(1 for _ in fam)
    n_bytes = ceil(N / 4)
    payload = bed_file.stat().st_size - 3
    assert payload % n_bytes == 0, "bim file doesn't match!"
    return N, payload // n_bytes, bed_file


def read_bed_packed(file: str, device, col0: int = 0, col1: Optional[int] = None,
                    chunk_snps: int = CHUNK_SNPS, allreduce=None) -> ops.PackedGenotypes:
    """Read SNP columns [col0, col1) of a PLINK .bed into a ``PackedGenotypes`` on ``device`` (a rank's SNP slice in
    sharded runs).  The flip test uses the mean over the columns read; pass ``allreduce`` (callable summing an int64
    tensor over the ranks, whose slices together cover the file) to decide it on the whole matrix as the reference
    does."""
    device = torch.device(device)
    if device.type != "cuda":
        raise NadmError("read_bed_packed writes the device-resident packed layout: a CUDA device is required")
    N, M, bed_file = bed_shape(file)
    col1 = M if col1 is None else col1
    assert 0 <= col0 < col1 <= M and chunk_snps % 128 == 0
    n_bytes = ceil(N / 4)
    mm = np.memmap(bed_file, dtype=np.uint8, mode="r", offset=3, shape=(M, n_bytes))
    Mloc = col1 - col0
    out = ops.PackedGenotypes.empty(N, Mloc, device)
    counts = torch.zeros(4, dtype=torch.int64, device=device)
    stage_h = torch.empty((min(chunk_snps, Mloc), n_bytes), dtype=torch.uint8).pin_memory()
    stage_d = torch.empty_like(stage_h, device=device)
    for s0 in range(0, Mloc, chunk_snps):
        s1 = min(Mloc, s0 + chunk_snps)
        stage_h[: s1 - s0].copy_(torch.from_numpy(np.ascontiguousarray(mm[col0 + s0:col0 + s1])))
        stage_d[: s1 - s0].copy_(stage_h[: s1 - s0], non_blocking=True)
        ops.bed_to_packed(stage_d[: s1 - s0], N, out, snp0=s0, counts=counts)
        torch.cuda.current_stream(device).synchronize()         # the pinned staging buffer is reused
    cells = float(N) * Mloc
    if allreduce is not None:
        allreduce(counts)
        cells = float(N) * M
    c = counts.cpu().numpy().astype(np.float64)
    n1, n2, n3 = c[1], c[2], c[3]
    # reference: assert G.min() == 0 and G.max() in (2, 3)  (snp_reader.py:109)
    assert n1 + n2 + n3 < cells and (n2 > 0 or n3 > 0), \
        "Only biallelic SNPs are supported. Please make sure multiallelic sites have been removed."
    mean = (n1 + 2 * n2 + 3 * n3) / cells
    if not mean < 1:
        ops.flip_packed(out)
    log.info(f"    Data contains {N} samples and {Mloc} SNPs.")
    return out
