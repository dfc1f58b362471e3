"""``neural-admixture infer`` on the packed path ("next" row f1): mirror of the reference's ``src/inference.py``.

Same inputs (``{save_dir}/{name}.pt`` + ``{name}_config.json`` written by training, a genotype file) and the same
``{out_name}.{K}.Q`` outputs, but the genotypes never exist as an N x M uint8 device tensor (reference :65): a PLINK
.bed goes straight into the 2-bit packed layout (``snp_reader.read_bed_packed``), Q is preallocated instead of grown
with ``torch.cat`` (reference :73-77), and with ``torch.distributed`` initialised the SNP axis is sharded over the
ranks (the reference refuses ``num_gpus > 1``, :20-21): each rank reads its own columns of the file and of V, the
B x C partial projections are all-reduced, rank 0 writes the outputs."""
from __future__ import annotations

import json
import logging
import sys
import time
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from .. import ops
from .._lib import NadmError
from ..model.neural_admixture import Q_P
from ..model.train import snp_slice
from . import snp_reader, utils

logging.basicConfig(stream=sys.stdout, level=logging.INFO, format="%(message)s")
log = logging.getLogger(__name__)


def load_model(save_dir: str, name: str, device: torch.device, col0: int = 0, col1: Optional[int] = None) -> Q_P:
    """Reference :40-61: config JSON + state dict (saved without the decoders, src/main.py:41) -> ``Q_P`` in inference
    mode.  ``col0:col1`` keeps only a SNP slice of V (sharded inference)."""
    with open(f"{save_dir}/{name}_config.json", "r") as fb:
        config = json.load(fb)
    state_dict = torch.load(f"{save_dir}/{name}.pt", map_location="cpu", weights_only=True)
    V = state_dict.get("V")
    if col1 is not None:
        V = V[col0:col1].contiguous()
        state_dict = dict(state_dict, V=V)
    model = Q_P(int(config["hidden_size"]), int(config["num_features"]), ks_list=config["ks"], V=V, is_train=False)
    model.load_state_dict(state_dict)
    model.to(device)
    model.bind()
    return model


def infer_packed(model: Q_P, pg: ops.PackedGenotypes, batch_size: int = 2048) -> List[np.ndarray]:
    sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    allreduce = (lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)) if sharded else None
    Qs = model.infer_packed(pg, batch_size, allreduce=allreduce)
    return [Q.cpu().numpy() for Q in Qs]


def main(args, t0: float):
    """Inference entry point (reference :16-99).  ``args``: data_path, out_name, save_dir, name, seed, batch_size,
    num_gpus."""
    if not torch.cuda.is_available():
        raise NadmError("the B200 engine runs on CUDA devices only (no CPU path)")
    sharded = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    rank = dist.get_rank() if sharded else 0
    world = dist.get_world_size() if sharded else 1
    device = torch.device("cuda", torch.cuda.current_device())
    if not str(args.data_path).endswith(".bed"):
        raise NadmError("the packed inference path reads PLINK .bed files")
    N, M, _ = snp_reader.bed_shape(args.data_path)
    c0, c1 = snp_slice(M, rank, world) if sharded else (0, M)
    try:
        model = load_model(args.save_dir, args.name, device, c0, c1 if sharded else None)
    except FileNotFoundError:
        log.error(f"    Config file ({args.save_dir}/{args.name}_config.json) not found. Make sure it is in the correct "
                  "directory and with the correct name.")
        return 1
    log.info("    Model weights loaded.")
    pg = snp_reader.read_bed_packed(args.data_path, device, c0, c1,
                                    allreduce=(lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)) if sharded else None)
    log.info("    Running inference...")
    Qs = infer_packed(model, pg, max(int(args.batch_size), 1))
    if rank == 0:
        log.info("    Inference run successfully! Writing outputs...!")
        ks = model.ks_list
        if len(ks) == 1:
            utils.write_outputs(Qs, args.out_name, ks[0], None, None, args.save_dir)
        else:
            utils.write_outputs(Qs, args.out_name, None, ks[0], ks[-1], args.save_dir)
        log.info("")
        log.info(f"    Total elapsed time: {time.time() - t0:.2f} seconds.")
        log.info("")
    return 0
