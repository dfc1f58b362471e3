"""Writers of the ``.Q`` / ``.P`` text outputs (file names and number format of the reference's ``write_outputs``,
src/utils.py:36-67: ``{run_name}.{K}.Q`` is N x K, ``{run_name}.{K}.P`` is M x K, space-delimited ``%.18e``)."""
from __future__ import annotations

import logging
import sys
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

logging.basicConfig(stream=sys.stdout, level=logging.INFO, format="%(message)s")
log = logging.getLogger(__name__)


def _heads(K, min_k, max_k) -> Sequence[int]:
    return [K] if K is not None else list(range(min_k, max_k + 1))


def write_outputs(Qs, run_name: str, K, min_k, max_k, out_path, Ps=None) -> None:
    """One ``.Q`` (and, when ``Ps`` is given, one ``.P``) file per head; ``K`` selects single-head naming, otherwise the
    heads are ``min_k..max_k`` in order.  Same call signature as the reference's writer."""
    folder = Path(out_path)
    folder.mkdir(parents=True, exist_ok=True)
    ks = _heads(K, min_k, max_k)
    for i, k in enumerate(ks):
        for ext, mats in (("Q", Qs), ("P", Ps)):
            if mats is not None:
                np.savetxt(folder / f"{run_name}.{k}.{ext}", np.asarray(mats[i]), delimiter=" ")
    what = "Q and P matrices" if Ps is not None else ("Q matrix" if len(ks) == 1 else "Q matrices")
    log.info(f"    {what} saved{'' if len(ks) == 1 else ' for all K'}.")
