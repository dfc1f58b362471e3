"""Output writers with the reference's file names and formats (src/utils.py:36-67)."""
from __future__ import annotations

import logging
import sys
from pathlib import Path

import numpy as np

logging.basicConfig(stream=sys.stdout, level=logging.INFO, format="%(message)s")
log = logging.getLogger(__name__)


def write_outputs(Qs, run_name: str, K, min_k, max_k, out_path, Ps=None) -> None:
    """``{run_name}.{K}.Q`` (N x K) and, when given, ``{run_name}.{K}.P`` (M x K), space-delimited ``np.savetxt``
    (reference src/utils.py:36-67)."""
    out_path = Path(out_path)
    out_path.mkdir(parents=True, exist_ok=True)
    if K is not None:
        np.savetxt(out_path / f"{run_name}.{K}.Q", Qs[0], delimiter=" ")
        if Ps is not None:
            np.savetxt(out_path / f"{run_name}.{K}.P", Ps[0], delimiter=" ")
            log.info("    Q and P matrices saved.")
        else:
            log.info("    Q matrix saved.")
    else:
        for i, k in enumerate(range(min_k, max_k + 1)):
            np.savetxt(out_path / f"{run_name}.{k}.Q", Qs[i], delimiter=" ")
            if Ps is not None:
                np.savetxt(out_path / f"{run_name}.{k}.P", Ps[i], delimiter=" ")
        log.info("    Q and P matrices saved for all K." if Ps is not None else "    Q matrices saved for all K.")
