"""Shared helpers for the parity tests: fixture loading and oracle-state construction."""
from pathlib import Path

import numpy as np

import nadm_oracle as orc

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_golden(name):
    z = np.load(GOLDEN / name, allow_pickle=False)
    return {k: z[k] for k in z.files}


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def state_from_sd(sd, ks, dtype=np.float64):
    """Reference state_dict keys -> OracleState."""
    n = len(ks)
    f = lambda a: np.array(a, dtype=dtype)
    return orc.OracleState(
        V=f(sd["V"]), w_rms=f(sd["batch_norm.weight"]), W1=f(sd["common_encoder.0.weight"]),
        b1=f(sd["common_encoder.0.bias"]),
        W2=[f(sd[f"multihead_encoder.heads.{i}.weight"]) for i in range(n)],
        b2=[f(sd[f"multihead_encoder.heads.{i}.bias"]) for i in range(n)],
        P=[f(sd[f"decoders.decoders.{i}.weight"]) for i in range(n)], ks=[int(k) for k in ks])


ORACLE_TO_SD = {"V": "V", "w_rms": "batch_norm.weight", "W1": "common_encoder.0.weight",
                "b1": "common_encoder.0.bias"}


def sd_name(oracle_name):
    if oracle_name in ORACLE_TO_SD:
        return ORACLE_TO_SD[oracle_name]
    kind, i = oracle_name.split(".")
    return {"W2": f"multihead_encoder.heads.{i}.weight", "b2": f"multihead_encoder.heads.{i}.bias",
            "P": f"decoders.decoders.{i}.weight"}[kind]


def relF(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
