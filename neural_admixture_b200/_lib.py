"""ctypes binding of libnadm_b200.so (include/nadm_b200.h).  There is no fallback: if the library is missing or a
call fails, this module raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# NADM_LIB=<file name under csrc/> selects an A/B build variant of the SAME library (tools/gpu_round.sh); never a fallback
LIB_PATH = Path(__file__).resolve().parent / "csrc" / os.environ.get("NADM_LIB", "libnadm_b200.so")

c_u8p = C.c_void_p
c_f32p = C.c_void_p
c_i64p = C.c_void_p


class NadmError(RuntimeError):
    pass


class AdamHyper(C.Structure):
    """nadm_adam_t"""
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("step", C.c_int32), ("reserved", C.c_int32), ("device_coef", C.c_void_p)]


class Xchg(C.Structure):
    """nadm_xchg_t"""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("area", C.c_void_p * 8), ("seq", C.c_void_p),
                ("slot_floats", C.c_int64)]


class MlpParams(C.Structure):
    """nadm_mlp_params_t"""
    _names = ["w_rms", "W1", "b1", "W2", "b2"]
    _fields_ = [(f"{pre}{n}", C.c_void_p) for pre in ("", "m_", "v_", "g_") for n in ["w_rms", "W1", "b1", "W2", "b2"]]


# name -> (restype, argtypes): exactly the declarations of include/nadm_b200.h
SIGNATURES = {
    "nadm_version": (C.c_int, []),
    "nadm_last_error": (C.c_char_p, []),
    "nadm_launch_count": (C.c_int64, []),
    "nadm_generic_launch_count": (C.c_int64, []),
    "nadm_pack2bit": (C.c_int, [c_u8p, C.c_int64, C.c_int64, C.c_int64, c_u8p, C.c_int64, C.c_void_p]),
    "nadm_unpack2bit": (C.c_int, [c_u8p, C.c_int64, C.c_int64, C.c_int64, c_u8p, C.c_int64, C.c_void_p]),
    "nadm_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "nadm_encoder_fwd": (C.c_int, [c_u8p, C.c_int64, c_i64p, C.c_int64, C.c_int32, C.c_int64, c_f32p, C.c_int32,
                                   c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nadm_encoder_fwd_deferred": (C.c_int, [c_u8p, C.c_int64, c_i64p, C.c_int64, C.c_int32, C.c_int64, c_f32p, C.c_int32,
                                   c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nadm_mlp_fwd": (C.c_int, [c_f32p, C.c_int32, C.c_int32, C.c_int32, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p,
                               C.POINTER(C.c_int32), C.c_int32, c_f32p, c_f32p, c_f32p, C.POINTER(Xchg), C.c_void_p]),
    "nadm_decoder_step": (C.c_int, [c_u8p, C.c_int64, c_i64p, C.c_int64, C.c_int32, C.c_int64, c_f32p, c_f32p,
                                    C.c_int32, C.c_int32, C.c_int32, c_f32p, c_f32p, c_f32p, C.POINTER(AdamHyper),
                                    c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nadm_decoder_step_deferred": (C.c_int, [c_u8p, C.c_int64, c_i64p, C.c_int64, C.c_int32, C.c_int64, c_f32p, c_f32p,
                                    C.c_int32, C.c_int32, C.c_int32, c_f32p, c_f32p, c_f32p, C.POINTER(AdamHyper),
                                    c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nadm_mlp_bwd": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32,
                               C.POINTER(C.c_int32), C.c_int32, c_i64p, C.c_float, C.POINTER(MlpParams),
                               C.POINTER(AdamHyper), c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.POINTER(Xchg),
                               C.c_void_p]),
    "nadm_mlp_bwd_deferred": (C.c_int, [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int32, C.c_int32, C.c_int32,
                               C.POINTER(C.c_int32), C.c_int32, c_i64p, C.c_float, C.POINTER(MlpParams),
                               C.POINTER(AdamHyper), c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.POINTER(Xchg),
                               C.c_void_p]),
    "nadm_xchg_area_bytes": (C.c_size_t, [C.c_int64]),
    "nadm_ipc_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "nadm_ipc_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "nadm_ipc_close": (C.c_int, [C.c_void_p]),
    "nadm_ipc_free": (C.c_int, [C.c_void_p]),
    "nadm_encoder_bwd": (C.c_int, [c_u8p, C.c_int64, c_i64p, C.c_int64, C.c_int32, C.c_int64, c_f32p, C.c_int32,
                                   c_f32p, c_f32p, c_f32p, C.POINTER(AdamHyper), c_f32p, C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    "nadm_loglikelihood": (C.c_int, [c_u8p, C.c_int64, C.c_int64, C.c_int64, c_f32p, c_f32p, C.c_int32, C.c_double,
                                     C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nadm_bed_to_packed": (C.c_int, [c_u8p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, c_u8p, C.c_int64,
                                     C.c_void_p, C.c_void_p]),
    "nadm_flip_packed": (C.c_int, [c_u8p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "nadm_step_begin": (C.c_int, [c_i64p, C.c_int64, c_i64p, C.c_int64, C.c_int32, c_i64p, C.POINTER(AdamHyper),
                                  C.c_void_p, c_f32p, C.c_void_p]),
    "nadm_step_end": (C.c_int, [c_i64p, c_f32p, c_f32p, C.c_void_p]),
    "nadm_step_next": (C.c_int, [c_i64p, C.c_int64, c_i64p, C.c_int64, C.c_int32, c_i64p, C.POINTER(AdamHyper),
                                 C.c_void_p, c_f32p, C.c_int32, c_f32p, C.c_void_p]),
    "nadm_step_flush": (C.c_int, [c_i64p, c_f32p, C.c_void_p]),
    "nadm_geno_matmul": (C.c_int, [c_u8p, C.c_int64, C.c_int64, C.c_int64, c_f32p, C.c_int32, C.c_int32, c_f32p,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "nadm_geno_matmul_t": (C.c_int, [c_u8p, C.c_int64, C.c_int64, C.c_int64, c_f32p, C.c_int32, C.c_int32, c_f32p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the library (once).  Raises NadmError when it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NadmError(f"{LIB_PATH} is missing: build it with `python -m neural_admixture_b200.build` "
                            "(or __graft_entry__.build()); this package has no CPU or PyTorch fallback")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().nadm_last_error().decode(errors="replace")
        raise NadmError(f"libnadm_b200 error {rc}: {msg}")
