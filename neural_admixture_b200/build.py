"""Compile libnadm_b200.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.  No torch involvement."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "libnadm_b200.so"
SOURCES = ["nadm_stream.cu", "nadm_mlp.cu", "nadm_tc_enc.cu", "nadm_tc_dec.cu", "nadm_bed.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libnadm_b200.so")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [CSRC.parent.parent / "include" / "nadm_b200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB), *[str(CSRC / s) for s in SOURCES]]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    return LIB


def build_variant(name: str, extra_flags: list[str]) -> Path:
    """A/B variant of the library under csrc/<name> (selected at run time with NADM_LIB=<name>)."""
    out = CSRC / name
    res = subprocess.run([_nvcc(), *NVCC_FLAGS, *extra_flags, "-o", str(out), *[str(CSRC / s) for s in SOURCES]],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
