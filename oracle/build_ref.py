"""Recipe for oracle/_ref/: compile the reference's ONE native source on this path — the pybind11 pack / unpack module
neural_admixture/src/utils_c/pack2bit.cu — from where it lies under /root/reference, unmodified, for sm_100a, into
oracle/_ref/pack2bit_ref.so.  TEST INFRASTRUCTURE ONLY: tests/test_gpu_parity.py loads it on the GPU box to pin this
repo's 2-bit layout (nadm_pack2bit / nadm_unpack2bit and everything that reads the packed matrix) to the reference's own
kernels; nothing under neural_admixture_b200/ touches it.  No reference source is copied into the repository; the
output directory is git-ignored (it travels to the GPU box with the snapshot, like the library itself).
The reference JIT-builds this file with torch.utils.cpp_extension.load (model/train.py:122-125); this is the same
compilation as one explicit nvcc command (no ninja, no cache directory outside the repository).

usage: python oracle/build_ref.py        (needs /root/reference; a no-op with a message when it is absent)"""
from __future__ import annotations

import subprocess
import sys
import sysconfig
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/neural_admixture/src/utils_c/pack2bit.cu")
OUT = ROOT / "oracle" / "_ref" / "pack2bit_ref.so"


def build(verbose: bool = True) -> Path | None:
    if not SRC.exists():
        if verbose:
            print(f"oracle/_ref: {SRC} not present (GPU box): using the prebuilt {OUT.name} if it travelled")
        return OUT if OUT.exists() else None
    if OUT.exists() and OUT.stat().st_mtime >= SRC.stat().st_mtime and OUT.stat().st_mtime >= Path(__file__).stat().st_mtime:
        if verbose:
            print(f"oracle/_ref: {OUT.name} is up to date")
        return OUT
    import torch
    from torch.utils import cpp_extension as ce
    OUT.parent.mkdir(parents=True, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = str(Path(torch.__file__).resolve().parent / "lib")
    cmd = ["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC",
           "-DTORCH_EXTENSION_NAME=pack2bit_ref", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "--expt-relaxed-constexpr",
           *inc, str(SRC), "-o", str(OUT), f"-L{libdir}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
           "-ltorch_python", "-Xlinker", f"-rpath={libdir}"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    out = build()
    print(f"built {out}" if out else "nothing built")
    sys.exit(0)
