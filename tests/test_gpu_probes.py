"""Hardware probes (tests/cuda/*.cu) as GPU tests: the tcgen05 operand layouts csrc/nadm_tc.cuh assumes are checked on
the device against a host reference.  The binaries are built in-tree by __graft_entry__.build() (nvcc cross-compiles)."""
import subprocess
from pathlib import Path

import pytest

CUDA_DIR = Path(__file__).resolve().parent / "cuda"


@pytest.mark.gpu
def test_umma_operand_layout_probe():
    exe = CUDA_DIR / "umma_probe.bin"
    assert exe.exists(), "tests/cuda/umma_probe.bin missing: run __graft_entry__.build()"
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    print(res.stdout)
    assert res.returncode == 0 and "PROBE OK" in res.stdout, res.stdout + res.stderr
    assert "FAIL" not in res.stdout
