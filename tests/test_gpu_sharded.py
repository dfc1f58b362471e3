"""SNP-sharded engine over NCCL on 2 / 4 / 8 B200s: the product `NeuralAdmixture` loop (graph-replayed steps with
their two all-reduces) against the single-GPU engine, against the fp64 oracle, and the sharded train -> save -> infer
round trip.  Each case is `tools/check_sharded.py` under torchrun; skipped when the box has fewer GPUs."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,exchange", [(2, "peer"), (2, "nccl"), (4, "peer"), (8, "peer")])
@pytest.mark.timeout(600)
def test_sharded_engine_matches_single_gpu_and_oracle(world, exchange):
    """exchange = peer: the two per-step messages are exchanged inside the MLP kernels over peer-mapped memory;
    nccl: two NCCL all-reduces per step (NADM_XCHG=nccl), the baseline path."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", NADM_XCHG=exchange)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                          str(ROOT / "tools" / "check_sharded.py")], capture_output=True, text=True, timeout=560, env=env)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-3000:]
    assert f"SHARDED PARITY ({world} GPUs): PASS" in out, out[-3000:]
    assert ("exchange = fused peer exchange" in out) == (exchange == "peer"), out[-3000:]
