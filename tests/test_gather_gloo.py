"""`gather_rows` (what `NeuralAdmixture.gather_P / gather_V` use to put the SNP shards back together) on two gloo
ranks with unequal slice sizes.  CPU only."""
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    from neural_admixture_b200.model.neural_admixture import gather_rows
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = torch.arange(7 * 3, dtype=torch.float32).view(7, 3)
    cuts = [0, 4, 7]                                   # rank 0 holds 4 rows, rank 1 holds 3
    got = gather_rows(full[cuts[rank]:cuts[rank + 1]].clone())
    torch.save(got, f"{out_dir}/r{rank}.pt")
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_rows_two_ranks(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    full = torch.arange(7 * 3, dtype=torch.float32).view(7, 3)
    for r in range(2):
        assert torch.equal(torch.load(tmp_path / f"r{r}.pt"), full)
