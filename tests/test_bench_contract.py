"""CPU checks of bench.py's own contract: the synthetic workload is the same at every world size (so the driver's
per-N lines describe ONE problem and `loss.first_timed_step` can be compared across N), the ranks' SNP slices are unions
of the generator's column blocks, and the reference arm prints the line the driver parses."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


@pytest.mark.parametrize("M", [64, 1000, 4099, 500_000])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_rank_slices_are_unions_of_generation_blocks(M, world):
    gb = bench.gen_bounds(M)
    assert gb[0] == 0 and gb[-1] == M and all(a <= b for a, b in zip(gb, gb[1:]))
    assert all(b % 64 == 0 for b in gb[:-1])
    covered = 0
    for r in range(world):
        c0, c1 = bench.snp_slice(M, r, world)
        assert c0 == covered and c0 % 64 == 0
        mine = [j for j in range(bench.GEN_BLOCKS) if gb[j] < gb[j + 1] and gb[j] >= c0 and gb[j + 1] <= c1]
        assert sum(gb[j + 1] - gb[j] for j in mine) == c1 - c0        # what synth_packed asserts on the device
        covered = c1
    assert covered == M


def test_initial_parameters_do_not_depend_on_the_sharding():
    M, ks, dev = 4099, [3, 5], torch.device("cpu")
    V, P = bench.synth_init(M, 0, M, ks, bench.SEED, dev)
    for world in (2, 4, 8):
        parts = [bench.synth_init(M, *bench.snp_slice(M, r, world), ks, bench.SEED, dev) for r in range(world)]
        assert torch.equal(torch.cat([p[0] for p in parts]), V)
        assert torch.equal(torch.cat([p[1] for p in parts], dim=1), P)


def test_genotype_columns_do_not_depend_on_the_sharding():
    """synth_packed's per-block generators, replayed on the CPU: the codes of a column block are a function of the block
    only, whichever rank draws them."""
    N, M, dev = 40, 4099, torch.device("cpu")
    Qt, Pt = bench.synth_params(N, M, 8, bench.SEED, dev)
    gb = bench.gen_bounds(M)

    def block(j):
        gen = torch.Generator(device=dev).manual_seed(bench.SEED * 100003 + j)
        return bench.synth_rows(Qt, Pt[:, gb[j]:gb[j + 1]], gen)

    live = [j for j in range(bench.GEN_BLOCKS) if gb[j] < gb[j + 1]]
    full = torch.cat([block(j) for j in live], dim=1)
    assert full.shape == (N, M) and set(full.unique().tolist()) <= {0, 1, 2, 3}
    for world in (2, 8):
        cols = []
        for r in range(world):
            c0, c1 = bench.snp_slice(M, r, world)
            cols += [block(j) for j in live if gb[j] >= c0 and gb[j + 1] <= c1]
        assert torch.equal(torch.cat(cols, dim=1), full)


@pytest.mark.timeout(300)
def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's op sequence on the host cores): one JSON line with the base
    contract's keys, the steps / warm-ups it was asked for, and a bounded sample."""
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--workload", "cfg2", "--cpu-budget", "4"], capture_output=True, text=True, timeout=280)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRICS["cfg2"] and line["unit"] == "samples/s"
    assert line["steps"] == 1 and line["warmup"] == 1 and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "1 timed steps (+1 warm-up)" in cb["sample"]
