"""The PyTorch CPU port used as bench.py's CPU baseline (oracle/nadm_torch_port.py) reproduces the reference-generated
golden fixtures and agrees with the numpy oracle.  CPU only."""
import numpy as np
import pytest
import torch

from helpers import load_golden, sub, relF
from nadm_torch_port import TorchPort


def _port_from(g, ks):
    init = {k: torch.as_tensor(v) for k, v in sub(g, "init/").items()}
    P = [init[f"decoders.decoders.{i}.weight"] for i in range(len(ks))]
    return TorchPort(init["V"], P, init["common_encoder.0.weight"].shape[0], lr=float(g["lr"]), state=init)


def test_port_two_steps_match_reference():
    g = load_golden("step_k5.npz")
    port = _port_from(g, [5])
    x = torch.as_tensor(g["G"])
    for step, key in enumerate(["after1/", "after2/"]):
        loss = port.step(x)
        assert abs(loss - g["losses"][step]) < 1e-6 * abs(g["losses"][step])
        sd = port.state_dict()
        for name, ref in sub(g, key).items():
            assert relF(sd[name].numpy(), ref) < 1e-6, (key, name)


@pytest.mark.parametrize("fixture", ["train_k3.npz", "train_k3to5.npz", "train_sup_k3.npz"])
def test_port_training_matches_reference(fixture):
    g = load_golden(fixture)
    ks = [int(k) for k in g["ks"]]
    port = _port_from(g, ks)
    data = torch.as_tensor(g["G"])
    y = torch.as_tensor(g["pops"], dtype=torch.int64) if "pops" in g else None
    B = int(g["batch"])
    for e, order in enumerate(g["orders"]):
        acc = 0.0
        for s in range(0, len(order), B):
            idx = torch.as_tensor(order[s:s + B])
            acc += port.step(TorchPort.gather(data, idx), None if y is None else y[idx])
        assert abs(acc - g["epoch_losses"][e]) < 1e-5 * abs(g["epoch_losses"][e])
    Qs = port.infer(data, min(data.shape[0], 1024))
    for i in range(len(ks)):
        assert relF(Qs[i].numpy(), g[f"Q/{i}"]) < 1e-5
        assert relF(port.P[i].detach().numpy(), g[f"P/{i}"]) < 1e-5
