"""CPU port of the reference's per-minibatch training step in PyTorch eager ops.  TEST / BASELINE INFRASTRUCTURE ONLY.

Why this exists next to nadm_oracle.py: the reference (AI-sandbox/neural-admixture) is Python and executes this path
through multi-threaded PyTorch CPU kernels (``--num_gpus 0``).  The reference's sources cannot travel to the GPU box
and may not be copied into this repository, so ``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference``
time THIS port: the same sequence of torch ops the reference issues for one step, on all host threads
(``cpu_baseline.kind = "port"``).  Only ``tests/`` and ``bench.py`` import it; nothing under ``neural_admixture_b200/``
does.  It is checked against the fp64 numpy oracle and the reference-generated golden fixtures in
``tests/test_torch_port.py``.

Op sequence per step (citations: /root/reference/neural_admixture/):
  loaders.py:70-72 + default collate   rows of the host uint8 matrix gathered one by one and stacked
  model/neural_admixture.py:169-170    X = g.float()/2 ; X = where(X == 1.5, 0, X)
  :172                                 X @ V
  :135,173                             RMSNorm(C, eps=1e-8)
  :138-140,174 ; :29,175 ; :176        Linear+ReLU ; per-head Linear ; softmax(dim=1)
  :96-97                               clamp_(Q @ P_k^T, 0, 1)      (nn.Linear(k, M, bias=False), weight = P_k, M x k)
  :288,431                             BCELoss(reduction='sum') summed over heads
  :429,410,411,412,414                 zero_grad(set_to_none) ; backward ; Adam(betas (0.9,0.95), fused) step ;
                                       P.clamp_(0,1) ; loss.item()
``torch.set_float32_matmul_precision('medium')`` (:349) is applied when ``as_shipped=True`` (the timing legs) and
left at 'highest' for parity checks.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F


class TorchPort:
    def __init__(self, V: torch.Tensor, P_list: Sequence[torch.Tensor], hidden: int, lr: float = 2e-3,
                 seed: int = 42, state: Optional[dict] = None, as_shipped: bool = False, device=None):
        """V: M x C; P_list[i]: M x k_i.  MLP parameters are drawn from torch's default nn.Linear initialisation
        (on the CPU generator, whatever ``device`` is) unless ``state`` (reference state_dict key -> tensor) is
        given.  ``device`` (default: V's) is where the parameters live and the ops run: ``bench.py --impl
        reference-cuda`` times this same op sequence on cuda:0."""
        if as_shipped:
            torch.set_float32_matmul_precision("medium")
            torch.set_flush_denormal(True)          # :350
        g = torch.Generator().manual_seed(seed)
        dev = torch.device(device) if device is not None else V.device
        C = V.shape[1]
        self.ks = [int(p.shape[1]) for p in P_list]

        def lin(out_f, in_f):
            bound = 1.0 / in_f ** 0.5
            w = (torch.rand((out_f, in_f), generator=g) * 2 - 1) * bound
            b = (torch.rand((out_f,), generator=g) * 2 - 1) * bound
            return w.to(dev), b.to(dev)

        self.V = V.clone().float().to(dev)
        self.w_rms = torch.ones(C, device=dev)
        self.W1, self.b1 = lin(hidden, C)
        self.W2, self.b2 = [], []
        for k in self.ks:
            w, b = lin(k, hidden)
            self.W2.append(w)
            self.b2.append(b)
        self.P = [p.clone().float().to(dev) for p in P_list]
        if state is not None:
            self.V = state["V"].clone().float().to(dev)
            self.w_rms = state["batch_norm.weight"].clone().float().to(dev)
            self.W1 = state["common_encoder.0.weight"].clone().float().to(dev)
            self.b1 = state["common_encoder.0.bias"].clone().float().to(dev)
            for i in range(len(self.ks)):
                self.W2[i] = state[f"multihead_encoder.heads.{i}.weight"].clone().float().to(dev)
                self.b2[i] = state[f"multihead_encoder.heads.{i}.bias"].clone().float().to(dev)
                self.P[i] = state[f"decoders.decoders.{i}.weight"].clone().float().to(dev)
        for t in self.parameters():
            t.requires_grad_(True)
        # five groups, one lr (:197-204); fused=True as the reference asks for
        self.opt = torch.optim.Adam([{"params": self.W2 + self.b2}, {"params": [self.W1, self.b1]},
                                     {"params": [self.w_rms]}, {"params": [self.V]}, {"params": self.P}],
                                    lr=lr, betas=(0.9, 0.95), fused=True)
        self.loss_fn = torch.nn.BCELoss(reduction="sum")

    def parameters(self) -> List[torch.Tensor]:
        return [self.V, self.w_rms, self.W1, self.b1, *self.W2, *self.b2, *self.P]

    def state_dict(self) -> dict:
        sd = {"V": self.V, "batch_norm.weight": self.w_rms, "common_encoder.0.weight": self.W1,
              "common_encoder.0.bias": self.b1}
        for i in range(len(self.ks)):
            sd[f"multihead_encoder.heads.{i}.weight"] = self.W2[i]
            sd[f"multihead_encoder.heads.{i}.bias"] = self.b2[i]
            sd[f"decoders.decoders.{i}.weight"] = self.P[i]
        return {k: v.detach().clone() for k, v in sd.items()}

    def encode(self, g: torch.Tensor):
        X = g.float() / 2
        X = torch.where(X == 1.5, 0.0, X)
        Z = X @ self.V
        Zn = F.rms_norm(Z, (Z.shape[1],), self.w_rms, 1e-8)
        Hh = torch.relu_(F.linear(Zn, self.W1, self.b1))
        return [torch.softmax(F.linear(Hh, w, b), dim=1) for w, b in zip(self.W2, self.b2)], X

    @staticmethod
    def gather(data: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """Dataset_admixture.__getitem__ per row + default collate (loaders.py:70-72)."""
        return torch.stack([data[int(i)] for i in idx])

    def step(self, g: torch.Tensor, y: Optional[torch.Tensor] = None, sup_weight: float = 100.0) -> float:
        self.opt.zero_grad(set_to_none=True)
        Qs, X = self.encode(g)
        loss = sum(self.loss_fn(torch.clamp_(F.linear(Q, P), 0, 1), X) for Q, P in zip(Qs, self.P))
        if y is not None:
            loss = loss + sup_weight * F.cross_entropy(Qs[0], y, reduction="sum")
        loss.backward()
        self.opt.step()
        with torch.no_grad():
            for P in self.P:
                P.clamp_(0.0, 1.0)
        return loss.item()

    @torch.inference_mode()
    def infer(self, data: torch.Tensor, batch: int = 1024) -> List[torch.Tensor]:
        outs = [[] for _ in self.ks]
        for s in range(0, data.shape[0], batch):
            Qs, _ = self.encode(data[s:s + batch])
            for o, q in zip(outs, Qs):
                o.append(q)
        return [torch.cat(o) for o in outs]
