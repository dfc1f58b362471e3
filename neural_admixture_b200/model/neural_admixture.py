"""B200 mirror of the reference's ``neural_admixture/model/neural_admixture.py`` for the per-minibatch hot path.

Same public names and argument meaning as the reference (``Q_P``, ``NeuralAdmixture`` with ``launch_training`` /
``process_results`` / ``display_divergences``; reference file cited per method), but none of its arithmetic runs in
PyTorch: every step is a fixed sequence of calls into ``libnadm_b200.so`` (``include/nadm_b200.h``) operating directly
on the 2-bit packed genotype matrix.  torch provides device memory, streams and ``torch.distributed`` only.

Differences a caller can see (all deliberate):
  * CUDA only.  There is no CPU / MPS path and no PyTorch fallback: a missing library or a non-CUDA device raises.
  * ``torch.set_float32_matmul_precision('medium')`` (reference :349) is not applied: all contractions are fp32.
  * multi-GPU is SNP-axis sharding with the single-process sampler and batch size (reference :287,:315-319 uses
    sample-axis DDP, which changes the gradient scale and the sampler); every rank passes its own column slice.
  * ``Q_P.forward`` in training mode is not materialised (the B x M reconstruction never exists); the training step
    is ``NeuralAdmixture``'s fused path.  In inference mode ``forward`` takes the reference's uint8 B x M input.
"""
from __future__ import annotations

import json
import logging
import os
import sys
import warnings
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import torch

from .. import ops
from .._lib import MlpParams, NadmError

logging.basicConfig(stream=sys.stdout, level=logging.INFO, format="%(message)s")
log = logging.getLogger(__name__)


class NeuralEncoder(torch.nn.Module):
    """Holder of the per-K heads ``Linear(H, k)`` (reference :15-49); state_dict keys ``heads.{i}.weight|bias``."""

    def __init__(self, input_size: int, ks: List[int]):
        super().__init__()
        self.ks = sorted(ks)
        self.min_k_val = min(self.ks)
        self.heads = torch.nn.ModuleList(torch.nn.Linear(input_size, k, bias=True) for k in self.ks)


class _DecoderWeight(torch.nn.Module):
    def __init__(self, w: torch.Tensor):
        super().__init__()
        self.weight = torch.nn.Parameter(w, requires_grad=False)


class NeuralDecoder(torch.nn.Module):
    """Holder of the per-K allele-frequency matrices: ``decoders[i].weight`` is P_k, M x k (reference :51-98 keeps it
    as the weight of ``Linear(k, M, bias=False)``).  ``inits`` is the (sum K) x M initial P, sliced per head as in
    reference :69-76; here each slice is stored contiguous M x k, the layout the kernels stream."""

    def __init__(self, inits: torch.Tensor, ks: List[int]):
        super().__init__()
        self.ks = sorted(ks)
        self.min_k_val = min(self.ks)
        mods, ini = [], 0
        for k in self.ks:
            mods.append(_DecoderWeight(inits[ini:ini + k].T.contiguous().clone()))
            ini += k
        self.decoders = torch.nn.ModuleList(mods)


class FusedAdamState:
    """What ``create_custom_adam`` returns: Adam(betas=(0.9, 0.95), eps=1e-8) state for every parameter of a
    ``Q_P`` (reference :187-204).  The update itself is fused into the kernels that produce each gradient."""

    def __init__(self, model: "Q_P", lr: float, betas=(0.9, 0.95), eps: float = 1e-8):
        self.lr, self.betas, self.eps, self.step_count = float(lr), (float(betas[0]), float(betas[1])), float(eps), 0
        model.bind()
        z = torch.zeros_like
        self.m = {"V": z(model.V), "w_rms": z(model.batch_norm.weight), "W1": z(model.common_encoder[0].weight),
                  "b1": z(model.common_encoder[0].bias), "W2": z(model.W2cat), "b2": z(model.b2cat),
                  "P": [z(d.weight) for d in model.decoders.decoders]}
        self.v = {k: ([z(t) for t in val] if isinstance(val, list) else z(val)) for k, val in self.m.items()}

    def hyper(self):
        return ops.adam_hyper(self.lr, self.step_count, self.betas[0], self.betas[1], self.eps)


class Q_P(torch.nn.Module):
    """Same constructor, attributes and state_dict keys as the reference ``Q_P`` (:100-230):
    ``V`` (M x C), ``batch_norm.weight`` (C), ``common_encoder.0.{weight,bias}``,
    ``multihead_encoder.heads.{i}.{weight,bias}``, ``decoders.decoders.{i}.weight`` (M x k)."""

    def __init__(self, hidden_size: int, num_features: int, V: Optional[torch.Tensor] = None,
                 P: Optional[torch.Tensor] = None, ks_list: List[int] = [], is_train: bool = True) -> None:
        super().__init__()
        self.ks_list = list(ks_list)
        self.V = torch.nn.Parameter(V.contiguous(), requires_grad=False) if V is not None else None
        self.num_features = num_features
        self.hidden_size = hidden_size
        # construction order = the reference's (:135-143), so the same torch seed gives the same initial weights
        self.batch_norm = torch.nn.RMSNorm(self.num_features, eps=1e-8)
        self.encoder_activation = torch.nn.ReLU(inplace=True)
        self.common_encoder = torch.nn.Sequential(
            torch.nn.Linear(self.num_features, self.hidden_size, bias=True), self.encoder_activation)
        self.multihead_encoder = NeuralEncoder(self.hidden_size, ks=self.ks_list)
        if P is not None:
            self.decoders = NeuralDecoder(P, ks=self.ks_list)
        for p in self.parameters():
            p.requires_grad_(False)
        self.is_train = is_train
        self.return_func = self._return_training if is_train else self._return_infer
        self.W2cat = None
        self.b2cat = None
        self._bufs: Dict[int, dict] = {}
        self.bind_epoch = 0        # bumped whenever bind() re-allocates the flat buffers (captured graphs key on it)

    # ---- binding: flat device buffers the kernels read; the per-head nn.Parameters become views of them ----------
    def bind(self) -> None:
        dev = self.batch_norm.weight.device
        if dev.type != "cuda":
            raise NadmError("Q_P runs on CUDA only (no CPU/MPS path): move the module to a cuda device first")
        heads = self.multihead_encoder.heads
        if self.W2cat is not None and self.W2cat.device == dev and heads[0].weight.data_ptr() == self.W2cat.data_ptr():
            return
        ks = self.multihead_encoder.ks
        self.W2cat = torch.cat([h.weight.data for h in heads], dim=0).contiguous()
        self.b2cat = torch.cat([h.bias.data for h in heads], dim=0).contiguous()
        off = 0
        for h, k in zip(heads, ks):
            h.weight.data = self.W2cat[off:off + k]
            h.bias.data = self.b2cat[off:off + k]
            off += k
        for t in (self.batch_norm.weight, self.common_encoder[0].weight, self.common_encoder[0].bias):
            t.data = t.data.contiguous()
        if self.V is not None:
            self.V.data = self.V.data.contiguous()
        self._bufs = {}
        self.bind_epoch += 1

    def _fwd_buffers(self, B: int) -> dict:
        buf = self._bufs.get(B)
        if buf is None:
            dev, H, C = self.W2cat.device, self.hidden_size, self.num_features
            sumK = self.W2cat.shape[0]
            f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
            M = self.V.shape[0]
            buf = {"Z": f(B, C), "rinv": f(B), "Hh": f(B, H), "Q": f(B, sumK),
                   "ws": torch.empty(ops.workspace_bytes(B, M, C, H, sumK), dtype=torch.uint8, device=dev)}
            self._bufs[B] = buf
        return buf

    def encode_packed(self, pg: ops.PackedGenotypes, *, row_idx: Optional[torch.Tensor] = None, row0: int = 0,
                      B: Optional[int] = None, allreduce=None, xchg=None) -> Tuple[List[torch.Tensor], dict]:
        """Forward of the encoder on rows of a packed matrix: Z = X V -> RMSNorm -> MLP -> per-head softmax
        (reference :169-176).  SNP-sharded runs sum the partial projection over the ranks either inside the MLP kernel
        (``xchg``: an ``ops.PeerExchange.xchg``, stores over NVLink into the peers' exchange areas) or with
        ``allreduce`` (a callable on the tensor, e.g. NCCL)."""
        self.bind()
        B = row_idx.numel() if row_idx is not None else B
        buf = self._fwd_buffers(B)
        # the sum over the encoder's CTAs is left to the MLP kernel unless something (NCCL) must read Z in between
        via_nccl = allreduce is not None and xchg is None
        ops.encoder_fwd(pg, self.V.data, buf["Z"], buf["ws"], row_idx=row_idx, row0=row0, B=B, deferred=not via_nccl)
        if via_nccl:
            allreduce(buf["Z"])
        ks = self.multihead_encoder.ks
        ops.mlp_fwd(buf["Z"], self.batch_norm.weight.data, self.common_encoder[0].weight.data,
                    self.common_encoder[0].bias.data, self.W2cat, self.b2cat, ks, buf["rinv"], buf["Hh"], buf["Q"],
                    xchg=xchg)
        probs, off = [], 0
        for k in ks:
            probs.append(buf["Q"][:, off:off + k])
            off += k
        return probs, buf

    def infer_packed(self, pg: ops.PackedGenotypes, batch: int = 2048, allreduce=None, xchg=None) -> List[torch.Tensor]:
        """Q for every row of a packed matrix, sequential row batches, outputs preallocated (the reference's
        inference loop, src/inference.py:71-77, and post-training Q pass, :369-383, grow Q with ``torch.cat`` per
        batch).  The batch size does not change any row's result.  ``allreduce`` sums the partial projections of the
        SNP shards (sharded runs: every rank passes its own column slice and receives the full Q)."""
        self.bind()
        ks = self.multihead_encoder.ks
        outs = [torch.empty((pg.N, k), dtype=torch.float32, device=pg.storage.device) for k in ks]
        for r0 in range(0, pg.N, batch):
            nb = min(batch, pg.N - r0)
            probs, _ = self.encode_packed(pg, row0=r0, B=nb, allreduce=allreduce, xchg=xchg)
            for o, p in zip(outs, probs):
                o[r0:r0 + nb].copy_(p)
        return outs

    def _return_training(self, probs):
        raise NadmError("Q_P.forward in training mode is not materialised by the B200 engine (the B x M reconstruction "
                        "never exists in memory); train through NeuralAdmixture.launch_training")

    def _return_infer(self, probs):
        return probs

    def forward(self, X: torch.Tensor):
        """Inference-mode forward with the reference's signature (:157-177): ``X`` is a uint8 B x M tensor of
        genotype codes on the device; returns ``(probs_list, X)``."""
        if self.return_func != self._return_infer:
            return self._return_training(None), X
        if X.dtype != torch.uint8 or X.dim() != 2 or not X.is_cuda:
            raise NadmError("Q_P.forward expects a uint8 B x M CUDA tensor of genotype codes")
        B, M = X.shape
        pg = ops.PackedGenotypes.empty(B, M, X.device)
        ops.pack2bit(X.contiguous(), pg.storage, M)
        probs, _ = self.encode_packed(pg, row0=0, B=B)
        return self.return_func([p.clone() for p in probs]), X

    @torch.no_grad()
    def restrict_P(self):
        """P in [0,1] (reference :179-185).  The fused decoder step already clamps; kept for API parity."""
        for dec in self.decoders.decoders:
            dec.weight.data.clamp_(0., 1.)

    def create_custom_adam(self, device: torch.device, lr: float = 1e-5) -> FusedAdamState:
        """Reference :187-204: one Adam, betas (0.9, 0.95), same lr for every group."""
        return FusedAdamState(self, lr)

    def save_config(self, name: str, save_dir: str) -> None:
        """``{name}_config.json`` with the reference's keys (:206-230)."""
        cfg = {"ks": self.ks_list, "num_features": self.num_features, "hidden_size": self.hidden_size,
               "activation": "relu"}
        with open(Path(save_dir) / f"{name}_config.json", "w") as fb:
            json.dump(cfg, fb)
        log.info("    Configuration file saved.")


def gather_rows(t: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate, in rank order, the row blocks ``t`` (rows x cols, row counts may differ between ranks) that the
    ranks of ``group`` hold: the full M x k matrix out of the SNP shards.  Every rank receives the result.  Works on
    any backend (NCCL on the device, gloo in the CPU tests)."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return t
    world = torch.distributed.get_world_size(group)
    if world == 1:
        return t
    rows = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(rows) for _ in range(world)]
    torch.distributed.all_gather(sizes, rows, group=group)
    sizes = [int(x.item()) for x in sizes]
    pad = torch.zeros((max(sizes),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    torch.distributed.all_gather(parts, pad, group=group)
    return torch.cat([p_[:n] for p_, n in zip(parts, sizes)], dim=0)


class NeuralAdmixture:
    """Reference ``NeuralAdmixture`` (:232-553) on the fused B200 path.  Constructor arguments are the reference's
    (:248-249)."""

    def __init__(self, k: int, epochs: int, batch_size: int, learning_rate: float, device: torch.device, seed: int,
                 num_gpus: int, master: bool, pack2bit, min_k: int, max_k: int,
                 supervised_loss_weight: Optional[float] = 100):
        self.k, self.min_k, self.max_k = k, min_k, max_k
        self.ks_list = [k] if k is not None else list(range(min_k, max_k + 1))
        self.num_gpus = num_gpus
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NadmError("the B200 engine runs on CUDA devices only (no CPU path)")
        self.master = master
        self.seed = seed
        self.generator = torch.Generator().manual_seed(self.seed)   # reference :283
        self.epochs = epochs
        # SNP-axis sharding keeps the global batch on every rank (the reference's sample-axis DDP divides it, :287)
        self.batch_size = batch_size
        self.lr = learning_rate
        self.supervised_loss_weight = supervised_loss_weight
        self.pack2bit = pack2bit
        self.sharded = bool(num_gpus > 1 and torch.distributed.is_available() and torch.distributed.is_initialized())
        self.loss_history: List[float] = []
        self.exchange: Optional[ops.PeerExchange] = None    # fused peer exchange of the sharded step (else NCCL)

    # ---- model ---------------------------------------------------------------------------------------------------
    def initialize_model(self, P: torch.Tensor, hidden_size: int, num_features: int, V: torch.Tensor,
                         ks_list: List[int]) -> None:
        """Reference :298-322 (no DDP wrapper: the only exchange is two small all-reduces per step)."""
        self.base_model = Q_P(hidden_size, num_features, V, P, ks_list).to(self.device)
        self.base_model.bind()
        if self.sharded:
            # identical replicated parameters on every rank (DDP's constructor broadcast, reference :317)
            for t in (self.base_model.batch_norm.weight, self.base_model.common_encoder[0].weight,
                      self.base_model.common_encoder[0].bias, self.base_model.W2cat, self.base_model.b2cat):
                torch.distributed.broadcast(t.data, src=0)
        self.model = self.base_model
        self.raw_model = self.base_model

    def _allreduce(self, t: torch.Tensor) -> None:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)

    def open_exchange(self, max_rows: int = 2048) -> None:
        """Sharded runs: set up the peer-mapped exchange areas (collective).  ``NADM_XCHG=nccl`` keeps the two NCCL
        all-reduces per step instead (the baseline the fused exchange is measured against); so does a node where CUDA
        IPC between the ranks' devices is not possible."""
        if not self.sharded or self.exchange is not None or os.environ.get("NADM_XCHG", "peer") == "nccl":
            return
        if torch.distributed.get_world_size() > 8:
            return
        width = max(self.raw_model.num_features, sum(self.ks_list))
        ex = ops.PeerExchange(max(max_rows, self.batch_size) * width + 1, self.device)
        if ex.ok:
            self.exchange = ex
        else:
            warnings.warn("nadm_b200: CUDA IPC between the ranks' devices is not available; the sharded step uses NCCL "
                          "all-reduces instead of the fused peer exchange", RuntimeWarning, stacklevel=2)
            ex.close()

    def close_exchange(self) -> None:
        """Collective: drop the step graphs that reference the exchange areas, then unmap / free them."""
        if self.exchange is not None:
            self.release_graphs()
            self.exchange.close()
            self.exchange = None

    def comm(self) -> dict:
        """Keyword arguments that make ``Q_P.encode_packed`` / ``infer_packed`` sum over the SNP shards."""
        if not self.sharded:
            return {}
        return {"xchg": self.exchange.xchg} if self.exchange is not None else {"allreduce": self._allreduce}

    def _step_buffers(self, B: int) -> dict:
        buf = self._train_bufs.get(B)
        if buf is None:
            sumK = sum(self.ks_list)
            dev = self.device
            dq_loss = torch.zeros(B * sumK + 1, dtype=torch.float32, device=dev)
            buf = {"dq_loss": dq_loss, "dQ": dq_loss[:B * sumK].view(B, sumK), "loss": dq_loss[B * sumK:],
                   "dZ": torch.empty((B, self.raw_model.num_features), dtype=torch.float32, device=dev)}
            self._train_bufs[B] = buf
        return buf

    def _mlp_params(self) -> MlpParams:
        m, o = self.raw_model, self.optimizer
        p = MlpParams()
        tensors = {"w_rms": m.batch_norm.weight.data, "W1": m.common_encoder[0].weight.data,
                   "b1": m.common_encoder[0].bias.data, "W2": m.W2cat, "b2": m.b2cat}
        for n, t in tensors.items():
            setattr(p, n, t.data_ptr())
            setattr(p, "m_" + n, o.m[n].data_ptr())
            setattr(p, "v_" + n, o.v[n].data_ptr())
            setattr(p, "g_" + n, None)
        return p

    def _train_step(self, row_idx: Optional[torch.Tensor], labels: Optional[torch.Tensor],
                    loss_out: Optional[torch.Tensor], pg: Optional[ops.PackedGenotypes] = None, hyper=None,
                    managed_loss: bool = False) -> None:
        """One minibatch: the body of the reference's ``_run_epoch`` loop (:403-414) — forward, loss, backward,
        Adam on every parameter, P clamp — as 5 library calls.  ``loss_out`` (1 float on device) receives the step's
        loss; nothing is synchronised with the host.  ``loss_out=None`` skips the evaluation of the reconstruction loss
        (its value never feeds the backward; the reference only logs it, :414-417).  The batch is rows ``row_idx`` of
        the resident matrix, or all rows of ``pg`` (a staged batch, see ``train_from_host``).  ``hyper``: Adam
        hyper-parameters whose coefficients live on the device (graph replay, see ``train_steps``); by default the
        host-side step count is advanced and its coefficients are passed by value.  ``managed_loss``: the caller
        (``nadm_step_begin`` / ``nadm_step_end``) zeroes and collects the step's loss accumulator itself."""
        m, o = self.raw_model, self.optimizer
        if pg is None:
            pg = self.packed
            B = row_idx.numel()
        else:
            row_idx, B = None, pg.N
        xc = self.exchange.xchg if self.exchange is not None else None
        probs, fb = m.encode_packed(pg, row_idx=row_idx, row0=0, B=B, **self.comm())
        sb = self._step_buffers(B)
        if not managed_loss:
            sb["loss"].zero_()
        if hyper is None:
            o.step_count += 1
            hyper = o.hyper()
        off = 0
        ks = m.multihead_encoder.ks
        via_nccl = self.sharded and xc is None
        for i, k in enumerate(ks):
            # the last head's sum over the decoder's CTAs is left to the MLP backward kernel (unless NCCL reads dQ first)
            ops.decoder_step(pg, fb["Q"], sb["dQ"], off, k, m.decoders.decoders[i].weight.data, o.m["P"][i], o.v["P"][i],
                             hyper, sb["loss"] if loss_out is not None else None, fb["ws"], row_idx=row_idx,
                             deferred=(i == len(ks) - 1) and not via_nccl)
            off += k
        if self.sharded and xc is None:
            self._allreduce(sb["dq_loss"])
        ops.mlp_bwd(sb["dQ"], fb["Q"], fb["Hh"], fb["Z"], fb["rinv"], m.multihead_encoder.ks, self._mlp_params(), hyper,
                    sb["dZ"], sb["loss"], fb["ws"], labels=labels,
                    sup_weight=float(self.supervised_loss_weight) if labels is not None else 0.0, xchg=xc,
                    deferred_apply=True)         # (the update rides along on the encoder backward below)
        ops.encoder_bwd(pg, sb["dZ"], m.V.data, o.m["V"], o.v["V"], hyper, fb["ws"], row_idx=row_idx)
        if loss_out is not None and not managed_loss:
            loss_out.copy_(sb["loss"])

    # ---- CUDA-graph replayed steps ------------------------------------------------------------------------------------
    use_graph = os.environ.get("NADM_NO_GRAPH", "0") != "1"
    graph_fallback: Optional[str] = None   # why this instance stopped replaying graphs (None: it has not)
    graph_kernel_launches = 0     # library kernels executed through graph replays (they bypass nadm_launch_count)
    generic_kernel_launches = 0   # of which first-generation CUDA-core kernels (shapes outside the tensor-core path)

    def _graph_state(self, order_len: int) -> dict:
        gs = getattr(self, "_gs", None)
        if gs is None or gs["order"].numel() < order_len:
            dev = self.device
            gs = {"order": torch.zeros(order_len, dtype=torch.int64, device=dev),
                  "counters": torch.zeros(4, dtype=torch.int64, device=dev),
                  "coef": torch.zeros(8, dtype=torch.float32, device=dev),
                  "loss1": torch.zeros(1, dtype=torch.float32, device=dev),
                  "losses": torch.zeros(order_len, dtype=torch.float32, device=dev), "idx": {}, "graphs": {},
                  "pops": None}   # "pops": persistent device copy of the labels (its ADDRESS is baked into the graphs)
            self._gs = gs
        return gs

    def _get_graph(self, gs: dict, Bs: int, want_loss: bool, sup: bool):
        """The whole step for a minibatch of ``Bs`` rows as ONE replayable CUDA graph: ``nadm_step_next`` (finishes the
        previous step — its loss into the per-step array, the counters advanced — then the rows of this minibatch out of
        the device-resident permutation + this step's Adam coefficients, both indexed by a device counter) and the
        five hot-path calls (sharded: with their exchanges inside).  Nothing in the graph depends on host state, so an
        epoch is ``nsteps`` graph launches + one ``nadm_step_flush`` for the last step."""
        key = (Bs, want_loss, sup, self.batch_size, self.raw_model.bind_epoch)
        g = gs["graphs"].get(key)
        if g is not None:
            return g
        o = self.optimizer
        self.raw_model.bind()
        self.raw_model._fwd_buffers(Bs)
        loss_acc = self._step_buffers(Bs)["loss"]
        idx = gs["idx"].setdefault(Bs, torch.zeros(Bs, dtype=torch.int64, device=self.device))
        h_host = ops.adam_hyper(o.lr, 0, o.betas[0], o.betas[1], o.eps)
        h_dev = ops.adam_hyper(o.lr, 0, o.betas[0], o.betas[1], o.eps, device_coef=gs["coef"])
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        before, gen_before = ops.launch_count(), ops.generic_launch_count()
        with torch.cuda.graph(g):
            ops.step_next(gs["order"], gs["counters"], self.batch_size, Bs, idx, h_host, gs["coef"], loss_acc, want_loss,
                          gs["losses"])
            labels = gs["pops"][idx] if sup else None
            self._train_step(idx, labels, loss_acc if want_loss else None, hyper=h_dev, managed_loss=True)
        g.nadm_kernels = ops.launch_count() - before          # library kernels per replay (bench.py's gpu_launches)
        g.nadm_generic = ops.generic_launch_count() - gen_before
        self._warn_generic(g.nadm_generic)
        gs["graphs"][key] = g
        return g

    def release_graphs(self) -> None:
        """Drop the captured step graphs (they hold the sharded step's NCCL kernels: release them before the process
        group is destroyed)."""
        if getattr(self, "_gs", None) is not None or getattr(self, "_host_stage", None) is not None:
            torch.cuda.synchronize(self.device)
            self._gs = None
            self._host_stage = None

    def train_steps(self, order_dev: torch.Tensor, nsteps: int, want_loss: bool = True,
                    pops: Optional[torch.Tensor] = None, first: int = 0) -> Optional[torch.Tensor]:
        """``nsteps`` consecutive minibatches of the row order ``order_dev`` (int64, device): minibatch ``s`` is
        ``order_dev[(first + s) * batch_size : (first + s + 1) * batch_size]`` (the last one may be ragged).  The body of
        the reference's epoch loop (:403-414) per minibatch; returns the per-step losses (device tensor) or None.
        Steps are replayed CUDA graphs unless ``NADM_NO_GRAPH=1`` (or capture is impossible), then eager calls."""
        o, Bfull, n = self.optimizer, self.batch_size, order_dev.numel()
        if self.use_graph:
            try:
                gs = self._graph_state(n)
                if gs["order"].data_ptr() != order_dev.data_ptr():
                    gs["order"][:n].copy_(order_dev)
                if pops is not None:
                    # the graphs read the labels through THIS buffer: copy, never rebind (a rebound tensor would leave
                    # the captured gather pointing at freed memory)
                    if gs["pops"] is None or gs["pops"].shape != pops.shape:
                        gs["pops"] = torch.empty_like(pops)
                        gs["graphs"] = {k_: g_ for k_, g_ in gs["graphs"].items() if not k_[2]}
                    gs["pops"].copy_(pops)
                gs["counters"].copy_(torch.tensor([first, o.step_count, 0, 0], dtype=torch.int64))
                graphs = [self._get_graph(gs, min(Bfull, n - (first + s) * Bfull), want_loss, pops is not None)
                          for s in range(nsteps)]
            except Exception as e:  # capture not possible (e.g. a collective that cannot be captured): eager steps
                self.graph_fallback = f"{type(e).__name__}: {e}"
                warnings.warn(f"nadm_b200: CUDA-graph capture of the training step failed ({self.graph_fallback}); "
                              "this engine falls back to eager launches (about 0.16 ms more host time per step)",
                              RuntimeWarning, stacklevel=2)
                self.use_graph = False           # this instance only
                graphs = None
            if graphs is not None:
                for g in graphs:
                    g.replay()
                    self.graph_kernel_launches += g.nadm_kernels
                    self.generic_kernel_launches += g.nadm_generic
                ops.step_flush(gs["counters"], gs["losses"])          # the last step's loss, the counters' last advance
                o.step_count += nsteps
                return gs["losses"][first:first + nsteps] if want_loss else None
        losses = torch.zeros(nsteps, dtype=torch.float32, device=self.device) if want_loss else None
        gen_before = ops.generic_launch_count()
        for s in range(nsteps):
            idx = order_dev[(first + s) * Bfull:(first + s + 1) * Bfull]
            labels = pops[idx].contiguous() if pops is not None else None
            self._train_step(idx, labels, losses[s:s + 1] if want_loss else None)
        gen = ops.generic_launch_count() - gen_before
        self.generic_kernel_launches += gen
        self._warn_generic(gen)
        return losses

    _warned_generic = False

    def _warn_generic(self, n: int) -> None:
        if n > 0 and not self._warned_generic:
            self._warned_generic = True
            warnings.warn(f"nadm_b200: this configuration (heads K={self.ks_list}, C={self.raw_model.num_features}, "
                          f"batch {self.batch_size}) leaves the tensor-core kernels for {n} launch(es) per step and runs "
                          "the first-generation CUDA-core kernels instead (several times slower; see DESIGN.md)",
                          RuntimeWarning, stacklevel=3)

    def epoch_order(self, N: int) -> torch.Tensor:
        """Row order of one epoch: exactly what the reference's ``RandomSampler(dataset, generator=self.generator)``
        yields (src/loaders.py:29-30, generator from :283)."""
        sampler = torch.utils.data.RandomSampler(range(N), generator=self.generator)
        return torch.tensor(list(sampler), dtype=torch.int64)

    def _run_epoch(self, epoch: int, order_dev: torch.Tensor, pops: Optional[torch.Tensor]) -> None:
        """Reference :394-417 / :434-458.  The per-step ``loss.item()`` host sync of the reference is replaced by one
        device->host read per epoch of the per-step losses, summed in the same order."""
        N = order_dev.numel()
        nsteps = (N + self.batch_size - 1) // self.batch_size
        every = 2 if pops is not None else 5
        want_loss = (epoch % every == 0) or self.keep_loss_history     # the epochs whose loss the reference prints
        losses = self.train_steps(order_dev, nsteps, want_loss, pops)
        loss_acc = float(losses.double().sum().item()) if want_loss else None
        if loss_acc is not None:
            self.loss_history.append(loss_acc)
        if epoch % every == 0 and self.master:
            log.info(f"            Loss in epoch {epoch:3d} on device {self.device} is {loss_acc:,.0f}")

    keep_loss_history = False

    def prepare(self, P: torch.Tensor, data, hidden_size: int, num_features: int, V: torch.Tensor, M: int, N: int):
        """Everything ``launch_training`` does before its epoch loop (reference :343-364): adopt the packed matrix,
        build the model on the device, create the optimizer state."""
        self.M, self.N = M, N
        if isinstance(data, ops.PackedGenotypes) or data is None:
            self.packed = data
        else:
            if not (torch.is_tensor(data) and data.is_cuda and data.dtype == torch.uint8):
                raise NadmError("launch_training needs the 2-bit packed genotype matrix on the CUDA device")
            self.packed = ops.PackedGenotypes.from_reference_layout(data, M)
        if self.packed is not None and (self.packed.N != N or self.packed.M != M):
            raise NadmError(f"packed matrix is {self.packed.N} x {self.packed.M}, expected {N} x {M}")
        self._train_bufs: Dict[int, dict] = {}
        self.initialize_model(P.to(self.device, torch.float32), hidden_size, num_features,
                              V.to(self.device, torch.float32), self.ks_list)
        self.optimizer = self.raw_model.create_custom_adam(device=self.device, lr=self.lr)
        self.open_exchange()

    def launch_training(self, P: torch.Tensor, data, hidden_size: int, num_features: int, V: torch.Tensor, M: int,
                        N: int, pops: Optional[torch.Tensor] = None):
        """Reference :324-392.  ``data`` is the device-resident 2-bit packed matrix: either a
        ``ops.PackedGenotypes`` or an N x ceil(M/4) uint8 CUDA tensor in the reference's layout (model/train.py:121).
        ``P`` is (sum K) x M, ``V`` is M x C.  In sharded mode M / P / V / data are this rank's SNP slice.
        Returns ``(Qs, Ps, raw_model)`` like the reference (numpy lists on the master rank)."""
        if data is None:
            raise NadmError("launch_training needs the 2-bit packed genotype matrix on the CUDA device")
        self.prepare(P, data, hidden_size, num_features, V, M, N)
        if pops is not None:
            pops = pops.to(self.device, torch.int64)

        if self.master:
            log.info("")
            log.info("    Starting training...")
            log.info("")
        for epoch in range(self.epochs):
            order = self.epoch_order(N).to(self.device, non_blocking=True)
            self._run_epoch(epoch, order, pops)

        # inference of Q for every sample, sequential batches of min(N, 1024) (reference :368-383)
        Qs = self.infer_Q(min(N, 1024))
        self.last_Q = Qs                     # device tensors, full N x k on every rank
        self.release_graphs()
        self.close_exchange()
        if self.master:
            log.info("")
            log.info("    Training finished!")
            log.info("")
        self.display_divergences(self.k)
        return self.process_results(Qs)

    def train_from_host(self, host_batches, labels=None) -> List[float]:
        """Out-of-core feed: train on minibatches whose 2-bit packed rows live in (pinned) HOST memory — each element
        of ``host_batches`` is a uint8 B x pitch tensor in the ``PackedGenotypes`` row layout.  Per step: the batch is
        copied host->device on a side stream into one of two staging buffers (overlapping the previous step's
        kernels), the fused step runs on it, and the step's loss is copied back to the host (the reference's
        ``loss.item()``, :414).  Returns the per-step losses.  Requires ``prepare`` (or ``launch_training``) first.

        With uniform batch shapes and no labels the step on each staging buffer is ONE replayed CUDA graph and nothing
        on the host waits for the device until the last step: the copies, the steps and the 4-byte loss read-backs are
        ordered by events only.  Otherwise (supervised labels, ragged shapes, ``NADM_NO_GRAPH=1``) every step is issued
        eagerly and its loss is read synchronously."""
        n = len(host_batches)
        if n == 0:
            return []
        main = torch.cuda.current_stream(self.device)
        copy = getattr(self, "_copy_stream", None)
        if copy is None:
            copy = self._copy_stream = torch.cuda.Stream(self.device)
        shape = tuple(host_batches[0].shape)
        uniform = all(tuple(hb.shape) == shape for hb in host_batches)
        hs = getattr(self, "_host_stage", None)
        if hs is None or hs["shape"] != shape:
            hs = self._host_stage = {
                "shape": shape, "graphs": {},
                "stage": [ops.PackedGenotypes(torch.empty(shape, dtype=torch.uint8, device=self.device), shape[0], self.M)
                          for _ in range(2)]}
        stage = hs["stage"]
        ready, free = [torch.cuda.Event(), torch.cuda.Event()], [None, None]

        def prefetch(i):
            hb, s_ = host_batches[i], i & 1
            st = stage[s_] if tuple(hb.shape) == shape else \
                ops.PackedGenotypes(torch.empty(hb.shape, dtype=torch.uint8, device=self.device), hb.shape[0], self.M)
            with torch.cuda.stream(copy):
                if free[s_] is not None:
                    copy.wait_event(free[s_])
                st.storage.copy_(hb, non_blocking=True)
                ready[s_].record(copy)
            return st

        if self.use_graph and uniform and labels is None:
            o = self.optimizer
            gs = self._graph_state(max(n, 1024))
            gs["counters"].copy_(torch.tensor([0, o.step_count, 0, 0], dtype=torch.int64))
            try:
                graphs = [self._get_host_graph(gs, hs, s_) for s_ in range(min(2, n))]
            except Exception as e:
                self.graph_fallback = f"{type(e).__name__}: {e}"
                warnings.warn(f"nadm_b200: CUDA-graph capture of the host-fed step failed ({self.graph_fallback}); "
                              "falling back to eager launches", RuntimeWarning, stacklevel=2)
                self.use_graph = False
                graphs = None
            if graphs is not None:
                loss_host = torch.zeros(n, dtype=torch.float32).pin_memory()
                prefetch(0)
                for i in range(n):
                    s_ = i & 1
                    if i + 1 < n:
                        prefetch(i + 1)
                    main.wait_event(ready[s_])
                    graphs[s_].replay()
                    self.graph_kernel_launches += graphs[s_].nadm_kernels
                    free[s_] = torch.cuda.Event()
                    free[s_].record(main)
                    loss_host[i:i + 1].copy_(gs["losses"][i:i + 1], non_blocking=True)   # this step's loss -> host
                main.synchronize()
                o.step_count += n
                return loss_host.tolist()

        loss_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        out: List[float] = []
        nxt = prefetch(0)
        for i in range(n):
            s_, cur = i & 1, nxt
            if i + 1 < n:
                nxt = prefetch(i + 1)
            main.wait_event(ready[s_])
            self._train_step(None, None if labels is None else labels[i], loss_dev, pg=cur)
            free[s_] = torch.cuda.Event()
            free[s_].record(main)
            loss_host.copy_(loss_dev, non_blocking=True)
            main.synchronize()
            out.append(float(loss_host[0]))
        return out

    def _get_host_graph(self, gs: dict, hs: dict, s_: int):
        """The host-fed step on staging buffer ``s_`` as one replayable graph (see ``_get_graph``; the minibatch is all
        rows of the staging buffer, so ``nadm_step_begin`` only provides the step's Adam coefficients and zeroes the
        loss accumulator)."""
        pg = hs["stage"][s_]
        key = (s_, pg.N, self.raw_model.bind_epoch, gs["counters"].data_ptr())
        g = hs["graphs"].get(key)
        if g is not None:
            return g
        o = self.optimizer
        self.raw_model.bind()
        self.raw_model._fwd_buffers(pg.N)
        loss_acc = self._step_buffers(pg.N)["loss"]
        idx = gs["idx"].setdefault(pg.N, torch.zeros(pg.N, dtype=torch.int64, device=self.device))
        h_host = ops.adam_hyper(o.lr, 0, o.betas[0], o.betas[1], o.eps)
        h_dev = ops.adam_hyper(o.lr, 0, o.betas[0], o.betas[1], o.eps, device_coef=gs["coef"])
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        before = ops.launch_count()
        with torch.cuda.graph(g):
            ops.step_begin(gs["order"], gs["counters"], pg.N, pg.N, idx, h_host, gs["coef"], loss_acc)
            self._train_step(None, None, loss_acc, pg=pg, hyper=h_dev, managed_loss=True)
            ops.step_end(gs["counters"], loss_acc, gs["losses"])
        g.nadm_kernels = ops.launch_count() - before
        hs["graphs"][key] = g
        return g

    def infer_Q(self, batch: int) -> List[torch.Tensor]:
        return self.raw_model.infer_packed(self.packed, batch, **self.comm())

    def gather_P(self) -> List[torch.Tensor]:
        """Full M x k P per head on every rank (concatenating the SNP shards in rank order)."""
        Ps = [d.weight.data for d in self.raw_model.decoders.decoders]
        return [gather_rows(P) for P in Ps] if self.sharded else Ps

    def gather_V(self) -> torch.Tensor:
        """Full M x C projection matrix on every rank (the ranks' row slices in rank order)."""
        V = self.raw_model.V.data
        return gather_rows(V) if self.sharded else V

    def display_divergences(self, k) -> None:
        """Hudson's Fst between estimated populations (reference :476-509)."""
        Ps = self.gather_P()
        if not self.master:
            return
        for P, k in zip(Ps, self.ks_list):
            header = "\t".join(f"Pop{p}" for p in range(k - 1))
            log.info("    Results:")
            log.info(f"\n            Fst divergences between estimated populations: (K = {k})")
            log.info("")
            log.info(f"                \t{header}")
            log.info("            Pop0")
            for j in range(1, k):
                out = f"            Pop{j}"
                for l in range(j):
                    out += f"\t{self._hudsons_fst(P[:, l], P[:, j]):0.3f}"
                log.info(out)
            log.info("\n")

    def process_results(self, Qs: List[torch.Tensor]):
        """Reference :511-530.  Sharded runs: the returned model carries the FULL M x C ``V`` (every rank trained only
        its row slice), so that ``state_dict()`` minus the decoders — what the reference saves, src/main.py:41 — is the
        same checkpoint a single-GPU run writes and ``src.inference.load_model`` can read it.  Training is over at
        this point: the per-shard optimizer state is not gathered."""
        Ps = self.gather_P()
        if self.sharded:
            self.shard_V = self.raw_model.V.data                 # this rank's trained slice (kept for inspection)
            self.raw_model.V = torch.nn.Parameter(self.gather_V().contiguous(), requires_grad=False)
            self.raw_model._bufs = {}
        if self.master:
            return [Q.cpu().numpy() for Q in Qs], [P.detach().cpu().numpy() for P in Ps], self.raw_model
        return [], [], self.raw_model

    @staticmethod
    def _hudsons_fst(pop1: torch.Tensor, pop2: torch.Tensor) -> float:
        """mean((p1-p2)^2) / (mean(p1(1-p2) + p2(1-p1)) + 1e-7)  (reference :532-553)."""
        try:
            num = torch.mean((pop1 - pop2) ** 2)
            den = torch.mean(pop1 * (1 - pop2) + pop2 * (1 - pop1)) + 1e-7
            return (num / den).item()
        except Exception as e:  # pragma: no cover
            log.info(f"            Error computing Hudson's Fst: {e}")
            return float("nan")
