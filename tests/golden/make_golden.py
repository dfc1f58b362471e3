#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference in the build container.

Usage (build container only; /root/reference does not exist on the GPU box):

    mkdir -p /tmp/refbuild && cp -r /root/reference/. /tmp/refbuild/
    printf '__version__ = "0.0.0"\n__version_tuple__ = (0, 0, 0)\n' > /tmp/refbuild/neural_admixture/_version.py
    (cd /tmp/refbuild && python setup.py build_ext --inplace)       # Cython read_bed / rsvd / loglikelihood
    python tests/golden/make_golden.py /tmp/refbuild

The reference's modules are imported as they are (``neural_admixture.model.neural_admixture``: ``Q_P``,
``NeuralAdmixture``; ``neural_admixture.model.train.train``; ``neural_admixture.src.svd.RSVD``;
``neural_admixture.src.snp_reader.SNPReader``).  Two things are neutralised and recorded in every fixture's
``meta``: (1) ``torch.set_float32_matmul_precision('medium')`` (model/neural_admixture.py:349) is made a no-op, so
that the fp32 oracle is pinned at fp32 ('highest'); on this container's AMX CPU 'medium' silently runs ``X@V`` in
bf16 (hard part 1).  (2) nothing else.

Fixtures (all small npz):
  step_k5.npz        one optimisation step of Q_P (+ fused Adam + restrict_P) with missing codes and exact 0/1 P
  train_k3.npz       NeuralAdmixture.launch_training, 4 epochs, ragged last batch
  train_k3to5.npz    multi-head K=3..5
  train_sup_k3.npz   supervised mode
  bed_demo_slices.npz  1500 SNP rows of the demo .bed (+ injected missing fields), as stored and with the homozygous
                     fields exchanged, each with the N x M uint8 matrix SNPReader.read_data returns (flip in case b)
  rsvd_demo_slices.npz  the reference's rsvd.multiply_A_omega / multiply_QT_A / RSVD on the two matrices of
                     bed_demo_slices.npz (missing = 3 and, flipped, 255)
  train_k4to12.npz   multi-head K=4..12 (sum K = 72: BASELINE configs[3]'s head range), 2 epochs
  loglik.npz         utils_c.loglikelihood (utils.pyx:17-40) on fp64 Q / P with missing codes, exact 0/1 entries of P
                     and reconstructions outside [eps, 1 - eps]: values for K = 3 and K = 8
  demo_k7.npz        the shipped demo BED through read_bed -> RSVD -> GMM init -> 5 epochs (K=7, seed 42), plus the
                     shipped demo_run.7.{Q,P}.expected for reference
"""
import argparse
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent


def _import_reference(root: str):
    sys.path.insert(0, root)
    torch.set_float32_matmul_precision("highest")
    torch.set_float32_matmul_precision = lambda *_a, **_k: None  # neutralise neural_admixture.py:349
    from neural_admixture.model import neural_admixture as ref_model  # noqa
    return ref_model


def _sd_np(sd):
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def _flat(prefix, d):
    return {f"{prefix}{k}": v for k, v in d.items()}


def make_step(ref_model, out: Path):
    """One step exactly as ``_run_epoch`` does it (neural_admixture.py:403-414), on CPU uint8 data."""
    torch.manual_seed(123)
    rng = np.random.default_rng(123)
    B, M, K, C, H = 48, 203, 5, 8, 64
    G = rng.integers(0, 3, size=(B, M), dtype=np.uint8)
    G[rng.random((B, M)) < 0.03] = 3
    V = (rng.standard_normal((M, C)) / np.sqrt(M)).astype(np.float32)
    P = rng.uniform(0.02, 0.98, size=(K, M)).astype(np.float32)
    P[:, 5] = 0.0          # R == 0 exactly: the -X * 1e12 branch of the BCELoss backward
    # (a column with P == 1 for every k makes R = sum(Q) = 1 +- 1 ulp: the reference's own result then depends on
    #  fp32 summation order (mask 0 vs gradient 1e12), so that case is deliberately NOT part of the fixture)
    P[0, 17] = 0.0
    P[3, 23] = 1.0
    model = ref_model.Q_P(H, C, torch.tensor(V), torch.tensor(P), [K])
    init = _sd_np(model.state_dict())
    opt = model.create_custom_adam(device=torch.device("cpu"), lr=2e-3)
    loss_fn = torch.nn.BCELoss(reduction="sum")
    x = torch.tensor(G)
    losses = []
    grads = None
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        (recs, probs), xt = model(x)
        loss = sum(loss_fn(r, xt) for r in recs)
        loss.backward()
        if grads is None:
            grads = {n: p.grad.detach().numpy().copy() for n, p in model.named_parameters()}
            Q1 = probs[0].detach().numpy().copy()
            Z1 = (xt @ model.V).detach().numpy().copy()
        opt.step()
        model.restrict_P()
        losses.append(loss.item())
        if len(losses) == 1:
            after1 = _sd_np(model.state_dict())
    after2 = _sd_np(model.state_dict())
    np.savez_compressed(out, G=G, ks=np.array([K]), lr=2e-3, losses=np.array(losses), Q1=Q1, Z1=Z1,
                        **_flat("init/", init), **_flat("grad/", grads), **_flat("after1/", after1),
                        **_flat("after2/", after2),
                        meta="reference Q_P two steps; matmul precision pinned highest")


def _run_training(ref_model, out: Path, *, N, M, ks, k, min_k, max_k, batch, epochs, H, seed, lr=2e-3, pops=None,
                  data=None, V=None, P=None, extra=None):
    """Drive ``NeuralAdmixture.launch_training`` (neural_admixture.py:324-392) on CPU (pack2bit=None) and record the
    initial parameters (captured right after ``initialize_model``), per-epoch loss sums, the sampler orders, and the
    final Qs / Ps."""
    C = V.shape[1]
    captured = {}
    orig_init = ref_model.NeuralAdmixture.initialize_model

    def init_and_capture(self, *a, **kw):
        orig_init(self, *a, **kw)
        captured["init"] = _sd_np(self.raw_model.state_dict())

    epoch_losses = []
    orig_epoch = ref_model.NeuralAdmixture._run_epoch
    orig_epoch_sup = ref_model.NeuralAdmixture._run_epoch_supervised

    class _Tap:
        def __init__(self):
            self.acc = 0.0

    def tap_loss(fn):
        def run(self, epoch, dataloader):
            # re-implements nothing: wraps loss.item via the optimizer step count is not possible, so we recompute the
            # epoch loss sum by intercepting torch.Tensor.item for scalars produced in this epoch.
            tap = _Tap()
            orig_item = torch.Tensor.item

            def item(t):
                val = orig_item(t)
                if t.requires_grad or t.grad_fn is not None:
                    tap.acc += val
                return val
            torch.Tensor.item = item
            try:
                fn(self, epoch, dataloader)
            finally:
                torch.Tensor.item = orig_item
            epoch_losses.append(tap.acc)
        return run

    ref_model.NeuralAdmixture.initialize_model = init_and_capture
    ref_model.NeuralAdmixture._run_epoch = tap_loss(orig_epoch)
    ref_model.NeuralAdmixture._run_epoch_supervised = tap_loss(orig_epoch_sup)
    try:
        torch.manual_seed(seed)
        na = ref_model.NeuralAdmixture(k, epochs, batch, lr, torch.device("cpu"), seed, 0, True, None, min_k, max_k)
        # the sampler stream the run will consume (same generator construction as neural_admixture.py:283)
        gen = torch.Generator().manual_seed(seed)
        orders = []
        for _ in range(epochs):
            orders.append(torch.randperm(N, generator=gen).numpy().copy())
            torch.randperm(N, generator=gen)  # RandomSampler draws a second, discarded permutation per epoch
        pops_t = None if pops is None else torch.as_tensor(pops, dtype=torch.int64)
        Qs, Ps, raw = na.launch_training(torch.tensor(P), torch.tensor(data), H, C, torch.tensor(V), M, N, pops_t)
    finally:
        ref_model.NeuralAdmixture.initialize_model = orig_init
        ref_model.NeuralAdmixture._run_epoch = orig_epoch
        ref_model.NeuralAdmixture._run_epoch_supervised = orig_epoch_sup
    final = _sd_np(raw.state_dict())
    payload = dict(G=data, ks=np.array(ks), lr=lr, batch=batch, epochs=epochs, seed=seed,
                   orders=np.stack(orders), epoch_losses=np.array(epoch_losses),
                   **_flat("init/", captured["init"]), **_flat("final/", final))
    for i, (q, p) in enumerate(zip(Qs, Ps)):
        payload[f"Q/{i}"] = q
        payload[f"P/{i}"] = p
    if pops is not None:
        payload["pops"] = np.asarray(pops)
    if extra:
        payload.update(extra)
    payload["meta"] = "reference NeuralAdmixture.launch_training on CPU; matmul precision pinned highest"
    np.savez_compressed(out, **payload)
    return Qs, Ps


def _synthetic(rng, N, M, Ktrue, miss=0.01):
    Pt = rng.uniform(0.05, 0.95, size=(Ktrue, M))
    Qt = rng.dirichlet(0.2 * np.ones(Ktrue), size=N)
    G = rng.binomial(2, Qt @ Pt).astype(np.uint8)
    G[rng.random((N, M)) < miss] = 3
    return G


def make_trainings(ref_model):
    rng = np.random.default_rng(7)
    N, M, C = 150, 403, 8
    G = _synthetic(rng, N, M, 3)
    V = np.linalg.qr(rng.standard_normal((M, C)))[0].astype(np.float32)
    P3 = rng.uniform(0.05, 0.95, size=(3, M)).astype(np.float32)
    _run_training(ref_model, HERE / "train_k3.npz", N=N, M=M, ks=[3], k=3, min_k=None, max_k=None, batch=64,
                  epochs=4, H=32, seed=7, data=G, V=V, P=P3)
    P345 = rng.uniform(0.05, 0.95, size=(3 + 4 + 5, M)).astype(np.float32)
    _run_training(ref_model, HERE / "train_k3to5.npz", N=N, M=M, ks=[3, 4, 5], k=None, min_k=3, max_k=5, batch=64,
                  epochs=3, H=32, seed=11, data=G, V=V, P=P345)
    pops = rng.integers(0, 3, size=N)
    _run_training(ref_model, HERE / "train_sup_k3.npz", N=N, M=M, ks=[3], k=3, min_k=None, max_k=None, batch=64,
                  epochs=3, H=32, seed=5, data=G, V=V, P=P3, pops=pops)


def make_k4to12(ref_model):
    """The head range of BASELINE.json configs[3] (multi-head K=4..12) at a size the reference finishes in seconds."""
    rng = np.random.default_rng(21)
    N, M, C = 160, 515, 8
    G = _synthetic(rng, N, M, 6)
    V = np.linalg.qr(rng.standard_normal((M, C)))[0].astype(np.float32)
    ks = list(range(4, 13))
    P = rng.uniform(0.05, 0.95, size=(sum(ks), M)).astype(np.float32)
    _run_training(ref_model, HERE / "train_k4to12.npz", N=N, M=M, ks=ks, k=None, min_k=4, max_k=12, batch=64,
                  epochs=2, H=32, seed=13, data=G, V=V, P=P)


def make_loglik(out: Path):
    """The reference's Cython log-likelihood (src/utils_c/utils.pyx:17-40, called at model/train.py:139,145)."""
    from neural_admixture.src.utils_c import utils as ref_utils_c
    rng = np.random.default_rng(33)
    N, M = 97, 1013
    res = {}
    G = rng.integers(0, 3, size=(N, M), dtype=np.uint8)
    G[rng.random((N, M)) < 0.04] = 3
    res["G"] = G
    for K in (3, 8):
        Q = rng.dirichlet(0.3 * np.ones(K), size=N)
        P = rng.uniform(0.0, 1.0, size=(M, K))
        P[rng.random((M, K)) < 0.15] = 0.0          # reconstructions at / below eps
        P[rng.random((M, K)) < 0.15] = 1.0          # ... and at / above 1 - eps
        Q32, P32 = Q.astype(np.float32), P.astype(np.float32)    # what the engine holds; the reference gets them as fp64
        ll = ref_utils_c.loglikelihood(np.ascontiguousarray(G), np.ascontiguousarray(P32.astype(np.float64)),
                                       np.ascontiguousarray(Q32.astype(np.float64)), K)
        res[f"Q{K}"], res[f"P{K}"], res[f"ll{K}"] = Q32, P32, np.float64(ll)
    np.savez_compressed(out, **res, meta="reference utils_c.loglikelihood (Cython, fp64), eps = 1e-6")


def make_demo(ref_model, ref_root: str):
    """The reference's only integration test (demo/run_demo.sh:4 + demo/run_diagnostics.py:9-26): K=7, 5 epochs,
    seed 42, through the reference's own read_bed, RSVD and GMM initialisation (model/train.py:47-69)."""
    from neural_admixture.src import utils as ref_utils
    from neural_admixture.src.svd import RSVD
    from neural_admixture.model import train as ref_train
    demo = Path(ref_root) / "demo"
    ref_utils.set_seed(42)
    data, pops, N, M = ref_utils.read_data(str(demo / "data" / "demo_data.bed"), None)
    Vt = RSVD(data, N, M, 8, 42)
    cap = {}
    orig = ref_model.NeuralAdmixture.launch_training

    def tap(self, P, dat, hidden, nfeat, V, M_, N_, pops_=None):
        cap.update(P=P.numpy().copy(), V=V.numpy().copy())
        return orig(self, P, dat, hidden, nfeat, V, M_, N_, pops_)

    # run the reference's init (PCA projection + GMM) but route the training through _run_training's recorder
    ref_model.NeuralAdmixture.launch_training = tap
    try:
        Ps, Qs, _ = ref_train.train(5, 800, 2e-3, 7, 42, torch.as_tensor(data), torch.device("cpu"), 0, 1024, True,
                                    Vt, None, None, None, 8)
    finally:
        ref_model.NeuralAdmixture.launch_training = orig
    exp_Q = np.loadtxt(demo / "outputs" / "demo_run.7.Q.expected")
    exp_P = np.loadtxt(demo / "outputs" / "demo_run.7.P.expected")
    print("demo: reference run vs shipped .expected  relF Q=%.3e P=%.3e" % (
        np.linalg.norm(Qs[0] - exp_Q) / np.linalg.norm(exp_Q), np.linalg.norm(Ps[0] - exp_P) / np.linalg.norm(exp_P)))
    # now the recorded run (same P_init / V; MLP initialisation pinned by torch.manual_seed(42) in _run_training)
    from oracle_pack import pack2bit  # noqa  (numpy restatement; only used to shrink the fixture)
    Qs2, Ps2 = _run_training(ref_model, HERE / "demo_k7.npz", N=N, M=M, ks=[7], k=7, min_k=None, max_k=None,
                             batch=800, epochs=5, H=1024, seed=42, data=np.ascontiguousarray(data), V=cap["V"],
                             P=cap["P"], extra=dict(expected_Q=exp_Q.astype(np.float32),
                                                    ref_full_pipeline_Q=Qs[0], ref_full_pipeline_P=Ps[0]))
    # shrink: store genotypes packed, P/V as float32 (they are), drop the unpacked matrix
    z = dict(np.load(HERE / "demo_k7.npz", allow_pickle=False))
    z["G_packed"] = pack2bit(z.pop("G"))
    z["M"] = np.array(M)
    np.savez_compressed(HERE / "demo_k7.npz", **z)


def make_bed(ref_root: str, out: Path):
    """PLINK .bed files through the reference's own reader (src/snp_reader.py: SNPReader.read_data -> Cython
    utils_c.read_bed + the `2 - G` allele flip).  Two cases: a slice of the shipped demo BED (mean < 1: no flip) and
    the same slice with the two homozygous fields exchanged in the file (mean >= 1: the reference flips it), both with
    N = 105 (N % 4 = 1: padding fields in the last byte of every SNP row)."""
    import shutil
    import tempfile
    from neural_admixture.src.snp_reader import SNPReader
    demo = Path(ref_root) / "demo" / "data" / "demo_data"
    N = sum(1 for _ in open(str(demo) + ".fam"))
    nb = (N + 3) // 4
    raw = np.fromfile(str(demo) + ".bed", dtype=np.uint8)
    magic, payload = raw[:3], raw[3:].reshape(-1, nb)
    M = 1500
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(payload.shape[0], size=M, replace=False))
    bed_a = payload[rows].copy()
    miss = rng.random(bed_a.shape) < 0.01                       # sprinkle missing fields (01) into sample 4b
    bed_a[miss] = (bed_a[miss] & 0xFC) | 0x01
    # exchange fields 00 <-> 11 everywhere (hom A1 <-> hom A2): the matrix mean rises above 1
    hi, lo = (bed_a >> 1) & 0x55, bed_a & 0x55
    same = ~(hi ^ lo) & 0x55
    bed_b = bed_a ^ (same | (same << 1))
    res = {"N": np.int64(N), "M": np.int64(M)}
    tmp = Path(tempfile.mkdtemp())
    try:
        for tag, bed in (("a", bed_a), ("b", bed_b)):
            base = tmp / f"case_{tag}"
            with open(str(base) + ".bed", "wb") as f:
                f.write(magic.tobytes())
                f.write(bed.tobytes())
            shutil.copy(str(demo) + ".fam", str(base) + ".fam")
            G = SNPReader().read_data(str(base) + ".bed")
            res[f"bed_{tag}"] = bed
            res[f"G_{tag}"] = np.asarray(G, dtype=np.uint8)
            res[f"flipped_{tag}"] = np.bool_(not (np.asarray(G) == _lut_read(bed, N)).all())
    finally:
        shutil.rmtree(tmp)
    assert not res["flipped_a"] and res["flipped_b"]
    np.savez_compressed(out, **res)


def _lut_read(bed, N):
    lut = np.array([2, 3, 1, 0], dtype=np.uint8)
    f = np.stack([(bed >> (2 * i)) & 3 for i in range(4)], axis=2).reshape(bed.shape[0], -1)[:, :N]
    return lut[f].T


def make_rsvd(ref_root: str, out: Path):
    """The reference's randomized-SVD building blocks on genotype matrices that contain missing values as the reference's
    reader leaves them (3, and 255 after its `2 - G` flip): the two Cython products (rsvd.pyx) and the full RSVD
    (svd.py:39-83, k = 8, seed 42)."""
    from neural_admixture.src.utils_c import rsvd as ref_rsvd
    from neural_admixture.src.svd import RSVD
    g = np.load(HERE / "bed_demo_slices.npz")
    rng = np.random.default_rng(11)
    res = {}
    for tag in "ab":
        A = np.ascontiguousarray(g[f"G_{tag}"])                  # 105 x 1500 uint8; case b holds 255 for missing
        N, M = A.shape
        Om = rng.standard_normal((M, 20)).astype(np.float32)
        QT = rng.standard_normal((20, N)).astype(np.float32)
        res[f"Omega_{tag}"], res[f"QT_{tag}"] = Om, QT
        res[f"Y_{tag}"] = ref_rsvd.multiply_A_omega(A, Om)
        res[f"B_{tag}"] = ref_rsvd.multiply_QT_A(QT, A)
        res[f"Vt_{tag}"] = RSVD(A, N, M, 8, 42).astype(np.float32)
    np.savez_compressed(out, **res)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ref_root", help="built scratch copy of /root/reference (see module docstring)")
    ap.add_argument("--only", default="", help="comma-separated subset: step,train,k4to12,loglik,demo,bed,rsvd")
    args = ap.parse_args()
    want = set(filter(None, args.only.split(",")))
    todo = lambda name: not want or name in want
    ref_model = _import_reference(args.ref_root)
    # numpy pack restatement without importing the oracle package path twice
    sys.path.insert(0, str(HERE.parent.parent / "oracle"))
    mod = types.ModuleType("oracle_pack")
    import nadm_oracle
    mod.pack2bit = nadm_oracle.pack2bit
    sys.modules["oracle_pack"] = mod
    if todo("step"):
        make_step(ref_model, HERE / "step_k5.npz")
    if todo("train"):
        make_trainings(ref_model)
    if todo("k4to12"):
        make_k4to12(ref_model)
    if todo("loglik"):
        make_loglik(HERE / "loglik.npz")
    if todo("demo"):
        make_demo(ref_model, args.ref_root)
    if todo("bed"):
        make_bed(args.ref_root, HERE / "bed_demo_slices.npz")
    if todo("rsvd"):
        make_rsvd(args.ref_root, HERE / "rsvd_demo_slices.npz")
    for f in sorted(HERE.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
