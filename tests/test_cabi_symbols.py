"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that include/nadm_b200.h
declares, the ctypes table matches the header, and the product package never imports the oracle.  No compute calls."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "nadm_b200.h"


def declared_symbols():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nadm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for need in ["nadm_pack2bit", "nadm_unpack2bit", "nadm_encoder_fwd", "nadm_mlp_fwd", "nadm_decoder_step",
                 "nadm_mlp_bwd", "nadm_encoder_bwd", "nadm_loglikelihood", "nadm_workspace_bytes", "nadm_last_error"]:
        assert need in syms


def test_library_exports_every_declared_symbol():
    from neural_admixture_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(str(lib_path))
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in nadm_b200.h but not exported by {lib_path.name}"
    assert lib.nadm_version() >= 100


def test_ctypes_table_matches_header():
    from neural_admixture_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.nadm_last_error() is not None


def test_argument_validation_without_a_gpu():
    """Bad arguments are rejected before any CUDA call: error code + message, like TORCH_CHECK in pack2bit.cu:67-76."""
    from neural_admixture_b200 import _lib
    lib = _lib.load()
    rc = lib.nadm_pack2bit(None, 4, 16, 16, None, 4, None)
    assert rc == -1 and b"NULL" in lib.nadm_last_error()
    rc = lib.nadm_encoder_fwd(ctypes.c_void_p(256), 20, None, 0, 4, 64, ctypes.c_void_p(256), 8, ctypes.c_void_p(256),
                              ctypes.c_void_p(256), 1 << 20, None)
    assert rc == -1 and b"pitch" in lib.nadm_last_error()
    with pytest.raises(_lib.NadmError):
        _lib.check(rc)


def test_product_never_imports_the_oracle():
    for py in (ROOT / "neural_admixture_b200").rglob("*.py"):
        src = py.read_text()
        assert "nadm_oracle" not in src and "import oracle" not in src and "from oracle" not in src, py


def test_no_cpu_path():
    import torch
    from neural_admixture_b200._lib import NadmError
    from neural_admixture_b200.model.neural_admixture import NeuralAdmixture, Q_P
    with pytest.raises(NadmError):
        NeuralAdmixture(3, 1, 8, 1e-3, torch.device("cpu"), 0, 0, True, None, None, None)
    m = Q_P(16, 8, torch.zeros(12, 8), torch.zeros(3, 12), [3])
    with pytest.raises(NadmError):
        m.bind()


def test_compiled_reference_checker_loads_and_exports_the_reference_surface():
    """oracle/_ref/pack2bit_ref.so (the reference's pack2bit.cu compiled unmodified by oracle/build_ref.py): present in the
    build container, loads without a GPU and exports the two functions of the reference's pybind module
    (pack2bit.cu:144-147).  Calling them needs a device: that is tests/test_gpu_parity.py."""
    import importlib.util
    from pathlib import Path

    import pytest
    so = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "pack2bit_ref.so"
    if not so.exists():
        pytest.skip("oracle/_ref not built here (python oracle/build_ref.py needs /root/reference)")
    import torch  # noqa: F401  (its libraries must be loaded first)
    spec = importlib.util.spec_from_file_location("pack2bit_ref", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert callable(mod.pack2bit_cpu_to_gpu) and callable(mod.unpack2bit_gpu_to_gpu)
