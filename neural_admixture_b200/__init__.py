"""neural_admixture_b200 — B200-native (sm_100a) engine for the Neural ADMIXTURE per-minibatch hot path.

Mirrors the reference's module surface for that path (``neural_admixture.model.neural_admixture.{Q_P,
NeuralAdmixture}``, ``neural_admixture.model.train.train``, the ``pack2bit`` extension) on top of a C-ABI CUDA
library (``csrc/libnadm_b200.so``, declared in ``include/nadm_b200.h``).  No CPU fallback."""
__version__ = "0.1.0"
