"""B200-native graph-augmented ViT hot path: hand-written sm_100a kernels behind a C ABI (``include/gvit.h``),
mirrored on the host by drop-in ``nn.Module``s with the reference's signatures (``modules``) and a
batch-sharded data-parallel gradient synchroniser (``dp``).  Importing the package does not load the CUDA
library; the first operator call does, and raises if it has not been built (there is no fallback path)."""
from . import _lib  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name in ("ops", "modules", "dp", "step", "optim", "checkpoint", "losses"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
