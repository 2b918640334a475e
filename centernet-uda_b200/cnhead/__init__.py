"""cnhead: B200-native CenterNet head path (detection loss, UDA losses, decode) behind a C ABI.

The plugin modules next to this package (``losses``, ``backends``, ``utils``) mirror the reference's
import names; this package holds the ctypes binding, the autograd wrappers, the batch-sharded
multi-GPU schedule and the synthetic-input generator.  Importing it does not load the shared
library -- the first kernel call does, and fails loudly if it is missing.
"""
from . import synthetic  # noqa: F401

__all__ = ["synthetic", "functional", "sharded"]


def __getattr__(name):
    if name in ("functional", "sharded", "_lib"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
