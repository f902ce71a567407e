"""Shim package: `torchdyn.core.NeuralODE` -> the B200 sampler (see shims/README.md)."""
from . import core  # noqa: F401

__version__ = "1.0.6+s2s_b200"
