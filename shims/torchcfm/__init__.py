"""Shim package: the two torchcfm entry points the reference uses -> the B200 engine (see shims/README.md)."""
from . import conditional_flow_matching, models  # noqa: F401

__version__ = "1.0.7+s2s_b200"
