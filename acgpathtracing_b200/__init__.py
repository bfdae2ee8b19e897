"""B200-native path tracer for the hot path of fallinbryan/ACGPathTracing.

``acgpathtracing_b200.api`` binds the C-ABI shared library (include/ptb200.h); importing it raises if the
CUDA library has not been built -- there is no CPU fallback in the product path.
"""
from . import types, scenes  # noqa: F401

__all__ = ["types", "scenes"]
