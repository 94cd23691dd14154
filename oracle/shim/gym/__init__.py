"""Minimal gym stand-in (only used when the real gym is absent; oracle only)."""
from . import spaces  # noqa: F401
from .spaces import Space  # noqa: F401
