"""dynenv_b200 — B200-native batched simulator for DynEnv's rollout hot path."""
from . import layout  # noqa: F401
