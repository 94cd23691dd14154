"""dynenv_b200 — B200-native batched simulator for DynEnv's rollout hot path.

Importing the package never touches CUDA; constructing an environment does, and fails loudly
when the C-ABI library or a CUDA device is missing (there is no CPU fallback)."""
from . import layout  # noqa: F401
from .enums import DynEnvType, NoiseType, ObservationType  # noqa: F401


def __getattr__(name):
    # torch is imported lazily so that `import dynenv_b200.layout` stays light (oracle, fixtures)
    if name in ("BatchedRoboCup", "RoboCupOutputs", "BatchedDriving", "DrivingOutputs"):
        from . import batched
        return getattr(batched, name)
    if name in ("RoboCupEnvironment", "DrivingEnvironment", "make_dyn_env", "BatchedVecEnv"):
        from . import api
        return getattr(api, name)
    if name == "DeviceRolloutStorage":
        from . import rollout
        return rollout.DeviceRolloutStorage
    raise AttributeError(name)
