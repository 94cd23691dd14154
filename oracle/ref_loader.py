"""Run the reference's OWN env files unmodified, on top of the restated physics shim.

TEST INFRASTRUCTURE ONLY (oracle).  Works only where /root/reference exists (this
container); it is how tests/golden/*.npz are produced (tests/golden/gen_golden.py).
Nothing on the GPU box imports this module.

How: pre-seed sys.modules['DynEnv'] with an empty package whose __path__
is /root/reference/DynEnv, so `DynEnv/__init__.py` (which drags in matplotlib, imageio,
h5py, cloudpickle) is skipped, and put the stand-ins for pymunk / pygame / gym
(oracle/shim) on sys.path.  The reference's rules, rewards, vision, noise and encoders
then execute from its own files: RoboCupEnvironment.py, DrivingEnvironment.py,
environment_base.py, cutils.py, Robot.py, Ball.py, Goalpost.py, Car.py, Pedestrian.py,
Obstacle.py, Road.py.

RNG tap: the reference draws from the global `random` / `np.random`.  `UniformTape`
replaces random.random / random.randint / np.random.permutation / np.random.randint /
np.random.rand by derivations from ONE stream of uniforms which is recorded, so the C
oracle (oracle/dynenv_oracle.c, tape mode) can replay exactly the same draws in the
same order.
"""
import contextlib
import importlib
import os
import random as _random
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("DYNENV_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "DynEnv"))


def load_reference():
    """Returns a namespace with the reference classes (RoboCupEnvironment, DrivingEnvironment,
    cutils module, ...)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if "DynEnv" not in sys.modules:
        pkg = types.ModuleType("DynEnv")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "DynEnv")]
        pkg.__package__ = "DynEnv"
        sys.modules["DynEnv"] = pkg
    ns = types.SimpleNamespace()
    ns.cutils = importlib.import_module("DynEnv.cutils")
    ns.environment_base = importlib.import_module("DynEnv.environment_base")
    ns.RoboCupEnvironment = importlib.import_module("DynEnv.RoboCupEnvironment").RoboCupEnvironment
    ns.DrivingEnvironment = importlib.import_module("DynEnv.DrivingEnvironment").DrivingEnvironment
    ns.Robot = importlib.import_module("DynEnv.Robot").Robot
    ns.Car = importlib.import_module("DynEnv.Car").Car
    ns.ObservationType = ns.cutils.ObservationType
    ns.NoiseType = ns.cutils.NoiseType
    ns.pymunk = importlib.import_module("pymunk")
    return ns


class UniformTape(object):
    """One recorded stream of U[0,1) doubles feeding every RNG entry point the reference uses.

    Derivations (the C oracle's tape mode uses the same ones):
      random.random()            -> u
      random.randint(a, b)       -> a + floor(u * (b - a + 1))
      np.random.randint(lo, hi, n) -> lo + floor(u_i * (hi - lo)), i < n
      np.random.rand(n)          -> u_0 .. u_{n-1}
      np.random.permutation(n)   -> Fisher-Yates from the top: for i = n-1 .. 1: j = floor(u * (i + 1)); swap
    """

    def __init__(self, seed=0, source=None):
        self._gen = np.random.Generator(np.random.PCG64(seed))
        self._source = None if source is None else list(source)
        self._pos = 0
        self.count = 0  # uniforms handed out so far
        self.log = []

    def u(self):
        if self._source is not None:
            v = self._source[self._pos]
            self._pos += 1
        else:
            v = float(self._gen.random())
        self.log.append(v)
        self.count += 1
        return v

    def take_log(self):
        out = np.asarray(self.log, dtype=np.float64)
        self.log = []
        return out

    # --- entry points
    def random(self):
        return self.u()

    def randint(self, a, b):
        return a + int(self.u() * (b - a + 1))

    def np_randint(self, low, high=None, size=None):
        if high is None:
            low, high = 0, low
        if size is None:
            return low + int(self.u() * (high - low))
        n = int(np.prod(size))
        out = np.array([low + int(self.u() * (high - low)) for _ in range(n)], dtype=np.int64)
        return out.reshape(size)

    def np_rand(self, *shape):
        if not shape:
            return self.u()
        n = int(np.prod(shape))
        return np.array([self.u() for _ in range(n)], dtype=np.float64).reshape(shape)

    def np_permutation(self, n):
        arr = np.arange(n) if np.isscalar(n) else np.array(n)
        for i in range(len(arr) - 1, 0, -1):
            j = int(self.u() * (i + 1))
            arr[i], arr[j] = arr[j], arr[i]
        return arr

    @staticmethod
    def regenerate(seed, start, stop):
        """The uniforms [start, stop) of the stream UniformTape(seed) hands out."""
        g = np.random.Generator(np.random.PCG64(int(seed)))
        return np.array([float(g.random()) for _ in range(int(stop))][int(start):], dtype=np.float64)

    @contextlib.contextmanager
    def installed(self):
        saved = (_random.random, _random.randint, np.random.randint, np.random.rand,
                 np.random.permutation)
        _random.random = self.random
        _random.randint = self.randint
        np.random.randint = self.np_randint
        np.random.rand = self.np_rand
        np.random.permutation = self.np_permutation
        try:
            yield self
        finally:
            (_random.random, _random.randint, np.random.randint, np.random.rand,
             np.random.permutation) = saved
