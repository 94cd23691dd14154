"""gym.spaces when gym is installed, else minimal stand-ins with the same names and fields (the reference
only stores spaces and exposes them as attributes: DynEnv/RoboCupEnvironment.py:338-430)."""
try:  # pragma: no cover - depends on the environment
    from gym import Space
    from gym.spaces import Box, Dict, MultiBinary, MultiDiscrete, Tuple
except Exception:  # gym is not part of this image
    import numpy as np

    class Space(object):
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = dtype

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            Space.__init__(self, shape, dtype)
            self.low = np.full(self.shape, low, dtype=dtype)
            self.high = np.full(self.shape, high, dtype=dtype)

    class MultiDiscrete(Space):
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            Space.__init__(self, self.nvec.shape, np.int64)

    class MultiBinary(Space):
        def __init__(self, n):
            self.n = n
            Space.__init__(self, (n,), np.int8)

    class Dict(Space):
        def __init__(self, spaces):
            Space.__init__(self)
            self.spaces = dict(spaces)

    class Tuple(Space):
        def __init__(self, spaces):
            Space.__init__(self)
            self.spaces = tuple(spaces)

        def __getitem__(self, i):
            return self.spaces[i]

        def __len__(self):
            return len(self.spaces)
