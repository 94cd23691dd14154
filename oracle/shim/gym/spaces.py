"""Minimal gym.spaces stand-in: the class names and fields the reference env files touch."""
from collections import OrderedDict

import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = dtype


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        Space.__init__(self, shape, dtype)
        self.low = low
        self.high = high


class Discrete(Space):
    def __init__(self, n):
        Space.__init__(self, (), np.int64)
        self.n = n


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        Space.__init__(self, self.nvec.shape, np.int64)


class MultiBinary(Space):
    def __init__(self, n):
        Space.__init__(self, (n,), np.int8)
        self.n = n


class Tuple(Space):
    def __init__(self, spaces):
        Space.__init__(self, None, None)
        self.spaces = tuple(spaces)

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]


class Dict(Space):
    def __init__(self, spaces=None, **kw):
        Space.__init__(self, None, None)
        self.spaces = OrderedDict(spaces if spaces is not None else kw)

    def __getitem__(self, k):
        return self.spaces[k]
