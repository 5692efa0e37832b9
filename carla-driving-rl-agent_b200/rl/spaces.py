"""Minimal stand-ins for the gym spaces the agent inspects (gym itself is optional)."""
import numpy as np


class Space:
    pass


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()
        self.dtype = dtype

    def is_bounded(self):
        return bool(np.all(np.isfinite(self.low)) and np.all(np.isfinite(self.high)))

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class Discrete(Space):
    def __init__(self, n):
        self.n = n
        self.shape = ()


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec)
        self.shape = self.nvec.shape


class Dict(Space):
    def __init__(self, spaces=None, **kwargs):
        self.spaces = dict(sorted((spaces or kwargs).items()))      # gym.spaces.Dict sorts its keys


def kind(space):
    """'box' | 'discrete' | 'multidiscrete' | 'dict' for either these classes or real gym spaces."""
    name = type(space).__name__.lower()
    if name in ('box', 'discrete', 'multidiscrete', 'dict'):
        return name
    raise ValueError('space must be one of Box, Discrete, MultiDiscrete, or Dict')
