from collections import OrderedDict

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def sample(self): raise NotImplementedError
    def contains(self, x): raise NotImplementedError
    def __contains__(self, x): return self.contains(x)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low, high = np.asarray(low), np.asarray(high)
            shape = low.shape
        else:
            low, high = np.full(shape, low), np.full(shape, high)
        self.low, self.high = low.astype(dtype), high.astype(dtype)
        super().__init__(shape, dtype)

    def sample(self):
        high = self.high if self.dtype.kind == "f" else self.high.astype("int64") + 1
        lo = np.where(np.isfinite(self.low), self.low, -1e3)
        hi = np.where(np.isfinite(high), high, 1e3)
        return self.np_random.uniform(low=lo, high=hi, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)

    def __repr__(self): return "Box" + str(self.shape)
    def __eq__(self, other): return isinstance(other, Box) and self.shape == other.shape and np.allclose(self.low, other.low) and np.allclose(self.high, other.high)


class Discrete(Space):
    def __init__(self, n):
        self.n = n
        super().__init__((), np.int64)

    def sample(self): return self.np_random.randint(self.n)
    def contains(self, x): return 0 <= int(x) < self.n
    def __eq__(self, other): return isinstance(other, Discrete) and self.n == other.n


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        super().__init__(self.nvec.shape, np.int64)

    def sample(self): return (self.np_random.random_sample(self.nvec.shape) * self.nvec).astype(self.dtype)
    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and (0 <= x).all() and (x < self.nvec).all()
    def __eq__(self, other): return isinstance(other, MultiDiscrete) and np.all(self.nvec == other.nvec)


class Dict(Space):
    def __init__(self, spaces=None, **kw):
        if spaces is None:
            spaces = kw
        if isinstance(spaces, dict) and not isinstance(spaces, OrderedDict):
            spaces = OrderedDict(sorted(list(spaces.items())))
        if isinstance(spaces, list):
            spaces = OrderedDict(spaces)
        self.spaces = spaces
        super().__init__(None, None)

    def seed(self, seed=None): return [s.seed(seed) for s in self.spaces.values()]
    def sample(self): return OrderedDict([(k, s.sample()) for k, s in self.spaces.items()])
    def contains(self, x): return isinstance(x, dict) and all(k in x and s.contains(x[k]) for k, s in self.spaces.items())
    def __getitem__(self, key): return self.spaces[key]
    def __iter__(self): return iter(self.spaces)
    def __eq__(self, other): return isinstance(other, Dict) and self.spaces == other.spaces


class Tuple(Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)
        super().__init__(None, None)

    def sample(self): return tuple(s.sample() for s in self.spaces)
    def contains(self, x): return len(x) == len(self.spaces) and all(s.contains(p) for s, p in zip(self.spaces, x))
