"""Minimal stand-ins for the two gym space classes the reference's consumers touch
(``gym`` is not a dependency of this build).  The onpolicy code inspects
``space.__class__.__name__`` ('Box' / 'Discrete'), ``.shape``, ``.n`` and ``.dtype``
(onpolicy/utils/util.py:35-43, graph_buffer.py:45-165), so the class names matter."""
from __future__ import annotations

import numpy as np


class Discrete:
    def __init__(self, n: int):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64

    def sample(self):
        return int(np.random.randint(self.n))

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, other):
        return isinstance(other, Discrete) and other.n == self.n


class Box:
    def __init__(self, low=-np.inf, high=np.inf, shape=None, dtype=np.float32):
        self.low, self.high = low, high
        self.shape = tuple(shape)
        self.dtype = dtype

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {np.dtype(self.dtype).name})"

    def __eq__(self, other):
        return isinstance(other, Box) and other.shape == self.shape
