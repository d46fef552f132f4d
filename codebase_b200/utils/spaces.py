"""Minimal stand-ins for the gymnasium space classes the reference touches (gymnasium is not a dependency here):
`Box(shape, dtype)`, `Discrete(n)`, `Tuple(spaces)` and `flatdim` (used at marlbase/dqn/model.py:32-33,
marlbase/ac/model.py:40-41, marlbase/ac/train.py:33)."""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {np.dtype(self.dtype).name})"


class Discrete:
    def __init__(self, n):
        self.n, self.shape, self.dtype = int(n), (), np.int64

    def __repr__(self):
        return f"Discrete({self.n})"


class Tuple(tuple):
    def __new__(cls, spaces):
        return super().__new__(cls, tuple(spaces))

    @property
    def spaces(self):
        return tuple(self)


def flatdim(space) -> int:
    if isinstance(space, (tuple, list)):
        return sum(flatdim(s) for s in space)
    if getattr(space, "n", None) is not None:
        return int(space.n)
    return int(np.prod(space.shape))
