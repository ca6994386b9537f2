"""ConvergenceHistory -- host bookkeeping with the counts of reference src/history.jl:54-66,238-252."""
from __future__ import annotations

import math
from dataclasses import dataclass, field


@dataclass
class ConvergenceHistory:
    mvps: int = 0
    mtvps: int = 0
    iters: int = 0
    restart: int | None = None
    isconverged: bool = False
    data: dict = field(default_factory=dict)

    def __getitem__(self, key):
        return self.data[key]

    def __setitem__(self, key, value):
        self.data[key] = value

    @property
    def niters(self) -> int:            # niters(history)   src/history.jl:245
        return self.iters

    @property
    def nprods(self) -> int:            # nprods(history)   src/history.jl:238
        return self.mvps + self.mtvps

    @property
    def nrests(self) -> int:            # nrests(history)   src/history.jl:252
        if self.restart is None:
            raise ValueError("not a restarted method")
        return int(math.ceil(self.iters / self.restart))


def niters(history: ConvergenceHistory) -> int:
    """niters(history) -- reference src/history.jl:245."""
    return history.niters


def nprods(history: ConvergenceHistory) -> int:
    """nprods(history) -- reference src/history.jl:238."""
    return history.nprods


def nrests(history: ConvergenceHistory) -> int:
    """nrests(history) -- reference src/history.jl:252."""
    return history.nrests
