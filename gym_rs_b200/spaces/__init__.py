"""gym_rs::spaces -- Space / Discrete / BoxR (reference: src/spaces/{space,discrete,box_r}.rs)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Generic, TypeVar

from .. import _capi

T = TypeVar("T")


class Space:
    """src/spaces/space.rs:2-10"""

    def contains(self, value) -> bool:
        raise NotImplementedError


@dataclass(frozen=True)
class Discrete(Space):
    """`Discrete(pub usize)`, src/spaces/discrete.rs:12.  contains = value < n (:14-20)."""
    n: int

    def contains(self, value: int) -> bool:
        if value < 0:  # not representable as usize
            return False
        return bool(_capi.load().gymrs_discrete_contains(self.n, int(value)))


@dataclass(frozen=True)
class BoxR(Space, Generic[T]):
    """`BoxR<T> { low, high }`, src/spaces/box_r.rs:5-13 -- a plain pair; the reference gives it
    no `contains` impl, and neither does this."""
    low: T
    high: T
