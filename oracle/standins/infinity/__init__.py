"""Stand-in for the un-vendored, unpinned third-party package ``infinity``.

TEST INFRASTRUCTURE ONLY (oracle side).  The reference declares ``infinity`` in
pyproject.toml:20 but it is not installed here and there is no network.  This
module provides the surface the reference uses:
``inf``, ``Infinity``, ``is_infinite`` with int comparison and absorbing
``+``/``-`` (call sites: utils/dynamic_programming.py:18,174,191,228-238,445;
compute/unordered_super_reconciliation.py:15,184).
"""
from functools import total_ordering


@total_ordering
class Infinity:
    __slots__ = ("positive",)

    def __init__(self, positive=True):
        self.positive = positive

    def __neg__(self):
        return Infinity(not self.positive)

    def __eq__(self, other):
        if isinstance(other, Infinity):
            return self.positive == other.positive
        if isinstance(other, float):
            return other == (float("inf") if self.positive else float("-inf"))
        return False

    def __ne__(self, other):
        return not self == other

    def __lt__(self, other):
        if self == other:
            return False
        return not self.positive

    def __gt__(self, other):
        if self == other:
            return False
        return self.positive

    def __le__(self, other):
        return self == other or self < other

    def __ge__(self, other):
        return self == other or self > other

    def __hash__(self):
        return hash(float("inf") if self.positive else float("-inf"))

    def __add__(self, other):
        if is_infinite(other) and other != self:
            raise TypeError("inf + -inf is undefined")
        return self

    __radd__ = __add__

    def __sub__(self, other):
        if is_infinite(other) and other == self:
            raise TypeError("inf - inf is undefined")
        return self

    def __rsub__(self, other):
        return -self

    def __mul__(self, other):
        if other == 0:
            raise TypeError("inf * 0 is undefined")
        if isinstance(other, Infinity):
            return Infinity(self.positive == other.positive)
        return Infinity(self.positive == (other > 0))

    __rmul__ = __mul__

    def __bool__(self):
        return True

    def __repr__(self):
        return "inf" if self.positive else "-inf"

    __str__ = __repr__

    def __float__(self):
        return float("inf") if self.positive else float("-inf")


inf = Infinity()


def is_infinite(value):
    return value == inf or value == -inf
