"""Solution-retention policy of the reconciliation entry points.

The reference's generic ``Table``/``Entry`` machinery
(``/root/reference/src/superrec2/utils/dynamic_programming.py:149-278``) is
replaced on the device by int32 rows and packed arg-min keys; the only part of
that module which is visible at the drop-in boundary is the ``RetentionPolicy``
enum (``utils/dynamic_programming.py:89-101``) that ``cli/reconcile.py:96-100``
passes as the second argument of ``reconcile_thl`` / ``usreconcile_extended_uspfs``.
"""
from enum import Enum, auto


class RetentionPolicy(Enum):
    """Policy on retaining solutions."""

    # Only the minimum cost is of interest
    NONE = auto()

    # Return any one minimum-cost solution: the first one in the reference's
    # candidate order (``utils/dynamic_programming.py:228-238``)
    ANY = auto()

    # Return every minimum-cost solution
    ALL = auto()
