"""superrec2_b200 -- B200-native reconciliation DP behind superrec2's entry points.

One hot path of UdeM-LBIT/superrec2, rebuilt for NVIDIA B200 (sm_100a): the
min-cost object-tree x species-tree dynamic program behind
``superrec2 reconcile`` for the ``dtl`` (``thl``) and ``superdtl`` algorithms.
Host code is Python and mirrors the reference's module layout
(``model``, ``utils``, ``compute``, ``cli``); the DP tables, arg-min traceback
and solution selection run in hand-written CUDA kernels behind the C ABI
declared in ``include/sr2.h`` (``libsr2.so``, loaded with ctypes).

There is no CPU implementation of the DP in this package: importing it works
anywhere, but every entry point raises ``Sr2Error`` when ``libsr2.so`` or a
B200-class GPU is missing.
"""
__version__ = "0.1.0"
