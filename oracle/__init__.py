"""CPU oracle for the reconciliation DP -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  ``superrec2_b200`` never does.
"""
