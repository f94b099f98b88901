"""Unordered synteny-labelled reconciliation on the GPU (``superdtl``).

Drop-in for ``usreconcile_extended_uspfs`` of the reference
(``/root/reference/src/superrec2/compute/unordered_super_reconciliation.py:523-542``).
Syntenies travel to the device as bitsets: bit *i* of a node's words is the
*i*-th gene family of the instance in ``sort_synteny`` order
(``model/synteny.py:21-33``), so a bitset reads back directly as the sorted list
the reference stores in its output.
"""
from __future__ import annotations

from typing import Dict, Iterable, List

import numpy as np

from ..model.synteny import sort_synteny


def family_index(leaf_syntenies) -> Dict[str, int]:
    """``{family: bit position}``, positions in canonical (natural-sort) order."""
    families = set()
    for synteny in leaf_syntenies.values():
        families.update(synteny)
    return {family: bit for bit, family in enumerate(sort_synteny(families))}


def leaf_bitsets(nodes: List, leaf_syntenies, families: Dict[str, int]) -> np.ndarray:
    """uint32 [len(nodes), W] bitsets of the leaves' families (zero rows for internal nodes)."""
    words = max(1, (len(families) + 31) // 32)
    bits = np.zeros((len(nodes), words), np.uint32)
    for u, node in enumerate(nodes):
        if node.children:
            continue
        for family in leaf_syntenies.get(node, ()):
            bit = families[family]
            bits[u, bit // 32] |= np.uint32(1 << (bit % 32))
    return bits


def bits_to_families(bits: Iterable[int], names: List[str]) -> List[str]:
    return [names[f] for f in range(len(names)) if int(bits[f // 32]) >> (f % 32) & 1]
