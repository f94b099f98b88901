"""Syntenies: sets of gene families attached to the nodes of an object tree.

Mirrors ``/root/reference/src/superrec2/model/synteny.py``: families are plain
strings; the canonical order splits names on digit runs and compares the digit
runs numerically (``model/synteny.py:21-33``).  On the device a synteny is a
bitset whose bit *i* is the *i*-th family of the instance in that canonical
order, so a bitset prints directly as the sorted list.
"""
import re
from typing import Dict, Iterable, List, Mapping, Sequence, Set

from .tree import Tree, TreeNode

GeneFamily = str
Synteny = Iterable[GeneFamily]
OrderedSynteny = Sequence[GeneFamily]
UnorderedSynteny = Set[GeneFamily]
SyntenyMapping = Mapping[TreeNode, Synteny]

_DIGIT_RUNS = re.compile(r"([0-9]+)")


def _natural_key(family: str):
    return [int(piece) if piece.isdigit() else piece
            for piece in _DIGIT_RUNS.split(family)]


def sort_synteny(synteny: Synteny) -> List[GeneFamily]:
    """Sort the families of a synteny in canonical (natural) order."""
    return sorted(synteny, key=_natural_key)


def parse_synteny_mapping(tree: Tree, data: Dict[str, Synteny]) -> SyntenyMapping:
    """Turn ``{node name: families}`` into ``{node: families}``."""
    return {tree & name: synteny for name, synteny in data.items()}


def serialize_synteny_mapping(mapping: SyntenyMapping) -> Dict[str, List[GeneFamily]]:
    """Turn ``{node: families}`` into ``{node name: list}`` (sets get sorted)."""
    return {
        node.name: sort_synteny(syn) if isinstance(syn, (set, frozenset)) else list(syn)
        for node, syn in mapping.items()
    }
