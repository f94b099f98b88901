"""Mappings from the nodes of one tree to the nodes of another.

Mirrors ``/root/reference/src/superrec2/model/tree_mapping.py``: plain
``{name: name}`` dictionaries on disk, ``{node: node}`` in memory, and the
``<species>_<suffix>`` leaf naming convention (``model/tree_mapping.py:25-57``).
"""
from typing import Dict, Mapping

from .tree import Tree, TreeNode

TreeMapping = Mapping[TreeNode, TreeNode]


def parse_tree_mapping(from_tree: Tree, to_tree: Tree,
                       data: Mapping[str, str]) -> TreeMapping:
    """Turn ``{name: name}`` into ``{node: node}`` (first level-order match)."""
    src = _first_by_name(from_tree)
    dst = _first_by_name(to_tree)
    result = {}
    for from_name, to_name in data.items():
        try:
            result[src[str(from_name)]] = dst[str(to_name)]
        except KeyError as exc:
            from .tree import TreeError
            raise TreeError(f"Node not found: {exc.args[0]!r}") from None
    return result


def _first_by_name(tree: Tree) -> Dict[str, TreeNode]:
    index: Dict[str, TreeNode] = {}
    for node in tree.traverse():
        index.setdefault(node.name, node)
    return index


def get_species_mapping(tree: Tree, species_tree: Tree) -> TreeMapping:
    """
    Map each leaf named ``<species>_<suffix>`` to the leaf species of that name
    (case-insensitive; shortest matching prefix followed by an underscore).
    """
    by_lower = {}
    for leaf in species_tree:
        if leaf.name:
            by_lower[leaf.name.lower()] = leaf
    result = {}
    for leaf in tree:
        pieces = leaf.name.split("_")
        for cut in range(1, len(pieces)):
            found = by_lower.get("_".join(pieces[:cut]).lower())
            if found is not None:
                result[leaf] = found
                break
    return result


def serialize_tree_mapping(mapping: TreeMapping) -> Dict[str, str]:
    """Turn ``{node: node}`` into ``{name: name}``."""
    return {src.name: dst.name for src, dst in mapping.items()}
