"""Flatten host-side trees into the integer arrays of ``include/sr2.h``.

Species nodes are numbered by level-order (BFS) rank: that is the order in
which the reference offers candidates (``tree.traverse()``, e.g.
``/root/reference/src/superrec2/compute/reconciliation.py:79,93,136,286``), so
"smallest rank wins ties" reproduces its first-candidate-wins rule.  Object
nodes are numbered in preorder, which is also the key order of the reference's
output dictionaries (``compute/reconciliation.py:257-262``).

Works on any node type with ``children``/``name`` (this package's ``TreeNode``
or ``ete3.TreeNode``).
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass
from typing import Dict, List

import numpy as np


@dataclass
class SpeciesLayout:
    nodes: List          # node objects by BFS rank
    rank: Dict           # node -> BFS rank
    parent: np.ndarray
    child0: np.ndarray
    child1: np.ndarray


@dataclass
class ObjectLayout:
    nodes: List          # node objects in preorder
    index: Dict          # node -> preorder index
    left: np.ndarray
    right: np.ndarray
    leaf_species: np.ndarray


def flatten_species(tree) -> SpeciesLayout:
    nodes = []
    queue = deque((tree,))
    while queue:
        node = queue.popleft()
        nodes.append(node)
        queue.extend(node.children)
    rank = {node: i for i, node in enumerate(nodes)}
    n = len(nodes)
    parent = np.full(n, -1, np.int32)
    child0 = np.full(n, -1, np.int32)
    child1 = np.full(n, -1, np.int32)
    for i, node in enumerate(nodes):
        kids = node.children
        if kids:
            if len(kids) != 2:
                raise ValueError(f"species node {node.name!r} has {len(kids)} children; "
                                 "the species tree must be binary")
            child0[i], child1[i] = rank[kids[0]], rank[kids[1]]
            parent[child0[i]] = parent[child1[i]] = i
    return SpeciesLayout(nodes, rank, parent, child0, child1)


def flatten_objects(tree, leaf_object_species, species_rank) -> ObjectLayout:
    nodes = []
    stack = [tree]
    while stack:
        node = stack.pop()
        nodes.append(node)
        stack.extend(reversed(node.children))
    index = {node: i for i, node in enumerate(nodes)}
    n = len(nodes)
    left = np.full(n, -1, np.int32)
    right = np.full(n, -1, np.int32)
    leaf_species = np.full(n, -1, np.int32)
    for i, node in enumerate(nodes):
        kids = node.children
        if kids:
            if len(kids) != 2:
                raise ValueError(f"object node {node.name!r} has {len(kids)} children; "
                                 "resolve polytomies first")
            left[i], right[i] = index[kids[0]], index[kids[1]]
        else:
            try:
                leaf_species[i] = species_rank[leaf_object_species[node]]
            except KeyError:
                raise KeyError(f"leaf {node.name!r} has no species in leaf_object_species") from None
    return ObjectLayout(nodes, index, left, right, leaf_species)


@dataclass
class PolytomyLayout:
    """A (possibly multifurcating) object tree in the CSR form of ``sr2_resolutions_*``."""

    nodes: List          # node objects in preorder
    child_off: np.ndarray
    child_idx: np.ndarray
    leaf_species: np.ndarray


def flatten_polytomies(tree, leaf_object_species, species_rank) -> PolytomyLayout:
    nodes = []
    stack = [tree]
    while stack:
        node = stack.pop()
        nodes.append(node)
        stack.extend(reversed(node.children))
    index = {node: i for i, node in enumerate(nodes)}
    child_off = np.zeros(len(nodes) + 1, np.int32)
    child_idx = []
    leaf_species = np.full(len(nodes), -1, np.int32)
    for i, node in enumerate(nodes):
        child_idx.extend(index[kid] for kid in node.children)
        child_off[i + 1] = len(child_idx)
        if not node.children:
            leaf_species[i] = species_rank[leaf_object_species[node]]
    return PolytomyLayout(nodes, child_off, np.asarray(child_idx, np.int32), leaf_species)
