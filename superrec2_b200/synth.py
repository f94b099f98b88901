"""Synthetic reconciliation inputs of the shapes named in BASELINE.json.

SURVEY.md section 8(d): "Yule random-join" trees -- start from n leaves and
repeatedly join two uniformly chosen roots under a new parent -- a surjective
leaf -> species map for ``dtl`` (the reference raises ``KeyError`` unless every
leaf species hosts a gene, ``compute/reconciliation.py:234-236``), and random
leaf syntenies for ``superdtl``.  Everything is produced directly as the flat
arrays of ``include/sr2.h`` (vectorised over the batch with numpy) so that a
10,000-tree batch is generated in well under a second; ``to_newick`` turns one
instance back into the JSON form that the reference and the CLI read.

Flat numbering of one random-join tree with n leaves (G = 2n-1 nodes): node 0
is the root, internal nodes are 0..n-2 in REVERSE order of creation, leaves
are n-1..2n-2; every parent precedes its children as the ABI requires.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np


def random_join_batch(n_leaves: int, batch: int, seed: int):
    """``(left, right)`` int32 arrays of shape [batch, 2n-1] (-1 for leaves)."""
    n = n_leaves
    size = 2 * n - 1
    left = np.full((batch, size), -1, np.int32)
    right = np.full((batch, size), -1, np.int32)
    if n == 1:
        return left, right
    rng = np.random.default_rng(seed)
    # current roots of every tree of the batch, as flat node ids
    roots = np.tile(np.arange(n - 1, size, dtype=np.int32), (batch, 1))
    rows = np.arange(batch)
    for step in range(n - 1):
        alive = n - step
        new_id = n - 2 - step          # later joins get smaller ids: the last join is the root
        i = rng.integers(0, alive, batch)
        a = roots[rows, i]
        roots[rows, i] = roots[rows, alive - 1]
        j = rng.integers(0, alive - 1, batch)
        b = roots[rows, j]
        roots[rows, j] = new_id
        left[:, new_id] = a
        right[:, new_id] = b
    return left, right


def bfs_species(left: np.ndarray, right: np.ndarray):
    """Renumber one flat tree in level order -> (parent, child0, child1, old_of_new)."""
    size = len(left)
    order = [0]
    head = 0
    while head < len(order):
        node = order[head]
        head += 1
        if left[node] >= 0:
            order.append(int(left[node]))
            order.append(int(right[node]))
    new_of_old = np.empty(size, np.int32)
    new_of_old[order] = np.arange(size, dtype=np.int32)
    parent = np.full(size, -1, np.int32)
    child0 = np.full(size, -1, np.int32)
    child1 = np.full(size, -1, np.int32)
    for new, old in enumerate(order):
        if left[old] >= 0:
            child0[new] = new_of_old[left[old]]
            child1[new] = new_of_old[right[old]]
            parent[child0[new]] = parent[child1[new]] = new
    return parent, child0, child1, np.asarray(order, np.int32)


@dataclass
class SynthBatch:
    """A batch of object trees over one species tree, as ``include/sr2.h`` arrays."""

    # species tree, BFS rank indexed
    parent: np.ndarray
    child0: np.ndarray
    child1: np.ndarray
    species_names: List[str]
    # objects, concatenated instances
    node_off: np.ndarray
    left: np.ndarray
    right: np.ndarray
    leaf_species: np.ndarray
    costs: np.ndarray
    n_object_leaves: int
    leaf_family_bits: Optional[np.ndarray] = None  # [N, words] uint32, superdtl only
    n_families: int = 0

    @property
    def n_species(self) -> int:
        return len(self.parent)

    @property
    def n_instances(self) -> int:
        return len(self.node_off) - 1

    @property
    def cells(self) -> int:
        """DP cells = internal object nodes x species nodes, over the whole batch."""
        return int((self.left >= 0).sum()) * self.n_species

    def instance(self, b: int):
        lo, hi = int(self.node_off[b]), int(self.node_off[b + 1])
        return self.left[lo:hi], self.right[lo:hi], self.leaf_species[lo:hi]

    def to_dict(self, b: int) -> Dict:
        """Instance ``b`` in the reference's input JSON form (all nodes named)."""
        left, right, leaf_species = self.instance(b)
        lo = int(self.node_off[b])
        n = self.n_object_leaves
        names = [f"o{i}" for i in range(n - 1)] + [f"g{i}" for i in range(n)]
        data = {
            "object_tree": to_newick(left, right, names),
            "species_tree": to_newick(self.child0, self.child1, self.species_names),
            "leaf_object_species": {
                names[u]: self.species_names[leaf_species[u]]
                for u in range(len(left)) if left[u] < 0
            },
            "costs": dict(zip(
                ["SPECIATION", "DUPLICATION", "HORIZONTAL_TRANSFER", "FULL_LOSS", "SEGMENTAL_LOSS"],
                (int(c) for c in self.costs))),
        }
        if self.leaf_family_bits is not None:
            syn = {}
            for u in range(len(left)):
                if left[u] < 0:
                    bits = self.leaf_family_bits[lo + u]
                    syn[names[u]] = [f"f{f}" for f in range(self.n_families)
                                     if bits[f // 32] >> (f % 32) & 1]
            data["leaf_syntenies"] = syn
        return data


def to_newick(left, right, names) -> str:
    """Newick text of a flat tree rooted at node 0 (iterative)."""
    out: List[str] = []
    # states: 0 = enter, 1 = between children, 2 = leave
    stack = [(0, 0)]
    while stack:
        node, state = stack.pop()
        if left[node] < 0:
            out.append(names[node])
        elif state == 0:
            out.append("(")
            stack.append((node, 1))
            stack.append((int(left[node]), 0))
        elif state == 1:
            out.append(",")
            stack.append((node, 2))
            stack.append((int(right[node]), 0))
        else:
            out.append(")" + names[node])
    return "".join(out) + ";"


def make_batch(n_species_leaves: int, n_object_leaves: int, batch: int, seed: int,
               costs=(0, 1, 1, 1, 1), n_families: int = 0) -> SynthBatch:
    """
    ``batch`` random-join object trees with ``n_object_leaves`` leaves over one
    random-join species tree with ``n_species_leaves`` leaves.  The first
    ``n_species_leaves`` object leaves of every tree get a random permutation of
    the species leaves (surjective map), the others a uniform species leaf.
    """
    if n_object_leaves < n_species_leaves:
        raise ValueError("need at least as many object leaves as species leaves (surjective map)")
    s_left, s_right = random_join_batch(n_species_leaves, 1, seed)
    parent, child0, child1, old_of_new = bfs_species(s_left[0], s_right[0])
    n_s = n_species_leaves
    species_names = []
    for old in old_of_new:
        species_names.append(f"S{old - (n_s - 1)}" if old >= n_s - 1 else f"s{old}")
    species_leaf_ranks = np.flatnonzero(child0 < 0).astype(np.int32)

    left, right = random_join_batch(n_object_leaves, batch, seed + 1)
    size = 2 * n_object_leaves - 1
    rng = np.random.default_rng(seed + 2)
    leaf_species = np.full((batch, size), -1, np.int32)
    perm = rng.permuted(np.tile(species_leaf_ranks, (batch, 1)), axis=1)
    first = n_object_leaves - 1
    leaf_species[:, first:first + n_s] = perm
    extra = n_object_leaves - n_s
    if extra:
        leaf_species[:, first + n_s:] = species_leaf_ranks[rng.integers(0, n_s, (batch, extra))]
    node_off = np.arange(batch + 1, dtype=np.int64) * size
    result = SynthBatch(
        parent, child0, child1, species_names, node_off,
        left.reshape(-1), right.reshape(-1), leaf_species.reshape(-1),
        np.asarray(costs, np.int32), n_object_leaves)
    if n_families:
        words = (n_families + 31) // 32
        keep = rng.random((batch, size, n_families)) < 0.6
        empty = ~keep.any(axis=2)
        forced = rng.integers(0, n_families, (batch, size))
        keep[empty, forced[empty]] = True
        bits = np.zeros((batch, size, words), np.uint32)
        for f in range(n_families):
            bits[:, :, f // 32] |= keep[:, :, f].astype(np.uint32) << np.uint32(f % 32)
        bits[left >= 0] = 0
        result.leaf_family_bits = bits.reshape(-1, words)
        result.n_families = n_families
    return result


def make_polytomy_input(n_species_leaves: int = 30, arities=(5, 6), subtree_leaves: int = 3,
                        n_families: int = 32, seed: int = 5000, costs=(0, 1, 1, 1, 1)) -> Dict:
    """
    BASELINE config #5: a binary root over one node per entry of
    ``arities``; every child of those nodes is a ``subtree_leaves``-leaf random-join subtree.
    The defaults give 33 leaves and 105 x 945 = 99,225 binary resolutions.  Returned in the
    reference's input JSON form (only leaves are named, as in data/crispr-class1.in.json).
    """
    species = make_batch(n_species_leaves, n_species_leaves, 1, seed)
    rng = np.random.default_rng(seed + 1)
    n_sub = sum(arities)
    subs_left, subs_right = random_join_batch(subtree_leaves, n_sub, seed + 1)
    leaf_counter = 0
    leaf_names: List[str] = []

    def subtree_newick(k):
        nonlocal leaf_counter
        size = 2 * subtree_leaves - 1
        names = [""] * size
        for u in range(size):
            if subs_left[k][u] < 0:
                names[u] = f"g{leaf_counter}"
                leaf_names.append(names[u])
                leaf_counter += 1
        return to_newick(subs_left[k], subs_right[k], names)[:-1]

    groups, k = [], 0
    for arity in arities:
        groups.append("(" + ",".join(subtree_newick(k + i) for i in range(arity)) + ")")
        k += arity
    if len(groups) != 2:
        raise ValueError("the root is binary: give exactly two arities")
    object_tree = "(" + ",".join(groups) + ");"
    species_leaves = [species.species_names[r] for r in np.flatnonzero(species.child0 < 0)]
    perm = list(rng.permutation(len(species_leaves)))
    leaf_map = {}
    for i, name in enumerate(leaf_names):
        pick = perm[i] if i < len(perm) else int(rng.integers(0, len(species_leaves)))
        leaf_map[name] = species_leaves[pick]
    syn = {}
    for name in leaf_names:
        keep = [f"f{f}" for f in range(n_families) if rng.random() < 0.6]
        syn[name] = keep or [f"f{int(rng.integers(0, n_families))}"]
    return {
        "object_tree": object_tree,
        "species_tree": to_newick(species.child0, species.child1, species.species_names),
        "leaf_object_species": leaf_map,
        "leaf_syntenies": syn,
        "costs": dict(zip(
            ["SPECIATION", "DUPLICATION", "HORIZONTAL_TRANSFER", "FULL_LOSS", "SEGMENTAL_LOSS"],
            (int(c) for c in costs))),
    }
