"""Unordered synteny-labelled reconciliation on the GPU (``superdtl``).

Drop-in for ``usreconcile_extended_uspfs`` of the reference
(``/root/reference/src/superrec2/compute/unordered_super_reconciliation.py:523-542``).
Syntenies travel to the device as bitsets: bit *i* of a node's words is the
*i*-th gene family of the instance in ``sort_synteny`` order
(``model/synteny.py:21-33``), so a bitset reads back directly as the sorted list
the reference stores in its output.
"""

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


# ---------------------------------------------------------------------------
# entry point
# ---------------------------------------------------------------------------
from typing import Set  # noqa: E402

from .. import _lib  # noqa: E402
from ..model.reconciliation import (  # noqa: E402
    SuperReconciliationInput,
    SuperReconciliationOutput,
)
from ..utils.dynamic_programming import RetentionPolicy  # noqa: E402
from .flatten import flatten_objects, flatten_species  # noqa: E402
from .reconciliation import cost_vector  # noqa: E402

#: resolutions flattened and sent to the device per batch by the Python driver
RESOLUTION_BATCH = 4096


def flatten_super(srec_input: SuperReconciliationInput, species=None, families=None):
    """(species layout, object layout, family index, leaf bitsets) of one BINARY input."""
    if species is None:
        species = flatten_species(srec_input.species_lca.tree)
    objects = flatten_objects(srec_input.object_tree, srec_input.leaf_object_species, species.rank)
    if families is None:
        families = family_index(srec_input.leaf_syntenies)
    bits = leaf_bitsets(objects.nodes, srec_input.leaf_syntenies, families)
    return species, objects, families, bits


def usreconcile_extended_uspfs(
    srec_input: SuperReconciliationInput,
    policy: RetentionPolicy,
) -> Set[SuperReconciliationOutput]:
    """
    Compute a minimum-cost unordered super-reconciliation using the
    SuperDTL algorithm on an NVIDIA B200 GPU.  This is an extension of the
    Unordered Small-Phylogeny-for-Syntenies (EUSPFS) algorithm that can handle
    all kinds of events (including transfers) and non-negative integer cost
    values.  Multifurcations of the object tree are resolved by trying every
    binary resolution; the solution returned is the one the reference
    implementation returns (same tie-breaking, same synteny labelling).

    :param srec_input: objects of the super-reconciliation
    :param policy: ``RetentionPolicy.ANY`` (one solution) or ``ALL`` (every minimum-cost solution; the value
        tables come from the GPU, the enumeration runs on the host, ``compute/unordered_super_all.py``)
    :returns: any minimum-cost super-reconciliation, if there is one, or all of them
    """
    return usreconcile_extended_uspfs_sharded(srec_input, policy)


#: widths of the packed best key ``cost << 40 | resolution << 16 | root`` (one 8-byte word for the
#: cross-GPU min-reduce): the library refuses uploads whose ``cost()`` bound reaches 2^23 and
#: instance numbers from 2^24 on
KEY_COST_BITS, KEY_RESOLUTION_BITS, KEY_ROOT_BITS = 23, 24, 16


def _batches(ctx, upload, first, last):
    """Solve resolutions [first, last) in as few batches as fit in device memory; yields the
    (cost, object resolution number, root) of every batch that has a solution."""
    start, step = first, max(1, last - first)
    while start < last:
        count = min(step, last - start)
        try:
            upload(start, count)
        except _lib.Sr2Error as err:
            if err.code != -4 or count == 1:   # SR2_ERR_NOMEM: halve the batch
                raise
            step = (count + 1) // 2
            continue
        ctx.superdtl_solve()
        cost, number, root, key = ctx.superdtl_best()
        if root >= 0:
            yield cost, number, root
        start += count


def solve_resolution_range(srec_input: SuperReconciliationInput, lo: int, hi: int, device: int = 0,
                           allowed=None, prepared=None):
    """
    First strict minimum of ``cost()`` over (resolution number, root species in level order) for
    the resolutions ``[lo, hi)`` of ``srec_input`` -- the running ``Entry(MIN, ANY)`` of the
    reference's ``_uspfs`` (``:454-497``) restricted to a block of its outer loop.  Returns
    ``(cost, resolution number, root species rank)`` or ``None`` when no resolution of the block
    has a solution.  The tuple is compared lexicographically across blocks / ranks.
    """
    from ..utils.trees import count_resolutions, is_binary
    costs = cost_vector(srec_input.costs)
    ctx = _lib.default_context(device)
    families = family_index(srec_input.leaf_syntenies)
    total = srec_input.count_resolutions()
    best = None
    if hi <= lo:
        return None
    if is_binary(srec_input.species_lca.tree):
        # one species tree for every resolution: unranking happens inside the library
        species = flatten_species(srec_input.species_lca.tree)
        ctx.set_species(species.parent, species.child0, species.child1)
        if total == 1:
            _, objects, _, bits = flatten_super(srec_input, species, families)
            ctx.superdtl_upload([0, len(objects.nodes)], objects.left, objects.right,
                                objects.leaf_species, bits, costs)
            if allowed is not None:
                ctx.superdtl_set_allowed(allowed(srec_input, objects, species))
            ctx.superdtl_solve()
            cost, number, root, _ = ctx.superdtl_best()
            return (cost, number, root) if root >= 0 else None
        if allowed is not None:
            raise ValueError("a per-node species restriction needs a binary object tree")
        from .flatten import flatten_polytomies
        poly = flatten_polytomies(srec_input.object_tree, srec_input.leaf_object_species, species.rank)
        bits = leaf_bitsets(poly.nodes, srec_input.leaf_syntenies, families)
        for found in _batches(ctx, lambda start, count: ctx.resolutions_upload(
                poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, start, count), lo, hi):
            best = found if best is None else min(best, found)
        return best
    # species polytomies: resolution number = object_number * n_species + species_number
    # (reference model/reconciliation.py:171-174).  One batch of object resolutions per species
    # resolution; the library orders a batch by (cost, object_number, root), which for a fixed
    # species_number is the order of (cost, resolution number, root); resolution numbers are
    # Python integers here, so products of large polytomies cannot overflow a packed field.
    if allowed is not None:
        raise ValueError("a per-node species restriction needs binary trees")
    from ..utils.trees import unrank_resolution
    from .flatten import flatten_polytomies
    n_species = count_resolutions(srec_input.species_lca.tree)
    leaf_species_nodes = set(srec_input.leaf_object_species.values())
    for number in range(n_species):
        first = (lo - number + n_species - 1) // n_species      # object numbers [first, last)
        last = (hi - number + n_species - 1) // n_species
        if last <= first:
            continue
        species = flatten_species(unrank_resolution(srec_input.species_lca.tree, number))
        by_name = {node.name: rank for node, rank in species.rank.items()}
        species_rank = {node: by_name[node.name] for node in leaf_species_nodes}
        ctx.set_species(species.parent, species.child0, species.child1)
        poly = flatten_polytomies(srec_input.object_tree, srec_input.leaf_object_species, species_rank)
        bits = leaf_bitsets(poly.nodes, srec_input.leaf_syntenies, families)
        for cost, obj_number, root in _batches(ctx, lambda start, count: ctx.resolutions_upload(
                poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, start, count), first, last):
            found = (cost, obj_number * n_species + number, root)
            best = found if best is None else min(best, found)
    return best


def materialise_resolution(srec_input: SuperReconciliationInput, found, device: int = 0,
                           allowed=None) -> SuperReconciliationOutput:
    """
    The solution behind ``found = (cost, resolution number, root)``: that one resolution is
    rebuilt exactly as the reference does (re-parsed trees, ``O#``/``S#`` labels) and decoded in
    its own preorder numbering; the device must reproduce the cost and the root it reported.
    """
    cost, index, root_rank = found
    costs = cost_vector(srec_input.costs)
    ctx = _lib.default_context(device)
    families = family_index(srec_input.leaf_syntenies)
    names = list(families)
    winner = srec_input.resolution(index)
    winner.label_internal()
    species, objects, _, bits = flatten_super(winner, None, families)
    ctx.set_species(species.parent, species.child0, species.child1)
    ctx.superdtl_upload([0, len(objects.nodes)], objects.left, objects.right, objects.leaf_species,
                        bits, costs)
    if allowed is not None:
        ctx.superdtl_set_allowed(allowed(winner, objects, species))
    ctx.superdtl_solve()
    min_cost, root, mapping, _, syn = ctx.superdtl_results()
    if (int(min_cost[0]), int(root[0])) != (cost, root_rank):
        raise _lib.Sr2Error(-5, "winning resolution does not reproduce its key")
    object_species = {node: species.nodes[int(mapping[u])] for u, node in enumerate(objects.nodes)}
    syntenies = {node: bits_to_families(syn[u], names) for u, node in enumerate(objects.nodes)}
    return SuperReconciliationOutput(
        input=winner, object_species=object_species, syntenies=syntenies, ordered=False)


def usreconcile_extended_uspfs_sharded(
    srec_input: SuperReconciliationInput,
    policy: RetentionPolicy = RetentionPolicy.ANY,
    device: int = 0,
    shard: int = 0,
    n_shards: int = 1,
    reduce_key=None,
    allowed=None,
) -> Set[SuperReconciliationOutput]:
    """
    ``usreconcile_extended_uspfs`` restricted to one contiguous block of resolution
    numbers: shard ``i`` of ``n`` handles ``[i*R/n, (i+1)*R/n)``.  ``reduce_key(found) -> found``
    is the cross-shard lexicographic min-reduce of ``(cost, resolution, root)`` (``None`` = this
    shard has no solution; an NCCL all-reduce of the packed 8-byte key in
    ``superrec2_b200.dist``); the shard that owns the winning resolution returns the solution,
    the others return an empty set.

    ``allowed(bin_input, objects, species) -> int32[nodes]`` optionally restricts every object
    node of a BINARY input to one species rank (-1 = any): the ``allowed_species`` argument of the
    reference's ``_compute_uspfs_table`` (``:300-358``).
    """
    if policy is RetentionPolicy.ALL:
        if n_shards != 1:
            raise NotImplementedError("--solutions all enumerates on the host: one GPU, no shards")
        from .unordered_super_all import usreconcile_extended_uspfs_all
        return usreconcile_extended_uspfs_all(srec_input, device, allowed)
    if policy is not RetentionPolicy.ANY:
        raise ValueError(f"unsupported retention policy {policy!r}")
    total = srec_input.count_resolutions()
    lo, hi = total * shard // n_shards, total * (shard + 1) // n_shards
    best = solve_resolution_range(srec_input, lo, hi, device, allowed)
    winner = reduce_key(best) if reduce_key is not None else best
    if winner is None or winner != best:
        return set()
    return {materialise_resolution(srec_input, winner, device, allowed)}


def usreconcile_base_uspfs(
    srec_input: SuperReconciliationInput,
    policy: RetentionPolicy,
) -> Set[SuperReconciliationOutput]:
    """
    Compute a minimum-cost unordered super-reconciliation using the original
    Unordered Small-Phylogeny-for-Syntenies (USPFS) algorithm, on an NVIDIA
    B200 GPU. This cannot generate horizontal gene transfer events and
    assumes that the cost of duplications is equal to the cost of losses:
    every object node is mapped to the LCA of the species below it and only
    the synteny labelling is optimised.

    Mirrors ``usreconcile_base_uspfs`` of the reference
    (``compute/unordered_super_reconciliation.py:500-520``): the same table as
    ``superdtl`` with ``allowed_species`` reduced to the LCA mapping.  Like the
    reference it needs a binary object tree (``reconcile_lca`` unpacks two
    children per node).

    :param srec_input: objects of the super-reconciliation
    :param policy: ``RetentionPolicy.ANY`` (one solution) or ``ALL`` (every minimum-cost solution)
    :returns: any minimum-cost super-reconciliation, if there is one, or all of them
    """
    from ..utils.trees import is_binary
    from .reconciliation import reconcile_lca
    if not is_binary(srec_input.object_tree) or not is_binary(srec_input.species_lca.tree):
        # the reference needs both: reconcile_lca unpacks two children per object node, and its LCA
        # mapping is keyed by the nodes of the ORIGINAL trees, which binarize() only keeps when
        # there is nothing to resolve (model/reconciliation.py:167-169)
        raise ValueError("base_uspfs needs binary object and species trees")

    def lca_only(bin_input, objects, species):
        mapping = reconcile_lca(bin_input).object_species
        only = np.full(len(objects.nodes), -1, np.int32)
        for u, node in enumerate(objects.nodes):
            if node.children:
                only[u] = species.rank[mapping[node]]
        return only

    return usreconcile_extended_uspfs_sharded(srec_input, policy, allowed=lca_only)
