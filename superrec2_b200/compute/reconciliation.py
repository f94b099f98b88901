"""Reconciliation with arbitrary event costs on the GPU (``dtl`` / ``thl``).

Drop-in for ``reconcile_thl`` of the reference
(``/root/reference/src/superrec2/compute/reconciliation.py:266-294``): same
signature, same return type, same solution.  The table fill, the per-root
selection by ``cost()`` and the traceback run in ``libsr2.so``
(``csrc/sr2_dtl.cu``); this module only flattens the trees and rebuilds the
``ReconciliationOutput``.
"""

from typing import List, Sequence, Set

import numpy as np

from .. import _lib
from ..model.reconciliation import COST_ORDER, ReconciliationInput, ReconciliationOutput
from ..utils.dynamic_programming import RetentionPolicy
from .flatten import flatten_objects, flatten_species


def cost_vector(costs) -> np.ndarray:
    """``[spe, dup, hgt, floss, sloss]`` as int32; rejects what the DP cannot hold."""
    values = []
    for event in COST_ORDER:
        value = costs[event]
        if isinstance(value, bool) or int(value) != value:
            raise ValueError(f"cost of {event.name} must be an integer, got {value!r}")
        values.append(int(value))
    return np.asarray(values, np.int32)


def reconcile_lca(rec_input: ReconciliationInput) -> ReconciliationOutput:
    """
    Compute the minimum-cost reconciliation between two trees by mapping each
    node of the object tree to the lowest common ancestor (LCA) of all species
    that correspond to leaves of that node's subtree. This yields the only
    optimal solution, provided only speciations, duplications, and losses are
    allowed and duplications cost the same as losses.

    Mirrors ``reconcile_lca`` of the reference
    (``/root/reference/src/superrec2/compute/reconciliation.py:23-44``).  It is a
    single bottom-up pass with no dynamic program behind it, so it stays on the
    host; ``usreconcile_base_uspfs`` uses it to restrict the GPU table.

    :param rec_input: input for the reconciliation
    :returns: reconciliation result
    """
    rec = {}
    stack = [(rec_input.object_tree, False)]
    while stack:  # postorder, iteratively (deep trees)
        node, done = stack.pop()
        if not node.children:
            rec[node] = rec_input.leaf_object_species[node]
        elif done:
            left, right = node.children
            rec[node] = rec_input.species_lca(rec[left], rec[right])
        else:
            stack.append((node, True))
            stack.extend((kid, False) for kid in reversed(node.children))
    # the reference builds the dict in postorder; keep that key order
    return ReconciliationOutput(rec_input, rec)


def reconcile_thl(
    rec_input: ReconciliationInput,
    policy: RetentionPolicy,
) -> Set[ReconciliationOutput]:
    """
    Compute a minimum-cost reconciliation between two trees, using the
    Tofigh-Hallett-Lagergren dynamic program with duplications, losses and
    horizontal transfers, on an NVIDIA B200 GPU.  Any non-negative integer
    cost values are supported.  The solution returned is the one the
    reference implementation returns (same tie-breaking).  Solutions are not
    guaranteed to be time-consistent.

    :param rec_input: reconciliation input
    :param policy: ``RetentionPolicy.ANY`` (one solution) or ``ALL`` (every solution the
        reference enumerates; the table comes from the GPU, the enumeration runs on the host)
    :returns: any minimum-cost reconciliation, or all of them
    """
    if policy is RetentionPolicy.ALL:
        return reconcile_thl_all(rec_input)
    outputs = reconcile_thl_batch([rec_input], policy)
    return outputs[0]


def reconcile_thl_all(rec_input: ReconciliationInput, device: int = 0) -> Set[ReconciliationOutput]:
    """
    ``reconcile_thl(rec_input, RetentionPolicy.ALL)`` of the reference
    (``compute/reconciliation.py:266-294`` with ``_decode_thl_table :228-263``): every
    reconciliation reachable through ALL minimum-achieving tags, filtered to those of minimum
    ``cost()``.

    The O(|G||S|^2) part -- the value table -- is computed on the GPU (the values do not depend
    on the retention policy: with non-negative loss cost the first candidate in level order is
    also the cheapest).  The tag SETS are rebuilt on the host only for
    the cells the enumeration actually visits, with the reference's rules
    (``utils/dynamic_programming.py:215-260``: an equal value adds a tag; minima over a child
    row keep every species that reaches the minimum), and checked against the GPU value.
    The number of solutions can be exponential, exactly as in the reference.
    """
    species = flatten_species(rec_input.species_lca.tree)
    layout = flatten_objects(rec_input.object_tree, rec_input.leaf_object_species, species.rank)
    costs = cost_vector(rec_input.costs)
    n, S = len(layout.nodes), len(species.nodes)
    ctx = _lib.default_context(device)
    ctx.set_species(species.parent, species.child0, species.child1)
    ctx.set_policy(_lib.SR2_POLICY_ALL)
    try:
        ctx.dtl_run(np.asarray([0, n], np.int64), layout.left, layout.right, layout.leaf_species, costs)
        value, _ = ctx.dtl_tables(0, n)
    finally:
        ctx.set_policy(_lib.SR2_POLICY_ANY)
    INF = _lib.SR2_INF
    T = value.astype(np.int64)

    # species geometry by BFS rank
    depth = np.zeros(S, np.int64)
    for y in range(1, S):
        depth[y] = depth[species.parent[y]] + 1
    size = np.ones(S, np.int64)
    for y in range(S - 1, 0, -1):
        size[species.parent[y]] += size[y]
    tin = np.zeros(S, np.int64)
    for y in range(S):
        if species.child0[y] >= 0:
            tin[species.child0[y]] = tin[y] + 1
            tin[species.child1[y]] = tin[y] + 1 + size[species.child0[y]]
    tout = tin + size
    ranks = np.arange(S)
    _, dup, hgt, loss, _ = (int(c) for c in costs)

    def below(x):            # species in the subtree of x (x included), level order
        return ranks[(tin >= tin[x]) & (tin < tout[x])]

    def separate(x):         # species neither below nor above x
        return ranks[~((tin >= tin[x]) & (tin < tout[x])) & ~((tin <= tin[x]) & (tin[x] < tout))]

    def minima(row, subset):
        """(minimum, every species of `subset` reaching it) -- Entry.update under ALL."""
        if len(subset) == 0:
            return INF, subset
        vals = row[subset]
        low = vals.min()
        return int(low), subset[vals == low]

    tag_cache = {}

    def tags(u, x):
        """The MappingInfo set of table[u][x] under RetentionPolicy.ALL, as (left, right) ranks."""
        key = (u, x)
        if key in tag_cache:
            return tag_cache[key]
        l, r = layout.left[u], layout.right[u]
        found = {}

        def offer(left_min, right_min, extra):
            (vl, sl), (vr, sr) = left_min, right_min
            if vl >= INF or vr >= INF:
                return
            for a in sl:
                for b in sr:
                    found.setdefault(vl + vr + extra(int(a), int(b)), set()).add((int(a), int(b)))

        dx = int(depth[x])
        if species.child0[x] >= 0:
            x0, x1 = int(species.child0[x]), int(species.child1[x])
            spe = lambda a, b: loss * (int(depth[a]) + int(depth[b]) - 2 * dx - 2)  # noqa: E731
            offer(minima(T[l], below(x0)), minima(T[r], below(x1)), spe)
            offer(minima(T[l], below(x1)), minima(T[r], below(x0)), spe)
        sub, sep = below(x), separate(x)
        offer(minima(T[l], sub), minima(T[r], sub),
              lambda a, b: dup + loss * (int(depth[a]) + int(depth[b]) - 2 * dx))
        offer(minima(T[l], sep), minima(T[r], sub), lambda a, b: hgt + loss * (int(depth[b]) - dx))
        offer(minima(T[l], sub), minima(T[r], sep), lambda a, b: hgt + loss * (int(depth[a]) - dx))
        result = ()
        if found:
            best = min(found)
            if best != T[u][x]:
                raise _lib.Sr2Error(-5, f"host tag sets disagree with the device table at cell ({u}, {x})")
            result = tuple(sorted(found[best]))
        tag_cache[key] = result
        return result

    import sys
    from itertools import product

    def decode(u, x):
        """_decode_thl_table: dicts in the reference's key order (node, left subtree, right subtree)."""
        node = layout.nodes[u]
        options = tags(u, x) if layout.left[u] >= 0 else ()
        if not options:
            yield {node: species.nodes[x]}
            return
        for a, b in options:
            for left_map, right_map in product(decode(int(layout.left[u]), a), decode(int(layout.right[u]), b)):
                yield {node: species.nodes[x], **left_map, **right_map}

    best_cost, best = None, set()
    old_limit = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old_limit, 4 * n + 1000))
    try:
        for x in range(S):
            if layout.left[0] >= 0 and T[0][x] >= INF:
                continue  # no entry: the reference raises KeyError here (SURVEY.md Hazard B); skipped
            for mapping in decode(0, x):
                output = ReconciliationOutput(rec_input, mapping)
                cost = output.cost()
                if best_cost is None or cost < best_cost:
                    best_cost, best = cost, {output}
                elif cost == best_cost:
                    best.add(output)
    finally:
        sys.setrecursionlimit(old_limit)
    return best


def reconcile_thl_batch(
    rec_inputs: Sequence[ReconciliationInput],
    policy: RetentionPolicy = RetentionPolicy.ANY,
    device: int = 0,
) -> List[Set[ReconciliationOutput]]:
    """
    Reconcile many object trees against ONE species tree in a single batch
    (all inputs must share ``species_lca`` and ``costs``).  Returns one result
    set per input, each identical to ``reconcile_thl(input, policy)``.
    """
    if policy is RetentionPolicy.ALL:
        return [reconcile_thl_all(item, device) for item in rec_inputs]
    if policy is not RetentionPolicy.ANY:
        raise ValueError(f"unsupported retention policy {policy!r}")
    if not rec_inputs:
        return []
    first = rec_inputs[0]
    for other in rec_inputs[1:]:
        if other.species_lca is not first.species_lca or other.costs != first.costs:
            raise ValueError("a batch must share one species tree and one cost vector")
    species = flatten_species(first.species_lca.tree)
    layouts = [flatten_objects(item.object_tree, item.leaf_object_species, species.rank)
               for item in rec_inputs]
    node_off = np.zeros(len(layouts) + 1, np.int64)
    np.cumsum([len(layout.nodes) for layout in layouts], out=node_off[1:])
    ctx = _lib.default_context(device)
    ctx.set_species(species.parent, species.child0, species.child1)
    min_cost, root, mapping = ctx.dtl_run(
        node_off,
        np.concatenate([layout.left for layout in layouts]),
        np.concatenate([layout.right for layout in layouts]),
        np.concatenate([layout.leaf_species for layout in layouts]),
        cost_vector(first.costs),
    )
    results: List[Set[ReconciliationOutput]] = []
    for i, (item, layout) in enumerate(zip(rec_inputs, layouts)):
        if root[i] < 0:
            results.append(set())
            continue
        ranks = mapping[node_off[i]:node_off[i + 1]]
        object_species = {node: species.nodes[ranks[k]] for k, node in enumerate(layout.nodes)}
        results.append({ReconciliationOutput(item, object_species)})
    return results


def reconcile_thl_flat(species_parent, species_child0, species_child1, node_off, left, right, leaf_species,
                       costs=(0, 1, 1, 1, 1), device: int = 0, out=None):
    """
    ``reconcile_thl`` for a batch that is ALREADY in the flat form of ``include/sr2.h`` -- the entry point for
    callers that keep their trees as arrays (a 10,000-tree batch is reconciled in 8 ms this way, while building
    and flattening 10^7 Python tree nodes for :func:`reconcile_thl_batch` takes seconds).

    Species nodes are numbered in level order (root 0, the two children of a node consecutive), object nodes of
    every instance parents-first with the root at 0 and instance-local child indices; ``node_off[b]`` is the
    first node of instance ``b``.  Returns ``(min_cost[B], root_species[B], map_species[N])``; an instance
    without a solution has root -1.  Trees and species tree of at most 32,767 nodes take the 16-bit form of
    the call (``sr2_dtl_run16``: half the bytes over PCIe) and return int16 species arrays.
    """
    node_off = np.ascontiguousarray(node_off, dtype=np.int64)
    ctx = _lib.default_context(device)
    ctx.set_species(species_parent, species_child0, species_child1)
    sizes = np.diff(node_off)
    narrow = len(species_parent) <= 32767 and (len(sizes) == 0 or int(sizes.max()) <= 32767)
    costs = np.asarray(costs, np.int32)
    if narrow:
        return ctx.dtl_run16(node_off, left, right, leaf_species, costs, out=out)
    return ctx.dtl_run(node_off, left, right, leaf_species, costs, out=out)
