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
    :param policy: ``RetentionPolicy.ANY`` (one solution)
    :returns: any minimum-cost reconciliation
    """
    outputs = reconcile_thl_batch([rec_input], policy)
    return outputs[0]


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
    if policy is not RetentionPolicy.ANY:
        raise NotImplementedError(
            "only RetentionPolicy.ANY is implemented on the GPU path so far "
            "(SURVEY.md section 8f item 1)")
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
