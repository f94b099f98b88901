"""``superdtl --solutions all``: every minimum-cost unordered super-reconciliation.

``usreconcile_extended_uspfs(srec_input, RetentionPolicy.ALL)`` of the reference
(``/root/reference/src/superrec2/compute/unordered_super_reconciliation.py:449-497`` with
``_decode_uspfs_table :361-446`` and the ALL-retention rules of
``utils/dynamic_programming.py:215-260``).

Split of the work, as for ``dtl`` (``reconcile_thl_all``):

* the O(|G||S|^2) part -- the (object, species, kind) VALUE table and the gain / LCA sets of every
  resolution -- comes from the GPU (values do not depend on the retention policy: they are true
  minima);
* the tag SETS (every ``ChildrenAssignment`` that reaches the minimum of a cell) are rebuilt on the
  host, only for the cells the enumeration visits, from the child rows of the device table, and
  checked against the device value of the cell;
* the product decode, the synteny labelling and ``cost()`` run on the host, because the number of
  solutions is exponential and every one of them becomes a Python object anyway.

What is and is not reproducible in the reference.  Under ALL the reference iterates
``Entry.infos()``, a Python ``set`` of tuples of ``TreeNode`` objects that hash by address, so its
iteration order changes from run to run; every visited INHERIT node unions its gain set IN PLACE
into the live LCA set of its nearest LCA-kind ancestor (``:386-391``, SURVEY.md Hazard A) and
snapshots the set at visit time.  The set of decoded (object -> species, kind) assignments and the
table are deterministic; the synteny labels -- and through ``cost()`` the set of solutions that is
kept -- can depend on that iteration order when a polluted set changes a label.  This module fixes
the order (tags sorted by (left species rank, left kind, right species rank, right kind), everything
else exactly as the reference: level-order roots, generators consumed by ``itertools.product`` left
side first, sets live across the roots of one resolution), so its answer is the reference's answer
for one admissible iteration order, and THE reference answer whenever that answer does not depend
on the order (``tests/golden/superdtl_all.json.gz`` records which fixtures are order-independent).
"""
from __future__ import annotations

from itertools import product
from typing import List, Set

import numpy as np

from .. import _lib
from ..model.reconciliation import SuperReconciliationInput, SuperReconciliationOutput
from .reconciliation import cost_vector
from .unordered_super_reconciliation import bits_to_families, family_index, flatten_super

BIG = 1 << 40          # stands for infinity.inf in 64-bit arithmetic (sums of it stay >= BIG)
LCA, INH = 0, 1


class _Resolution:
    """One binary resolution: device value table + host-side ALL-retention tag sets."""

    def __init__(self, ctx, bin_input, families, costs, allowed=None):
        species, objects, _, bits = flatten_super(bin_input, None, families)
        self.species, self.objects = species, objects
        ctx.set_species(species.parent, species.child0, species.child1)
        n = len(objects.nodes)
        ctx.set_policy(_lib.SR2_POLICY_ALL)
        try:
            ctx.superdtl_upload(np.asarray([0, n], np.int64), objects.left, objects.right, objects.leaf_species,
                                bits, costs)
            if allowed is not None:
                # cells of other species are never created (allowed_species of _compute_uspfs_table,
                # :340-356): their values stay infinite, so neither a tag nor a root can reach them
                ctx.superdtl_set_allowed(allowed(bin_input, objects, species))
            ctx.superdtl_solve()
            value, _, gain, lca = ctx.superdtl_tables(0)
        finally:
            ctx.set_policy(_lib.SR2_POLICY_ANY)
        table = value.astype(np.int64)
        table[value >= _lib.SR2_INF] = BIG
        self.T = table                                  # [node, species, kind]
        words = gain.shape[1]
        as_int = lambda row: sum(int(row[w]) << (32 * w) for w in range(words))  # noqa: E731
        self.gain = [as_int(gain[u]) for u in range(n)]
        self.clean = [as_int(lca[u]) for u in range(n)]
        self.costs = tuple(int(c) for c in costs)
        S = len(species.nodes)
        parent = species.parent
        depth = np.zeros(S, np.int64)
        for y in range(1, S):
            depth[y] = depth[parent[y]] + 1
        size = np.ones(S, np.int64)
        for y in range(S - 1, 0, -1):
            size[parent[y]] += size[y]
        tin = np.zeros(S, np.int64)
        for y in range(S):
            if species.child0[y] >= 0:
                tin[species.child0[y]] = tin[y] + 1
                tin[species.child1[y]] = tin[y] + 1 + size[species.child0[y]]
        self.depth, self.tin, self.tout = depth, tin, tin + size
        self._tags = {}

    # species sets, as boolean masks over level-order ranks
    def below(self, x):
        return (self.tin >= self.tin[x]) & (self.tin < self.tout[x])

    def separate(self, x):
        return ~self.below(x) & ~((self.tin <= self.tin[x]) & (self.tin[x] < self.tout))

    def _choices(self, child, mask, rel_depth, lca_add, inh_add):
        """Entry(MIN, ALL) over the two candidates per species of `mask`:
        (value, [(species, kind), ...]) -- reference :203-283."""
        if not mask.any():
            return BIG, []
        ranks = np.flatnonzero(mask)
        dist = (self.depth[ranks] - rel_depth) * self.costs[3] if rel_depth is not None else 0
        a = self.T[child, ranks, LCA] + dist + lca_add
        b = self.T[child, ranks, INH] + dist + inh_add
        low = min(int(a.min()), int(b.min()))
        if low >= BIG:
            return BIG, []
        picks = [(int(s), LCA) for s in ranks[a == low]] + [(int(s), INH) for s in ranks[b == low]]
        return low, picks

    def tags(self, u, x, kind):
        """Every ChildrenAssignment of table[u][x][kind] under RetentionPolicy.ALL, sorted."""
        key = (u, x, kind)
        if key in self._tags:
            return self._tags[key]
        if self.T[u, x, kind] >= BIG:      # the entry was never created (no finite candidate, or a species the
            self._tags[key] = ()           # node is not allowed at): nothing to decode
            return ()
        spe, dup, hgt, floss, sloss = self.costs
        kids = (int(self.objects.left[u]), int(self.objects.right[u]))
        dx = int(self.depth[x])
        sub, sep = self.below(x), self.separate(x)
        x0, x1 = int(self.species.child0[x]), int(self.species.child1[x])
        choice = []
        for child in kids:
            subset = (self.clean[u] & ~self.clean[child]) == 0          # lca_sets[u] <= lca_sets[child] (:182)
            ll, li = (0, BIG) if subset else (sloss, 0)
            if kind == INH:
                cons_add, seg_add, sep_add = (sloss, 0), (0, 0), (0, 0)
            else:
                cons_add, seg_add, sep_add = (ll, li), (0, li), (0, li)
            entry = {
                "conserved": self._choices(child, sub, dx, *cons_add),
                "segment": self._choices(child, sub, dx, *seg_add),
                "separate": self._choices(child, sep, None, *sep_add),
            }
            if x0 >= 0:
                # left / right: conserved restricted to one child subtree, distance from that child
                entry["left"] = self._choices(child, self.below(x0), dx + 1, *cons_add)
                entry["right"] = self._choices(child, self.below(x1), dx + 1, *cons_add)
            choice.append(entry)
        events = [("conserved", "segment", dup), ("segment", "conserved", dup),
                  ("conserved", "separate", hgt), ("separate", "conserved", hgt)]
        if x0 >= 0:
            events = [("left", "right", spe), ("right", "left", spe)] + events
        best, found = BIG, set()
        for first, second, event in events:
            (va, pa), (vb, pb) = choice[0][first], choice[1][second]
            if va >= BIG or vb >= BIG:
                continue
            total = va + vb + event
            if total < best:
                best, found = total, set()
            if total == best:
                found.update((a, ka, b, kb) for a, ka in pa for b, kb in pb)
        if best != self.T[u, x, kind]:
            raise _lib.Sr2Error(-5, f"host tag sets disagree with the device table at cell ({u}, {x}, {kind})")
        result = tuple(sorted(found)) if best < BIG else ()
        self._tags[key] = result
        return result


def usreconcile_extended_uspfs_all(srec_input: SuperReconciliationInput, device: int = 0, allowed=None
                                   ) -> Set[SuperReconciliationOutput]:
    """
    All minimum-cost unordered super-reconciliations over every binary resolution of the input,
    as ``usreconcile_extended_uspfs(srec_input, RetentionPolicy.ALL)`` of the reference returns
    them (module docstring: which part of that answer is well defined).  ``allowed`` restricts every
    object node to one species (``usreconcile_base_uspfs`` under ALL, binary inputs).
    """
    import sys

    ctx = _lib.default_context(device)
    costs = cost_vector(srec_input.costs)
    families = family_index(srec_input.leaf_syntenies)
    names = list(families)
    best_cost, best = None, []

    for index in range(srec_input.count_resolutions()):
        bin_input = srec_input.resolution(index)
        bin_input.label_internal()
        res = _Resolution(ctx, bin_input, families, costs, allowed)
        nodes, sp_nodes = res.objects.nodes, res.species.nodes
        left, right = res.objects.left, res.objects.right
        live = list(res.clean)        # lca_sets: LIVE across the roots of this resolution (:463-464)

        def decode(u, x, kind, owner):
            """_decode_uspfs_table (:361-446) on flat ids; yields lists of (node, species, synteny bits)."""
            if kind == LCA:
                owner = u
            else:
                live[owner] |= res.gain[u]          # in place, persists (:390)
            mine = (u, x, live[owner])
            if left[u] < 0:
                if res.T[u, x, kind] < BIG:
                    yield [mine]
                return
            for a, ka, b, kb in res.tags(u, x, kind):
                pairs = product(decode(int(left[u]), a, ka, owner), decode(int(right[u]), b, kb, owner))
                for part_left, part_right in pairs:
                    yield [mine] + part_left + part_right

        old_limit = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old_limit, 6 * len(nodes) + 1000))
        try:
            for x in range(len(sp_nodes)):
                for solution in decode(0, x, LCA, 0):
                    output = SuperReconciliationOutput(
                        input=bin_input,
                        object_species={nodes[u]: sp_nodes[s] for u, s, _ in solution},
                        syntenies={nodes[u]: bits_to_families(_words(bits, len(names)), names)
                                   for u, _, bits in solution},
                        ordered=False)
                    cost = output.cost()
                    if best_cost is None or cost < best_cost:
                        best_cost, best = cost, [output]
                    elif cost == best_cost:
                        best.append(output)
        finally:
            sys.setrecursionlimit(old_limit)
    return set(best)


def _words(bits: int, n_families: int) -> List[int]:
    return [(bits >> (32 * w)) & 0xFFFFFFFF for w in range((n_families + 31) // 32 or 1)]
