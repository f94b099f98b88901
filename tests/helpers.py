"""Shared helpers of the test-suite: fixtures -> flat arrays (through the product's
own host model) and name-keyed views of flat results."""
import numpy as np

from superrec2_b200.compute.flatten import flatten_objects, flatten_species
from superrec2_b200.compute.reconciliation import cost_vector
from superrec2_b200.model.reconciliation import ReconciliationInput, SuperReconciliationInput

INF = 0x3FFFFFFF


class FlatDtl:
    def __init__(self, data):
        self.rec_input = ReconciliationInput.from_dict(data)
        self.species = flatten_species(self.rec_input.species_lca.tree)
        self.objects = flatten_objects(
            self.rec_input.object_tree, self.rec_input.leaf_object_species, self.species.rank)
        self.costs = cost_vector(self.rec_input.costs)

    @property
    def species_args(self):
        return self.species.parent, self.species.child0, self.species.child1

    @property
    def object_args(self):
        return self.objects.left, self.objects.right, self.objects.leaf_species

    def table_by_name(self, value, tag):
        """{object name: {species name: [value, [left sp, right sp] | None]}} (finite or tagged only)"""
        out = {}
        for u, node in enumerate(self.objects.nodes):
            row = {}
            for x, sp in enumerate(self.species.nodes):
                v = int(value[u, x])
                has_tag = tag[u, x, 0] >= 0
                if v < INF or has_tag:
                    names = ([self.species.nodes[tag[u, x, 0]].name, self.species.nodes[tag[u, x, 1]].name]
                             if has_tag else None)
                    row[sp.name] = [None if v >= INF else v, names]
            out[node.name] = row
        return out

    def mapping_by_name(self, mapping):
        return {node.name: self.species.nodes[int(mapping[u])].name
                for u, node in enumerate(self.objects.nodes)}


def flat_cost(batch, b, mapping):
    """cost() of a mapping given as species ranks, on flat arrays -- the host-side
    formulas of ReconciliationOutput.cost() (model/reconciliation.py of this package),
    vectorised with numpy so that 10^4-node instances are checked in milliseconds."""
    lo, hi = int(batch.node_off[b]), int(batch.node_off[b + 1])
    left, right, leaf_species = batch.left[lo:hi], batch.right[lo:hi], batch.leaf_species[lo:hi]
    m = np.asarray(mapping[lo:hi], np.int64)
    S = len(batch.parent)
    depth = np.zeros(S, np.int64)
    for y in range(1, S):
        depth[y] = depth[batch.parent[y]] + 1
    size = np.ones(S, np.int64)
    for y in range(S - 1, 0, -1):
        size[batch.parent[y]] += size[y]
    tin = np.zeros(S, np.int64)
    for y in range(S):
        if batch.child0[y] >= 0:
            tin[batch.child0[y]] = tin[y] + 1
            tin[batch.child1[y]] = tin[y] + 1 + size[batch.child0[y]]
    tout = tin + size
    spe, dup, hgt, floss = (int(c) for c in batch.costs[:4])
    internal = np.flatnonzero(left >= 0)
    if ((m[left < 0]) != leaf_species[left < 0]).any():
        return INF
    p, a, b_ = m[internal], m[left[internal]], m[right[internal]]

    def anc(x, y):
        return (tin[x] <= tin[y]) & (tin[y] < tout[x])

    if ((anc(a, p) & (a != p)) | (anc(b_, p) & (b_ != p))).any():
        return INF
    pa, pb = anc(p, a), anc(p, b_)
    if (~pa & ~pb).any():
        return INF
    both = pa & pb
    c0 = batch.child0[p]
    safe = np.where(c0 >= 0, c0, 0)
    below0_a = anc(safe, a) & (c0 >= 0)
    below0_b = anc(safe, b_) & (c0 >= 0)
    spec = both & (a != p) & (b_ != p) & (below0_a != below0_b)
    da, db = depth[a] - depth[p], depth[b_] - depth[p]
    total = np.where(spec, spe + floss * (da + db - 2),
                     np.where(both, dup + floss * (da + db),
                              np.where(pa, hgt + floss * da, hgt + floss * db)))
    return int(total.sum())


KINDS = ("LCA", "INHERIT")


class FlatSuper:
    """One BINARY super-reconciliation input (already labelled) as flat arrays."""

    def __init__(self, srec_input):
        from superrec2_b200.compute.unordered_super_reconciliation import family_index, leaf_bitsets
        self.srec_input = srec_input
        self.species = flatten_species(srec_input.species_lca.tree)
        self.objects = flatten_objects(
            srec_input.object_tree, srec_input.leaf_object_species, self.species.rank)
        self.costs = cost_vector(srec_input.costs)
        self.families = family_index(srec_input.leaf_syntenies)
        self.leaf_bits = leaf_bitsets(self.objects.nodes, srec_input.leaf_syntenies, self.families)

    @classmethod
    def from_dict(cls, data):
        srec_input = SuperReconciliationInput.from_dict(data)
        (binary,) = list(srec_input.binarize())
        binary.label_internal()
        return cls(binary)

    @property
    def species_args(self):
        return self.species.parent, self.species.child0, self.species.child1

    @property
    def object_args(self):
        return self.objects.left, self.objects.right, self.objects.leaf_species

    def bits_to_names(self, bits):
        names = list(self.families)
        return [names[f] for f in range(len(names)) if bits[f // 32] >> (f % 32) & 1]

    def sets_by_name(self, bitsets):
        return {node.name: sorted(self.bits_to_names(bitsets[u]))
                for u, node in enumerate(self.objects.nodes)}

    def table_by_name(self, value, tag):
        out = {}
        sp = self.species.nodes
        for u, node in enumerate(self.objects.nodes):
            row = {}
            for x in range(len(sp)):
                cell = {}
                for k, kname in enumerate(KINDS):
                    v = int(value[u, x, k])
                    t = tag[u, x, k]
                    has_tag = t[0] >= 0
                    if v < INF or has_tag:
                        names = ([[sp[t[0]].name, KINDS[t[1]]], [sp[t[2]].name, KINDS[t[3]]]]
                                 if has_tag else None)
                        cell[kname] = [None if v >= INF else v, names]
                if cell:
                    row[sp[x].name] = cell
            out[node.name] = row
        return out

    def result_dict(self, mapping, syn):
        """The reference's output JSON for this (labelled) input."""
        from superrec2_b200.model.reconciliation import SuperReconciliationOutput
        nodes = self.objects.nodes
        output = SuperReconciliationOutput(
            input=self.srec_input,
            object_species={node: self.species.nodes[int(mapping[u])] for u, node in enumerate(nodes)},
            syntenies={node: self.bits_to_names(syn[u]) for u, node in enumerate(nodes)},
            ordered=False)
        return output.to_dict()
