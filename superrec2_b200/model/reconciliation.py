"""Inputs and outputs of the (super-)reconciliation problem.

Same public surface, JSON shape and cost semantics as
``/root/reference/src/superrec2/model/reconciliation.py`` (cited per item
below), rewritten without ete3/infinity and with iterative tree walks:

* events and default costs: reference ``:28-73``;
* ``ReconciliationInput`` (``to_dict``/``from_dict``/``binarize``/
  ``label_internal``): reference ``:79-230``;
* ``ReconciliationOutput`` (``node_event``/``cost``): reference ``:233-377``;
* ``SuperReconciliationInput``/``SuperReconciliationOutput`` (unordered
  labelling cost): reference ``:380-572``.

``cost()`` is the objective the reference minimises when it selects the
returned solution and the number the CLI prints; the device kernels evaluate
the same formulas on flat arrays (see ``csrc/``), this module is the host-side
statement used by the CLI and by the tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from enum import Enum, auto
from itertools import product
from typing import Dict, Iterator, Mapping, Union

from .synteny import SyntenyMapping, parse_synteny_mapping, serialize_synteny_mapping
from .tree import Tree, TreeNode
from .tree_mapping import (
    TreeMapping,
    get_species_mapping,
    parse_tree_mapping,
    serialize_tree_mapping,
)
from ..utils.trees import (
    LowestCommonAncestor,
    count_resolutions,
    is_binary,
    unrank_resolution,
)

inf = math.inf


class NodeEvent(Enum):
    """Evolutionary events at a node of the object tree."""

    LEAF = auto()
    INVALID = auto()
    SPECIATION = auto()
    DUPLICATION = auto()
    HORIZONTAL_TRANSFER = auto()


class EdgeEvent(Enum):
    """Evolutionary events on a branch of the object tree."""

    FULL_LOSS = auto()
    SEGMENTAL_LOSS = auto()


Event = Union[NodeEvent, EdgeEvent]
CostValues = Mapping[Event, int]

#: order of the cost vector handed to the device library (``include/sr2.h``)
COST_ORDER = (
    NodeEvent.SPECIATION,
    NodeEvent.DUPLICATION,
    NodeEvent.HORIZONTAL_TRANSFER,
    EdgeEvent.FULL_LOSS,
    EdgeEvent.SEGMENTAL_LOSS,
)


def get_default_cost() -> CostValues:
    """Default event costs (reference ``model/reconciliation.py:65-73``)."""
    return {
        NodeEvent.SPECIATION: 0,
        NodeEvent.DUPLICATION: 1,
        NodeEvent.HORIZONTAL_TRANSFER: 1,
        EdgeEvent.FULL_LOSS: 1,
        EdgeEvent.SEGMENTAL_LOSS: 1,
    }


def _event_from_name(name) -> Event:
    if isinstance(name, (NodeEvent, EdgeEvent)):
        return name
    if name in NodeEvent.__members__:
        return NodeEvent[name]
    return EdgeEvent[name]


_NEWICK_OPTS = dict(format=8, format_root_node=True, features=["color"])


@dataclass(frozen=True)
class ReconciliationInput:
    """Input to the reconciliation problem."""

    object_tree: Tree
    species_lca: LowestCommonAncestor
    leaf_object_species: TreeMapping
    costs: CostValues = field(default_factory=get_default_cost)

    def to_dict(self):
        return {
            "object_tree": self.object_tree.write(**_NEWICK_OPTS),
            "species_tree": self.species_lca.tree.write(**_NEWICK_OPTS),
            "leaf_object_species": serialize_tree_mapping(self.leaf_object_species),
            "costs": {event.name: value for event, value in self.costs.items()},
        }

    def __repr__(self):
        body = ",\n".join(f"  {k}={v!r}" for k, v in self.to_dict().items())
        return f"{self.__class__.__name__}(\n{body}\n)"

    @classmethod
    def _from_dict(cls, data):
        object_tree = Tree(data["object_tree"], format=1)
        species_tree = Tree(data["species_tree"], format=1)
        if "leaf_object_species" in data:
            leaf_map = parse_tree_mapping(
                object_tree, species_tree, data["leaf_object_species"])
        else:
            leaf_map = get_species_mapping(object_tree, species_tree)
        if "costs" in data:
            costs = {_event_from_name(k): v for k, v in data["costs"].items()}
        else:
            costs = get_default_cost()
        return {
            "object_tree": object_tree,
            "species_lca": LowestCommonAncestor(species_tree),
            "leaf_object_species": leaf_map,
            "costs": costs,
        }

    @classmethod
    def from_dict(cls, data):
        return cls(**cls._from_dict(data))

    # -- polytomies -----------------------------------------------------------
    def count_resolutions(self) -> int:
        """Number of inputs that :meth:`binarize` yields."""
        return (count_resolutions(self.object_tree)
                * count_resolutions(self.species_lca.tree))

    def resolution(self, index: int) -> "ReconciliationInput":
        """
        ``index``-th input of :meth:`binarize` (object-tree resolutions vary
        slowest, reference ``model/reconciliation.py:171-174``), rebuilt through
        the serialised form exactly as the reference does (``:175-189``).
        """
        if is_binary(self.object_tree) and is_binary(self.species_lca.tree):
            if index != 0:
                raise IndexError("resolution index out of range")
            return self
        n_species = count_resolutions(self.species_lca.tree)
        object_tree = unrank_resolution(self.object_tree, index // n_species)
        species_tree = unrank_resolution(self.species_lca.tree, index % n_species)
        return self.__class__.from_dict({
            **self.to_dict(),
            "object_tree": object_tree.write(**_NEWICK_OPTS),
            "species_tree": species_tree.write(**_NEWICK_OPTS),
        })

    def binarize(self) -> Iterator["ReconciliationInput"]:
        """All inputs obtained by resolving polytomies (self if already binary)."""
        for index in range(self.count_resolutions()):
            yield self.resolution(index)

    def label_internal(self) -> None:
        """Name unnamed nodes ``O<n>`` / ``S<n>`` in preorder, skipping used names."""
        for tree, prefix in ((self.object_tree, "O"), (self.species_lca.tree, "S")):
            used = {node.name for node in tree.traverse("preorder")}
            counter = 0
            for node in tree.traverse("preorder"):
                if not node.name or node.name == "NoName":
                    while f"{prefix}{counter}" in used:
                        counter += 1
                    node.name = f"{prefix}{counter}"
                    used.add(node.name)

    def __hash__(self):
        return hash((
            id(self.object_tree),
            id(self.species_lca),
            tuple(sorted(serialize_tree_mapping(self.leaf_object_species).items())),
            tuple(sorted((e.name, v) for e, v in self.costs.items())),
        ))


@dataclass(frozen=True)
class ReconciliationOutput:
    """A reconciliation: every object node mapped to a species node."""

    input: ReconciliationInput
    object_species: TreeMapping

    def to_dict(self):
        return {
            "input": self.input.to_dict(),
            "object_species": serialize_tree_mapping(self.object_species),
        }

    def __repr__(self):
        items = self.to_dict()
        items["input"] = self.input
        body = ",\n".join(f"  {k}={v!r}" for k, v in items.items())
        return f"{self.__class__.__name__}(\n{body}\n)"

    @classmethod
    def _from_dict(cls, data):
        rec_input = ReconciliationInput.from_dict(data["input"])
        return {
            "input": rec_input,
            "object_species": parse_tree_mapping(
                rec_input.object_tree, rec_input.species_lca.tree,
                data["object_species"]),
        }

    @classmethod
    def from_dict(cls, data):
        return cls(**cls._from_dict(data))

    def node_event(self, node: TreeNode) -> NodeEvent:
        """Classify a node from its species and its children's species."""
        anc = self.input.species_lca
        rec = self.object_species
        if node.is_leaf():
            same = rec[node] is self.input.leaf_object_species[node]
            return NodeEvent.LEAF if same else NodeEvent.INVALID
        left, right = node.children
        here, sp_l, sp_r = rec[node], rec[left], rec[right]
        if anc.is_strict_ancestor_of(sp_l, here) or anc.is_strict_ancestor_of(sp_r, here):
            return NodeEvent.INVALID
        above_l = anc.is_ancestor_of(here, sp_l)
        above_r = anc.is_ancestor_of(here, sp_r)
        if above_l and above_r:
            if here is anc(sp_l, sp_r) and not anc.is_comparable(sp_l, sp_r):
                return NodeEvent.SPECIATION
            return NodeEvent.DUPLICATION
        if above_l or above_r:
            return NodeEvent.HORIZONTAL_TRANSFER
        return NodeEvent.INVALID

    def cost(self):
        """Total event cost of this reconciliation (``inf`` if invalid)."""
        anc = self.input.species_lca
        costs = self.input.costs
        rec = self.object_species
        floss = costs[EdgeEvent.FULL_LOSS]
        total = 0
        for node in self.input.object_tree.traverse("preorder"):
            event = self.node_event(node)
            if event is NodeEvent.INVALID:
                return inf
            if event is NodeEvent.LEAF:
                continue
            left, right = node.children
            dist_l = anc.distance(rec[node], rec[left])
            dist_r = anc.distance(rec[node], rec[right])
            if event is NodeEvent.SPECIATION:
                total += costs[NodeEvent.SPECIATION] + floss * (dist_l + dist_r - 2)
            elif event is NodeEvent.DUPLICATION:
                total += costs[NodeEvent.DUPLICATION] + floss * (dist_l + dist_r)
            else:
                kept = dist_l if anc.is_ancestor_of(rec[node], rec[left]) else dist_r
                total += costs[NodeEvent.HORIZONTAL_TRANSFER] + floss * kept
        return total

    def __hash__(self):
        return hash((
            self.input,
            tuple(sorted(serialize_tree_mapping(self.object_species).items())),
        ))


@dataclass(frozen=True, repr=False)
class SuperReconciliationInput(ReconciliationInput):
    """Input to the super-reconciliation problem (adds the leaf syntenies)."""

    leaf_syntenies: SyntenyMapping = field(default_factory=dict)

    def to_dict(self):
        return {
            **super().to_dict(),
            "leaf_syntenies": serialize_synteny_mapping(self.leaf_syntenies),
        }

    @classmethod
    def _from_dict(cls, data):
        base = super()._from_dict(data)
        return {
            **base,
            "leaf_syntenies": parse_synteny_mapping(
                base["object_tree"], data["leaf_syntenies"]),
        }

    def __hash__(self):
        return hash((
            super().__hash__(),
            tuple(sorted(
                (name, tuple(syn)) for name, syn in
                serialize_synteny_mapping(self.leaf_syntenies).items())),
        ))


@dataclass(frozen=True, repr=False)
class SuperReconciliationOutput(ReconciliationOutput):
    """A super-reconciliation: species mapping plus a synteny for every node."""

    input: SuperReconciliationInput
    syntenies: SyntenyMapping
    ordered: bool

    def to_dict(self):
        return {
            **super().to_dict(),
            "syntenies": serialize_synteny_mapping(self.syntenies),
            "ordered": self.ordered,
        }

    @classmethod
    def _from_dict(cls, data):
        rec_input = SuperReconciliationInput.from_dict(data["input"])
        return {
            "input": rec_input,
            "object_species": parse_tree_mapping(
                rec_input.object_tree, rec_input.species_lca.tree,
                data["object_species"]),
            "syntenies": parse_synteny_mapping(
                rec_input.object_tree, data["syntenies"]),
            "ordered": data.get("ordered", True),
        }

    def reconciliation_cost(self):
        return ReconciliationOutput.cost(self)

    def labeling_cost(self):
        """Segmental-loss cost of the labelling (unordered model only)."""
        if self.ordered:
            raise NotImplementedError(
                "ordered syntenies belong to the SPFS algorithms, which are "
                "outside the dtl/superdtl path (SURVEY.md section 2, row 6)")
        anc = self.input.species_lca
        rec = self.object_species
        sloss = self.input.costs[EdgeEvent.SEGMENTAL_LOSS]
        total = 0
        for node in self.input.object_tree.traverse("preorder"):
            if node.is_leaf():
                continue
            event = self.node_event(node)
            left, right = node.children
            here = set(self.syntenies[node])
            pay_l = 0 if here <= set(self.syntenies[left]) else sloss
            pay_r = 0 if here <= set(self.syntenies[right]) else sloss
            if event is NodeEvent.SPECIATION:
                total += pay_l + pay_r
            elif event is NodeEvent.DUPLICATION:
                total += min(pay_l, pay_r)
            elif event is NodeEvent.HORIZONTAL_TRANSFER:
                total += pay_l if anc.is_comparable(rec[node], rec[left]) else pay_r
            else:
                raise AssertionError("labelling cost of an invalid reconciliation")
        return total

    def cost(self):
        base = self.reconciliation_cost()
        if base == inf:
            return inf
        return base + self.labeling_cost()

    def __hash__(self):
        return hash((
            super().__hash__(),
            tuple(sorted(
                (name, tuple(syn)) for name, syn in
                serialize_synteny_mapping(self.syntenies).items())),
            self.ordered,
        ))
