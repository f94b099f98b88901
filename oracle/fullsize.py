"""BASELINE.json configurations at FULL size for the oracle (TEST INFRASTRUCTURE ONLY).

Shared by ``oracle/make_fullsize_golden.py`` (writes ``tests/golden/fullsize_*.npz`` from the CPU
oracle), the ``-m gpu`` parity tests that compare the CUDA path with those fixtures, and
``bench.py``'s CPU-baseline legs.  The workloads are the ones of SURVEY.md section 8(d).

Tables of the big single instances (config #3: 9,999 x 3,999 cells) do not fit in a fixture, so
they are compared through one 64-bit checksum per object node: ``sum_x (a[u, x] + 2) * w[x]`` over
fixed pseudo-random odd 64-bit weights (wrapping arithmetic), separately for every table field.
A row that differs in a single cell changes its checksum.
"""
import numpy as np

from superrec2_b200 import _lib, synth
from superrec2_b200.compute.flatten import flatten_polytomies, flatten_species
from superrec2_b200.compute.reconciliation import cost_vector
from superrec2_b200.compute.unordered_super_reconciliation import family_index, leaf_bitsets
from superrec2_b200.model.reconciliation import SuperReconciliationInput


def weights(n: int, stream: int = 0) -> np.ndarray:
    rng = np.random.default_rng(977 + stream)
    return rng.integers(0, 1 << 63, n, dtype=np.uint64) * np.uint64(2) + np.uint64(1)


def row_checksums(table: np.ndarray, stream: int = 0) -> np.ndarray:
    """uint64[G] checksums of an integer array [G, ...] (everything after axis 0 is one row)."""
    flat = np.ascontiguousarray(table).reshape(table.shape[0], -1)
    w = weights(flat.shape[1], stream)
    out = np.empty(flat.shape[0], np.uint64)
    step = max(1, (1 << 24) // max(1, flat.shape[1]))
    with np.errstate(over="ignore"):
        for lo in range(0, flat.shape[0], step):
            block = (flat[lo:lo + step].astype(np.int64) + 2).astype(np.uint64)
            out[lo:lo + step] = (block * w).sum(axis=1, dtype=np.uint64)
    return out


def config2():
    """dtl, 10,000 random 500-leaf gene trees vs one 200-leaf species tree (seed 2000)."""
    return synth.make_batch(200, 500, 10000, 2000)


def config3():
    """dtl, one 5,000-leaf object tree vs a 2,000-leaf species tree (seed 3000)."""
    return synth.make_batch(2000, 5000, 1, seed=3000)


def config4():
    """superdtl, one 1,000-leaf synteny tree (32 families) vs a 500-leaf species tree (seed 4000)."""
    return synth.make_batch(500, 1000, 1, seed=4000, n_families=32)


def config5():
    """superdtl, 33-leaf multifurcating tree with 105 x 945 = 99,225 resolutions (seed 5000)."""
    return synth.make_polytomy_input()


class ResolutionSet:
    """The resolutions of one multifurcating input as flat binary instances (library numbering:
    ``sr2_resolutions_tree``, pinned to the reference's ``binarize()`` order by
    ``tests/test_host_model.py``)."""

    def __init__(self, data):
        self.srec_input = SuperReconciliationInput.from_dict(data)
        self.species = flatten_species(self.srec_input.species_lca.tree)
        self.poly = flatten_polytomies(self.srec_input.object_tree, self.srec_input.leaf_object_species,
                                       self.species.rank)
        self.families = family_index(self.srec_input.leaf_syntenies)
        self.bits = leaf_bitsets(self.poly.nodes, self.srec_input.leaf_syntenies, self.families)
        self.costs = cost_vector(self.srec_input.costs)
        self.total, self.nodes_per_tree = _lib.resolutions_count(self.poly.child_off, self.poly.child_idx)

    @property
    def species_args(self):
        return self.species.parent, self.species.child0, self.species.child1

    def batch(self, first: int, count: int):
        """(node_off, left, right, leaf_species, leaf_bits) of resolutions [first, first+count)."""
        G, W = self.nodes_per_tree, self.bits.shape[1]
        left = np.empty((count, G), np.int32)
        right = np.empty((count, G), np.int32)
        leaf_species = np.full((count, G), -1, np.int32)
        bits = np.zeros((count, G, W), np.uint32)
        for i in range(count):
            l, r, origin = _lib.resolutions_tree(self.poly.child_off, self.poly.child_idx, first + i)
            left[i], right[i] = l, r
            leaves = l < 0
            leaf_species[i, leaves] = self.poly.leaf_species[origin[leaves]]
            bits[i, leaves] = self.bits[origin[leaves]]
        node_off = np.arange(count + 1, dtype=np.int64) * G
        return node_off, left.reshape(-1), right.reshape(-1), leaf_species.reshape(-1), bits.reshape(-1, W)
