"""Parity of the CUDA superdtl path (through the C ABI) with the CPU oracle and with
the fixtures produced by the unmodified reference: gain/LCA sets, the full
(object, species, kind) table with tags, the Hazard-A decode, cost() and the
returned solution -- all bit-exact."""
import numpy as np
import pytest

from oracle import binding as oracle
from superrec2_b200 import _lib, synth
from superrec2_b200.compute.unordered_super_reconciliation import usreconcile_extended_uspfs
from superrec2_b200.model.reconciliation import SuperReconciliationInput
from superrec2_b200.utils.dynamic_programming import RetentionPolicy
from tests.helpers import FlatSuper

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _lib.default_context(0)


def _run_one(ctx, flat, keep=True):
    ctx.set_species(*flat.species_args)
    node_off = np.array([0, len(flat.objects.left)], np.int64)
    out = ctx.superdtl_run(node_off, *flat.object_args, flat.leaf_bits, flat.costs,
                           flags=_lib.SR2_KEEP_TABLES if keep else 0)
    return out


def test_golden_small_tables_sets_results(ctx, golden):
    for fixture in golden("superdtl_small.json.gz") + golden("superdtl_reference.json.gz"):
        flat = FlatSuper.from_dict(fixture["input"])
        min_cost, root, mapping, kind, syn = _run_one(ctx, flat)
        value, tag, gain, lca = ctx.superdtl_tables(0)
        assert flat.sets_by_name(gain) == {k: sorted(v) for k, v in fixture["gain_sets"].items()}
        assert flat.sets_by_name(lca) == {k: sorted(v) for k, v in fixture["lca_sets"].items()}
        assert flat.table_by_name(value, tag) == fixture["table"]
        assert min_cost[0] == fixture["min_cost"]
        assert flat.result_dict(mapping, syn) == fixture["result"]


def test_entry_point_polytomies_and_crispr(golden):
    """usreconcile_extended_uspfs(...).to_dict() equals the reference's output JSON, including
    the winning resolution's newick with O#/S# labels (BASELINE config #1 = crispr)."""
    fixtures = golden("superdtl_poly.json.gz") + golden("crispr.json.gz") + golden("superdtl_small.json.gz")[:30]
    for fixture in fixtures:
        srec_input = SuperReconciliationInput.from_dict(fixture["input"])
        (result,) = usreconcile_extended_uspfs(srec_input, RetentionPolicy.ANY)
        got = result.to_dict()
        assert got == fixture["result"]
        assert list(got["object_species"]) == list(fixture["result"]["object_species"])
        assert list(got["syntenies"]) == list(fixture["result"]["syntenies"])
        assert result.cost() == fixture["min_cost"]


@pytest.mark.parametrize("n_s,n_o,fams,costs", [
    (5, 8, 4, (0, 1, 1, 1, 1)),
    (12, 20, 9, (0, 1, 1, 1, 1)),
    (12, 20, 33, (1, 2, 1, 1, 2)),
    (30, 40, 32, (0, 0, 0, 0, 0)),
    (40, 64, 32, (0, 1, 2, 1, 3)),
])
def test_random_batches_against_oracle(ctx, n_s, n_o, fams, costs):
    batch = synth.make_batch(n_s, n_o, 5, seed=400 + n_s, costs=costs, n_families=fams)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    min_cost, root, mapping, kind, syn = ctx.superdtl_run(
        batch.node_off, batch.left, batch.right, batch.leaf_species, batch.leaf_family_bits,
        batch.costs, flags=_lib.SR2_KEEP_TABLES)
    for b in range(batch.n_instances):
        lo, hi = int(batch.node_off[b]), int(batch.node_off[b + 1])
        left, right, leaf = batch.instance(b)
        ref = oracle.superdtl(batch.parent, batch.child0, batch.child1, left, right, leaf,
                              batch.leaf_family_bits[lo:hi], batch.costs)
        assert ref["rc"] == 0
        value, tag, gain, lca = ctx.superdtl_tables(b)
        assert (gain == ref["gain"]).all() and (lca == ref["lca"]).all()
        assert (value == ref["value"]).all()
        assert (tag == ref["tag"]).all()
        assert min_cost[b] == ref["min_cost"] and root[b] == ref["root_species"]
        assert (mapping[lo:hi] == ref["mapping"]).all()
        assert (kind[lo:hi] == ref["kind"]).all()
        assert (syn[lo:hi] == ref["syn"]).all()


def test_cta_per_row_path(ctx):
    """|S| > 256 uses one CTA per row (config #4 shape, reduced): checked against the oracle."""
    batch = synth.make_batch(140, 150, 1, seed=4000, n_families=32)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    min_cost, root, mapping, kind, syn = ctx.superdtl_run(
        batch.node_off, batch.left, batch.right, batch.leaf_species, batch.leaf_family_bits,
        batch.costs, flags=_lib.SR2_KEEP_TABLES)
    ref = oracle.superdtl(batch.parent, batch.child0, batch.child1, batch.left, batch.right,
                          batch.leaf_species, batch.leaf_family_bits, batch.costs)
    value, tag, gain, lca = ctx.superdtl_tables(0)
    assert (value == ref["value"]).all() and (tag == ref["tag"]).all()
    assert min_cost[0] == ref["min_cost"] and root[0] == ref["root_species"]
    assert (mapping == ref["mapping"]).all() and (syn == ref["syn"]).all()


def test_best_key_orders_by_cost_then_instance(ctx):
    batch = synth.make_batch(6, 9, 40, seed=9, n_families=5)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    ctx.superdtl_upload(batch.node_off, batch.left, batch.right, batch.leaf_species,
                        batch.leaf_family_bits, batch.costs, index_base=1000)
    ctx.superdtl_solve()
    min_cost, root, *_ = ctx.superdtl_results()
    cost, index, root_species, key = ctx.superdtl_best()
    first = int(np.flatnonzero(min_cost == min_cost.min())[0])
    assert (cost, index, root_species) == (int(min_cost.min()), 1000 + first, int(root[first]))
    assert key == (cost << 40) | (index << 16) | root_species
