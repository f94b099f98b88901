"""Oracle-backed parity at the FULL sizes of BASELINE.json configs #2-#5.

The CPU oracle needs minutes for these workloads, so its answers are committed fixtures
(``tests/golden/fullsize_config*.npz``, written by ``oracle/make_fullsize_golden.py``): results of
every instance and, for the two big single instances, one checksum per table row (value and tag
tables separately, ``oracle/fullsize.py``).  Here the CUDA path runs the same seeded inputs through
the C ABI and must reproduce every number."""
import os

import numpy as np
import pytest

from oracle import fullsize
from superrec2_b200 import _lib

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture(number):
    return np.load(os.path.join(GOLDEN, f"fullsize_config{number}.npz"))


@pytest.fixture(scope="module")
def ctx():
    return _lib.default_context(0)


def test_config2_every_tree_against_the_oracle(ctx):
    """10,000 x 500-leaf trees vs 200 species leaves: min cost, root species and the whole mapping
    (as a per-tree checksum) of EVERY tree, through sr2_dtl_run and through upload/solve/results."""
    want = _fixture(2)
    batch = fullsize.config2()
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
    size = int(batch.node_off[1])
    for how in ("run", "stepwise"):
        if how == "run":
            min_cost, root, mapping = ctx.dtl_run(*args)
        else:
            ctx.dtl_upload(*args)
            ctx.dtl_solve()
            min_cost, root, mapping = ctx.dtl_results()
        assert np.array_equal(min_cost, want["min_cost"]), how
        assert np.array_equal(root, want["root"].astype(np.int32)), how
        assert np.array_equal(fullsize.row_checksums(mapping.reshape(-1, size)), want["map_hash"]), how
        assert int(min_cost.astype(np.int64).sum()) == int(want["checksum"])


def test_config3_full_table_against_the_oracle(ctx):
    """5,000-leaf object tree x 2,000-leaf species tree: all 9,999 x 3,999 values and tags (row
    checksums), the minimum cost, the root and the mapping; then the same call without kept tables
    (the product configuration) must return the same solution."""
    want = _fixture(3)
    batch = fullsize.config3()
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
    min_cost, root, mapping = ctx.dtl_run(*args, flags=_lib.SR2_KEEP_TABLES)
    value, tag = ctx.dtl_tables(0, len(batch.left))
    assert int((value < _lib.SR2_INF).sum()) == int(want["finite_cells"])
    assert np.array_equal(fullsize.row_checksums(value, 1), want["value_rows"])
    assert np.array_equal(fullsize.row_checksums(tag, 2), want["tag_rows"])
    for got in ((min_cost, root, mapping), ctx.dtl_run(*args)):
        assert int(got[0][0]) == int(want["min_cost"])
        assert int(got[1][0]) == int(want["root"])
        assert np.array_equal(got[2], want["mapping"].astype(np.int32))


def test_config4_full_table_and_labelling_against_the_oracle(ctx):
    """superdtl, 1,000-leaf synteny tree (32 families) x 500-leaf species tree: gain / LCA sets, all
    1,999 x 999 x 2 values and tags (row checksums), min cost, root, mapping, kinds and syntenies."""
    want = _fixture(4)
    batch = fullsize.config4()
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.leaf_family_bits, batch.costs)
    min_cost, root, mapping, kind, syn = ctx.superdtl_run(*args, flags=_lib.SR2_KEEP_TABLES)
    value, tag, gain, lca = ctx.superdtl_tables(0)
    assert np.array_equal(gain, want["gain"]) and np.array_equal(lca, want["lca"])
    assert int((value < _lib.SR2_INF).sum()) == int(want["finite_cells"])
    assert np.array_equal(fullsize.row_checksums(value, 1), want["value_rows"])
    assert np.array_equal(fullsize.row_checksums(tag, 2), want["tag_rows"])
    for got in ((min_cost, root, mapping, kind, syn), ctx.superdtl_run(*args)):
        assert int(got[0][0]) == int(want["min_cost"])
        assert int(got[1][0]) == int(want["root"])
        assert np.array_equal(got[2], want["mapping"].astype(np.int32))
        assert np.array_equal(got[3], want["kind"].astype(np.int32))
        assert np.array_equal(got[4], want["syn"])


@pytest.mark.parametrize("unshared", [False, True])
def test_config5_every_resolution_against_the_oracle(ctx, monkeypatch, unshared):
    """All 99,225 resolutions of the 33-leaf tree: the decoded minimum cost and the chosen root
    species of EVERY resolution (Hazard-A labelling included), and the packed best key."""
    want = _fixture(5)
    if unshared:
        monkeypatch.setenv("SR2_RESOLUTIONS_UNSHARED", "1")
    res = fullsize.ResolutionSet(fullsize.config5())
    assert res.total == 99225 == len(want["min_cost"])
    ctx.set_species(*res.species_args)
    step = 33075 if unshared else res.total
    for first in range(0, res.total, step):
        ctx.resolutions_upload(res.poly.child_off, res.poly.child_idx, res.poly.leaf_species, res.bits,
                               res.costs, first, step)
        ctx.superdtl_solve()
        min_cost, root = ctx.superdtl_results(costs_only=True)
        assert np.array_equal(min_cost, want["min_cost"][first:first + step].astype(np.int32))
        assert np.array_equal(root, want["root"][first:first + step].astype(np.int32))
        cost, inst, best_root, _ = ctx.superdtl_best()
        order = np.lexsort((root, np.arange(step), min_cost))
        assert (cost, inst, best_root) == (int(min_cost[order[0]]), first + int(order[0]), int(root[order[0]]))
    if not unshared:
        assert [cost, inst, best_root] == want["best"].tolist()
