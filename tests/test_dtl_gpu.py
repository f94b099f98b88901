"""Parity of the CUDA dtl path (through the C ABI) with the CPU oracle and with
the fixtures produced by the unmodified reference.  Bit-exact: integer costs,
same tags, same returned mapping."""
import numpy as np
import pytest

from oracle import binding as oracle
from superrec2_b200 import _lib, synth
from superrec2_b200.compute.reconciliation import reconcile_thl, reconcile_thl_batch
from superrec2_b200.model.reconciliation import ReconciliationInput
from superrec2_b200.utils.dynamic_programming import RetentionPolicy
from tests.helpers import FlatDtl

pytestmark = pytest.mark.gpu
INF = _lib.SR2_INF


@pytest.fixture(scope="module")
def ctx():
    return _lib.default_context(0)


def _gpu_tables(ctx, flat):
    ctx.set_species(*flat.species_args)
    node_off = np.array([0, len(flat.objects.left)], np.int64)
    min_cost, root, mapping = ctx.dtl_run(node_off, *flat.object_args, flat.costs,
                                          flags=_lib.SR2_KEEP_TABLES)
    value, tag = ctx.dtl_tables(0, len(flat.objects.left))
    return int(min_cost[0]), int(root[0]), mapping, value, tag


def test_golden_small_tables(ctx, golden):
    """Cell by cell (value and tag) against the reference's own tables."""
    for fixture in (golden("thl_small.json.gz") + golden("thl_reference_tests.json.gz") +
                    golden("thl_internal_species.json.gz")):
        flat = FlatDtl(fixture["input"])
        min_cost, root, mapping, value, tag = _gpu_tables(ctx, flat)
        assert flat.table_by_name(value, tag) == fixture["table"]
        if fixture.get("error") == "KeyError":
            # the reference crashes here (a root without an entry); the library skips such
            # roots -- same answer as the oracle's non-strict mode
            ref = oracle.dtl(*flat.species_args, *flat.object_args, flat.costs, strict=False)
            assert (min_cost, root) == (ref["min_cost"], ref["root_species"])
            assert (mapping == ref["mapping"]).all()
            continue
        assert min_cost == fixture["min_cost"]
        assert flat.mapping_by_name(mapping) == fixture["result"]["object_species"]


def test_golden_results_through_entry_point(golden):
    """reconcile_thl(...).to_dict() equals the reference's output JSON."""
    for fixture in golden("thl_medium.json.gz") + golden("thl_small.json.gz")[:40]:
        if fixture.get("error"):
            continue
        rec_input = ReconciliationInput.from_dict(fixture["input"])
        (result,) = reconcile_thl(rec_input, RetentionPolicy.ANY)
        assert result.to_dict() == fixture["result"]
        assert list(result.to_dict()["object_species"]) == list(fixture["result"]["object_species"])
        assert result.cost() == fixture["min_cost"]


@pytest.mark.parametrize("n_s,n_o,costs", [
    (8, 12, (0, 1, 1, 1, 1)),
    (20, 33, (0, 1, 1, 1, 1)),
    (20, 33, (0, 0, 0, 0, 0)),
    (33, 40, (2, 3, 1, 2, 0)),
    (64, 64, (0, 2, 3, 1, 1)),
    (100, 150, (1, 1, 1, 1, 1)),
])
def test_random_batches_against_oracle(ctx, n_s, n_o, costs):
    """Ragged-free batches of random trees: tables, min cost, root and mapping."""
    batch = synth.make_batch(n_s, n_o, 6, seed=1000 + n_s, costs=costs)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    min_cost, root, mapping = ctx.dtl_run(batch.node_off, batch.left, batch.right,
                                          batch.leaf_species, batch.costs, flags=_lib.SR2_KEEP_TABLES)
    for b in range(batch.n_instances):
        left, right, leaf = batch.instance(b)
        ref = oracle.dtl(batch.parent, batch.child0, batch.child1, left, right, leaf, batch.costs,
                         strict=False)
        value, tag = ctx.dtl_tables(b, len(left))
        assert (value == ref["value"]).all()
        # the reference creates no tag for cells without an entry
        assert (tag == ref["tag"]).all()
        assert min_cost[b] == ref["min_cost"] and root[b] == ref["root_species"]
        lo, hi = batch.node_off[b], batch.node_off[b + 1]
        assert (mapping[lo:hi] == ref["mapping"]).all()


def test_ragged_batch_and_single_leaf(ctx):
    """Instances of different sizes in one batch, including a one-node object tree."""
    species = synth.make_batch(6, 6, 1, seed=5)
    parts = [synth.make_batch(6, n, 1, seed=5) for n in (6, 9, 17)]
    lefts = [p.left for p in parts] + [np.array([-1], np.int32)]
    rights = [p.right for p in parts] + [np.array([-1], np.int32)]
    leaves = [p.leaf_species for p in parts] + [np.array([3], np.int32)]
    node_off = np.cumsum([0] + [len(x) for x in lefts]).astype(np.int64)
    ctx.set_species(species.parent, species.child0, species.child1)
    min_cost, root, mapping = ctx.dtl_run(node_off, np.concatenate(lefts), np.concatenate(rights),
                                          np.concatenate(leaves), species.costs)
    for b in range(4):
        ref = oracle.dtl(species.parent, species.child0, species.child1, lefts[b], rights[b], leaves[b],
                         species.costs, strict=False, tables=False)
        assert min_cost[b] == ref["min_cost"] and root[b] == ref["root_species"]
        assert (mapping[node_off[b]:node_off[b + 1]] == ref["mapping"]).all()
    assert min_cost[3] == 0 and root[3] == 3


def test_large_species_tree_uses_cta_per_row(ctx):
    """|S| > 1024 switches to one CTA per row; checked against the oracle."""
    batch = synth.make_batch(600, 610, 1, seed=77)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    min_cost, root, mapping = ctx.dtl_run(batch.node_off, batch.left, batch.right,
                                          batch.leaf_species, batch.costs, flags=_lib.SR2_KEEP_TABLES)
    ref = oracle.dtl(batch.parent, batch.child0, batch.child1, batch.left, batch.right,
                     batch.leaf_species, batch.costs, strict=False)
    value, tag = ctx.dtl_tables(0, len(batch.left))
    assert (value == ref["value"]).all() and (tag == ref["tag"]).all()
    assert min_cost[0] == ref["min_cost"] and (mapping == ref["mapping"]).all()


def test_full_size_single_instance_properties(ctx):
    """BASELINE config #3 (5,000-leaf object tree x 2,000-leaf species tree): too large for
    the O(G S^2) oracle, so check size-independent properties: the host-side cost() of the
    returned mapping equals the reported minimum, every leaf sits at its species, and the
    result is reproducible."""
    batch = synth.make_batch(2000, 5000, 1, seed=3000)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
    min_cost, root, mapping = ctx.dtl_run(*args)
    assert 0 < min_cost[0] < INF
    leaf = batch.left < 0
    assert (mapping[leaf] == batch.leaf_species[leaf]).all()
    from tests.helpers import flat_cost
    assert flat_cost(batch, 0, mapping) == min_cost[0]
    again = ctx.dtl_run(*args)
    assert again[0][0] == min_cost[0] and (again[2] == mapping).all()


def test_batch_entry_point_matches_single(golden):
    fixture = golden("thl_medium.json.gz")[0]
    rec_input = ReconciliationInput.from_dict(fixture["input"])
    outs = reconcile_thl_batch([rec_input, rec_input], RetentionPolicy.ANY)
    for out in outs:
        (result,) = out
        assert result.to_dict() == fixture["result"]


def test_bad_input_is_rejected(ctx):
    batch = synth.make_batch(4, 5, 1, seed=1)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    left = batch.left.copy()
    left[0] = 0  # a node cannot be its own child
    with pytest.raises(_lib.Sr2Error):
        ctx.dtl_run(batch.node_off, left, batch.right, batch.leaf_species, batch.costs)
    with pytest.raises(_lib.Sr2Error):
        ctx.dtl_run(batch.node_off, batch.left, batch.right, batch.leaf_species,
                    np.array([0, -1, 1, 1, 1], np.int32))


@pytest.mark.parametrize("n_s,n_o,n_trees,costs", [
    (2, 3, 40, (0, 1, 1, 1, 1)),
    (9, 14, 37, (0, 1, 1, 1, 1)),
    (33, 50, 19, (0, 2, 1, 3, 1)),
    (100, 130, 5, (0, 0, 0, 0, 0)),
    (200, 210, 3, (0, 1, 1, 1, 1)),
    (240, 250, 2, (0, 1, 2, 1, 1)),   # 479 species: the largest pitch (480) the batched kernels take
])
@pytest.mark.parametrize("variant", ["flat2", "flat2-waves", "flat2-dense-cells", "flat"])
def test_batched_wave_kernels_against_oracle(ctx, monkeypatch, n_s, n_o, n_trees, costs, variant):
    """Force the batched-wave kernels (32-bit keys) on small batches, including partial last
    groups, and compare every cell and tag with the oracle: "flat2" (the default: sweep schedule in
    registers, sparse cells pass for rows with few root-path species; at these sizes every wave above
    the cherries joins the chained launch and the trees are traced back one CTA per tree),
    "flat2-waves" (one launch per wave for the sparse / dense row lists, traceback one launch per
    height), flat2 with the sparse pass switched off, and "flat", the first form, kept as a
    cross-check."""
    monkeypatch.setenv("SR2_DTL_SIMD_MIN_ROWS", "1")
    monkeypatch.setenv("SR2_DTL_VARIANT", variant.split("-")[0])
    if variant.endswith("dense-cells"):
        monkeypatch.setenv("SR2_FLAT2_SPARSE_MAX", "0")
    if variant.endswith("waves"):
        monkeypatch.setenv("SR2_DTL_CHAIN_ROWS", "0")
        monkeypatch.setenv("SR2_DTL_TRACE_BY_WAVE", "1")
    batch = synth.make_batch(n_s, n_o, n_trees, seed=31 + n_s, costs=costs)
    if n_s in (9, 33, 240):
        # the leaf map may name ANY species node: put a third of the object leaves at random
        # (mostly internal) species
        rng = np.random.default_rng(n_s)
        leaves = np.flatnonzero(batch.left < 0)
        moved = leaves[rng.random(len(leaves)) < 0.33]
        batch.leaf_species[moved] = rng.integers(0, batch.n_species, len(moved))
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    min_cost, root, mapping = ctx.dtl_run(batch.node_off, batch.left, batch.right,
                                          batch.leaf_species, batch.costs, flags=_lib.SR2_KEEP_TABLES)
    for b in range(batch.n_instances):
        left, right, leaf = batch.instance(b)
        ref = oracle.dtl(batch.parent, batch.child0, batch.child1, left, right, leaf, batch.costs,
                         strict=False)
        value, tag = ctx.dtl_tables(b, len(left))
        assert (value == ref["value"]).all()
        assert (tag == ref["tag"]).all()
        assert min_cost[b] == ref["min_cost"] and root[b] == ref["root_species"]
        lo, hi = batch.node_off[b], batch.node_off[b + 1]
        assert (mapping[lo:hi] == ref["mapping"]).all()
    # without SR2_KEEP_TABLES the cherry rows are never written: their readers evaluate the closed form
    again = ctx.dtl_run(batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
    for a, b in zip((min_cost, root, mapping), again):
        assert (a == b).all()


@pytest.mark.parametrize("mode", ["chained", "waves"])
def test_virtual_cherries_with_two_leaf_trees(ctx, monkeypatch, mode):
    """Cherry rows are not materialised (no SR2_KEEP_TABLES): trees that ARE a cherry (selection from
    the closed form), trees with one and with two cherry children at the root, leaves mapped to
    internal species (one leaf species an ancestor of the other); against the oracle."""
    monkeypatch.setenv("SR2_DTL_SIMD_MIN_ROWS", "1")
    if mode == "waves":
        monkeypatch.setenv("SR2_DTL_CHAIN_ROWS", "0")
    species = synth.make_batch(12, 12, 1, seed=3)
    rng = np.random.default_rng(8)
    shapes = {
        2: ([1, -1, -1], [2, -1, -1]),
        3: ([1, 2, -1, -1, -1], [4, 3, -1, -1, -1]),
        4: ([1, 2, -1, -1, 5, -1, -1], [4, 3, -1, -1, 6, -1, -1]),
    }
    lefts, rights, leaves = [], [], []
    for k in range(90):
        left, right = (np.array(v, np.int32) for v in shapes[2 + k % 3])
        leaf = np.where(left < 0, rng.integers(0, species.n_species, len(left)), -1).astype(np.int32)
        lefts.append(left), rights.append(right), leaves.append(leaf)
    node_off = np.cumsum([0] + [len(x) for x in lefts]).astype(np.int64)
    ctx.set_species(species.parent, species.child0, species.child1)
    for costs in ((0, 1, 1, 1, 1), (0, 3, 2, 1, 1), (0, 0, 0, 0, 0)):
        costs = np.array(costs, np.int32)
        min_cost, root, mapping = ctx.dtl_run(node_off, np.concatenate(lefts), np.concatenate(rights),
                                              np.concatenate(leaves), costs)
        for b in range(len(lefts)):
            ref = oracle.dtl(species.parent, species.child0, species.child1, lefts[b], rights[b], leaves[b],
                             costs, strict=False, tables=False)
            assert min_cost[b] == ref["min_cost"] and root[b] == ref["root_species"], (b, costs)
            assert (mapping[node_off[b]:node_off[b + 1]] == ref["mapping"]).all(), (b, costs)


def test_batch_of_cherries_and_single_nodes_only(ctx):
    """Every tree is a cherry or a single node: no wave kernel runs at all, selection and traceback
    work from the closed form."""
    species = synth.make_batch(9, 9, 1, seed=4)
    rng = np.random.default_rng(2)
    lefts, rights, leaves = [], [], []
    for k in range(70):
        if k % 5 == 4:
            left, right = np.array([-1], np.int32), np.array([-1], np.int32)
        else:
            left, right = np.array([1, -1, -1], np.int32), np.array([2, -1, -1], np.int32)
        leaf = np.where(left < 0, rng.integers(0, species.n_species, len(left)), -1).astype(np.int32)
        lefts.append(left), rights.append(right), leaves.append(leaf)
    node_off = np.cumsum([0] + [len(x) for x in lefts]).astype(np.int64)
    ctx.set_species(species.parent, species.child0, species.child1)
    costs = np.array((0, 2, 3, 1, 1), np.int32)
    min_cost, root, mapping = ctx.dtl_run(node_off, np.concatenate(lefts), np.concatenate(rights),
                                          np.concatenate(leaves), costs)
    for b in range(len(lefts)):
        ref = oracle.dtl(species.parent, species.child0, species.child1, lefts[b], rights[b], leaves[b],
                         costs, strict=False, tables=False)
        assert min_cost[b] == ref["min_cost"] and root[b] == ref["root_species"], b
        assert (mapping[node_off[b]:node_off[b + 1]] == ref["mapping"]).all(), b


def test_kernel_variants_agree_on_a_large_batch(ctx, monkeypatch):
    """Size-independent property at bench scale (config #2 shape, 600 trees): the batched-wave
    kernel and the row-per-warp kernel produce identical costs, roots and mappings."""
    batch = synth.make_batch(200, 500, 600, seed=2000)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
    monkeypatch.setenv("SR2_DTL_NO_SIMD", "1")
    plain = ctx.dtl_run(*args)
    monkeypatch.setenv("SR2_DTL_NO_SIMD", "0")
    monkeypatch.setenv("SR2_DTL_SIMD_MIN_ROWS", "1")
    simd = ctx.dtl_run(*args)
    for a, b in zip(plain, simd):
        assert (a == b).all()
    # wave 1: the sparse cherry kernel (root paths from the ancestor table, default) against the
    # dense closed-form kernel
    monkeypatch.setenv("SR2_DTL_DENSE_CHERRIES", "1")
    dense = ctx.dtl_run(*args)
    for a, b in zip(simd, dense):
        assert (a == b).all()
    monkeypatch.delenv("SR2_DTL_DENSE_CHERRIES")
    # cherry rows written by the wave-1 kernel instead of being evaluated by their readers
    monkeypatch.setenv("SR2_DTL_REAL_CHERRIES", "1")
    real = ctx.dtl_run(*args)
    for a, b in zip(simd, real):
        assert (a == b).all()
    monkeypatch.delenv("SR2_DTL_REAL_CHERRIES")
    # top waves: one chained launch for all waves above the cherries / one launch per wave;
    # traceback: one launch per height instead of one CTA per tree
    for chain_rows, by_wave in (("1000000000", "0"), ("0", "1")):
        monkeypatch.setenv("SR2_DTL_CHAIN_ROWS", chain_rows)
        if by_wave == "1":
            monkeypatch.setenv("SR2_DTL_TRACE_BY_WAVE", "1")
        other = ctx.dtl_run(*args)
        for a, b in zip(simd, other):
            assert (a == b).all()
    from tests.helpers import flat_cost
    for b in (0, 299, 599):
        assert flat_cost(batch, b, simd[2]) == simd[0][b]


def test_all_solutions_against_the_reference(golden):
    """reconcile_thl(..., RetentionPolicy.ALL) (`--solutions all`): the value
    table comes from the GPU, the reference's ALL-retention tag sets are rebuilt on the host for
    the visited cells.  Set equality with the reference's full solution sets on 160 instances
    (up to 3,000 solutions each), incl. tests/compute/test_reconciliation.py:67-123."""
    from superrec2_b200.compute.reconciliation import reconcile_thl
    from superrec2_b200.model.reconciliation import ReconciliationInput
    from superrec2_b200.utils.dynamic_programming import RetentionPolicy
    for fixture in golden("thl_all.json.gz"):
        rec_input = ReconciliationInput.from_dict(fixture["input"])
        results = reconcile_thl(rec_input, RetentionPolicy.ALL)
        got = sorted(sorted([node.name, species.name] for node, species in out.object_species.items())
                     for out in results)
        assert got == fixture["solutions"]
        assert {out.cost() for out in results} == {fixture["min_cost"]}


@pytest.mark.parametrize("pinned", [False, True])
def test_device_side_bucketing_equals_host_side(ctx, monkeypatch, pinned):
    """sr2_dtl_run buckets big batches on the device (rows_height_kernel / rows_fill_kernel, one
    thread per tree).  Same costs, roots and mappings as the host-side bucketing, from pageable and
    from pinned input arrays; a malformed tree is reported with its instance number; a tree higher
    than the device tables falls back to the host path."""
    batch = synth.make_batch(40, 90, 3000, seed=77)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    arrays = [batch.left, batch.right, batch.leaf_species]
    if pinned:
        import torch
        arrays = [torch.from_numpy(a.copy()).pin_memory().numpy() for a in arrays]
    monkeypatch.setenv("SR2_DEVICE_ROWS", "0")
    host = ctx.dtl_run(batch.node_off, *arrays, batch.costs)
    monkeypatch.setenv("SR2_DEVICE_ROWS", "1")
    out = None
    if pinned:  # pinned result arrays are filled by the device-to-host copies directly
        out = tuple(torch.from_numpy(np.full_like(a, -7)).pin_memory().numpy() for a in host)
    dev = ctx.dtl_run(batch.node_off, *arrays, batch.costs, out=out)
    for a, b in zip(host, dev):
        assert (a == b).all()
    # malformed: a child numbered before its parent
    bad = batch.left.copy()
    off = int(batch.node_off[1234])
    bad[off + 5], bad[off + 0] = bad[off + 0], 0
    with pytest.raises(_lib.Sr2Error, match="instance 1234"):
        ctx.dtl_run(batch.node_off, bad, batch.right, batch.leaf_species, batch.costs)
    # caterpillar trees (height = leaves - 1 = 199 > 127): handled by the host path
    n = 200
    left = np.full(2 * n - 1, -1, np.int32)
    right = np.full(2 * n - 1, -1, np.int32)
    for u in range(n - 1):
        left[u] = u + 1 if u + 1 < n - 1 else 2 * n - 2
        right[u] = n - 1 + u
    species_leaves = np.flatnonzero(batch.child0 < 0).astype(np.int32)
    leaf = np.full(2 * n - 1, -1, np.int32)
    leaf[n - 1:] = species_leaves[np.arange(n) % len(species_leaves)]
    count = 1100
    node_off = np.arange(count + 1, dtype=np.int64) * (2 * n - 1)
    tall = ctx.dtl_run(node_off, np.tile(left, count), np.tile(right, count), np.tile(leaf, count), batch.costs)
    ref = oracle.dtl(batch.parent, batch.child0, batch.child1, left, right, leaf, batch.costs, strict=False)
    assert (tall[0] == ref["min_cost"]).all() and (tall[1] == ref["root_species"]).all()
    assert (tall[2].reshape(count, -1) == ref["mapping"]).all()


def test_full_size_batch_pipelined_run_equals_stepwise(ctx):
    """BASELINE config #2 at full size (10,000 x 500-leaf trees vs 200 leaves, 1.99e9 cells):
    sr2_dtl_run (pipelined chunks, device-side bucketing, two compute streams) returns exactly what
    upload + solve + results (one chunk, host-side bucketing) returns; the checksum of the minimum
    costs is the one bench.py prints; a sample of trees is re-checked with the host cost()."""
    from tests.helpers import flat_cost
    batch = synth.make_batch(200, 500, 10000, seed=2000)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
    ctx.dtl_upload(*args)
    ctx.dtl_solve()
    stepwise = ctx.dtl_results()
    run = ctx.dtl_run(*args)
    for a, b in zip(stepwise, run):
        assert (a == b).all()
    assert int(run[0].astype(np.int64).sum()) == 4774426
    for b in (0, 4999, 9999):
        assert flat_cost(batch, b, run[2]) == run[0][b]
    # an empty batch is a valid call
    empty = ctx.dtl_run(np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32),
                        np.zeros(0, np.int32), batch.costs)
    assert all(len(part) == 0 for part in empty)


def _caterpillar_species(n_leaves):
    """A maximally unbalanced species tree in level-order numbering: node 2k is internal with
    children 2k+1 (a leaf) and 2k+2."""
    size = 2 * n_leaves - 1
    parent = np.full(size, -1, np.int32)
    child0 = np.full(size, -1, np.int32)
    child1 = np.full(size, -1, np.int32)
    for k in range(n_leaves - 1):
        child0[2 * k], child1[2 * k] = 2 * k + 1, 2 * k + 2
        parent[2 * k + 1] = parent[2 * k + 2] = 2 * k
    return parent, child0, child1


@pytest.mark.parametrize("n_species_leaves", [40, 120])
def test_deep_species_tree_against_oracle(ctx, monkeypatch, n_species_leaves):
    """A caterpillar species tree (depth = leaves - 1): more levels than the flat2 schedule holds
    (the generic wave kernels take over), long root paths in the sparse cherry kernel and ancestor
    table.  Every cell and tag against the oracle."""
    monkeypatch.setenv("SR2_DTL_SIMD_MIN_ROWS", "1")
    parent, child0, child1 = _caterpillar_species(n_species_leaves)
    base = synth.make_batch(n_species_leaves, n_species_leaves + 12, 6, seed=5)
    leaves = np.flatnonzero(child0 < 0).astype(np.int32)
    rng = np.random.default_rng(3)
    leaf_species = np.where(base.left < 0, leaves[rng.integers(0, len(leaves), len(base.left))], -1).astype(np.int32)
    ctx.set_species(parent, child0, child1)
    costs = np.asarray([0, 2, 3, 1, 1], np.int32)
    min_cost, root, mapping = ctx.dtl_run(base.node_off, base.left, base.right, leaf_species, costs,
                                          flags=_lib.SR2_KEEP_TABLES)
    for b in range(base.n_instances):
        lo, hi = int(base.node_off[b]), int(base.node_off[b + 1])
        ref = oracle.dtl(parent, child0, child1, base.left[lo:hi], base.right[lo:hi], leaf_species[lo:hi], costs,
                         strict=False)
        value, tag = ctx.dtl_tables(b, hi - lo)
        assert (value == ref["value"]).all()
        assert (tag == ref["tag"]).all()
        assert min_cost[b] == ref["min_cost"] and root[b] == ref["root_species"]
        assert (mapping[lo:hi] == ref["mapping"]).all()


def test_device_bucketing_ragged_batch_with_single_leaf_trees(ctx, monkeypatch):
    """More than 1024 instances of different sizes, one-node trees included: the pipelined run
    (device-side bucketing, chunks, both compute streams) against the oracle, tree by tree."""
    parts = [synth.make_batch(3, n_o, 400, seed=60 + n_o) for n_o in (3, 4, 6, 9)]
    species = parts[0]
    lefts, rights, leaves, sizes = [], [], [], []
    rng = np.random.default_rng(1)
    species_leaves = np.flatnonzero(species.child0 < 0).astype(np.int32)
    for k in range(1600 + 300):
        if k % 6 == 5:  # a one-node tree
            lefts.append(np.array([-1], np.int32))
            rights.append(np.array([-1], np.int32))
            leaves.append(species_leaves[rng.integers(0, len(species_leaves), 1)])
            sizes.append(1)
            continue
        part = parts[k % 4]
        left, right, leaf = part.instance(k % 400)
        lefts.append(left)
        rights.append(right)
        leaves.append(leaf)
        sizes.append(len(left))
    node_off = np.zeros(len(sizes) + 1, np.int64)
    np.cumsum(sizes, out=node_off[1:])
    left, right, leaf = np.concatenate(lefts), np.concatenate(rights), np.concatenate(leaves)
    ctx.set_species(species.parent, species.child0, species.child1)
    monkeypatch.setenv("SR2_PIPELINE_CHUNKS", "5")
    min_cost, root, mapping = ctx.dtl_run(node_off, left, right, leaf, species.costs)
    for b in range(0, len(sizes), 7):
        lo, hi = int(node_off[b]), int(node_off[b + 1])
        ref = oracle.dtl(species.parent, species.child0, species.child1, left[lo:hi], right[lo:hi], leaf[lo:hi],
                         species.costs, strict=False)
        assert min_cost[b] == ref["min_cost"] and root[b] == ref["root_species"], b
        assert (mapping[lo:hi] == ref["mapping"]).all(), b


def test_large_costs_leave_the_32_bit_kernels(ctx, monkeypatch):
    """Costs too large for 32-bit (value, rank, depth) keys: the batched kernels step aside for the
    64-bit row-per-warp kernel; tables and results still equal the oracle's."""
    monkeypatch.setenv("SR2_DTL_SIMD_MIN_ROWS", "1")
    costs = (0, 700000, 900000, 40000, 1)
    batch = synth.make_batch(20, 30, 6, seed=12, costs=costs)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    min_cost, root, mapping = ctx.dtl_run(batch.node_off, batch.left, batch.right, batch.leaf_species,
                                          batch.costs, flags=_lib.SR2_KEEP_TABLES)
    for b in range(batch.n_instances):
        left, right, leaf = batch.instance(b)
        ref = oracle.dtl(batch.parent, batch.child0, batch.child1, left, right, leaf, batch.costs, strict=False)
        value, tag = ctx.dtl_tables(b, len(left))
        assert (value == ref["value"]).all() and (tag == ref["tag"]).all()
        assert min_cost[b] == ref["min_cost"] and root[b] == ref["root_species"]


def test_traceback_of_a_large_tree_per_tree_and_per_wave(ctx, monkeypatch):
    """A 15,000-leaf object tree: the per-tree traceback needs a 120 KB frontier (opt-in shared
    memory, 1,024 threads); same mapping as the launch-per-height traceback, and its cost() is
    the reported minimum."""
    batch = synth.make_batch(50, 15000, 1, seed=91)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
    per_tree = ctx.dtl_run(*args)
    monkeypatch.setenv("SR2_DTL_TRACE_BY_WAVE", "1")
    per_wave = ctx.dtl_run(*args)
    for a, b in zip(per_tree, per_wave):
        assert (a == b).all()
    from tests.helpers import flat_cost
    assert flat_cost(batch, 0, per_tree[2]) == per_tree[0][0] < INF


def test_batch_of_one_node_trees_only(ctx):
    """No internal object node at all (zero DP rows), through the device-bucketing path."""
    species = synth.make_batch(6, 6, 1, seed=2)
    ctx.set_species(species.parent, species.child0, species.child1)
    count = 1500
    leaves = np.flatnonzero(species.child0 < 0).astype(np.int32)
    leaf = leaves[np.arange(count) % len(leaves)]
    none = np.full(count, -1, np.int32)
    min_cost, root, mapping = ctx.dtl_run(np.arange(count + 1, dtype=np.int64), none, none, leaf, species.costs)
    assert (min_cost == 0).all() and (root == leaf).all() and (mapping == leaf).all()


@pytest.mark.parametrize("pinned", [False, True])
def test_sixteen_bit_run_equals_the_32_bit_run(ctx, pinned):
    """sr2_dtl_run16 (int16 arrays over PCIe, widened / narrowed on the device) returns exactly what
    sr2_dtl_run returns: on the device-bucketing path (big batch, pageable and pinned buffers), on the
    host path (small batch), and it refuses trees that do not fit 16 bits."""
    import torch
    for n_s, n_o, n_trees in [(60, 90, 3000), (12, 20, 7)]:
        batch = synth.make_batch(n_s, n_o, n_trees, seed=1600 + n_trees)
        ctx.set_species(batch.parent, batch.child0, batch.child1)
        want = ctx.dtl_run(batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
        arrays = [np.ascontiguousarray(a, dtype=np.int16) for a in (batch.left, batch.right, batch.leaf_species)]
        out = None
        if pinned:
            arrays = [torch.from_numpy(a).pin_memory().numpy() for a in arrays]
            out = (torch.zeros(n_trees, dtype=torch.int32).pin_memory().numpy(),
                   torch.zeros(n_trees, dtype=torch.int16).pin_memory().numpy(),
                   torch.zeros(len(batch.left), dtype=torch.int16).pin_memory().numpy())
        got = ctx.dtl_run16(batch.node_off, *arrays, batch.costs, out=out)
        assert got[1].dtype == np.int16 and got[2].dtype == np.int16
        for a, b in zip(got, want):
            assert np.array_equal(a.astype(np.int32), b)
    big = synth.make_batch(10, 20000, 1, seed=3)
    ctx.set_species(big.parent, big.child0, big.child1)
    with pytest.raises(_lib.Sr2Error) as err:
        ctx.dtl_run16(big.node_off, big.left % 30000, big.right % 30000, big.leaf_species, big.costs)
    assert err.value.code == -3


def test_very_large_species_tree_uses_the_two_row_kernel(ctx):
    """More than ~5,900 species nodes: the four key rows of dtl_wave_kernel no longer fit one CTA's shared
    memory and the two-row in-place form runs instead.  6,399 species nodes, a 24-leaf object tree whose
    leaves sit at random species leaves (not surjective: roots without an entry are skipped, as in the
    oracle's non-strict mode); full table, tags, result against the oracle."""
    species = synth.make_batch(3200, 3200, 1, seed=61)
    small = synth.make_batch(24, 24, 1, seed=62)
    rng = np.random.default_rng(63)
    leaves = np.flatnonzero(species.child0 < 0)
    leaf_species = np.where(small.left < 0, leaves[rng.integers(0, len(leaves), len(small.left))], -1).astype(np.int32)
    ctx.set_species(species.parent, species.child0, species.child1)
    args = (small.node_off, small.left, small.right, leaf_species, small.costs)
    min_cost, root, mapping = ctx.dtl_run(*args, flags=_lib.SR2_KEEP_TABLES)
    ref = oracle.dtl(species.parent, species.child0, species.child1, small.left, small.right, leaf_species,
                     small.costs, strict=False)
    value, tag = ctx.dtl_tables(0, len(small.left))
    assert (value == ref["value"]).all() and (tag == ref["tag"]).all()
    assert min_cost[0] == ref["min_cost"] and root[0] == ref["root_species"] and (mapping == ref["mapping"]).all()


def test_flat_entry_point_and_empty_sixteen_bit_batch(ctx):
    """reconcile_thl_flat (arrays in, arrays out; picks the 16-bit call when it fits) equals the object API on
    the same instances, and an empty batch is a valid 16-bit call."""
    from superrec2_b200.compute.reconciliation import reconcile_thl_flat
    batch = synth.make_batch(20, 30, 6, seed=91)
    min_cost, root, mapping = reconcile_thl_flat(batch.parent, batch.child0, batch.child1, batch.node_off,
                                                 batch.left, batch.right, batch.leaf_species, batch.costs)
    assert mapping.dtype == np.int16
    for b in range(batch.n_instances):
        rec_input = ReconciliationInput.from_dict(batch.to_dict(b))
        (result,) = reconcile_thl(rec_input, RetentionPolicy.ANY)
        assert result.cost() == min_cost[b]
        flat = FlatDtl(batch.to_dict(b))
        lo, hi = int(batch.node_off[b]), int(batch.node_off[b + 1])
        ref = oracle.dtl(*flat.species_args, *flat.object_args, flat.costs, strict=False, tables=False)
        # the flat batch and the object API number nodes differently; compare through the oracle on each side
        assert ref["min_cost"] == min_cost[b]
        own = oracle.dtl(batch.parent, batch.child0, batch.child1, *batch.instance(b), batch.costs, strict=False,
                         tables=False)
        assert (own["mapping"] == mapping[lo:hi]).all() and own["root_species"] == root[b]
    empty = ctx.dtl_run16(np.zeros(1, np.int64), np.zeros(0, np.int16), np.zeros(0, np.int16),
                          np.zeros(0, np.int16), batch.costs)
    assert all(len(part) == 0 for part in empty)


_THREAD_TEST_ORACLE = {}


@pytest.mark.parametrize("tiers", [None, "6,13,64", "64", "9,10,11,12,40"])
@pytest.mark.parametrize("costs", [(0, 1, 1, 1, 1), (0, 5, 2, 3, 1), (0, 0, 4, 0, 2)])
def test_thread_per_row_kernel_against_oracle(ctx, monkeypatch, tiers, costs):
    """dtl_wave_thread_kernel (a thread per sparse row, rows stored compact) with its readers (itself, flat2's
    scatter of a compact child, the per-tree traceback): 700 trees of 40 leaves over 150 species leaves, a third
    of the object leaves at random -- mostly internal -- species (a leaf species that is an ancestor of its
    sibling's: the flags of the generation are not exact there), tier limits that put the cuts at odd sizes and
    a child with exactly as many entries as the tier holds.  Costs, roots and mappings of every tree against
    the oracle (no SR2_KEEP_TABLES: compact rows and virtual cherries are only used when nobody reads the tables)."""
    monkeypatch.setenv("SR2_DTL_SIMD_MIN_ROWS", "1")
    monkeypatch.setenv("SR2_DTL_CHAIN_ROWS", "300")   # (by default every wave of a batch this small is chained)
    if tiers:
        monkeypatch.setenv("SR2_THREAD_TIERS", tiers)
    species = synth.make_batch(150, 150, 1, seed=78)
    batch = synth.make_batch(40, 40, 700, seed=77)     # the shapes of the object trees
    rng = np.random.default_rng(5)
    species_leaves = np.flatnonzero(species.child0 < 0)
    n_nodes = len(batch.left)
    leaf_species = np.where(rng.random(n_nodes) < 0.33, rng.integers(0, species.n_species, n_nodes),
                            species_leaves[rng.integers(0, len(species_leaves), n_nodes)])
    leaf_species = np.where(batch.left < 0, leaf_species, -1).astype(np.int32)
    ctx.set_species(species.parent, species.child0, species.child1)
    args = (batch.node_off, batch.left, batch.right, leaf_species, np.asarray(costs, np.int32))
    got = ctx.dtl_run(*args)
    launches = ctx.timing().n_launches
    if costs not in _THREAD_TEST_ORACLE:   # the oracle's answer does not depend on the tiers
        _THREAD_TEST_ORACLE[costs] = oracle.dtl_batch(species.parent, species.child0, species.child1, *args)
    want = _THREAD_TEST_ORACLE[costs]
    for g, w in zip(got, want):
        assert np.array_equal(np.asarray(g), np.asarray(w))
    # the 4-lanes-per-row kernel with dense rows (the path before compact rows) gives the same answers
    monkeypatch.setenv("SR2_SPARSE_THREAD", "0")
    for g, w in zip(ctx.dtl_run(*args), want):
        assert np.array_equal(np.asarray(g), np.asarray(w))
    # (one launch per tier and wave against two: the first run did go through the thread kernel)
    n_tiers = len(tiers.split(",")) if tiers else 4
    assert n_tiers in (2, 3) or ctx.timing().n_launches != launches


@pytest.mark.parametrize("dup", [40000, 52000])
def test_compact_rows_next_to_their_value_limit(ctx, monkeypatch, dup):
    """A compact entry holds a 22-bit value.  With a 15-node species tree the 32-bit keys leave 25 bits for
    values, so the bound of the table values (nodes x largest event term) can sit on either side of 2^22:
    79 x (40,000 + ...) = 3.2 M keeps compact rows (table values in the millions go through the entries),
    79 x (52,000 + ...) = 4.1 M+ makes the library store dense rows.  Every tree against the oracle either way."""
    monkeypatch.setenv("SR2_DTL_SIMD_MIN_ROWS", "1")
    monkeypatch.setenv("SR2_DTL_CHAIN_ROWS", "300")
    species = synth.make_batch(8, 8, 1, seed=12)
    batch = synth.make_batch(40, 40, 500, seed=13)
    rng = np.random.default_rng(14)
    leaf_species = np.where(batch.left < 0, rng.integers(0, species.n_species, len(batch.left)), -1).astype(np.int32)
    costs = np.asarray((0, dup, 30000, 100, 1), np.int32)
    ctx.set_species(species.parent, species.child0, species.child1)
    args = (batch.node_off, batch.left, batch.right, leaf_species, costs)
    got = ctx.dtl_run(*args)
    launches = ctx.timing().n_launches
    want = oracle.dtl_batch(species.parent, species.child0, species.child1, *args)
    assert int(np.asarray(want[0]).max()) > 200000   # values far beyond what config #2 ever holds
    for g, w in zip(got, want):
        assert np.array_equal(np.asarray(g), np.asarray(w))
    monkeypatch.setenv("SR2_SPARSE_THREAD", "0")
    for g, w in zip(ctx.dtl_run(*args), want):
        assert np.array_equal(np.asarray(g), np.asarray(w))
    # below the limit the first run went through the thread kernel (four tiers per wave, not two)
    assert (ctx.timing().n_launches != launches) == (dup == 40000)


def test_thread_per_row_kernel_largest_species_tree(ctx, monkeypatch):
    """479 species nodes (the largest pitch, 480, of the batched kernels; 9-bit ranks, 15 words of path set): the
    thread-per-row kernel, flat2 reading its compact rows and the traceback, 300 trees of 70 leaves against the oracle."""
    monkeypatch.setenv("SR2_DTL_SIMD_MIN_ROWS", "1")
    monkeypatch.setenv("SR2_DTL_CHAIN_ROWS", "200")
    species = synth.make_batch(240, 240, 1, seed=90)
    batch = synth.make_batch(70, 70, 300, seed=91)
    rng = np.random.default_rng(92)
    species_leaves = np.flatnonzero(species.child0 < 0)
    n_nodes = len(batch.left)
    leaf_species = np.where(rng.random(n_nodes) < 0.2, rng.integers(0, species.n_species, n_nodes),
                            species_leaves[rng.integers(0, len(species_leaves), n_nodes)])
    leaf_species = np.where(batch.left < 0, leaf_species, -1).astype(np.int32)
    costs = np.asarray((0, 2, 3, 1, 1), np.int32)
    ctx.set_species(species.parent, species.child0, species.child1)
    args = (batch.node_off, batch.left, batch.right, leaf_species, costs)
    got = ctx.dtl_run(*args)
    want = oracle.dtl_batch(species.parent, species.child0, species.child1, *args)
    for g, w in zip(got, want):
        assert np.array_equal(np.asarray(g), np.asarray(w))
    got16 = ctx.dtl_run16(*args)
    for g, w in zip(got16, want):
        assert np.array_equal(np.asarray(g, np.int32), np.asarray(w))


@pytest.mark.parametrize("n_species_leaves", [17, 31])
def test_thread_per_row_kernel_on_a_caterpillar(ctx, monkeypatch, n_species_leaves):
    """A caterpillar species tree with 17 / 31 leaves (depths up to 16 / 30: the deepest a 5-bit depth field holds):
    path sets are long chains with one leaf hanging off every element, siblings alternate between "next element" and
    "previous element".  600 trees of 45 leaves, leaves at any species node, against the oracle."""
    monkeypatch.setenv("SR2_DTL_SIMD_MIN_ROWS", "1")
    monkeypatch.setenv("SR2_DTL_CHAIN_ROWS", "300")
    parent, child0, child1 = _caterpillar_species(n_species_leaves)
    batch = synth.make_batch(45, 45, 600, seed=101)
    rng = np.random.default_rng(102)
    leaf_species = np.where(batch.left < 0, rng.integers(0, len(parent), len(batch.left)), -1).astype(np.int32)
    ctx.set_species(parent, child0, child1)
    for costs in ((0, 1, 1, 1, 1), (0, 3, 2, 2, 0)):
        args = (batch.node_off, batch.left, batch.right, leaf_species, np.asarray(costs, np.int32))
        got = ctx.dtl_run(*args)
        want = oracle.dtl_batch(parent, child0, child1, *args)
        for g, w in zip(got, want):
            assert np.array_equal(np.asarray(g), np.asarray(w))
