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
    for fixture in (golden("superdtl_small.json.gz") + golden("superdtl_reference.json.gz") +
                    golden("superdtl_internal_species.json.gz")):
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
    fixtures = (golden("superdtl_poly.json.gz") + golden("crispr.json.gz") + golden("superdtl_small.json.gz")[:30]
                + golden("superdtl_species_poly.json.gz"))
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


def test_large_costs_are_exact_or_refused(ctx):
    """the packed best key keeps 23 bits of cost.  Costs whose cost() bound stays
    below 2^23 are exact (tables, best key and results against the oracle); beyond it the upload
    is refused with SR2_ERR_RANGE instead of truncating the key."""
    for costs in [(3000, 20000, 15000, 700, 9000), (0, 40000, 40000, 1000, 40000)]:
        batch = synth.make_batch(8, 12, 6, seed=77, costs=costs, n_families=6)
        ctx.set_species(batch.parent, batch.child0, batch.child1)
        ctx.superdtl_upload(batch.node_off, batch.left, batch.right, batch.leaf_species,
                            batch.leaf_family_bits, batch.costs, index_base=5)
        ctx.superdtl_solve()
        min_cost, root, mapping, kind, syn = ctx.superdtl_results()
        for b in range(batch.n_instances):
            lo, hi = int(batch.node_off[b]), int(batch.node_off[b + 1])
            ref = oracle.superdtl(batch.parent, batch.child0, batch.child1, *batch.instance(b),
                                  batch.leaf_family_bits[lo:hi], batch.costs, tables=False)
            assert (min_cost[b], root[b]) == (ref["min_cost"], ref["root_species"])
            assert (mapping[lo:hi] == ref["mapping"]).all() and (syn[lo:hi] == ref["syn"]).all()
        cost, index, root_species, key = ctx.superdtl_best()
        first = int(np.flatnonzero(min_cost == min_cost.min())[0])
        assert (cost, index, root_species) == (int(min_cost.min()), 5 + first, int(root[first]))
        assert cost > 0 and key == (cost << 40) | (index << 16) | root_species
    batch = synth.make_batch(8, 12, 2, seed=77, costs=(0, 1 << 20, 1 << 20, 1 << 18, 1 << 20), n_families=6)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    with pytest.raises(_lib.Sr2Error) as err:
        ctx.superdtl_upload(batch.node_off, batch.left, batch.right, batch.leaf_species,
                            batch.leaf_family_bits, batch.costs)
    assert err.value.code == -3 and "packed" in str(err.value)


# ------------------------------------------------------------------ resolutions at scale
def _poly_setup(ctx, data):
    from superrec2_b200.compute.flatten import flatten_polytomies, flatten_species
    from superrec2_b200.compute.reconciliation import cost_vector
    from superrec2_b200.compute.unordered_super_reconciliation import family_index, leaf_bitsets
    srec_input = SuperReconciliationInput.from_dict(data)
    species = flatten_species(srec_input.species_lca.tree)
    poly = flatten_polytomies(srec_input.object_tree, srec_input.leaf_object_species, species.rank)
    families = family_index(srec_input.leaf_syntenies)
    bits = leaf_bitsets(poly.nodes, srec_input.leaf_syntenies, families)
    ctx.set_species(species.parent, species.child0, species.child1)
    return srec_input, species, poly, bits, cost_vector(srec_input.costs)


def test_resolution_batches_against_oracle_per_resolution(ctx):
    """A 7,875-resolution input (5-ary x 4-ary... of 2-leaf subtrees): every 37th resolution is
    re-solved alone by the CPU oracle through the host-side unranker (pinned to the reference's
    enumeration order) and must give the same (cost, root) as the library-side unranked batch."""
    data = synth.make_polytomy_input(8, (5, 4), 2, 6, seed=77)
    srec_input, species, poly, bits, costs = _poly_setup(ctx, data)
    total = srec_input.count_resolutions()
    assert total == 105 * 15
    ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, 0, total)
    ctx.superdtl_solve()
    min_cost, root, *_ = ctx.superdtl_results()
    for index in list(range(0, total, 37)) + [total - 1]:
        flat = FlatSuper(srec_input.resolution(index))
        ref = oracle.superdtl(*flat.species_args, *flat.object_args, flat.leaf_bits, flat.costs, tables=False)
        assert (min_cost[index], root[index]) == (ref["min_cost"], ref["root_species"]), index
    cost, index, root_species, key = ctx.superdtl_best()
    first = int(np.flatnonzero(min_cost == min_cost.min())[0])
    assert (cost, index, root_species) == (int(min_cost.min()), first, int(root[first]))


def test_sharded_ranges_agree_with_single_range(ctx):
    """Sharding property: min over shards of the per-shard key == key of the unsharded run, for
    2, 3 and 8 shards (what the NCCL min-reduce computes across GPUs)."""
    from superrec2_b200.dist import shard_range
    data = synth.make_polytomy_input(10, (4, 5), 2, 8, seed=5)
    srec_input, species, poly, bits, costs = _poly_setup(ctx, data)
    total = srec_input.count_resolutions()
    ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, 0, total)
    ctx.superdtl_solve()
    whole = ctx.superdtl_best()[3]
    for world in (2, 3, 8):
        keys = []
        for rank in range(world):
            lo, hi = shard_range(total, rank, world)
            ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, lo, hi - lo)
            ctx.superdtl_solve()
            keys.append(ctx.superdtl_best()[3])
        assert min(keys) == whole


def test_species_polytomies_sharded_entry_point(golden):
    """Species-tree polytomies: resolution number = object_number * n_species + species_number.
    Shards of that numbering, reduced with min over the packed keys, return the reference's
    solution from exactly one shard (2, 3 and 5 shards)."""
    from superrec2_b200.compute.unordered_super_reconciliation import usreconcile_extended_uspfs_sharded
    for fixture in golden("superdtl_species_poly.json.gz")[:12]:
        srec_input = SuperReconciliationInput.from_dict(fixture["input"])
        for world in (2, 3, 5):
            keys = []
            for rank in range(world):
                usreconcile_extended_uspfs_sharded(
                    srec_input, shard=rank, n_shards=world,
                    reduce_key=lambda found: keys.append(found))
            keys = [key for key in keys if key is not None]
            winner = min(keys) if keys else None
            found = []
            for rank in range(world):
                found += list(usreconcile_extended_uspfs_sharded(
                    srec_input, shard=rank, n_shards=world, reduce_key=lambda key: winner))
            if fixture["result"] is None:
                assert not found
                continue
            (result,) = found
            assert result.to_dict() == fixture["result"]


def test_config5_full_enumeration_properties(ctx):
    """BASELINE config #5: 99,225 resolutions of a 33-leaf tree, 59 species, 32 families.  Too many
    for the oracle, so: the entry point's answer must (1) have the cost the host-side cost()
    computes from the returned mapping and syntenies, (2) be the resolution the packed key names,
    (3) be reproduced by the oracle on that one resolution, and (4) a sample of other resolutions
    must not beat it."""
    data = synth.make_polytomy_input()
    srec_input = SuperReconciliationInput.from_dict(data)
    assert srec_input.count_resolutions() == 99225
    (result,) = usreconcile_extended_uspfs(srec_input, RetentionPolicy.ANY)
    best_cost = result.cost()
    flat = FlatSuper(result.input)
    ref = oracle.superdtl(*flat.species_args, *flat.object_args, flat.leaf_bits, flat.costs, tables=False)
    assert ref["min_cost"] == best_cost
    assert flat.result_dict(ref["mapping"], ref["syn"]) == result.to_dict()
    rng = np.random.default_rng(0)
    for index in rng.integers(0, 99225, 40):
        other = FlatSuper(srec_input.resolution(int(index)))
        out = oracle.superdtl(*other.species_args, *other.object_args, other.leaf_bits, other.costs,
                              tables=False)
        assert out["min_cost"] >= best_cost


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [
    (8, (5, 4), 2, 6, 77),     # binary root over a 5-ary and a 4-ary node, 2-leaf subtrees
    (10, (4, 5), 2, 8, 5),
    (6, (3, 6), 1, 5, 9),      # leaves directly under the polytomies
    (12, (6, 2), 3, 40, 21),   # two 32-bit words of families
])
def test_shared_rows_equal_unshared_rows(ctx, monkeypatch, shape):
    """Resolutions are solved with one table row per DISTINCT subtree and trees unranked on the
    device (SharedPlan).  Every per-resolution result (cost, root, mapping, kinds, syntenies) must
    equal what the plain per-resolution batch (host unranking, one row per node) gives, for the
    whole enumeration and for a sub-range (the multi-GPU sharding case)."""
    n_s, arities, sub, fam, seed = shape
    data = synth.make_polytomy_input(n_s, arities, sub, fam, seed=seed)
    srec_input, species, poly, bits, costs = _poly_setup(ctx, data)
    total = srec_input.count_resolutions()
    for lo, hi in ((0, total), (total // 3, total // 3 + max(1, total // 5))):
        outs = []
        for unshared in ("1", "0"):
            monkeypatch.setenv("SR2_RESOLUTIONS_UNSHARED", unshared)
            ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, lo, hi - lo)
            ctx.superdtl_solve()
            fused = ctx.superdtl_results(costs_only=True)   # straight from the fused decode kernel
            outs.append((ctx.superdtl_results(), ctx.superdtl_best()))
            assert (fused[0] == outs[-1][0][0]).all() and (fused[1] == outs[-1][0][1]).all()
        (plain, plain_best), (shared, shared_best) = outs
        assert plain_best == shared_best
        for a, b in zip(plain, shared):
            assert (a == b).all()


def test_base_uspfs_against_the_reference(golden):
    """usreconcile_base_uspfs = the superdtl table restricted to the LCA mapping (`allowed_species`,
    reference :500-520).  Outputs equal the reference's on its own test instances
    (tests/compute/test_unordered_super_reconciliation.py:99-331) and on 120 random ones."""
    from superrec2_b200.compute.unordered_super_reconciliation import usreconcile_base_uspfs
    for fixture in golden("base_uspfs.json.gz"):
        srec_input = SuperReconciliationInput.from_dict(fixture["input"])
        results = usreconcile_base_uspfs(srec_input, RetentionPolicy.ANY)
        if fixture["result"] is None:
            assert not results
            continue
        (result,) = results
        assert result.to_dict() == fixture["result"]
        assert result.cost() == fixture["min_cost"]


def test_all_solutions_against_the_reference(golden):
    """``--solutions all`` (RetentionPolicy.ALL): the whole set of minimum-cost solutions -- winning
    resolutions, mappings AND synteny labels -- equals the set the unmodified reference returned, on
    90 instances (binary and multifurcating) whose reference answer was the same in three runs
    (``stable``; see compute/unordered_super_all.py for what can depend on the run in the reference)."""
    import json
    fixtures = golden("superdtl_all.json.gz")
    assert sum(len(fixture["solutions"]) > 1 for fixture in fixtures) >= 30
    for fixture in fixtures:
        srec_input = SuperReconciliationInput.from_dict(fixture["input"])
        results = usreconcile_extended_uspfs(srec_input, RetentionPolicy.ALL)
        got = sorted(json.dumps(out.to_dict(), sort_keys=True) for out in results)
        mappings = sorted({json.dumps([json.loads(doc)["input"]["object_tree"], json.loads(doc)["object_species"]],
                                      sort_keys=True) for doc in got})
        assert mappings == fixture["mappings"]          # the order-independent projection
        if fixture["stable"]:
            assert got == fixture["solutions"]
        if results:
            assert {out.cost() for out in results} == {fixture["min_cost"]}
        # ANY returns one of them
        (one,) = usreconcile_extended_uspfs(srec_input, RetentionPolicy.ANY) or (None,)
        if one is not None:
            assert json.dumps(one.to_dict(), sort_keys=True) in fixture["solutions"]


@pytest.mark.parametrize("wide", [False, True])
def test_fused_decode_kernels_agree_with_the_unfused_path(ctx, monkeypatch, wide):
    """The compact fused kernel (event words from the table, per-family time stamps), the general fused
    kernel and the per-(node, root) kernels in global memory return the same cost and root for every
    resolution, also with costs that make cost() re-classify duplications as speciations."""
    if wide:
        monkeypatch.setenv("SR2_SUPERDTL_FUSED_WIDE", "1")
    for seed, costs, families in [(5, (0, 1, 1, 1, 1), 7), (6, (3, 1, 2, 1, 2), 7), (7, (2, 0, 1, 0, 3), 7),
                                  (8, (1, 1, 0, 2, 0), 7), (9, (0, 1, 1, 1, 1), 40)]:   # 40 families: two synteny words
        data = synth.make_polytomy_input(9, (4, 4), 2, families, seed=seed, costs=costs)
        srec_input, species, poly, bits, cvec = _poly_setup(ctx, data)
        total = srec_input.count_resolutions()
        outs = []
        for unfused in ("0", "1"):
            monkeypatch.setenv("SR2_SUPERDTL_UNFUSED", unfused)
            ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, cvec, 0, total)
            ctx.superdtl_solve()
            outs.append((ctx.superdtl_results(costs_only=True), ctx.superdtl_best()))
        assert (outs[0][0][0] == outs[1][0][0]).all() and (outs[0][0][1] == outs[1][0][1]).all()
        assert outs[0][1] == outs[1][1]
        # per-node results on demand materialise the resolved trees; a second solve then reads them from memory
        # instead of unranking inside the kernel and must not change anything
        monkeypatch.setenv("SR2_SUPERDTL_UNFUSED", "0")
        ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, cvec, 0, total)
        ctx.superdtl_solve()
        first = (ctx.superdtl_results(), ctx.superdtl_best())
        ctx.superdtl_solve()
        again = (ctx.superdtl_results(), ctx.superdtl_best())
        assert first[1] == again[1] == outs[0][1]
        for a, b in zip(first[0], again[0]):
            assert (a == b).all()
        # the mapping / kinds / syntenies of every instance against the oracle on a few resolutions
        from oracle.fullsize import ResolutionSet
        res = ResolutionSet(data)
        for index in (0, total // 3, total - 1):
            node_off, left, right, leaf, rbits = res.batch(index, 1)
            ref = oracle.superdtl(*res.species_args, left, right, leaf, rbits, res.costs, tables=False)
            lo, hi = index * res.nodes_per_tree, (index + 1) * res.nodes_per_tree
            assert first[0][0][index] == ref["min_cost"] and first[0][1][index] == ref["root_species"]
            assert (first[0][2][lo:hi] == ref["mapping"]).all() and (first[0][3][lo:hi] == ref["kind"]).all()
            assert (first[0][4][lo:hi] == ref["syn"]).all()


def test_base_uspfs_all_solutions_against_the_reference(golden):
    """usreconcile_base_uspfs under RetentionPolicy.ALL (the reference's own test has an instance with two
    solutions, tests/compute/test_unordered_super_reconciliation.py:210-291): the exact solution set."""
    import json
    from superrec2_b200.compute.unordered_super_reconciliation import usreconcile_base_uspfs
    fixtures = golden("base_uspfs_all.json.gz")
    assert any(len(fixture["solutions"]) > 1 for fixture in fixtures)
    for fixture in fixtures:
        srec_input = SuperReconciliationInput.from_dict(fixture["input"])
        results = usreconcile_base_uspfs(srec_input, RetentionPolicy.ALL)
        assert sorted(json.dumps(out.to_dict(), sort_keys=True) for out in results) == fixture["solutions"]
        if results:
            assert {out.cost() for out in results} == {fixture["min_cost"]}


def test_every_resolution_of_random_polytomies_against_the_oracle(ctx, golden):
    """Nested polytomies, polytomies at the root, polytomies whose children are leaves (the 60 multifurcating
    fixtures plus larger random trees): EVERY resolution's decoded minimum cost and root -- the device-side
    unranking with its per-arrangement preorder tables, shared rows, Hazard-A time stamps -- against the oracle
    run on the host-side unranking of the same resolution number."""
    import random
    from oracle.fullsize import ResolutionSet
    from oracle.make_golden_inputs import random_polytomy_instance
    inputs = [fixture["input"] for fixture in golden("superdtl_poly.json.gz")]
    rng = random.Random(4242)
    inputs += [random_polytomy_instance(rng, n_species=rng.randint(3, 9), n_objects=rng.randint(6, 11),
                                        families=rng.randint(2, 9)) for _ in range(25)]
    checked = 0
    for data in inputs:
        res = ResolutionSet(data)
        if not 1 < res.total <= 20000:
            continue
        ctx.set_species(*res.species_args)
        ctx.resolutions_upload(res.poly.child_off, res.poly.child_idx, res.poly.leaf_species, res.bits, res.costs,
                               0, res.total)
        ctx.superdtl_solve()
        min_cost, root = ctx.superdtl_results(costs_only=True)
        step = max(1, res.total // 400)            # every resolution of small inputs, a stride of large ones
        picks = list(range(0, res.total, step))
        for index in picks:
            ref_cost, ref_root = oracle.superdtl_batch(*res.species_args, *res.batch(index, 1), res.costs)
            assert (min_cost[index], root[index]) == (ref_cost[0], ref_root[0]), (res.total, index)
        checked += len(picks)
    assert checked > 3000
