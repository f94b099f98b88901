"""The CPU oracle against fixtures produced by the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle; the GPU tests then
compare the CUDA path with the oracle and with the same fixtures."""
import pytest

from oracle import binding as oracle
from tests.helpers import FlatDtl


def _check_thl(fixture, with_table):
    flat = FlatDtl(fixture["input"])
    out = oracle.dtl(*flat.species_args, *flat.object_args, flat.costs, strict=True, tables=True)
    if with_table:
        assert flat.table_by_name(out["value"], out["tag"]) == fixture["table"]
    if fixture.get("error") == "KeyError":
        assert out["rc"] == oracle.ERR_KEYERROR
        return
    assert out["rc"] == 0
    assert out["min_cost"] == fixture["min_cost"]
    assert flat.mapping_by_name(out["mapping"]) == fixture["result"]["object_species"]
    # key order of the reference's output dict is the preorder of the object tree
    assert list(flat.mapping_by_name(out["mapping"])) == list(fixture["result"]["object_species"])


def test_thl_small_tables_and_results(golden):
    fixtures = golden("thl_small.json.gz") + golden("thl_internal_species.json.gz")
    assert len(fixtures) >= 100
    for fixture in fixtures:
        _check_thl(fixture, True)


def test_thl_medium_results(golden):
    for fixture in golden("thl_medium.json.gz"):
        _check_thl(fixture, False)


def test_thl_reference_test_instance(golden):
    # tests/compute/test_reconciliation.py:67-123 of the reference: cost 2, and the ANY
    # solution is one of the two mappings the reference test lists for ALL.
    (fixture,) = golden("thl_reference_tests.json.gz")
    _check_thl(fixture, True)
    assert fixture["min_cost"] == 2
    got = fixture["result"]["object_species"]
    assert {k: got[k] for k in "12345"} in (
        {"1": "XYZ", "2": "YZ", "3": "X", "4": "X", "5": "YZ"},
        {"1": "XYZ", "2": "YZ", "3": "X", "4": "YZ", "5": "YZ"},
    )


def test_thl_non_strict_skips_missing_roots(golden):
    """Where the reference raises KeyError the oracle's non-strict mode (the product's
    documented behaviour) still returns the best root that has an entry."""
    seen = 0
    for fixture in golden("thl_small.json.gz") + golden("thl_internal_species.json.gz"):
        if fixture.get("error") != "KeyError":
            continue
        flat = FlatDtl(fixture["input"])
        out = oracle.dtl(*flat.species_args, *flat.object_args, flat.costs, strict=False)
        assert out["rc"] == 0
        assert out["root_species"] >= 0
        seen += 1
    assert seen > 0


# ------------------------------------------------------------------ superdtl
from tests.helpers import FlatSuper  # noqa: E402
from superrec2_b200.model.reconciliation import SuperReconciliationInput  # noqa: E402


def _oracle_super(flat, tables=True):
    return oracle.superdtl(*flat.species_args, *flat.object_args, flat.leaf_bits, flat.costs, tables=tables)


def _check_superdtl_binary(fixture):
    flat = FlatSuper.from_dict(fixture["input"])
    out = _oracle_super(flat)
    assert out["rc"] == 0
    assert flat.sets_by_name(out["gain"]) == {k: sorted(v) for k, v in fixture["gain_sets"].items()}
    assert flat.sets_by_name(out["lca"]) == {k: sorted(v) for k, v in fixture["lca_sets"].items()}
    assert flat.table_by_name(out["value"], out["tag"]) == fixture["table"]
    assert out["min_cost"] == fixture["min_cost"]
    got = flat.result_dict(out["mapping"], out["syn"])
    assert got == fixture["result"]
    assert list(got["object_species"]) == list(fixture["result"]["object_species"])
    assert list(got["syntenies"]) == list(fixture["result"]["syntenies"])


def test_superdtl_small_tables_sets_and_results(golden):
    fixtures = golden("superdtl_small.json.gz") + golden("superdtl_internal_species.json.gz")
    assert len(fixtures) >= 100
    for fixture in fixtures:
        _check_superdtl_binary(fixture)


def test_superdtl_reference_instances(golden):
    # data/example.in.json -> data/example.out.json, and test_transfers
    # (tests/compute/test_unordered_super_reconciliation.py:294-355): both cost 2
    for fixture in golden("superdtl_reference.json.gz"):
        _check_superdtl_binary(fixture)
        assert fixture["min_cost"] == 2


def _solve_resolutions(data):
    """First strict minimum over (resolution, root) -- _uspfs :454-497 -- with the oracle."""
    srec_input = SuperReconciliationInput.from_dict(data)
    best = None
    for index in range(srec_input.count_resolutions()):
        binary = srec_input.resolution(index)
        binary.label_internal()
        flat = FlatSuper(binary)
        out = _oracle_super(flat, tables=False)
        assert out["rc"] == 0
        if out["root_species"] >= 0 and (best is None or out["min_cost"] < best[0]):
            best = (out["min_cost"], flat.result_dict(out["mapping"], out["syn"]))
    return best


def test_superdtl_polytomies(golden):
    for fixture in golden("superdtl_poly.json.gz") + golden("superdtl_species_poly.json.gz"):
        cost, result = _solve_resolutions(fixture["input"])
        assert cost == fixture["min_cost"]
        assert result == fixture["result"]


def test_superdtl_crispr(golden):
    """BASELINE config #1: data/crispr-class1.in.json, 27 resolutions, minimum cost 16."""
    (fixture,) = golden("crispr.json.gz")
    cost, result = _solve_resolutions(fixture["input"])
    assert cost == fixture["min_cost"] == 16
    assert result == fixture["result"]
