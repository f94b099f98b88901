"""Host-side model (no GPU): newick/JSON round trips, label_internal, resolution order,
cost() -- against the fixtures produced by the unmodified reference."""
import json
import os

import pytest

from superrec2_b200.model.reconciliation import (
    ReconciliationInput,
    ReconciliationOutput,
    SuperReconciliationInput,
    SuperReconciliationOutput,
)
from superrec2_b200.model.synteny import sort_synteny
from superrec2_b200.model.tree import Tree, TreeError
from superrec2_b200.model.tree_mapping import get_species_mapping
from superrec2_b200.utils.trees import LowestCommonAncestor, count_resolutions, is_binary


def test_newick_roundtrip_with_nhx_and_odd_names():
    text = ("((A b,(C+d,E)F g)H[&&NHX:color=0037DB],I-1)R;")
    tree = Tree(text, format=1)
    assert [n.name for n in tree.traverse()] == ["R", "H", "I-1", "A b", "F g", "C+d", "E"]
    assert (tree & "H").color == "0037DB"
    assert tree.write(format=8, format_root_node=True, features=["color"]) == text
    assert tree.write(format=9) == "((A b,(C+d,E)),I-1);"
    with pytest.raises(TreeError):
        Tree("((a,b);", format=1)


def test_multiline_newick_and_leaf_iteration():
    tree = Tree("""
        (
            ((x_1,z_1)3,(w_1,w_2)4)2,
            (y_1, t_1)5
        )1;
    """, format=1)
    assert [leaf.name for leaf in tree] == ["x_1", "z_1", "w_1", "w_2", "y_1", "t_1"]
    assert len(tree) == 6 and "4" in tree and "nope" not in tree


def test_natural_sort_of_families():
    assert sort_synteny({"10", "2", "3'", "3\"", "1", "11"}) == ["1", "2", "3\"", "3'", "10", "11"]
    assert sort_synteny(["f10", "f2", "f1"]) == ["f1", "f2", "f10"]


def test_species_mapping_from_names():
    objects = Tree("((x_1,xy_2),Y_3,z);", format=1)
    species = Tree("(X,(Y,X_Y)P);", format=1)
    mapping = get_species_mapping(objects, species)
    assert {k.name: v.name for k, v in mapping.items()} == {"x_1": "X", "Y_3": "Y"}


def test_lca_queries():
    tree = Tree("(((X,Y)XY,Z)XYZ,(W,T)WT)R;", format=1)
    lca = LowestCommonAncestor(tree)
    node = lambda name: tree & name  # noqa: E731
    assert lca(node("X"), node("Z")) is node("XYZ")
    assert lca(node("X"), node("Y"), node("T")) is tree
    assert lca.is_ancestor_of(node("XYZ"), node("Y")) and lca.is_ancestor_of(node("Y"), node("Y"))
    assert not lca.is_strict_ancestor_of(node("Y"), node("Y"))
    assert lca.is_comparable(node("X"), node("XYZ")) and not lca.is_comparable(node("X"), node("W"))
    assert lca.level(node("X")) == 3 and lca.distance(node("X"), node("W")) == 5


def test_resolution_order_matches_reference_binarize(golden):
    """Resolution number r of unrank_resolution == r-th element of the reference's binarize()."""
    for fixture in golden("binarize.json.gz"):
        data = {
            "object_tree": fixture["object_tree"],
            "species_tree": "(X,Y)R;",
            "leaf_object_species": {leaf.name: "X" for leaf in Tree(fixture["object_tree"], format=1)},
        }
        rec_input = ReconciliationInput.from_dict(data)
        assert rec_input.count_resolutions() == len(fixture["resolutions"])
        got = []
        for item in rec_input.binarize():
            assert is_binary(item.object_tree)
            item.label_internal()
            got.append(item.object_tree.write(format=8, format_root_node=True))
        assert got == fixture["resolutions"]
    assert count_resolutions(Tree("((a,b,c,d,e),(f,g,h,i,j,k));", format=1)) == 105 * 945


def test_cost_and_json_of_reference_results(golden):
    """Re-reading the reference's output JSON gives the same JSON back and the same cost()."""
    for fixture in golden("thl_small.json.gz") + golden("thl_medium.json.gz"):
        if fixture.get("error"):
            continue
        out = ReconciliationOutput.from_dict(fixture["result"])
        assert out.to_dict() == fixture["result"]
        assert out.cost() == fixture["min_cost"]
    for name in ("superdtl_small.json.gz", "superdtl_poly.json.gz", "crispr.json.gz"):
        for fixture in golden(name):
            out = SuperReconciliationOutput.from_dict(fixture["result"])
            assert json.dumps(out.to_dict()) == json.dumps(fixture["result"])
            assert out.cost() == fixture["min_cost"]


def test_label_internal_skips_used_names():
    data = {
        "object_tree": "((a,b),(O0,c)O1);",
        "species_tree": "(X,Y);",
        "leaf_object_species": {"a": "X", "b": "X", "O0": "Y", "c": "Y"},
        "leaf_syntenies": {"a": ["g"], "b": ["g"], "O0": ["g"], "c": ["g"]},
    }
    item = SuperReconciliationInput.from_dict(data)
    item.label_internal()
    assert item.object_tree.write() == "((a,b)O3,(O0,c)O1)O2;"
    assert item.species_lca.tree.write() == "(X,Y)S0;"


def test_library_unranking_equals_host_unranking(golden):
    """sr2_resolutions_tree (C++, what the GPU batches are built from) builds the same trees,
    with the same child order, as the Python unranker that is pinned to the reference order."""
    from superrec2_b200 import _lib
    from superrec2_b200.compute.flatten import flatten_polytomies
    from superrec2_b200.utils.trees import unrank_resolution

    def nested_host(node):
        return node.name if node.is_leaf() else tuple(nested_host(kid) for kid in node.children)

    def nested_lib(left, right, origin, names, at=0):
        if left[at] < 0:
            return names[origin[at]]
        return (nested_lib(left, right, origin, names, left[at]),
                nested_lib(left, right, origin, names, right[at]))

    trees = [f["object_tree"] for f in golden("binarize.json.gz")] + ["((a,b,c,d,e)p,(f,g,h)q,(i,j));"]
    for newick in trees:
        tree = Tree(newick, format=1)
        layout = flatten_polytomies(tree, {leaf: 0 for leaf in tree}, {0: 0})
        names = [node.name for node in layout.nodes]
        total, binary_nodes = _lib.resolutions_count(layout.child_off, layout.child_idx)
        assert total == count_resolutions(tree) and binary_nodes == 2 * len(tree) - 1
        step = max(1, total // 400)
        for index in list(range(0, total, step)) + [total - 1]:
            left, right, origin = _lib.resolutions_tree(layout.child_off, layout.child_idx, index)
            assert nested_lib(left, right, origin, names) == nested_host(unrank_resolution(tree, index))
            # ABI numbering rule: parents before children
            assert all(left[u] < 0 or (left[u] > u and right[u] > u) for u in range(len(left)))


def test_entry_point_signatures_satisfy_the_cli_introspection():
    """cli/reconcile.py:79-100 of the reference compares the ANNOTATIONS of the callable with the
    input class and RetentionPolicy and dedents its docstring: they must be real objects."""
    import inspect
    from superrec2_b200.compute.reconciliation import reconcile_thl
    from superrec2_b200.compute.unordered_super_reconciliation import usreconcile_extended_uspfs
    from superrec2_b200.utils.dynamic_programming import RetentionPolicy
    for fn, cls in ((reconcile_thl, ReconciliationInput), (usreconcile_extended_uspfs, SuperReconciliationInput)):
        params = list(inspect.signature(fn).parameters.values())
        assert len(params) == 2
        assert params[0].annotation is cls and params[1].annotation is RetentionPolicy
        assert fn.__doc__ and fn.__doc__.strip()


def test_reconcile_lca_against_the_reference(golden):
    """reconcile_lca (host only, reference compute/reconciliation.py:23-44): same mapping and
    cost() as the reference on the base_uspfs fixtures."""
    from superrec2_b200.compute.reconciliation import reconcile_lca
    from superrec2_b200.model.reconciliation import SuperReconciliationInput
    for fixture in golden("base_uspfs.json.gz"):
        rec_input = SuperReconciliationInput.from_dict(fixture["input"])
        output = reconcile_lca(rec_input)
        got = {node.name: species.name for node, species in output.object_species.items() if node.name}
        assert got == fixture["lca"]
        cost = output.cost()
        assert (None if cost == float("inf") else cost) == fixture["lca_cost"]


def _conforming_outputs(golden):
    """(fixture result JSON, expected cost) of instances whose leaves follow the reference's
    ``<species>_<suffix>`` naming, which its ``draw`` layout requires (render/layout.py:124)."""
    fixtures = (golden("crispr.json.gz") + golden("superdtl_reference.json.gz") +
                golden("thl_reference_tests.json.gz") + golden("base_uspfs.json.gz")[:6])
    return [(f["result"], f["min_cost"]) for f in fixtures if f.get("result")]


def test_emitted_json_round_trips_through_the_output_readers(golden):
    """the documents this package writes are what ``superrec2 draw`` reads
    (``cli/draw.py:18-29`` -> ``(Super)ReconciliationOutput.from_dict``): every key the reader needs is
    there, the parsed object re-serialises to the same document and has the printed cost."""
    from superrec2_b200.model.reconciliation import ReconciliationOutput, SuperReconciliationOutput
    outputs = _conforming_outputs(golden) + [
        (f["result"], f["min_cost"]) for f in golden("thl_medium.json.gz") + golden("superdtl_poly.json.gz")
        if f.get("result")]
    assert len(outputs) > 40
    for document, cost in outputs:
        assert {"input", "object_species"} <= set(document)
        assert {"object_tree", "species_tree", "leaf_object_species", "costs"} <= set(document["input"])
        cls = SuperReconciliationOutput if "syntenies" in document else ReconciliationOutput
        parsed = cls.from_dict(json.loads(json.dumps(document)))
        assert parsed.to_dict() == document
        assert list(parsed.to_dict()["object_species"]) == list(document["object_species"])
        assert parsed.cost() == cost


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference only exists in the build container")
def test_emitted_json_is_drawn_by_the_reference(golden):
    """The UNMODIFIED reference's ``draw`` pipeline (reader, layout, TikZ renderer; only the TeX call that
    measures labels is stubbed, ``oracle/draw_check.py``) accepts documents serialised by THIS package's
    output classes and sees the same solution and the same cost."""
    import subprocess
    import sys
    from superrec2_b200.model.reconciliation import ReconciliationOutput, SuperReconciliationOutput
    outputs = _conforming_outputs(golden)
    lines = []
    for document, _ in outputs:
        cls = SuperReconciliationOutput if "syntenies" in document else ReconciliationOutput
        lines.append(json.dumps(cls.from_dict(document).to_dict()))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, os.path.join(root, "oracle", "draw_check.py")],
                          input="\n".join(lines) + "\n", capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    reports = [json.loads(line) for line in proc.stdout.splitlines()]
    assert len(reports) == len(outputs)
    for report, (_, cost) in zip(reports, outputs):
        assert report["cost"] == cost and report["same_solution"] and report["same_trees"]
        assert report["tikz_chars"] > 500
