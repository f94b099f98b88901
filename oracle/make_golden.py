#!/usr/bin/env python3
"""Generate tests/golden/*.json.gz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Runs only in the build container, where the reference
source is mounted at /root/reference (it does not exist on the GPU box; the
fixtures are what travels).  The reference's two absent third-party packages
are replaced by the stand-ins under oracle/standins/ (SURVEY.md section 8c);
under them the reference passes 54/56 of its own tests/{compute,model,utils}
(the other two need TeX / ete3's expand_polytomies) and reproduces
data/example.out.json byte for byte.

Usage:  python oracle/make_golden.py            # rewrites every fixture file

Fixtures:
  thl_small.json.gz        random dtl instances with the reference's full table
  thl_medium.json.gz       larger dtl instances, result only
  thl_reference_tests.json.gz   the instance of tests/compute/test_reconciliation.py
  superdtl_small.json.gz   random binary superdtl instances with full tables and
                           the gain/LCA sets
  superdtl_poly.json.gz    multifurcating superdtl instances, result only
  superdtl_reference.json.gz  data/example.in.json, the reference test instances
  crispr.json.gz           data/crispr-class1.in.json (superdtl, default costs)
  binarize.json.gz         enumeration order of ReconciliationInput.binarize()
  thl_all.json.gz          dtl instances with the full RetentionPolicy.ALL solution set
  superdtl_species_poly.json.gz  multifurcating species trees (and object trees), result only
  thl_internal_species.json.gz, superdtl_internal_species.json.gz
                           instances whose object leaves sit at internal species nodes, full tables
  superdtl_all.json.gz     superdtl instances with the full RetentionPolicy.ALL solution set, solved
                           three times each (`stable`: the reference's answer did not depend on the run)
  base_uspfs.json.gz       binary instances solved by usreconcile_base_uspfs (superdtl table
                           restricted to the LCA mapping) and by reconcile_lca, incl. the
                           instances of tests/compute/test_unordered_super_reconciliation.py
"""
import gzip
import io
import json
import os
import random
import sys
import contextlib

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("SR2_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "standins"))
sys.path.insert(0, os.path.join(REFERENCE, "src"))
sys.setrecursionlimit(100000)

from infinity import is_infinite  # noqa: E402
from superrec2.compute import reconciliation as ref_rec  # noqa: E402
from superrec2.compute import unordered_super_reconciliation as ref_usr  # noqa: E402
from superrec2.model.reconciliation import (  # noqa: E402
    ReconciliationInput,
    SuperReconciliationInput,
)
from superrec2.utils.dynamic_programming import RetentionPolicy  # noqa: E402

GOLDEN = os.path.join(HERE, "..", "tests", "golden")
COST_NAMES = ["SPECIATION", "DUPLICATION", "HORIZONTAL_TRANSFER", "FULL_LOSS", "SEGMENTAL_LOSS"]


def quiet(fn, *args):
    """tqdm writes progress bars to stderr; keep the generator's output readable."""
    with contextlib.redirect_stderr(io.StringIO()):
        return fn(*args)


def num(value):
    return None if is_infinite(value) else int(value)


# --------------------------------------------------------------------------
# random instances (plain `random`, names unique)
# --------------------------------------------------------------------------
def random_tree(rng, leaves, internal_prefix, name_internal=True, polytomy=0.0):
    """Random-join tree; with probability `polytomy` a join takes 3-4 roots."""
    roots = list(leaves)
    counter = 0
    while len(roots) > 1:
        arity = 2
        if polytomy and len(roots) >= 3 and rng.random() < polytomy:
            arity = min(len(roots), rng.choice([3, 3, 4]))
        picked = [roots.pop(rng.randrange(len(roots))) for _ in range(arity)]
        label = f"{internal_prefix}{counter}" if name_internal else ""
        counter += 1
        roots.append("(" + ",".join(picked) + ")" + label)
    return roots[0] + ";"


def random_costs(rng, mode):
    if mode == "default":
        return [0, 1, 1, 1, 1]
    if mode == "zero":
        return [0, 0, 0, 0, 0]
    if mode == "spe":
        return [rng.randint(1, 3), rng.randint(0, 3), rng.randint(0, 3), rng.randint(0, 2), rng.randint(0, 3)]
    return [0, rng.randint(0, 4), rng.randint(0, 4), rng.randint(0, 3), rng.randint(0, 3)]


def random_instance(rng, n_species, n_objects, surjective, cost_mode, families=0,
                    polytomy=0.0, name_internal=True):
    sleaves = [f"S{i}" for i in range(n_species)]
    oleaves = [f"g{i}" for i in range(n_objects)]
    data = {
        "species_tree": random_tree(rng, sleaves, "s"),
        "object_tree": random_tree(rng, oleaves, "o", name_internal, polytomy),
    }
    if surjective and n_objects >= n_species:
        perm = sleaves[:]
        rng.shuffle(perm)
        leaf_map = {oleaves[i]: perm[i] for i in range(n_species)}
        for i in range(n_species, n_objects):
            leaf_map[oleaves[i]] = rng.choice(sleaves)
    else:
        leaf_map = {leaf: rng.choice(sleaves) for leaf in oleaves}
    data["leaf_object_species"] = leaf_map
    data["costs"] = dict(zip(COST_NAMES, random_costs(rng, cost_mode)))
    if families:
        fams = [f"f{i}" for i in range(families)]
        syn = {}
        for leaf in oleaves:
            keep = [f for f in fams if rng.random() < 0.6]
            if not keep:
                keep = [rng.choice(fams)]
            rng.shuffle(keep)
            syn[leaf] = keep
        data["leaf_syntenies"] = syn
    return data


# --------------------------------------------------------------------------
# dtl
# --------------------------------------------------------------------------
def thl_fixture(data, with_table):
    rec_input = ReconciliationInput.from_dict(data)
    fixture = {"kind": "thl", "input": data}
    if with_table:
        table = ref_rec._compute_thl_table(rec_input, RetentionPolicy.ANY)
        dump = {}
        for obj in rec_input.object_tree.traverse():
            row = {}
            for sp in rec_input.species_lca.tree.traverse():
                entry = table[obj][sp]
                infos = list(entry.infos())
                assert len(infos) <= 1
                tag = [infos[0].left.name, infos[0].right.name] if infos else None
                if not is_infinite(entry.value()) or tag:
                    row[sp.name] = [num(entry.value()), tag]
            dump[obj.name] = row
        fixture["table"] = dump
    try:
        results = list(ref_rec.reconcile_thl(rec_input, RetentionPolicy.ANY))
    except KeyError:
        fixture["error"] = "KeyError"
        return fixture
    assert len(results) == 1
    fixture["min_cost"] = num(results[0].cost())
    fixture["result"] = results[0].to_dict()
    return fixture


# --------------------------------------------------------------------------
# superdtl
# --------------------------------------------------------------------------
def superdtl_fixture(data, with_table):
    fixture = {"kind": "superdtl", "input": data}
    srec_input = SuperReconciliationInput.from_dict(data)
    if with_table:
        # same steps as _uspfs (unordered_super_reconciliation.py:456-472) on a binary input
        work = SuperReconciliationInput.from_dict(data)
        (work_bin,) = list(work.binarize())
        work_bin.label_internal()
        gain = ref_usr._compute_gain_sets(work_bin)
        lcas = ref_usr._compute_lca_sets(work_bin, gain)
        table = quiet(
            ref_usr._compute_uspfs_table, work_bin, lcas,
            lambda species, _: species.traverse("postorder"), RetentionPolicy.ANY)
        fixture["gain_sets"] = {n.name: sorted(v) for n, v in gain.items()}
        fixture["lca_sets"] = {n.name: sorted(v) for n, v in lcas.items()}
        dump = {}
        for obj in work_bin.object_tree.traverse():
            row = {}
            for sp in work_bin.species_lca.tree.traverse():
                cell = {}
                for kind in ref_usr.SyntenyAssignment:
                    entry = table[obj][sp][kind]
                    infos = list(entry.infos())
                    assert len(infos) <= 1
                    tag = None
                    if infos:
                        tag = [[infos[0].left.species.name, infos[0].left.synteny.name],
                               [infos[0].right.species.name, infos[0].right.synteny.name]]
                    if not is_infinite(entry.value()) or tag:
                        cell[kind.name] = [num(entry.value()), tag]
                if cell:
                    row[sp.name] = cell
            dump[obj.name] = row
        fixture["table"] = dump
    results = list(quiet(ref_usr.usreconcile_extended_uspfs, srec_input, RetentionPolicy.ANY))
    if not results:
        fixture["result"] = None
        return fixture
    assert len(results) == 1
    fixture["min_cost"] = num(results[0].cost())
    fixture["result"] = results[0].to_dict()
    return fixture


def thl_all_fixture(data):
    """reconcile_thl with RetentionPolicy.ALL: the whole solution set (compared as a set)."""
    rec_input = ReconciliationInput.from_dict(data)
    fixture = {"kind": "thl_all", "input": data}
    try:
        results = list(ref_rec.reconcile_thl(rec_input, RetentionPolicy.ALL))
    except KeyError:
        fixture["error"] = "KeyError"
        return fixture
    fixture["min_cost"] = num(results[0].cost()) if results else None
    fixture["solutions"] = sorted(
        sorted((node.name, species.name) for node, species in out.object_species.items())
        for out in results)
    return fixture


def superdtl_all_fixture(data, runs=3):
    """usreconcile_extended_uspfs with RetentionPolicy.ALL.  Under ALL the reference iterates Python
    sets of tuples of TreeNode objects (hashed by address) while unioning gain sets in place (Hazard A),
    so its synteny labels -- and through cost() the kept solutions -- can depend on the run.  Every
    instance is therefore solved `runs` times on freshly parsed inputs; `stable` says whether all
    runs returned the same set of output JSON documents."""
    fixture = {"kind": "superdtl_all", "input": data}
    seen = []
    for _ in range(runs):
        srec_input = SuperReconciliationInput.from_dict(data)
        results = list(quiet(ref_usr.usreconcile_extended_uspfs, srec_input, RetentionPolicy.ALL))
        seen.append(sorted(json.dumps(out.to_dict(), sort_keys=True) for out in results))
        cost = num(results[0].cost()) if results else None
    fixture["min_cost"] = cost
    fixture["stable"] = all(item == seen[0] for item in seen)
    fixture["solutions"] = seen[0]
    # the order-independent projection: which (resolution, object -> species) assignments are optimal
    fixture["mappings"] = sorted({json.dumps([json.loads(doc)["input"]["object_tree"],
                                              json.loads(doc)["object_species"]], sort_keys=True)
                                  for doc in seen[0]})
    return fixture


def base_uspfs_fixture(data):
    """usreconcile_base_uspfs (:500-520) and reconcile_lca (compute/reconciliation.py:23-44)."""
    fixture = {"kind": "base_uspfs", "input": data}
    srec_input = SuperReconciliationInput.from_dict(data)
    lca = ref_rec.reconcile_lca(srec_input)
    fixture["lca"] = {node.name: species.name for node, species in lca.object_species.items()
                      if node.name}
    fixture["lca_cost"] = num(lca.cost())
    results = list(quiet(ref_usr.usreconcile_base_uspfs, srec_input, RetentionPolicy.ANY))
    if not results:
        fixture["result"] = None
        return fixture
    assert len(results) == 1
    fixture["min_cost"] = num(results[0].cost())
    fixture["result"] = results[0].to_dict()
    return fixture


def binarize_fixture(newick):
    """Order of ReconciliationInput.binarize() on a multifurcating object tree."""
    from ete3 import Tree
    tree = Tree(newick, format=1)
    leaves = [leaf.name for leaf in tree]
    data = {
        "object_tree": newick,
        "species_tree": "(X,Y)R;",
        "leaf_object_species": {leaf: "X" for leaf in leaves},
    }
    rec_input = ReconciliationInput.from_dict(data)
    resolved = []
    for item in rec_input.binarize():
        item.label_internal()
        resolved.append(item.object_tree.write(format=8, format_root_node=True))
    return {"kind": "binarize", "object_tree": newick, "resolutions": resolved}


def write(name, fixtures):
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name)
    raw = json.dumps(fixtures, separators=(",", ":")).encode()
    with open(path, "wb") as handle:
        with gzip.GzipFile(fileobj=handle, mode="wb", mtime=0) as gz:
            gz.write(raw)
    print(f"{name}: {len(fixtures)} fixtures, {os.path.getsize(path) / 1024:.0f} KiB", flush=True)


def main():
    only = set(sys.argv[1:])   # optional: names of the fixture files to (re)write
    global write
    if only:
        real_write = write
        write = lambda name, fixtures: real_write(name, fixtures) if name in only else None  # noqa: E731
    rng = random.Random(20261019)

    # ---- dtl ----------------------------------------------------------------
    small = []
    for i in range(420):
        n_s = rng.randint(1, 7)
        n_o = rng.randint(1, 9)
        mode = ["default", "zero", "spe", "random"][i % 4]
        surj = i % 5 != 4
        if surj:
            n_o = max(n_o, n_s)
        small.append(thl_fixture(random_instance(rng, n_s, n_o, surj, mode), True))
    write("thl_small.json.gz", small)

    medium = []
    for i in range(10):
        n_s = rng.randint(10, 16)
        n_o = rng.randint(n_s, 26)
        mode = ["default", "random", "spe"][i % 3]
        medium.append(thl_fixture(random_instance(rng, n_s, n_o, True, mode), False))
    write("thl_medium.json.gz", medium)

    ref_tests = []
    # tests/compute/test_reconciliation.py:67-123 (thl, costs 1/1/1)
    ref_tests.append(thl_fixture({
        "object_tree": "((x_1,(x_2,(y_1,z_1)5)4)3,(y_2,z_2)2)1;",
        "species_tree": "(X,(Y,Z)YZ)XYZ;",
        "leaf_object_species": {"x_1": "X", "x_2": "X", "y_1": "Y", "z_1": "Z", "y_2": "Y", "z_2": "Z"},
        "costs": dict(zip(COST_NAMES, [0, 1, 1, 1, 1])),
    }, True))
    write("thl_reference_tests.json.gz", ref_tests)

    # ---- superdtl -------------------------------------------------------------
    ssmall = []
    for i in range(380):
        n_s = rng.randint(1, 6)
        n_o = rng.randint(1, 8)
        mode = ["default", "zero", "spe", "random"][i % 4]
        fams = rng.randint(1, 6)
        ssmall.append(superdtl_fixture(
            random_instance(rng, n_s, n_o, i % 3 != 2, mode, families=fams,
                            name_internal=i % 2 == 0), True))
    write("superdtl_small.json.gz", ssmall)

    spoly = []
    for i in range(60):
        n_s = rng.randint(2, 5)
        n_o = rng.randint(3, 7)
        mode = ["default", "random", "spe"][i % 3]
        spoly.append(superdtl_fixture(
            random_instance(rng, n_s, n_o, True, mode, families=rng.randint(2, 5),
                            polytomy=0.6, name_internal=i % 2 == 0), False))
    write("superdtl_poly.json.gz", spoly)

    sref = []
    with open(os.path.join(REFERENCE, "data", "example.in.json")) as handle:
        example = json.load(handle)
    example["costs"] = dict(zip(COST_NAMES, [0, 1, 1, 1, 1]))
    fixture = superdtl_fixture(example, True)
    with open(os.path.join(REFERENCE, "data", "example.out.json")) as handle:
        assert fixture["result"] == json.load(handle), "golden example.out.json not reproduced"
    sref.append(fixture)
    # tests/compute/test_unordered_super_reconciliation.py:294-355 (test_transfers)
    sref.append(superdtl_fixture({
        "object_tree": "((x_1,y_1)1,(x_2,y_2)2)3;",
        "species_tree": "(X,Y)XY;",
        "leaf_object_species": {"x_1": "X", "y_1": "Y", "x_2": "X", "y_2": "Y"},
        "leaf_syntenies": {"x_1": list("abc"), "y_1": list("ab"), "x_2": list("bc"), "y_2": list("abc")},
        "costs": dict(zip(COST_NAMES, [0, 1, 1, 1, 1])),
    }, True))
    write("superdtl_reference.json.gz", sref)

    with open(os.path.join(REFERENCE, "data", "crispr-class1.in.json")) as handle:
        crispr = json.load(handle)
    crispr["costs"] = dict(zip(COST_NAMES, [0, 1, 1, 1, 1]))
    write("crispr.json.gz", [superdtl_fixture(crispr, False)])

    write("binarize.json.gz", [
        binarize_fixture("(a,b,c)r;"),
        binarize_fixture("(a,b,c,d)r;"),
        binarize_fixture("(a,b,c,d,e)r;"),
        binarize_fixture("((a,b,c)p,(d,e,f,g)q,h)r;"),
        binarize_fixture("((a,b)p,(c,d,e)q,(f,(g,h,i)s)t);"),
    ])

    # ---- base_uspfs / lca (own generator: the fixtures above stay byte-identical) ------
    rng2 = random.Random(20261020)
    base = []
    default = dict(zip(COST_NAMES, [0, 1, 1, 1, 1]))
    # tests/compute/test_unordered_super_reconciliation.py:99-331
    for tree, species, syn in [
        ("((x_1,y_1)2,z_1)1;", "((X,Y)XY,Z)XYZ;", {"x_1": "a", "y_1": "a", "z_1": "ab"}),
        ("((x_1,y_1)2,z_1)1;", "((X,Y)XY,Z)XYZ;", {"x_1": "ab", "y_1": "a", "z_1": "b"}),
        ("((x_1,y_1)2,z_1)1;", "((X,Y)XY,Z)XYZ;", {"x_1": "abc", "y_1": "bd", "z_1": "ace"}),
        ("((x_1,x_2)2,x_3)1;", "((X,Y)XY,Z)XYZ;", {"x_1": "ab", "x_2": "b", "x_3": "a"}),
        ("((x_1,x_2)2,(x_3,y_1)3)1;", "((X,Y)XY,Z)XYZ;", {"x_1": "abc", "x_2": "ab", "x_3": "bc", "y_1": "abc"}),
        ("((x_1,y_1)1,(x_2,y_2)2)3;", "(X,Y)XY;", {"x_1": "abc", "y_1": "ab", "x_2": "bc", "y_2": "abc"}),
    ]:
        leaves = [name for name in syn]
        base.append(base_uspfs_fixture({
            "object_tree": tree, "species_tree": species,
            "leaf_object_species": {leaf: leaf.split("_")[0].upper() for leaf in leaves},
            "leaf_syntenies": {leaf: list(fams) for leaf, fams in syn.items()},
            "costs": default,
        }))
    for i in range(120):
        n_s = rng2.randint(1, 6)
        n_o = rng2.randint(1, 9)
        mode = ["default", "zero", "spe", "random"][i % 4]
        base.append(base_uspfs_fixture(
            random_instance(rng2, n_s, n_o, i % 3 != 2, mode, families=rng2.randint(1, 6),
                            name_internal=i % 2 == 0)))
    write("base_uspfs.json.gz", base)
    # usreconcile_base_uspfs under RetentionPolicy.ALL (tests/compute/test_unordered_super_reconciliation.py:210-291
    # has an instance with two solutions): the first 60 instances above, every solution
    base_all = []
    for fixture in base[:60]:
        srec_input = SuperReconciliationInput.from_dict(fixture["input"])
        results = list(quiet(ref_usr.usreconcile_base_uspfs, srec_input, RetentionPolicy.ALL))
        base_all.append({"kind": "base_uspfs_all", "input": fixture["input"],
                         "min_cost": num(results[0].cost()) if results else None,
                         "solutions": sorted(json.dumps(out.to_dict(), sort_keys=True) for out in results)})
    write("base_uspfs_all.json.gz", base_all)

    # ---- dtl, every solution (--solutions all) ------------------------------------------
    rng3 = random.Random(20261021)
    every = [thl_all_fixture({
        "object_tree": "((x_1,(x_2,(y_1,z_1)5)4)3,(y_2,z_2)2)1;",
        "species_tree": "(X,(Y,Z)YZ)XYZ;",
        "leaf_object_species": {"x_1": "X", "x_2": "X", "y_1": "Y", "z_1": "Z", "y_2": "Y", "z_2": "Z"},
        "costs": dict(zip(COST_NAMES, [0, 1, 1, 1, 1])),
    })]
    while len(every) < 160:
        i = len(every)
        n_s = rng3.randint(1, 6)
        n_o = max(n_s, rng3.randint(1, 8))
        mode = ["default", "zero", "spe", "random"][i % 4]
        fixture = thl_all_fixture(random_instance(rng3, n_s, n_o, True, mode))
        if len(fixture.get("solutions", ())) <= 3000:
            every.append(fixture)
    write("thl_all.json.gz", every)

    # ---- object leaves mapped to INTERNAL species nodes (allowed: the leaf map takes any named node) ----
    rng4 = random.Random(20261022)
    inner_thl, inner_super = [], []
    for i in range(90):
        n_s = rng4.randint(2, 7)
        n_o = rng4.randint(2, 9)
        mode = ["default", "zero", "spe", "random"][i % 4]
        data = random_instance(rng4, n_s, n_o, False, mode, families=rng4.randint(1, 5) if i % 3 == 0 else 0)
        internal_names = [f"s{k}" for k in range(n_s - 1)]
        for leaf in list(data["leaf_object_species"]):
            if rng4.random() < 0.4:
                data["leaf_object_species"][leaf] = rng4.choice(internal_names)
        if "leaf_syntenies" in data:
            inner_super.append(superdtl_fixture(data, True))
        else:
            inner_thl.append(thl_fixture(data, True))
    # ---- multifurcating SPECIES trees (every (object, species) resolution pair is tried, :171-174) ----
    rng5 = random.Random(20261023)
    species_poly = []
    for i in range(24):
        n_s = rng5.randint(3, 5)
        n_o = rng5.randint(3, 6)
        mode = ["default", "random", "spe"][i % 3]
        data = random_instance(rng5, n_s, n_o, True, mode, families=rng5.randint(2, 4),
                               polytomy=0.5 if i % 2 else 0.0, name_internal=i % 4 < 2)
        data["species_tree"] = random_tree(rng5, [f"S{k}" for k in range(n_s)], "s", i % 4 < 2, polytomy=0.7)
        species_poly.append(superdtl_fixture(data, False))
    write("superdtl_species_poly.json.gz", species_poly)
    # ---- superdtl, every solution (--solutions all) ---------------------------------------------
    rng6 = random.Random(20261024)
    every_super = []
    # tests/compute/test_unordered_super_reconciliation.py:294-355 (test_transfers) and data/example.in.json
    every_super.append(superdtl_all_fixture({
        "object_tree": "((x_1,y_1)1,(x_2,y_2)2)3;",
        "species_tree": "(X,Y)XY;",
        "leaf_object_species": {"x_1": "X", "y_1": "Y", "x_2": "X", "y_2": "Y"},
        "leaf_syntenies": {"x_1": list("abc"), "y_1": list("ab"), "x_2": list("bc"), "y_2": list("abc")},
        "costs": dict(zip(COST_NAMES, [0, 1, 1, 1, 1])),
    }))
    every_super.append(superdtl_all_fixture(example))
    while len(every_super) < 90:
        i = len(every_super)
        n_s = rng6.randint(1, 5)
        n_o = rng6.randint(1, 7)
        mode = ["default", "zero", "spe", "random"][i % 4]
        fixture = superdtl_all_fixture(random_instance(
            rng6, n_s, n_o, i % 3 != 2, mode, families=rng6.randint(1, 5),
            polytomy=0.5 if i % 5 == 0 else 0.0, name_internal=i % 2 == 0))
        if len(fixture["solutions"]) <= 2000:
            every_super.append(fixture)
    write("superdtl_all.json.gz", every_super)
    write("thl_internal_species.json.gz", inner_thl)
    write("superdtl_internal_species.json.gz", inner_super)


if __name__ == "__main__":
    main()
