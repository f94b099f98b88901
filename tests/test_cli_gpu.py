"""The command line (same flags and output as `superrec2 reconcile`) on a GPU box."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tmp_path, data, *args):
    src = tmp_path / "in.json"
    dst = tmp_path / "out.json"
    src.write_text(json.dumps(data))
    proc = subprocess.run(
        [sys.executable, "-m", "superrec2_b200.cli", "reconcile", "--input", str(src),
         "--output", str(dst), *args],
        cwd=ROOT, capture_output=True, text=True, timeout=300)
    return proc, dst


def test_superdtl_example_matches_reference_output(tmp_path, golden):
    """data/example.in.json -> byte-identical data/example.out.json, 'Minimum cost: 2' on stderr."""
    fixture = golden("superdtl_reference.json.gz")[0]
    data = {k: v for k, v in fixture["input"].items() if k != "costs"}
    proc, dst = _run(tmp_path, data, "superdtl")
    assert proc.returncode == 0, proc.stderr
    assert "Minimum cost: 2" in proc.stderr
    assert dst.read_text() == json.dumps(fixture["result"]) + "\n"


def test_dtl_alias_and_cost_flags(tmp_path, golden):
    fixture = next(f for f in golden("thl_medium.json.gz") if f["input"]["costs"]["SPECIATION"] == 0)
    costs = fixture["input"]["costs"]
    data = {k: v for k, v in fixture["input"].items() if k != "costs"}
    flags = ["--cost-dup", str(costs["DUPLICATION"]), "--cost-hgt", str(costs["HORIZONTAL_TRANSFER"]),
             "--cost-floss", str(costs["FULL_LOSS"]), "--cost-sloss", str(costs["SEGMENTAL_LOSS"])]
    for name in ("dtl", "thl"):
        proc, dst = _run(tmp_path, data, name, *flags)
        assert proc.returncode == 0, proc.stderr
        assert f"Minimum cost: {fixture['min_cost']}" in proc.stderr
        assert json.loads(dst.read_text()) == fixture["result"]


def test_superdtl_without_syntenies_is_an_error(tmp_path, golden):
    fixture = golden("thl_medium.json.gz")[0]
    data = {k: v for k, v in fixture["input"].items() if k != "costs"}
    proc, dst = _run(tmp_path, data, "superdtl")
    assert proc.returncode == 1
    assert "you need to provide leaf syntenies" in proc.stderr


def _cost_flags(costs):
    return ["--cost-spe", str(costs["SPECIATION"]), "--cost-dup", str(costs["DUPLICATION"]),
            "--cost-hgt", str(costs["HORIZONTAL_TRANSFER"]), "--cost-floss", str(costs["FULL_LOSS"]),
            "--cost-sloss", str(costs["SEGMENTAL_LOSS"])]


def test_lca_and_base_uspfs_commands(tmp_path, golden):
    """`lca` (one-argument algorithm, syntenies ignored with a warning) and `base_uspfs`."""
    fixture = next(f for f in golden("base_uspfs.json.gz")[6:] if f["result"] is not None)
    data = {k: v for k, v in fixture["input"].items() if k != "costs"}
    flags = _cost_flags(fixture["input"]["costs"])
    proc, dst = _run(tmp_path, data, "base_uspfs", *flags)
    assert proc.returncode == 0, proc.stderr
    assert f"Minimum cost: {fixture['min_cost']}" in proc.stderr
    assert json.loads(dst.read_text()) == fixture["result"]
    proc, dst = _run(tmp_path, data, "lca", *flags)
    assert proc.returncode == 0, proc.stderr
    assert "declared leaf syntenies will be ignored" in proc.stderr
    got = json.loads(dst.read_text())["object_species"]
    assert {k: v for k, v in got.items()} == fixture["lca"]


def test_all_solutions_flag(tmp_path, golden):
    """`thl --solutions all` prints one JSON document per solution, the reference's set."""
    fixture = next(f for f in golden("thl_all.json.gz") if 3 <= len(f["solutions"]) <= 40)
    data = {k: v for k, v in fixture["input"].items() if k != "costs"}
    proc, dst = _run(tmp_path, data, "thl", "--solutions", "all", *_cost_flags(fixture["input"]["costs"]))
    assert proc.returncode == 0, proc.stderr
    lines = [json.loads(line) for line in dst.read_text().splitlines()]
    got = sorted(sorted([k, v] for k, v in doc["object_species"].items()) for doc in lines)
    assert got == fixture["solutions"]


def test_superdtl_all_solutions_flag(tmp_path, golden):
    """`superdtl --solutions all` prints one JSON document per solution: the reference's set."""
    fixture = next(f for f in golden("superdtl_all.json.gz") if 3 <= len(f["solutions"]) <= 40)
    data = {k: v for k, v in fixture["input"].items() if k != "costs"}
    proc, dst = _run(tmp_path, data, "superdtl", "--solutions", "all", *_cost_flags(fixture["input"]["costs"]))
    assert proc.returncode == 0, proc.stderr
    assert f"Minimum cost: {fixture['min_cost']}" in proc.stderr
    got = sorted(json.dumps(json.loads(line), sort_keys=True) for line in dst.read_text().splitlines())
    assert got == fixture["solutions"]
