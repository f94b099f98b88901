"""INTEGRATION.md's stub B -- the code a reference maintainer would add -- executed against the UNMODIFIED
reference's own classes up to the ctypes boundary (``oracle/stub_b_check.py``; build container only):
the flat arrays it hands to ``sr2_set_species`` / ``sr2_dtl_run`` are the ones this package's own binding
passes for the same input, and with the device's answer it rebuilds the reference's output document."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import FlatDtl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/src"),
                                reason="the reference only exists in the build container")


def test_stub_b_passes_the_same_arrays_as_this_package(golden):
    fixtures = [f for f in golden("thl_medium.json.gz") + golden("thl_small.json.gz")[:60] if f.get("result")]
    cases = "".join(json.dumps({"input": f["input"], "mapping": f["result"]["object_species"]}) + "\n"
                    for f in fixtures)
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "stub_b_check.py")], input=cases,
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    reports = [json.loads(line) for line in proc.stdout.splitlines()]
    assert len(reports) == len(fixtures) > 30
    for fixture, report in zip(fixtures, reports):
        flat = FlatDtl(fixture["input"])
        captured = report["captured"]
        for got, want in zip(captured["species"], flat.species_args):
            assert np.array_equal(got, want)
        for got, want in zip(captured["objects"], flat.object_args):
            assert np.array_equal(got, want)
        assert captured["costs"] == flat.costs.tolist()
        assert captured["node_off"] == [0, len(flat.objects.left)] and captured["flags"] == 0
        assert report["result"] == fixture["result"] and report["cost"] == fixture["min_cost"]
