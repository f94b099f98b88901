#!/usr/bin/env python3
"""Execute INTEGRATION.md's stub B (the code a reference maintainer would add) up to the ctypes
boundary (TEST INFRASTRUCTURE; build container only: it imports the UNMODIFIED reference under the
stand-ins for ete3 / infinity).

The python blocks of INTEGRATION.md section B are extracted and executed with ``ctypes.CDLL`` replaced
by a recorder: ``sr2_set_species`` and ``sr2_dtl_run`` capture the flat arrays they are given, and
``sr2_dtl_run`` answers with the mapping passed on stdin (the GPU's / fixture's answer), so the stub's
return value can be compared with the reference's own output.  stdin: one JSON line
``{"input": <reconcile input>, "mapping": {object name: species name}}`` per case; stdout: one JSON line
per case with the captured arrays and the stub's ``to_dict()``.
"""
import ctypes
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("SR2_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "standins"))
sys.path.insert(0, os.path.join(REFERENCE, "src"))
sys.setrecursionlimit(100000)

from superrec2.model.reconciliation import (  # noqa: E402
    EdgeEvent, NodeEvent, ReconciliationInput, ReconciliationOutput)
from superrec2.utils.dynamic_programming import RetentionPolicy  # noqa: E402

captured = {}
answer = {}


def as_array(pointer, n, dtype):
    return np.ctypeslib.as_array(ctypes.cast(pointer, ctypes.POINTER(dtype)), shape=(n,))


class FakeFunction:
    def __init__(self, name):
        self.name = name
        self.restype = ctypes.c_int

    def __call__(self, *args):
        if self.name == "sr2_create":
            return 0
        if self.name == "sr2_last_error":
            return b""
        if self.name == "sr2_set_species":
            _, S, parent, c0, c1 = args
            captured["species"] = [as_array(p, S, ctypes.c_int32).tolist() for p in (parent, c0, c1)]
            return 0
        if self.name == "sr2_dtl_run":
            _, B, off, left, right, leaf, costs, flags, out_cost, out_root, out_map = args
            G = int(as_array(off, B + 1, ctypes.c_int64)[B])
            captured["objects"] = [as_array(p, G, ctypes.c_int32).tolist() for p in (left, right, leaf)]
            captured["costs"] = as_array(costs, 5, ctypes.c_int32).tolist()
            captured["flags"], captured["node_off"] = int(flags), as_array(off, B + 1, ctypes.c_int64).tolist()
            as_array(out_map, G, ctypes.c_int32)[:] = answer["ranks"](G)
            as_array(out_cost, 1, ctypes.c_int32)[0] = 0
            as_array(out_root, 1, ctypes.c_int32)[0] = int(as_array(out_map, G, ctypes.c_int32)[0])
            return 0
        raise AssertionError(f"stub B called an entry point the header does not have: {self.name}")


class FakeLibrary:
    def __getattr__(self, name):
        fn = FakeFunction(name)
        setattr(self, name, fn)
        return fn


with open(os.path.join(HERE, "..", "INTEGRATION.md")) as handle:
    text = handle.read()
section = text[text.index("## B."):text.index("## C ABI summary")]
blocks = re.findall(r"```python\n(.*?)```", section, re.S)
assert len(blocks) == 2, "INTEGRATION.md section B: expected the _sr2.py block and the reconcile_thl block"
real_cdll = ctypes.CDLL
ctypes.CDLL = lambda *a, **k: FakeLibrary()
namespace = {"ReconciliationInput": ReconciliationInput, "ReconciliationOutput": ReconciliationOutput,
             "RetentionPolicy": RetentionPolicy, "NodeEvent": NodeEvent, "EdgeEvent": EdgeEvent,
             "Set": set}
try:
    exec(compile(blocks[0], "INTEGRATION.md:_sr2.py", "exec"), namespace)
    exec(compile(blocks[1], "INTEGRATION.md:reconcile_thl", "exec"), namespace)
finally:
    ctypes.CDLL = real_cdll

for line in sys.stdin:
    if not line.strip():
        continue
    case = json.loads(line)
    rec_input = ReconciliationInput.from_dict(case["input"])
    species_nodes = list(rec_input.species_lca.tree.traverse())
    rank = {node.name: i for i, node in enumerate(species_nodes)}
    object_nodes = list(rec_input.object_tree.traverse("preorder"))
    answer["ranks"] = lambda G: [rank[case["mapping"][node.name]] for node in object_nodes]
    captured.clear()
    (result,) = namespace["reconcile_thl"](rec_input, RetentionPolicy.ANY)
    print(json.dumps({"captured": captured, "result": result.to_dict(), "cost": int(result.cost())}), flush=True)
