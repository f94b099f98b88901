#!/usr/bin/env python3
"""Feed reconciliation output JSON (one document per line on stdin) to the UNMODIFIED reference's
``superrec2 draw`` reader and layout (TEST INFRASTRUCTURE; build container only).

``cli/draw.py:18-29`` parses the document with ``(Super)ReconciliationOutput.from_dict`` and hands it to
``render.layout.compute`` + ``render.tikz.render``.  The only thing replaced here is the TeX call that
measures node labels (``utils/tex.py:measure``; no TeX in this image): every box gets a fixed size.
Prints one JSON line per document: the cost the reference computes from it and the size of the TikZ code.
"""
import io
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("SR2_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "standins"))
sys.path.insert(0, os.path.join(REFERENCE, "src"))
sys.setrecursionlimit(100000)

from superrec2.cli import draw  # noqa: E402
from superrec2.model.reconciliation import ReconciliationOutput, SuperReconciliationOutput  # noqa: E402
from superrec2.utils import tex  # noqa: E402

tex.measure = lambda texts, preamble="": [tex.MeasureBox(24.0, 7.0, 2.0) for _ in texts]

for line in sys.stdin:
    if not line.strip():
        continue
    data = json.loads(line)
    cls = SuperReconciliationOutput if "syntenies" in data else ReconciliationOutput
    parsed = cls.from_dict(data)
    args = types.SimpleNamespace(input=io.StringIO(line), orientation="vertical")
    code = draw.generate_tikz(args)
    cost = parsed.cost()
    print(json.dumps({"cost": None if cost == float("inf") else int(cost), "tikz_chars": len(code),
                      "same_solution": all(parsed.to_dict().get(key) == data.get(key)
                                           for key in ("object_species", "syntenies", "ordered")),
                      "same_trees": all(parsed.to_dict()["input"][key] == data["input"][key]
                                        for key in ("object_tree", "species_tree", "leaf_object_species"))}),
          flush=True)
