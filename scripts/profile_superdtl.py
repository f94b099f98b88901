#!/usr/bin/env python3
"""For ncu: one upload + `--solves` resident solves of BASELINE config #4 or #5 (no timing output).

    python scripts/profile_superdtl.py 4|5|1 [--solves 1]
"""
import argparse
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from superrec2_b200 import _lib, synth  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("config", type=int, choices=[1, 4, 5])
    parser.add_argument("--solves", type=int, default=1)
    args = parser.parse_args()
    ctx = _lib.default_context(0)
    if args.config == 4:
        batch = synth.make_batch(500, 1000, 1, seed=4000, n_families=32)
        ctx.set_species(batch.parent, batch.child0, batch.child1)
        ctx.superdtl_upload(batch.node_off, batch.left, batch.right, batch.leaf_species,
                            batch.leaf_family_bits, batch.costs)
    else:
        sys.path.insert(0, ROOT)
        from bench import resolution_workload
        if args.config == 5:
            data = synth.make_polytomy_input()
        else:
            with gzip.open(os.path.join(ROOT, "tests", "golden", "crispr.json.gz"), "rt") as handle:
                data = json.load(handle)[0]["input"]
        srec_input, species, poly, bits, costs = resolution_workload(data)
        ctx.set_species(species.parent, species.child0, species.child1)
        ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, 0,
                               srec_input.count_resolutions())
    for _ in range(args.solves):
        ctx.superdtl_solve()
    t = ctx.timing()
    print(f"config {args.config}: solve {t.solve_ms:.3f} ms, waves {t.waves_ms:.3f} ms, {t.n_launches} launches")


if __name__ == "__main__":
    main()
