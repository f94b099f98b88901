#!/usr/bin/env python3
"""Randomised differential stress test on a GPU box (not part of the test-suite; minutes, not seconds):
random batch shapes, costs and paths through the library against the CPU oracle.

    python scripts/stress_parity.py [seconds]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import binding as oracle  # noqa: E402
from oracle.fullsize import ResolutionSet  # noqa: E402
from superrec2_b200 import _lib, synth  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    rng = np.random.default_rng(20261020)
    ctx = _lib.default_context(0)
    start = time.time()
    counts = {"dtl": 0, "dtl16": 0, "superdtl": 0, "resolutions": 0}
    while time.time() - start < budget:
        kind = rng.choice(["dtl", "dtl16", "superdtl", "resolutions"])
        costs = tuple(int(c) for c in rng.choice([(0, 1, 1, 1, 1), (0, 2, 3, 1, 2), (1, 1, 2, 0, 1), (0, 0, 0, 0, 0),
                                                   (2, 5, 3, 2, 4), (0, 1, 1, 0, 0)]))
        seed = int(rng.integers(1, 1 << 30))
        if kind in ("dtl", "dtl16"):
            n_s = int(rng.integers(1, 70))
            n_o = n_s + int(rng.integers(0, 60))
            n_trees = int(rng.choice([1, 3, 40, 1200, 2500]))
            batch = synth.make_batch(n_s, n_o, n_trees, seed, costs=costs)
            if rng.random() < 0.5:   # object leaves at internal species too (the reference allows any species node)
                leaves = np.flatnonzero(batch.left < 0)
                moved = leaves[rng.random(len(leaves)) < 0.3]
                batch.leaf_species[moved] = rng.integers(0, batch.n_species, len(moved))
            # tier limits of the thread-per-row kernel: the default, or cuts at odd sizes
            tiers = str(rng.choice(["", "5,9,14,64", "64", "7,8,30,31,50"]))
            os.environ.pop("SR2_THREAD_TIERS", None)
            if tiers:
                os.environ["SR2_THREAD_TIERS"] = tiers
            os.environ["SR2_DTL_CHAIN_ROWS"] = str(rng.choice(["300", "19000"]))
            ctx.set_species(batch.parent, batch.child0, batch.child1)
            args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
            got = ctx.dtl_run16(*args) if kind == "dtl16" else ctx.dtl_run(*args)
            want = oracle.dtl_batch(batch.parent, batch.child0, batch.child1, *args, threads=os.cpu_count())
            ok = all(np.array_equal(np.asarray(g, np.int32), w) for g, w in zip(got, want))
            what = f"{kind} n_s={n_s} n_o={n_o} trees={n_trees} costs={costs} seed={seed} tiers={tiers!r}"
        elif kind == "superdtl":
            n_s = int(rng.integers(1, 40))
            n_o = n_s + int(rng.integers(0, 30))
            fams = int(rng.integers(1, 45))
            n_trees = int(rng.choice([1, 5, 60]))
            batch = synth.make_batch(n_s, n_o, n_trees, seed, costs=costs, n_families=fams)
            ctx.set_species(batch.parent, batch.child0, batch.child1)
            got = ctx.superdtl_run(batch.node_off, batch.left, batch.right, batch.leaf_species,
                                   batch.leaf_family_bits, batch.costs)
            want = oracle.superdtl_batch(batch.parent, batch.child0, batch.child1, batch.node_off, batch.left,
                                         batch.right, batch.leaf_species, batch.leaf_family_bits, batch.costs,
                                         want_nodes=True, threads=os.cpu_count())
            ok = all(np.array_equal(g, w) for g, w in zip(got, want))
            what = f"superdtl n_s={n_s} n_o={n_o} fams={fams} trees={n_trees} costs={costs} seed={seed}"
        else:
            arities = tuple(int(a) for a in rng.choice([(3, 3), (4, 3), (4, 4), (5, 3), (3, 5), (5, 4)]))
            sub = int(rng.integers(1, 4))
            fams = int(rng.integers(2, 40))
            n_s = int(rng.integers(4, 25))
            data = synth.make_polytomy_input(n_s, arities, sub, fams, seed=seed, costs=costs)
            res = ResolutionSet(data)
            ctx.set_species(*res.species_args)
            ctx.resolutions_upload(res.poly.child_off, res.poly.child_idx, res.poly.leaf_species, res.bits,
                                   res.costs, 0, res.total)
            ctx.superdtl_solve()
            got = ctx.superdtl_results(costs_only=True)
            picks = sorted(set(int(i) for i in rng.integers(0, res.total, 40)))
            ok = True
            for index in picks:
                want = oracle.superdtl_batch(*res.species_args, *res.batch(index, 1), res.costs)
                ok &= (got[0][index], got[1][index]) == (want[0][0], want[1][0])
            what = f"resolutions arities={arities} sub={sub} fams={fams} n_s={n_s} costs={costs} seed={seed}"
        counts[kind] += 1
        if not ok:
            print("MISMATCH:", what, flush=True)
            return 1
    print(f"stress ok after {time.time() - start:.0f} s: {counts}", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
