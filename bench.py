#!/usr/bin/env python3
"""bench.py -- DP cells/s of the reconciliation hot path (``dtl`` / ``superdtl``) on B200.

Headline workload (BASELINE.json configs[1]): ``dtl`` on a synthetic batch of 10,000 random
binary gene trees (500 leaves) against one 200-leaf species tree, default costs.  One *step* = one
pass of the hot path over the whole batch: wavefront DP over every (internal object node, species)
cell, selection of the returned root and traceback of the mapping, for every tree.

  value     cells/s with the batch already resident in HBM (sr2_dtl_solve)
  e2e       the same through the C-ABI call a binding makes (sr2_dtl_run): host buffers in, host
            buffers out, H2D/D2H copies and bucketing inside the timed region
  roofline  dominant kernels (the dtl wave kernels): algorithmic bytes (20 B per cell, SURVEY.md
            section 8d) / CUDA-event time of the wave launches, against the measured HBM copy
            bandwidth in MEASURED_PEAKS.json; ``frac_evaluated`` counts only the cells of rows
            that are materialised, ``dram_frac`` the bytes ncu saw cross the HBM interface
  cpu_baseline  the CPU oracle port (plain C restatement of the reference's O(|G||S|^2) algorithm,
            OpenMP over trees) on a bounded sample of the same workload, timed on this box's host
            cores (rank 0, N=1 only); its answers are compared with the GPU's
  secondary the other BASELINE.json configurations, each with its own time, rate, roofline, e2e
            and CPU baseline: #1 superdtl on crispr-class1, #3 dtl 5,000 x 2,000 leaves (target:
            under 100 ms), #4 superdtl 1,000 x 500 leaves x 32 families, #5 superdtl over the
            99,225 resolutions of a multifurcating tree.  At N > 1 config #5 is STRONG-scaled:
            resolution ranges per rank, one NCCL all-reduce(MIN) of the packed key.

Multi-GPU (torchrun, one rank per GPU): instances of the headline batch are independent, so every
rank gets its own 10,000-tree batch (weak scaling), there is no data-path collective, and
value = cells of all ranks / max-over-ranks time.

``--impl reference`` times the reference algorithm's CPU port with all host threads on a bounded
sample per step (the reference itself is pure Python and cannot travel to the GPU box).
"""
import argparse
import gzip
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (species leaves, object leaves, trees per GPU, seed)
    "dtl_batch_10000x500_vs_200": (200, 500, 10000, 2000),
    "dtl_batch_small": (50, 80, 512, 2000),
}
BYTES_PER_CELL = 20  # 2 x 4 B child reads + 4 B value + 8 B traceback
# DRAM bytes per evaluated cell, measured: ncu dram__bytes_read.sum + dram__bytes_write.sum of EVERY launch of
# one step of the headline workload (profiles/r7h_dram_per_launch_config2.csv: 6.57 GB per step, 4.59 GB of
# it in the 66 wave launches; 17.59 GB before the sparse rows were stored compact, r4g) / cells of the rows that
# are materialised (1,325,775,654 on rank 0's batch)
DRAM_B_PER_CELL_STEP, DRAM_B_PER_CELL_WAVES = 6.572e9 / 1325775654, 4.586e9 / 1325775654


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as handle:
            return float(json.load(handle)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


class ClockSampler:
    """Polls SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
        0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._dev, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._dev)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self._nvml:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def host_cores():
    return os.cpu_count() or 1


def cpu_sample(batch, n_trees):
    """First n_trees instances of the batch, for the CPU port."""
    hi = int(batch.node_off[n_trees])
    return (batch.parent, batch.child0, batch.child1, batch.node_off[:n_trees + 1],
            batch.left[:hi], batch.right[:hi], batch.leaf_species[:hi], batch.costs)


def time_cpu_port(batch, n_trees):
    from oracle import binding as oracle
    args = cpu_sample(batch, n_trees)
    start = time.perf_counter()
    answers = oracle.dtl_batch(*args, want_map=True, threads=host_cores())  # torchrun exports OMP_NUM_THREADS=1
    seconds = time.perf_counter() - start
    cells = int((args[4] >= 0).sum()) * batch.n_species
    return cells / seconds, seconds, answers


def require_equal(what, got, want):
    """The CPU legs are also checkers: a GPU answer that differs from the oracle's stops the bench."""
    if not all(np.array_equal(np.asarray(g), np.asarray(w)) for g, w in zip(got, want)):
        raise SystemExit(f"bench.py: PARITY FAILURE in {what}: the GPU result differs from the CPU oracle")
    return "equal to the CPU oracle on the sample (min cost, root species, mapping)"


def run_reference(args, workload, rank, world):
    """--impl reference: the CPU port of the reference algorithm, all host threads."""
    if rank != 0:
        return
    from superrec2_b200 import synth
    n_s, n_o, n_trees, seed = WORKLOADS[workload]
    cores = host_cores()
    sample = min(n_trees, max(16, 2 * cores))
    batch = synth.make_batch(n_s, n_o, sample, seed)
    for _ in range(min(args.warmup, 1)):
        time_cpu_port(batch, sample)
    start = time.perf_counter()
    for _ in range(args.steps):
        time_cpu_port(batch, sample)
    seconds = time.perf_counter() - start
    cells = batch.cells * args.steps
    value = cells / seconds
    line = {
        "impl": "reference", "metric": "dp_cells_per_s", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": seconds / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload, "species_leaves": n_s, "object_leaves": n_o,
                   "trees_per_step": sample, "costs": [0, 1, 1, 1, 1]},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} of the workload's trees per step, OpenMP over trees; "
                                   "plain-C port of the reference's O(|G||S|^2) algorithm "
                                   "(the reference is pure Python and is not on this box)"},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# secondary configurations (BASELINE.json configs #1, #3, #4, #5)
# ----------------------------------------------------------------------------------------------
def roofline_entry(cells_evaluated, waves_ms, note):
    peak, peak_src = measured_peak()
    achieved = cells_evaluated * BYTES_PER_CELL / (waves_ms * 1e-3) / 1e9 if waves_ms > 0 else 0.0
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "cells_evaluated": int(cells_evaluated), "bytes_per_cell": BYTES_PER_CELL,
            "wave_ms": waves_ms, "peak_source": peak_src, "note": note}


def repeat_solve(solve, timing, repeats):
    """Median CUDA-event times (whole solve, wave kernels) and launch counts of `repeats` solves."""
    for _ in range(3):
        solve()
    solve_ms, waves_ms = [], []
    for _ in range(repeats):
        solve()
        t = timing()
        solve_ms.append(t.solve_ms)
        waves_ms.append(t.waves_ms)
    t = timing()
    return float(np.median(solve_ms)), float(np.median(waves_ms)), int(t.n_launches), int(t.n_waves), int(t.n_cells)


def repeat_wall(call, repeats):
    call()
    times = []
    for _ in range(repeats):
        start = time.perf_counter()
        out = call()
        times.append((time.perf_counter() - start) * 1e3)
    return out, float(np.median(times)), float(min(times))


def secondary_dtl_single(ctx, repeats, with_cpu):
    """Config #3: dtl, one 5,000-leaf object tree vs a 2,000-leaf species tree (seed 3000)."""
    from superrec2_b200 import synth
    batch = synth.make_batch(2000, 5000, 1, seed=3000)
    cells = batch.cells
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    ctx.dtl_upload(*args)
    solve_ms, waves_ms, launches, waves, _ = repeat_solve(ctx.dtl_solve, ctx.timing, repeats)
    got, e2e_ms, e2e_best = repeat_wall(lambda: ctx.dtl_run(*args), repeats)
    nbytes = lambda *arrays: int(sum(a.nbytes for a in arrays))  # noqa: E731
    entry = {
        "config": 3, "algorithm": "dtl (reconcile_thl)", "species_leaves": 2000, "object_leaves": 5000,
        "cells": cells, "device_ms": solve_ms, "value": cells / (solve_ms * 1e-3), "unit": "cells/s",
        "instances_per_s": 1e3 / solve_ms, "launches": launches, "wave_launches": waves,
        "min_cost": int(got[0][0]),
        "e2e": {"ms": e2e_ms, "ms_best": e2e_best, "value": cells / (e2e_ms * 1e-3), "unit": "cells/s",
                "api": "sr2_dtl_run, pageable host arrays in and out",
                "h2d_bytes": nbytes(batch.left, batch.right, batch.leaf_species, batch.node_off),
                "d2h_bytes": nbytes(*got)},
        "target_ms": 100.0, "meets_target": bool(e2e_ms < 100.0),
        "roofline": roofline_entry(cells, waves_ms,
                                   f"one instance: {waves} dependent wave launches, latency-bound "
                                   "(the tables are 160 MB; the HBM floor of 0.4 GB is 62 us)"),
    }
    def cpu_leg():
        from oracle import binding as oracle
        oracle.load().oracle_set_threads(host_cores())
        start = time.perf_counter()
        ref = oracle.dtl(batch.parent, batch.child0, batch.child1, batch.left, batch.right, batch.leaf_species,
                         batch.costs, strict=False, tables=False)
        seconds = time.perf_counter() - start
        entry["cpu_baseline"] = {
            "value": cells / seconds, "unit": "cells/s", "cores": host_cores(), "kind": "port",
            "sample": f"the whole instance, OpenMP over the species of a node, {seconds:.1f} s wall",
            "parity": require_equal("config #3", (got[0][0], got[1][0], got[2]),
                                    (ref["min_cost"], ref["root_species"], ref["mapping"]))}
    return entry, (cpu_leg if with_cpu else None)


def secondary_superdtl_single(ctx, repeats, with_cpu):
    """Config #4: superdtl, one 1,000-leaf synteny tree, 32 families, 500-leaf species tree (seed 4000)."""
    from superrec2_b200 import synth
    batch = synth.make_batch(500, 1000, 1, seed=4000, n_families=32)
    kinds = batch.cells * 2
    args = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.leaf_family_bits, batch.costs)
    ctx.set_species(batch.parent, batch.child0, batch.child1)
    ctx.superdtl_upload(*args)
    solve_ms, waves_ms, launches, waves, _ = repeat_solve(ctx.superdtl_solve, ctx.timing, repeats)
    got, e2e_ms, e2e_best = repeat_wall(lambda: ctx.superdtl_run(*args), repeats)
    nbytes = lambda *arrays: int(sum(a.nbytes for a in arrays))  # noqa: E731
    entry = {
        "config": 4, "algorithm": "superdtl (usreconcile_extended_uspfs)", "species_leaves": 500,
        "object_leaves": 1000, "families": 32, "cell_kinds": kinds, "device_ms": solve_ms,
        "value": kinds / (solve_ms * 1e-3), "unit": "cell-kinds/s", "instances_per_s": 1e3 / solve_ms,
        "launches": launches, "wave_launches": waves, "min_cost": int(got[0][0]),
        "e2e": {"ms": e2e_ms, "ms_best": e2e_best, "value": kinds / (e2e_ms * 1e-3), "unit": "cell-kinds/s",
                "api": "sr2_superdtl_run, pageable host arrays in and out",
                "h2d_bytes": nbytes(batch.left, batch.right, batch.leaf_species, batch.leaf_family_bits),
                "d2h_bytes": nbytes(*got)},
        "roofline": roofline_entry(kinds, waves_ms,
                                   f"one instance: {waves} dependent wave launches, latency-bound "
                                   "(HBM floor of 0.04 GB: 6 us)"),
    }
    def cpu_leg():
        from oracle import binding as oracle
        oracle.load().oracle_set_threads(host_cores())
        start = time.perf_counter()
        ref = oracle.superdtl(batch.parent, batch.child0, batch.child1, batch.left, batch.right,
                              batch.leaf_species, batch.leaf_family_bits, batch.costs, tables=False)
        seconds = time.perf_counter() - start
        entry["cpu_baseline"] = {
            "value": kinds / seconds, "unit": "cell-kinds/s", "cores": host_cores(), "kind": "port",
            "sample": f"the whole instance, OpenMP over the species of a node, {seconds:.1f} s wall",
            "parity": require_equal("config #4", (got[0][0], got[1][0], got[2], got[3], got[4]),
                                    (ref["min_cost"], ref["root_species"], ref["mapping"], ref["kind"], ref["syn"]))
            .replace("mapping)", "mapping, kinds, syntenies)")}
    return entry, (cpu_leg if with_cpu else None)


def resolution_workload(data):
    """Flat form of a multifurcating input: what sr2_resolutions_upload takes."""
    from superrec2_b200.compute.flatten import flatten_polytomies, flatten_species
    from superrec2_b200.compute.reconciliation import cost_vector
    from superrec2_b200.compute.unordered_super_reconciliation import family_index, leaf_bitsets
    from superrec2_b200.model.reconciliation import SuperReconciliationInput
    srec_input = SuperReconciliationInput.from_dict(data)
    species = flatten_species(srec_input.species_lca.tree)
    poly = flatten_polytomies(srec_input.object_tree, srec_input.leaf_object_species, species.rank)
    families = family_index(srec_input.leaf_syntenies)
    bits = leaf_bitsets(poly.nodes, srec_input.leaf_syntenies, families)
    return srec_input, species, poly, bits, cost_vector(srec_input.costs)


def cpu_resolutions(data, first, count):
    """The CPU oracle on resolutions [first, first+count): (rate inputs, per-resolution answers)."""
    from oracle import binding as oracle
    from oracle.fullsize import ResolutionSet
    res = ResolutionSet(data)
    flat = res.batch(first, count)          # unranking is not part of the timed region
    start = time.perf_counter()
    min_cost, root = oracle.superdtl_batch(*res.species_args, *flat, res.costs, threads=host_cores())
    return time.perf_counter() - start, min_cost, root


def secondary_resolutions(ctx, name, data, number, repeats, with_cpu, cpu_count, rank, world, local_rank):
    """Configs #1 / #5: superdtl over every binary resolution of a multifurcating object tree.
    N = 1: one GPU.  N > 1: strong scaling, resolution ranges per rank + NCCL all-reduce(MIN)."""
    import torch
    import torch.distributed as dist
    from superrec2_b200.compute.unordered_super_reconciliation import usreconcile_extended_uspfs
    from superrec2_b200.dist import allreduce_min_found, usreconcile_extended_uspfs_distributed
    from superrec2_b200.utils.dynamic_programming import RetentionPolicy

    srec_input, species, poly, bits, costs = resolution_workload(data)
    total = srec_input.count_resolutions()
    S = len(species.parent)
    n_leaves = int((np.diff(poly.child_off) == 0).sum())
    cell_kinds_ref = total * (n_leaves - 1) * S * 2       # what the reference evaluates: every tree in full
    entry = {"config": number, "algorithm": "superdtl (usreconcile_extended_uspfs)", "resolutions": total,
             "object_leaves": n_leaves, "species_nodes": S, "cell_kinds_reference_equivalent": cell_kinds_ref}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device time of this rank's block: rows resident, repeated solves --------------------
    lo, hi = total * rank // world, total * (rank + 1) // world
    ctx.set_species(species.parent, species.child0, species.child1)
    ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, lo, hi - lo)
    upload_ms = ctx.timing().upload_ms
    solve_ms, waves_ms, launches, waves, evaluated = repeat_solve(ctx.superdtl_solve, ctx.timing, repeats)
    gpu_costs = ctx.superdtl_results(costs_only=True)
    # ---- the C ABI end to end: flat host arrays in, the packed best key out (+ the reduce over ranks) ----
    device = torch.device("cuda", local_rank)
    abi = []
    for _ in range(repeats):
        barrier()
        start = time.perf_counter()
        ctx.resolutions_upload(poly.child_off, poly.child_idx, poly.leaf_species, bits, costs, lo, hi - lo)
        ctx.superdtl_solve()
        cost, number, root, _ = ctx.superdtl_best()
        allreduce_min_found((cost, number, root) if root >= 0 else None, device)
        abi.append((time.perf_counter() - start) * 1e3)
    abi_ms = float(np.median(abi))
    stats = torch.tensor([solve_ms, waves_ms, upload_ms, abi_ms], dtype=torch.float64, device="cuda")
    cells_t = torch.tensor([float(evaluated)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(cells_t, op=dist.ReduceOp.SUM)
    solve_ms, waves_ms, upload_ms, abi_ms = (float(v) for v in stats.tolist())
    evaluated_all = int(cells_t.item())

    # ---- end to end through the Python entry point (every rank calls it with the same input) --
    timings = {}
    call = lambda: usreconcile_extended_uspfs_distributed(srec_input, RetentionPolicy.ANY, local_rank, timings)  # noqa: E731
    call()
    walls, parts = [], []
    for _ in range(repeats):
        barrier()
        start = time.perf_counter()
        result = call()
        torch.cuda.synchronize()
        walls.append((time.perf_counter() - start) * 1e3)
        parts.append((timings["solve_ms"], timings["reduce_ms"], timings["materialise_ms"]))
    wall = torch.tensor([float(np.median(walls))] + [float(v) for v in np.median(np.asarray(parts), axis=0)],
                        dtype=torch.float64, device="cuda")
    owner = torch.tensor([1 if result else 0], device="cuda")
    if world > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
        dist.all_reduce(owner, op=dist.ReduceOp.SUM)
    e2e_ms, part_solve, part_reduce, part_mat = (float(v) for v in wall.tolist())
    winner = timings["winner"]
    if int(owner.item()) != 1:
        raise SystemExit(f"bench.py: {name}: {int(owner.item())} ranks returned the solution (exactly one must)")

    entry.update({
        "device_ms": solve_ms, "upload_ms": upload_ms, "value": total / (solve_ms * 1e-3), "unit": "resolutions/s",
        "cell_kinds_per_s_reference_equivalent": cell_kinds_ref / (solve_ms * 1e-3),
        "cell_kinds_evaluated": evaluated_all, "launches": launches, "wave_launches": waves,
        "min_cost": int(winner[0]), "winner": {"cost": int(winner[0]), "resolution": int(winner[1]), "root": int(winner[2])},
        "e2e": {"ms": e2e_ms, "value": total / (e2e_ms * 1e-3), "unit": "resolutions/s",
                "api": "usreconcile_extended_uspfs (Python entry point: host plan, unranking on the device, solve, "
                       "key reduce, winner rebuilt and decoded)" + (" sharded over ranks" if world > 1 else ""),
                "solve_ms": part_solve, "reduce_ms": part_reduce, "materialise_ms": part_mat,
                "abi_ms": abi_ms, "abi_value": total / (abi_ms * 1e-3),
                "abi": "sr2_resolutions_upload + sr2_superdtl_solve + sr2_superdtl_best through ctypes: flat host "
                       "arrays in, packed (cost, resolution, root) key out"
                       + (", + the all-reduce(MIN) over ranks" if world > 1 else "")
                       + "; the Python entry point adds tree flattening and the rebuilt winner (fixed cost per call)"},
        "roofline": roofline_entry(evaluated_all / world, waves_ms,
                                   "rows shared between resolutions: cells actually evaluated (reference-equivalent "
                                   "cells are reported separately); the solve is dominated by the per-resolution "
                                   "decode + cost() kernel, not by the table"),
        "sharding": "strong: contiguous resolution ranges per rank, one all-reduce(MIN) of the packed 8-byte key"
                    if world > 1 else "one GPU",
    })
    if world > 1:
        # latency of the collective alone, and equality with the single-GPU answer
        lat = []
        for _ in range(20):
            barrier()
            start = time.perf_counter()
            allreduce_min_found((7, rank, 1), device)
            lat.append((time.perf_counter() - start) * 1e6)
        entry["allreduce_us"] = float(np.median(lat))
        if rank == 0:
            from superrec2_b200.compute.unordered_super_reconciliation import solve_resolution_range
            start = time.perf_counter()
            alone = solve_resolution_range(srec_input, 0, total, local_rank)
            entry["single_gpu_solve_ms_same_process"] = (time.perf_counter() - start) * 1e3
            entry["equals_single_gpu"] = bool(alone == winner)
            if alone != winner:
                raise SystemExit(f"bench.py: {name}: sharded winner {winner} differs from the single-GPU {alone}")
    def cpu_leg():
        count = min(cpu_count, hi - lo)
        seconds, cpu_cost, cpu_root = cpu_resolutions(data, lo, count)
        per_res = (n_leaves - 1) * S * 2
        entry["cpu_baseline"] = {
            "value": count / seconds, "unit": "resolutions/s", "cell_kinds_per_s": count * per_res / seconds,
            "cores": host_cores(), "kind": "port",
            "sample": f"first {count} resolutions (every tree evaluated in full, as the reference does), "
                      f"OpenMP over resolutions, {seconds:.2f} s wall",
            "parity": require_equal(name, (gpu_costs[0][:count], gpu_costs[1][:count]), (cpu_cost, cpu_root))
            .replace("(min cost, root species, mapping)", "(min cost and root species of every sampled resolution)")}
    return entry, result, (cpu_leg if with_cpu and rank == 0 else None)


def secondary_crispr_check(entry, result):
    """Config #1 has a fixture produced by the UNMODIFIED reference: compare the returned JSON."""
    with gzip.open(os.path.join(ROOT, "tests", "golden", "crispr.json.gz"), "rt") as handle:
        (fixture,) = json.load(handle)
    if result:
        (output,) = result
        if output.to_dict() != fixture["result"] or output.cost() != fixture["min_cost"]:
            raise SystemExit("bench.py: PARITY FAILURE: crispr-class1 result differs from the reference's output")
        entry["reference_output"] = "identical JSON to the unmodified reference's (tests/golden/crispr.json.gz)"


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--workload", default="dtl_batch_10000x500_vs_200", choices=sorted(WORKLOADS))
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-secondary", action="store_true", help="only the headline workload")
    parser.add_argument("--profile", action="store_true",
                        help="for ncu runs only: no warm-up, no e2e, no CPU baseline (prints no bench line)")
    args = parser.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, args.workload, rank, world)
        return

    import torch
    import torch.distributed as dist
    from superrec2_b200 import _lib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 GPU: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n_s, n_o, n_trees, seed = WORKLOADS[args.workload]
    batch = synth.make_batch(n_s, n_o, n_trees, seed + 7919 * rank)
    cells = batch.cells
    inputs = (batch.node_off, batch.left, batch.right, batch.leaf_species, batch.costs)

    ctx = _lib.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)  # kernels launch on a torch stream: torch events see them
    ctx.set_species(batch.parent, batch.child0, batch.child1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: inputs resident in HBM -------------------------------------
    ctx.dtl_upload(*inputs)
    if args.profile:
        for _ in range(args.steps):
            ctx.dtl_solve()
        torch.cuda.synchronize()
        print("profile run done (no bench line)", flush=True)
        return
    for _ in range(max(args.warmup, 3)):
        ctx.dtl_solve()
    wave_ms, solve_ms, launches = 0.0, 0.0, 0
    sampler = ClockSampler(local_rank)
    barrier()
    with sampler, torch.cuda.stream(stream):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        for _ in range(args.steps):
            ctx.dtl_solve()
            t = ctx.timing()
            wave_ms += t.waves_ms
            solve_ms += t.solve_ms
            launches += t.n_launches
        stop.record(stream)
        barrier()
    elapsed_ms = start.elapsed_time(stop)
    min_cost, root, mapping = ctx.dtl_results()
    checksum = int(min_cost.astype(np.int64).sum())
    waves_per_step = max(1, ctx.timing().n_waves)

    # ---------------- e2e: host buffers through the C ABI --------------------------------
    # the caller's result buffers are allocated once (pinned) and reused, as a service solving batch
    # after batch would
    # this step's inputs sit in pinned host memory (the library copies them to the device chunk by
    # chunk and buckets them there; pageable arrays work too, through the library's own staging)
    def pin(array):
        return torch.from_numpy(np.ascontiguousarray(array)).pin_memory().numpy()
    pageable_inputs = inputs
    inputs = (inputs[0], pin(inputs[1]), pin(inputs[2]), pin(inputs[3])) + tuple(inputs[4:])
    out = (pin(np.zeros_like(min_cost)), pin(np.zeros_like(root)), pin(np.zeros_like(mapping)))
    for _ in range(2):
        ctx.dtl_run(*inputs, out=out)
    barrier()
    e2e_start = time.perf_counter()
    e2e_steps = max(1, min(2 * args.steps, 10))  # more steps than the device-timed loop: host jitter averages out
    e2e_parts = [0.0, 0.0, 0.0]
    for _ in range(e2e_steps):
        out[0][:] = -1
        out = ctx.dtl_run(*inputs, out=out)
        t = ctx.timing()
        e2e_parts[0] += t.upload_ms
        e2e_parts[1] += t.solve_ms
        e2e_parts[2] += t.results_ms
    barrier()
    e2e_s = time.perf_counter() - e2e_start
    assert int(out[0].astype(np.int64).sum()) == checksum
    h2d = int(batch.left.nbytes + batch.right.nbytes + batch.leaf_species.nbytes + batch.node_off.nbytes)
    d2h = int(mapping.nbytes + min_cost.nbytes + root.nbytes)
    # the 16-bit form of the call (sr2_dtl_run16: trees of 999 nodes and 399 species fit int16): half the
    # bytes over PCIe, widened / narrowed on the device
    inputs16 = (inputs[0],) + tuple(pin(np.ascontiguousarray(a, dtype=np.int16)) for a in inputs[1:4]) + (inputs[4],)
    out16 = (pin(np.zeros_like(min_cost)), pin(np.zeros(len(root), np.int16)), pin(np.zeros(len(mapping), np.int16)))
    for _ in range(2):
        ctx.dtl_run16(*inputs16, out=out16)
    barrier()
    e2e16_start = time.perf_counter()
    for _ in range(e2e_steps):
        out16[0][:] = -1
        ctx.dtl_run16(*inputs16, out=out16)
    barrier()
    e2e16_s = time.perf_counter() - e2e16_start
    assert int(out16[0].astype(np.int64).sum()) == checksum and np.array_equal(out16[2].astype(np.int32), out[2])
    h2d16 = int(sum(a.nbytes for a in inputs16[1:4]) + batch.node_off.nbytes)
    d2h16 = int(out16[0].nbytes + out16[1].nbytes + out16[2].nbytes)
    # the same call from a caller that does NOT pin anything (plain numpy arrays in and out)
    ctx.dtl_run(*pageable_inputs)
    barrier()
    pageable_start = time.perf_counter()
    for _ in range(e2e_steps):
        pageable_out = ctx.dtl_run(*pageable_inputs)
    barrier()
    pageable_s = time.perf_counter() - pageable_start
    assert int(pageable_out[0].astype(np.int64).sum()) == checksum

    # ---------------- reduce over ranks ----------------------------------------------------
    stats = torch.tensor([elapsed_ms, e2e_s * 1e3, wave_ms, pageable_s * 1e3, e2e16_s * 1e3], dtype=torch.float64,
                         device="cuda")
    total_cells = torch.tensor([cells], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_cells, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e32_ms, wave_ms_max, pageable_ms, e2e16_ms = (float(x) for x in stats.tolist())
    # the headline e2e figure is the call a binding makes for this workload: the 16-bit form
    e2e_ms, e2e_h2d, e2e_d2h = e2e16_ms, h2d16, d2h16
    all_cells = float(total_cells.item())

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = cells * args.steps * BYTES_PER_CELL / (wave_ms * 1e-3) / 1e9
        # cells of rows that are materialised: cherry rows (wave 1, a third of the rows) are virtual
        internal = batch.left >= 0
        safe_l = np.where(internal, batch.left, 0) + np.repeat(batch.node_off[:-1], np.diff(batch.node_off))
        safe_r = np.where(internal, batch.right, 0) + np.repeat(batch.node_off[:-1], np.diff(batch.node_off))
        cherry = internal & ~internal[safe_l] & ~internal[safe_r]
        n_cherry, n_other = int(cherry.sum()), int(internal.sum() - cherry.sum())
        cells_evaluated = n_other * batch.n_species
        achieved_evaluated = cells_evaluated * args.steps * BYTES_PER_CELL / (wave_ms * 1e-3) / 1e9
        dram_bytes_per_step = cells_evaluated * DRAM_B_PER_CELL_WAVES     # the wave launches
        dram_gbs = dram_bytes_per_step * args.steps / (wave_ms * 1e-3) / 1e9
        line = {
            "metric": "dp_cells_per_s",
            "value": all_cells * args.steps / (elapsed_ms * 1e-3),
            "unit": "cells/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "int32",
            "data": "synthetic",
            "config": {
                "workload": args.workload, "algorithm": "dtl (reconcile_thl)",
                "species_leaves": n_s, "species_nodes": batch.n_species, "object_leaves": n_o,
                "trees_per_gpu": n_trees, "cells_per_gpu": cells, "costs": [0, 1, 1, 1, 1],
                "instances_per_s": n_trees * world * args.steps / (elapsed_ms * 1e-3),
                "l2_policy": "inputs larger than L2: each step streams %.1f GB of DP rows "
                             "(126 MB L2)" % (cells * 8 / 1e9),
                "sharding": "independent instances per rank, no collective",
            },
            "e2e": {"value": all_cells * e2e_steps / (e2e_ms * 1e-3), "unit": "cells/s",
                    "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "api": "sr2_dtl_run16 (pinned int16 host arrays in, host arrays out; H2D, device-side bucketing, "
                           "solve and D2H pipelined over instance chunks)",
                    "int32_ms_per_step": e2e32_ms / e2e_steps,
                    "int32_value": all_cells * e2e_steps / (e2e32_ms * 1e-3),
                    "int32_h2d_bytes_per_step": h2d, "int32_d2h_bytes_per_step": d2h,
                    "int32_api": "sr2_dtl_run (the same with int32 arrays)",
                    "int32_host_ms": e2e_parts[0] / e2e_steps, "int32_device_span_ms": e2e_parts[1] / e2e_steps,
                    "int32_tail_ms": e2e_parts[2] / e2e_steps,
                    "pageable_ms_per_step": pageable_ms / e2e_steps,
                    "pageable_value": all_cells * e2e_steps / (pageable_ms * 1e-3),
                    "pageable_note": "the same call with plain (unpinned) numpy arrays in and freshly allocated "
                                     "arrays out: the library stages through its own pinned pools",
                    "python_entry_point_note": "reconcile_thl_batch (tree objects in, ReconciliationOutput objects "
                                               "out) additionally flattens and rebuilds ~1e7 Python tree nodes, "
                                               "which takes seconds; callers that keep trees as arrays use "
                                               "superrec2_b200.compute.reconciliation.reconcile_thl_flat"},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": dram_bytes_per_step / waves_per_step,
                         "frac_evaluated": achieved_evaluated / peak,
                         "dram_frac": dram_gbs / peak,
                         "dram_bytes_per_step_all_launches": cells_evaluated * DRAM_B_PER_CELL_STEP,
                         "reading": "frac: all cells x 20 B (the reference-equivalent figure, SURVEY 8d); "
                                    "frac_evaluated: only cells of rows that are materialised (cherry rows, a third of "
                                    "the rows, are evaluated by their readers and never stored); dram_frac: bytes that "
                                    "really cross the HBM interface (ncu, every launch of a step) / time / peak -- the kernels are bound by "
                                    "instruction latency and issue slots, not by HBM",
                         "kernel": "dtl wave kernels: dtl_wave_thread_kernel (rows with <= 64 root-path species: a thread "
                                   "per row, rows stored compact; 38 % of the serialised kernel time of a step), "
                                   "dtl_wave_flat2_kernel (the other rows, a warp per row, 39 %, and the chained launch "
                                   "of the small top waves, 7 %); wave 1 is virtual (readers evaluate the closed form of a cherry row)",
                         "peak_source": peak_src,
                         "bytes_per_cell": BYTES_PER_CELL,
                         "algorithmic_bytes_per_launch": cells * BYTES_PER_CELL / waves_per_step,
                         "wave_launches_per_step": waves_per_step,
                         "launch_definition": "a launch = one wave (one height of the whole batch): up to five concurrent "
                                              "kernel launches (four tiers of the thread kernel + the dense rows), 66 kernel "
                                              "launches in 14 waves for this workload; achieved and traffic are per wave",
                         "wave_ms_per_step": wave_ms / args.steps,
                         "cells_evaluated_per_step": cells_evaluated, "cherry_rows": n_cherry,
                         "traffic_source": "bytes per wave (average over the step's 14 waves) from ncu "
                                           "dram__bytes_read.sum + dram__bytes_write.sum of every launch of one step, "
                                           "profiles/r7h_dram_per_launch_config2.csv (cold caches, launches serialised), "
                                           "scaled by the cells this batch evaluates"},
            "checksum_min_costs": checksum,
        }
    cpu_legs = []   # every CPU leg runs after ALL GPU timing: a busy OpenMP team of another runtime in the
                    # process adds milliseconds of jitter to host-side calls (profiles/README.md r3h)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        def primary_cpu_leg():
            cores = host_cores()
            sample = min(n_trees, max(16, 6 * cores))  # about 20 core-seconds
            rate, seconds, answers = time_cpu_port(batch, sample)
            hi = int(batch.node_off[sample])
            line["cpu_baseline"] = {
                "value": rate, "unit": "cells/s", "cores": cores, "kind": "port",
                "sample": f"first {sample} trees of the batch, OpenMP over trees, {seconds:.1f} s wall; "
                          "plain-C port of the reference's O(|G||S|^2) algorithm",
                "parity": require_equal("config #2", (min_cost[:sample], root[:sample], mapping[:hi]), answers)}
        cpu_legs.append(primary_cpu_leg)
    ctx.close()

    # ---------------- the other BASELINE configurations ----------------------------------------
    if not args.no_secondary and args.workload == "dtl_batch_10000x500_vs_200":
        with_cpu = not args.no_cpu_baseline and world == 1
        repeats = max(args.steps, 5)
        sctx = _lib.default_context(local_rank)   # the context the Python entry points use
        secondary = {}
        if world == 1:
            with gzip.open(os.path.join(ROOT, "tests", "golden", "crispr.json.gz"), "rt") as handle:
                crispr = json.load(handle)[0]["input"]
            entry, result, leg = secondary_resolutions(sctx, "config #1", crispr, 1, repeats, with_cpu, 27,
                                                       rank, world, local_rank)
            secondary_crispr_check(entry, result)
            secondary["superdtl_crispr_class1"] = entry
            cpu_legs.append(leg)
            secondary["dtl_single_5000x2000"], leg = secondary_dtl_single(sctx, repeats, with_cpu)
            cpu_legs.append(leg)
            secondary["superdtl_single_1000x500x32"], leg = secondary_superdtl_single(sctx, repeats, with_cpu)
            cpu_legs.append(leg)
        entry, _, leg = secondary_resolutions(sctx, "config #5", synth.make_polytomy_input(), 5, repeats, with_cpu,
                                              2048, rank, world, local_rank)
        secondary["superdtl_resolutions_99225"] = entry
        cpu_legs.append(leg)
        # the same workload one size up (6-ary x 6-ary: 945 x 945 resolutions), where a GPU has ~50 ms of work
        # per call and sharding the enumeration pays; not a BASELINE config, reported for the scaling curve
        entry, _, leg = secondary_resolutions(sctx, "resolutions 6x6", synth.make_polytomy_input(arities=(6, 6)),
                                              "5 (6-ary x 6-ary variant)", max(3, repeats // 2), with_cpu, 512,
                                              rank, world, local_rank)
        secondary["superdtl_resolutions_893025"] = entry
        cpu_legs.append(leg)
        if line is not None:
            line["secondary"] = secondary
            if world > 1:
                line["secondary_note"] = ("configs #1, #3 and #4 are single instances (replicas only, SURVEY 8e): "
                                          "measured at N=1; config #5 is strong-scaled over the ranks")
    for leg in cpu_legs:
        if leg is not None:
            leg()
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
