#!/usr/bin/env python3
"""bench.py -- DP cells/s of the dtl reconciliation hot path on B200.

Workload (BASELINE.json configs[1]): ``dtl`` on a synthetic batch of 10,000 random
binary gene trees (500 leaves) against one 200-leaf species tree, default costs.
One *step* = one pass of the hot path over the whole batch: wavefront DP over
every (internal object node, species) cell, selection of the returned root and
traceback of the mapping, for every tree.

  value     cells/s with the batch already resident in HBM (sr2_dtl_solve)
  e2e       the same through the C-ABI call a binding makes (sr2_dtl_run): host
            buffers in, host buffers out, H2D/D2H copies and host-side bucketing
            inside the timed region
  roofline  dominant kernel (dtl_wave_kernel): algorithmic bytes (20 B per cell,
            SURVEY.md section 8d) / CUDA-event time of the wave launches, against
            the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle port (plain C restatement of the reference's
            O(|G||S|^2) algorithm, OpenMP over trees) on a bounded sample of the same
            workload, timed on this box's host cores (rank 0, N=1 only)

Multi-GPU (torchrun, one rank per GPU): instances are independent, so every rank
gets its own 10,000-tree batch (weak scaling), there is no data-path collective,
and value = cells of all ranks / max-over-ranks time.

``--impl reference`` times the reference algorithm's CPU port with all host threads
on a bounded sample per step (the reference itself is pure Python and cannot travel
to the GPU box).
"""
import argparse
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


def cpu_sample(batch, n_trees):
    """First n_trees instances of the batch, for the CPU port."""
    hi = int(batch.node_off[n_trees])
    return (batch.parent, batch.child0, batch.child1, batch.node_off[:n_trees + 1],
            batch.left[:hi], batch.right[:hi], batch.leaf_species[:hi], batch.costs)


def time_cpu_port(batch, n_trees):
    from oracle import binding as oracle
    args = cpu_sample(batch, n_trees)
    start = time.perf_counter()
    oracle.dtl_batch(*args, want_map=True, threads=os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
    seconds = time.perf_counter() - start
    cells = int((args[4] >= 0).sum()) * batch.n_species
    return cells / seconds, seconds


def run_reference(args, workload, rank, world):
    """--impl reference: the CPU port of the reference algorithm, all host threads."""
    if rank != 0:
        return
    from superrec2_b200 import synth
    n_s, n_o, n_trees, seed = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    sample = min(n_trees, max(16, 2 * cores))
    batch = synth.make_batch(n_s, n_o, sample, seed)
    for _ in range(min(args.warmup, 1)):
        time_cpu_port(batch, sample)
    start = time.perf_counter()
    for _ in range(args.steps):
        rate, _ = time_cpu_port(batch, sample)
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


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--workload", default="dtl_batch_10000x500_vs_200", choices=sorted(WORKLOADS))
    parser.add_argument("--no-cpu-baseline", action="store_true")
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

    # ---------------- e2e: host buffers through the C ABI --------------------------------
    # the caller's result buffers are allocated once (pinned) and reused, as a service solving batch
    # after batch would
    # this step's inputs sit in pinned host memory (the library copies them to the device chunk by
    # chunk and buckets them there; pageable arrays work too, through the library's own staging)
    def pin(array):
        return torch.from_numpy(np.ascontiguousarray(array)).pin_memory().numpy()
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

    # ---------------- reduce over ranks ----------------------------------------------------
    stats = torch.tensor([elapsed_ms, e2e_s * 1e3, wave_ms], dtype=torch.float64, device="cuda")
    total_cells = torch.tensor([cells], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_cells, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms, wave_ms_max = (float(x) for x in stats.tolist())
    all_cells = float(total_cells.item())

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = cells * args.steps * BYTES_PER_CELL / (wave_ms * 1e-3) / 1e9
        # DRAM traffic per wave launch, from the ncu --set full captures under profiles/
        # (r2o: dtl_wave_sparse_kernel<48> 0.69 GB read + 2.25 GB written for 1,000,032 rows of 399 species =
        #  7.4 B/cell; r1y / r2m: dtl_wave_flat2_kernel 12.1 B/cell; cherry rows (wave 1) are never written: their
        #  readers evaluate the closed form), scaled to this batch's rows (74 % of the other rows are sparse)
        internal = batch.left >= 0
        safe_l = np.where(internal, batch.left, 0) + np.repeat(batch.node_off[:-1], np.diff(batch.node_off))
        safe_r = np.where(internal, batch.right, 0) + np.repeat(batch.node_off[:-1], np.diff(batch.node_off))
        cherry = internal & ~internal[safe_l] & ~internal[safe_r]
        n_cherry, n_other = int(cherry.sum()), int(internal.sum() - cherry.sum())
        waves_per_step = max(1, ctx.timing().n_waves)
        traffic = n_other * batch.n_species * (0.74 * 7.4 + 0.26 * 12.1) / waves_per_step
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
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "api": "sr2_dtl_run (pinned host arrays in, host arrays out; H2D, device-side bucketing, solve "
                           "and D2H pipelined over instance chunks)",
                    "host_ms": e2e_parts[0] / e2e_steps, "device_span_ms": e2e_parts[1] / e2e_steps,
                    "tail_ms": e2e_parts[2] / e2e_steps},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "dtl wave kernels: dtl_wave_sparse_kernel<38|57> (rows with <= 57 root-path species, 55 % "
                                   "of the step), dtl_wave_flat2_kernel (the other rows, 32 %, and the chained launch "
                                   "of the small top waves, 5 %); wave 1 is virtual (readers evaluate the closed form of a cherry row)",
                         "peak_source": peak_src,
                         "bytes_per_cell": BYTES_PER_CELL,
                         "algorithmic_bytes_per_launch": cells * BYTES_PER_CELL / waves_per_step,
                         "wave_launches_per_step": waves_per_step,
                         "wave_ms_per_step": wave_ms / args.steps,
                         "traffic_source": "bytes per wave launch (average over the step's launches) from "
                                           "ncu dram__bytes_read+write per row, profiles/r2o_sparse44_ncu_details.csv, r1y_flat2_ncu_details.csv / r2m_flat2_ncu_details.csv"},
            "checksum_min_costs": checksum,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = min(n_trees, max(16, 6 * cores))  # about 20 core-seconds
            rate, seconds = time_cpu_port(batch, sample)
            line["cpu_baseline"] = {
                "value": rate, "unit": "cells/s", "cores": cores, "kind": "port",
                "sample": f"first {sample} trees of the batch, OpenMP over trees, {seconds:.1f} s wall; "
                          "plain-C port of the reference's O(|G||S|^2) algorithm"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
