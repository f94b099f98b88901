#!/usr/bin/env python3
"""Summarise an .ncu-rep here (no GPU needed): headline metrics of each captured launch and the
warp-stall samples grouped by CUDA source line.   usage: scripts/ncu_summary.py <file.ncu-rep> [top_n]"""
import collections
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
]


def ncu(path, page, *extra):
    out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv", *extra],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    path = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = ncu(path, "raw")
    head, units = rows[0], rows[1]
    for row in rows[2:]:
        print("==", row[head.index("Kernel Name")][:60])
        for name in WANT:
            if name in head:
                i = head.index(name)
                print(f"  {name:62s} {row[i]} {units[i]}")
    rows = ncu(path, "source", "--print-source", "cuda,sass")
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    if not heads:
        return
    per = []
    for n, s in enumerate(heads):   # one section per source file
        e = heads[n + 1] if n + 1 < len(heads) else len(rows)
        head = rows[s]
        if "Address" not in head or "# Samples" not in head:
            continue
        fname = ""
        for back in range(1, 4):
            if s - back >= 0 and rows[s - back] and rows[s - back][0] in ("File Path", "File Name"):
                fname = rows[s - back][1].rsplit("/", 1)[-1]
                break
        i_src, i_adr = 1, head.index("Address")
        i_smp, i_exe = head.index("# Samples"), head.index("Instructions Executed")
        stall = [(i, name[6:]) for i, name in enumerate(head) if name.startswith("stall_") and "Not Issued" not in name]
        for r in rows[s + 1:e]:
            if len(r) <= i_exe or not r[0].isdigit() or r[i_adr] != "-":
                continue  # SASS rows repeat what their source line row aggregates
            per.append((fname[:14] + ":" + r[0], r[i_src].strip()[:80], int(r[i_smp] or 0), int(r[i_exe] or 0),
                        collections.Counter({name: int(r[i] or 0) for i, name in stall})))
    total = sum(p[2] for p in per) or 1
    tot_exe = sum(p[3] for p in per) or 1
    print(f"== stall samples by source line (total {total} samples, {tot_exe / 1e6:.1f} M warp instructions)")
    for line, src, smp, exe, st in sorted(per, key=lambda p: -p[2])[:top_n]:
        top = ", ".join(f"{n}={100 * v / max(smp, 1):.0f}%" for n, v in st.most_common(3))
        print(f"  {100 * smp / total:5.1f}% smp {100 * exe / tot_exe:5.1f}% ins  {line:>20s} {src:80s} {top}")


if __name__ == "__main__":
    main()
