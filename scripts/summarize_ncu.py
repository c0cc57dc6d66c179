#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.md
    python scripts/summarize_ncu.py full gpurun_out/prof_r1_b.ncu-rep    > profiles/r1_ncu_full.md
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_grid, i_block = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        n = r[i_name].split("(")[0]
        t = tot.setdefault(n, [0, 0.0, r[i_grid], r[i_block]])
        t[0] += 1
        t[1] += float(r[i_val].replace(",", ""))
    s = sum(v[1] for v in tot.values())
    print("| kernel | launches | total ms | avg us | share of GPU time | grid (first) | block |")
    print("|---|---|---|---|---|---|---|")
    for n, (c, t, g, b) in sorted(tot.items(), key=lambda x: -x[1][1]):
        print(f"| `{n}` | {c} | {t / 1e6:.3f} | {t / c / 1e3:.1f} | {t / s:.3f} | {g} | {b} |")
    print(f"\n{len(rows) - 1} launches, {s / 1e6:.3f} ms of kernel time (ncu per-launch times are cold-cache and serialised).")


def full(path):
    # a report, or the `ncu -i report --page raw --csv` text made on the GPU box (reports are too big to travel)
    out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"### `{r[idx['Kernel Name']].split('(')[0]}`  (launch id {r[idx['ID']]})\n")
        print("| metric | value | unit |")
        print("|---|---|---|")
        for m in METRICS:
            if m in idx:
                print(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
