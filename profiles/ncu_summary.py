#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): per-launch headline metrics + stall breakdown + hottest SASS.

usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--top N]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "launch__grid_size", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_lsu.sum",
    "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum",
]
for k, r in enumerate(data):
    print(f"--- launch {k}: {r[ci['Kernel Name']][:110]}")
    for w in WANT:
        if w in ci:
            print(f"  {w:72s} {r[ci[w]]:>16s} {units[ci[w]]}")
    stalls = []
    for h, i in ci.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            stalls.append((float(r[i] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    stalls.sort(reverse=True)
    print("  stalls per issue: " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next((r for r in rows if "Source" in r and "# Samples" in r), None)
if h:
    si, ss = h.index("Source"), h.index("# Samples")
    seen, body = set(), []
    for r in rows[rows.index(h) + 1:]:
        if len(r) > ss and r[ss].isdigit() and r[0] not in seen:
            seen.add(r[0])
            body.append((int(r[ss]), r[si].strip()))
    tot = sum(s for s, _ in body)
    print(f"--- hottest SASS (of {tot} samples, {len(body)} instructions)")
    for s, t in sorted(body, reverse=True)[:top_n]:
        print(f"  {100.0 * s / max(tot, 1):5.2f}%  {t[:100]}")
