#!/usr/bin/env python
"""profiles/instruction_mix.json and the traffic stamp from ncu captures of the RK4 whole-step kernel.

usage: [WSB_MIX_RPC=256] python profiles/make_mix_json.py <strict.ncu-rep> [<folded.ncu-rep>]
       (captures taken with --import-source on; WSB_MIX_RPC = rows per chunk of the captured launches)

For each capture: executed warp instructions per opcode (ncu source page), divided by the strip-row iterations of
the launch (strips x chunks x (rows_per_chunk + 8)), and the FMA-pipe cycles per iteration they stand for (scalar
fp32 instruction = 1 cycle, packed FMUL2/FADD2/FFMA2 = 2). bench.py derives `roofline.fp32_pipe` from this file and
compares its source stamp with the kernel sources it runs. Also rewrites profiles/traffic.json's swe8192_rk4 entry
(dram__bytes_read.sum + dram__bytes_write.sum of the same capture) with the same stamp.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_stamp  # noqa: E402

W = H = 8192
COLS = 56
RPC = int(os.environ.get("WSB_MIX_RPC", "256"))  # rows per chunk of the captured launch (the shipped default at 8192^2)
ITER = -(-W // COLS) * -(-H // RPC) * (RPC + 8)


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", "--launch-count", "1"], capture_output=True,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def mix_of(rep):
    rows = page(rep, "source")
    h = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
    si, ie = h.index("Source"), h.index("Instructions Executed")
    mix, seen = collections.Counter(), set()
    for r in rows[rows.index(h) + 1:]:
        if len(r) <= ie or not r[ie].isdigit() or r[0] in seen:
            continue
        seen.add(r[0])
        toks = r[si].strip().split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        mix[op] += int(r[ie])
    per = {k: v / ITER for k, v in mix.items()}
    fp1 = sum(per.get(k, 0.0) for k in ("FADD", "FMUL", "FFMA"))
    fp2 = sum(per.get(k, 0.0) for k in ("FADD2", "FMUL2", "FFMA2"))
    raw = page(rep, "raw")
    ci = {k: i for i, k in enumerate(raw[0])}
    row = raw[2]
    traffic = None
    if "dram__bytes_read.sum" in ci:
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        traffic = sum(float(row[ci[k]]) * unit.get(raw[1][ci[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    return {"executed_warp_instructions_per_iteration": round(sum(per.values()), 2),
            "fp_scalar_per_iteration": round(fp1, 2), "fp_packed_per_iteration": round(fp2, 2),
            "fma_pipe_cycles_per_iteration": round(fp1 + 2.0 * fp2, 2),
            "top": {k: round(v, 2) for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:12]},
            "columns_per_strip": COLS, "rows_per_chunk": RPC, "iterations": ITER, "capture": os.path.basename(rep),
            "kernel_time_us_under_ncu": float(row[ci["gpu__time_duration.sum"]]) if "gpu__time_duration.sum" in ci else None,
            "stamp": source_stamp()}, traffic


out = {}
strict, traffic = mix_of(sys.argv[1])
out["strict"] = strict
if len(sys.argv) > 2:
    out["folded"], _ = mix_of(sys.argv[2])
with open(os.path.join(ROOT, "profiles", "instruction_mix.json"), "w") as f:
    json.dump(out, f, indent=1)
    f.write("\n")
tj_path = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(tj_path)) if os.path.exists(tj_path) else {}
if traffic:
    tj.setdefault("swe8192_rk4", {})["step_fused_tma"] = int(traffic)
    tj["_stamp"] = source_stamp()
    tj["_capture"] = os.path.basename(sys.argv[1])
    with open(tj_path, "w") as f:
        json.dump(tj, f, indent=1)
        f.write("\n")
print(json.dumps(out, indent=1))
