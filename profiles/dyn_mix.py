#!/usr/bin/env python
"""Dynamic (executed) instruction mix of a kernel from an .ncu-rep captured with --import-source on.

usage: python profiles/dyn_mix.py prof.ncu-rep [iterations]   (iterations: divide counts, e.g. strip-row iterations)
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
iters = float(sys.argv[2]) if len(sys.argv) > 2 else None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
si, ie, sm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
mix, samples, seen = collections.Counter(), collections.Counter(), set()
for r in rows[rows.index(h) + 1:]:
    if len(r) <= ie or not r[ie].isdigit() or r[0] in seen:
        continue
    seen.add(r[0])
    toks = r[si].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.split(".")[0] in ("LDS", "STS", "STG", "LDG", "SHFL") and "." in op else "")
    mix[op] += int(r[ie])
    samples[op] += int(r[sm]) if r[sm].isdigit() else 0
tot, ts = sum(mix.values()), sum(samples.values())
print(f"executed warp instructions: {tot}" + (f"  ({tot / iters:.1f} per iteration)" if iters else ""))
for op, n in mix.most_common(40):
    print(f"  {op:14s} {n:12d} {100.0 * n / tot:6.2f}%  samples {100.0 * samples[op] / max(ts, 1):5.2f}%" + (f"  {n / iters:7.2f}/iter" if iters else ""))
