#!/usr/bin/env python
"""Instruction mix of the hot loop of a kernel in a cuobjdump -sass listing.

usage: cuobjdump -sass lib.so | python profiles/sass_mix.py <kernel-name-substring> [iterations_per_loop]
Finds the largest backward-branch loop in the kernel and prints the opcode histogram inside it.
"""
import collections
import re
import sys

name = sys.argv[1]
per = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ins, on = [], False
for l in sys.stdin:
    if "Function :" in l:
        on = name in l
        if on and ins:
            break
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
for a, t in ins:
    if "BRA" in t:
        mm = re.search(r"BRA\s+0x([0-9a-f]+)", t)
        if mm:
            tgt = int(mm.group(1), 16)
            if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                best = (tgt, a)
print("kernel instructions:", len(ins))
if best:
    body = [t for a, t in ins if best[0] <= a <= best[1]]
    print("hot loop 0x%x..0x%x: %d instructions (%.1f per iteration at %d iterations/loop)" % (best[0], best[1], len(body), len(body) / per, per))
    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
    for k, v in c.most_common(25):
        print("  %-14s %5d  %6.1f/iter" % (k, v, v / per))
