#!/usr/bin/env python
"""Soak test of the step overlap: BASELINE config 2 (SWE 8192^2 fp32 RK4, f = 0.1) for N steps
  A  whole-step kernel, ONE enqueue of N steps (every launch chained to its predecessor),
  B  whole-step kernel, the same N steps as irregular batches with a synchronisation between them,
  C  per-stage kernel (no overlap, 4 launches per step),
and the three final states must be bit-identical: a race in the chunk-row counters would show up as a difference
between A and B or against C.   usage: python profiles/tools/soak.py [N] [grid]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path[:0] = [os.path.join(ROOT, "nvidia-jetson-workload_b200")]
from weather_sim import _capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
e = np.exp(-((np.arange(G, dtype=np.float64) - (G - 1) / 2.0) ** 2) / (2.0 * (0.1 * G) ** 2)).astype(np.float32)
h0 = e[:, None] * e[None, :] + np.float32(10.0)   # the bench's Gaussian bump (IC-A)
z = np.zeros_like(h0)


def run(kernel, batches):
    s = _capi.Simulation(G, G, integrator="rk4", coriolis_f=0.1, max_time=1e9, kernel=kernel)
    s.set_state(z, z, h0)
    t0 = time.perf_counter()
    for b in batches:
        s.step(b)
    dt = time.perf_counter() - t0
    out = s.state()
    assert s.steps == N
    s.close()
    return out, dt


rng = np.random.default_rng(7)
ragged = []
while sum(ragged) < N:
    ragged.append(int(min(rng.integers(1, 40), N - sum(ragged))))
a, ta = run("step_fused_tma", [N])
print(f"A  {N} steps in one enqueue      {ta / N * 1e3:.4f} ms/step (wall)", flush=True)
b, tb = run("step_fused_tma", ragged)
print(f"B  {len(ragged)} irregular batches        {tb / N * 1e3:.4f} ms/step (wall)", flush=True)
c, tc = run("stage_direct", [N])
print(f"C  per-stage kernel             {tc / N * 1e3:.4f} ms/step (wall)", flush=True)
ok = True
for f in ("u", "v", "h"):
    ab, ac = np.array_equal(a[f], b[f]), np.array_equal(a[f], c[f])
    fin = bool(np.isfinite(a[f]).all())
    print(f"{f}: A==B {ab}  A==C {ac}  finite {fin}  max|{f}| {np.abs(a[f]).max():.6g}")
    ok &= ab and ac and fin
print("soak", "ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
