#!/usr/bin/env python
"""Wall time of wsb_ic_apply for every initial condition of the reference on one grid size.

usage: python profiles/tools/ic_timing.py [N]      (default 32768: the BASELINE config 5 grid, 2^30 cells)
Separable conditions are expanded on the device from O(W + H) host values; vortex / mountain / breaking_wave / random
are evaluated by host threads (libm and the mt19937 stream are part of the reference's results) in 32 MiB row blocks
through two page-locked buffer sets. WSB_IC_HOST=1 forces the host path for everything (the round-1 behaviour)."""
import os
import sys
import time

sys.path[:0] = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "nvidia-jetson-workload_b200")]
from weather_sim import _capi  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
g = _capi.Grid(N, N)
print(f"grid {N}x{N} fp32, host threads {len(os.sched_getaffinity(0))}, WSB_IC_HOST={os.environ.get('WSB_IC_HOST')}")
for name, params, seed in (("uniform", (), 0), ("zonal_flow", (), 0), ("jet_stream", (), 0), ("front", (), 0),
                           ("standard_atmosphere", (), 0), ("vortex", (), 0), ("mountain", (), 0),
                           ("breaking_wave", (), 0), ("random", (0.5,), 42)):
    g.reset()
    t0 = time.perf_counter()
    g.apply_ic(name, params, seed)
    dt = time.perf_counter() - t0
    print(f"{name:22s} {dt * 1e3:10.1f} ms   {N * N / dt / 1e9:8.2f} G cells/s")
g.close()
