#!/usr/bin/env bash
# Round-1 GPU session J: row-sweep diagnostics kernel: parity (incl. the rest of the suite that reads vorticity) + timing.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_shim_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/diag_timing.txt
import sys, time, numpy as np
sys.path[:0] = ["nvidia-jetson-workload_b200"]
from weather_sim import _capi
for (W, H, dt) in ((8192, 8192, np.float32), (8192, 8192, np.float64)):
    s = _capi.Simulation(W, H, integrator=0, max_time=1e30, dtype=dt)
    s.grid.apply_ic("vortex", (), 0)
    s.grid.calculate_diagnostics(); s.synchronize()
    t = time.perf_counter()
    for _ in range(20):
        s.grid.calculate_diagnostics()
    s.synchronize()
    ms = (time.perf_counter() - t) / 20 * 1e3
    b = 4 * np.dtype(dt).itemsize * W * H
    print(f"diagnostics {W}x{H} {np.dtype(dt).name}: {ms:.3f} ms  {b / ms / 1e6:.0f} GB/s algorithmic (read u,v; write vorticity, divergence)")
    s.close()
PY
