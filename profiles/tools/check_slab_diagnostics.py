"""2-rank check (torchrun --nproc-per-node 2): vorticity and divergence of a row-slab run == the oracle's."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nvidia-jetson-workload_b200"), os.path.join(ROOT, "oracle")]
import torch.distributed as dist
from oracle_py import Oracle
from weather_sim import distributed as wd, synthetic as syn
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
bad = 0
for (W, H, dtype) in ((203, 77, "float32"), (515, 130, "float64"), (64, 9, "float32")):
    dt = np.dtype(dtype)
    u, v, h = syn.white_noise_state(W, H, dtype=dt, seed=W + H)
    sim = wd.slab_simulation(W, H, rank, world, device_id=rank, integrator=2, coriolis_f=0.1, max_time=1e30, dtype=dt)
    r0, n = sim.local_rows
    sim.set_state(u[r0:r0 + n].copy(), v[r0:r0 + n].copy(), h[r0:r0 + n].copy())
    sim.step(3)
    got = {k: wd.gather_rows(sim.get_field(k), 0) for k in ("vorticity", "divergence")}
    sim.close()
    if rank == 0:
        o = Oracle(W, H, 0, 2, coriolis_f=0.1, dtype=dt)
        o.set_state(u, v, h); o.step(3)
        for k in got:
            same = got[k].tobytes() == o.get_field(k).tobytes()
            bad += not same
            print(W, H, dtype, k, "ok" if same else "DIFFERS", flush=True)
if rank == 0:
    print("slab diagnostics:", "all bit-identical" if not bad else f"{bad} mismatches", flush=True)
