"""Multi-GPU parity (-m gpu, needs >= 2 devices): row slabs + NCCL ghost-row exchange inside the library.

One process per GPU; gloo carries the rendezvous (NCCL id, gather of results), libweather_b200.so does the
ncclSend/ncclRecv itself. The decomposed run must be bit-identical to the oracle's single-domain run.
"""
import os
import socket
import sys

import numpy as np
import pytest

from weather_sim import _capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _extended_primitive_slabs(rank, world, msgs):
    """Extended Primitive model on slabs (called inside an initialised gloo group; also by
    profiles/tools/check_slab_ext_primitive.py): p, T, q are transported, so THEIR ghost rows travel too."""
    from oracle_py import Oracle
    from weather_sim import distributed as wd
    from weather_sim import synthetic as syn

    ext = (0.02, 0.05, 0.03)
    W, H, Lv = 90, 53, 2
    rng = np.random.default_rng(4)
    for integ, dx, dy in ((1, 1.0, 1.0), (0, 0.8, 1.7)):
        u, v, h = syn.white_noise_state(W, H, seed=33)
        u3, v3, h3 = (np.stack([a * (1.0 + 0.5 * k) for k in range(Lv)]).astype(np.float32) for a in (u, v, h - 10.0))
        h3 += 10.0
        p3 = (1013.25 + rng.uniform(-5, 5, (Lv, H, W))).astype(np.float32)
        t3 = (288.15 + rng.uniform(-3, 3, (Lv, H, W))).astype(np.float32)
        q3 = rng.uniform(0.0, 0.02, (Lv, H, W)).astype(np.float32)
        sim = wd.slab_simulation(W, H, rank, world, device_id=rank, model="primitive", integrator=integ,
                                 num_levels=Lv, coriolis_f=0.1, dx=dx, dy=dy, max_time=1e30, extended=ext)
        r0, n = sim.local_rows
        sim.set_state(*(a[:, r0:r0 + n] for a in (u3, v3, h3)), p=p3[:, r0:r0 + n], t=t3[:, r0:r0 + n],
                      q=q3[:, r0:r0 + n])
        sim.step(2)
        sim.set_state(q=q3[:, r0:r0 + n])  # a field written between steps: its ghost rows must follow
        sim.step(2)
        got = {k: wd.gather_rows(sim.get_field(k), 0) for k in ("u", "h", "p", "t", "q")}
        sim.close()
        if rank == 0:
            for lev in range(Lv):
                o = Oracle(W, H, 2, integ, coriolis_f=0.1, dx=dx, dy=dy, extended=ext)
                o.set_state(u3[lev], v3[lev], h3[lev], p=p3[lev], t=t3[lev], q=q3[lev])
                o.step(2)
                o.set_state(q=q3[lev])
                o.step(2)
                for k in got:
                    if got[k][lev].tobytes() != o.get_field(k).tobytes():
                        msgs.append(f"extended primitive slabs integ{integ} dx{dx} level {lev} field {k} differs")
                o.close()


def _worker(rank, world, port, cases, out_q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "nvidia-jetson-workload_b200"), os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle_py import Oracle
    from weather_sim import _capi
    from weather_sim import distributed as wd
    from weather_sim import synthetic as syn

    dist.init_process_group("gloo", rank=rank, world_size=world)
    msgs = []
    try:
        for (W, H, integ, kernel, steps, dtype) in cases:
            dt = np.dtype(dtype)
            u, v, h = syn.white_noise_state(W, H, dtype=dt, seed=W + H)
            sim = wd.slab_simulation(W, H, rank, world, device_id=rank, integrator=integ, kernel=kernel,
                                     coriolis_f=0.1, max_time=1e30, dtype=dt)
            r0, n = sim.local_rows
            assert (r0, n) == wd.slab_rows(H, world, rank)
            sim.set_state(u[r0:r0 + n], v[r0:r0 + n], h[r0:r0 + n])
            sim.step(steps)
            got = {k: wd.gather_rows(sim.get_field(k), 0) for k in ("u", "v", "h", "vorticity")}
            halo_ms = sim.metrics.halo_time_ms
            sim.close()
            if rank == 0:
                o = Oracle(W, H, 0, integ, coriolis_f=0.1, dtype=dt)
                o.set_state(u, v, h)
                o.step(steps)
                for k in got:
                    a, b = got[k], o.get_field(k)
                    if a.tobytes() != b.tobytes():
                        msgs.append(f"{W}x{H} integ{integ} {kernel} {dtype} field {k}: differs "
                                    f"({int((a != b).sum())} cells)")
                o.close()
        # streamed host step on slabs: edge rows go up first and are exchanged while the slabs stream behind them;
        # a device-resident step after it must find the state (and re-exchange the ghost rows)
        for (W, H, integ, dtype) in ((200, 16 * world + 5, 2, "float32"), (1100, 1300, 2, "float32"), (530, 700, 0, "float64")):
            dt = np.dtype(dtype)
            u, v, h = syn.white_noise_state(W, H, dtype=dt, seed=W + H)
            sim = wd.slab_simulation(W, H, rank, world, device_id=rank, integrator=integ, kernel="step_fused_tma",
                                     coriolis_f=0.1, max_time=1e30, dtype=dt)
            r0, n = sim.local_rows
            ins = [a[r0:r0 + n].copy() for a in (u, v, h)]  # copies: the second call writes into them
            outs = [np.empty((n, W), dt) for _ in range(3)]
            sim.step_host(*ins, *outs)
            sim.step_host(*outs, *ins)      # second step fed from the first one's host output
            sim.step(1)                     # third, device resident
            got = {k: wd.gather_rows(sim.get_field(k), 0) for k in ("u", "v", "h")}
            two = {k: wd.gather_rows(a, 0) for k, a in zip(("u", "v", "h"), ins)}
            sim.close()
            if rank == 0:
                o = Oracle(W, H, 0, integ, coriolis_f=0.1, dtype=dt)
                o.set_state(u, v, h)
                o.step(2)
                for k in two:
                    if two[k].tobytes() != o.get_field(k).tobytes():
                        msgs.append(f"step_host x2 {W}x{H} integ{integ} {dtype} field {k} differs")
                o.step(1)
                for k in got:
                    if got[k].tobytes() != o.get_field(k).tobytes():
                        msgs.append(f"step_host x2 + step {W}x{H} integ{integ} {dtype} field {k} differs")
                o.close()
        # initial conditions on slabs: each rank evaluates ITS rows of the global field (incl. the sequential RNG)
        import ctypes
        W, H = 96, 61
        for name, params, seed in (("vortex", (0.4, 0.6, 0.2, 5.0, 10.0), 0), ("random", (0.5,), 42), ("front", (), 0)):
            sim = wd.slab_simulation(W, H, rank, world, device_id=rank, integrator=0, max_time=1e30)
            sim.grid.apply_ic(name, params, seed)
            got = {k: wd.gather_rows(sim.get_field(k), 0) for k in ("u", "v", "h", "t")}
            sim.close()
            if rank == 0:
                want = {k: np.full((H, W), np.nan, np.float32) for k in "uvhptq"}
                arr = (ctypes.c_double * max(len(params), 1))(*params)
                st = _capi.load_library().wsb_ic_fill_host(name.encode(), arr, len(params), seed, None, W, H, 1.0, 1.0,
                                                           *[want[k].ctypes.data for k in "uvhptq"])
                assert st == 0
                for k in got:
                    ref = want[k] if not np.isnan(want[k]).all() else np.full((H, W), {"h": 10.0, "t": 288.15}.get(k, 0.0), np.float32)
                    if got[k].tobytes() != ref.tobytes():
                        msgs.append(f"slab initial condition {name} field {k} differs")
        # Primitive model with independent levels on slabs: every level ≡ the 2-D oracle of that level, T/p drift included
        W, H, Lv = 96, 50, 3
        u, v, h = syn.white_noise_state(W, H, seed=5)
        u3, v3, h3 = (np.stack([a * (1.0 + 0.25 * k) for k in range(Lv)]).astype(np.float32) for a in (u, v, h - 10.0))
        h3 += 10.0
        sim = wd.slab_simulation(W, H, rank, world, device_id=rank, model="primitive", integrator="rk2", num_levels=Lv,
                                 coriolis_f=0.1, max_time=1e30)
        r0, n = sim.local_rows
        sim.set_state(u3[:, r0:r0 + n], v3[:, r0:r0 + n], h3[:, r0:r0 + n])
        sim.step(3)
        got = {k: wd.gather_rows(sim.get_field(k), 0) for k in ("u", "v", "h", "t", "p", "vorticity")}
        sim.close()
        if rank == 0:
            for lev in range(Lv):
                o = Oracle(W, H, 2, 1, coriolis_f=0.1)
                o.set_state(u3[lev], v3[lev], h3[lev])
                o.step(3)
                for k in got:
                    if got[k][lev].tobytes() != o.get_field(k).tobytes():
                        msgs.append(f"primitive slabs level {lev} field {k} differs")
                o.close()
        # extended physics on slabs: the beta plane reads the GLOBAL row index, viscosity the ghost rows
        W, H = 150, 97
        u, v, h = syn.white_noise_state(W, H, seed=21)
        ext = (0.02, 0.05, 0.03)
        sim = wd.slab_simulation(W, H, rank, world, device_id=rank, integrator="rk4", rk4_classical=True, coriolis_f=0.1,
                                 max_time=1e30, extended=ext)
        r0, n = sim.local_rows
        sim.set_state(u[r0:r0 + n], v[r0:r0 + n], h[r0:r0 + n])
        sim.step(4)
        got = {k: wd.gather_rows(sim.get_field(k), 0) for k in ("u", "v", "h")}
        sim.close()
        if rank == 0:
            o = Oracle(W, H, 0, 2, coriolis_f=0.1, rk4_classical=True, extended=ext)
            o.set_state(u, v, h)
            o.step(4)
            for k in got:
                if got[k].tobytes() != o.get_field(k).tobytes():
                    msgs.append(f"extended physics on slabs: field {k} differs")
            o.close()
        _extended_primitive_slabs(rank, world, msgs)
        # the drop-in surface on slabs: pyweather_sim.SimulationConfig.rank / nranks / nccl_unique_id
        import weather_sim.pyweather_sim as m
        W, H = 120, 90
        u, v, h = syn.white_noise_state(W, H, seed=9)
        c = m.SimulationConfig()
        c.grid_width, c.grid_height = W, H
        c.integration_method = m.IntegrationMethod.RungeKutta4
        c.coriolis_f, c.max_time = 0.1, 1.0e9
        c.device_id, c.rank, c.nranks = rank, rank, world
        c.nccl_unique_id = wd.broadcast_bytes(m.nccl_unique_id() if rank == 0 else None, 0)
        ps = m.WeatherSimulation(c)
        ps.initialize()
        r0, n = ps.get_local_rows()
        g = ps.get_current_grid()
        assert (g.get_height(), g.get_width()) == (n, W)
        g.set_velocity_field(u[r0:r0 + n], v[r0:r0 + n])
        g.set_height_field(h[r0:r0 + n])
        ps.run(4)
        ps.step()
        pu, pv = ps.get_current_grid().get_velocity_field()
        got = {"u": wd.gather_rows(pu, 0), "v": wd.gather_rows(pv, 0),
               "h": wd.gather_rows(ps.get_current_grid().get_height_field(), 0),
               "vorticity": wd.gather_rows(ps.get_current_grid().get_vorticity_field(), 0)}
        del ps
        if rank == 0:
            o = Oracle(W, H, 0, 2, coriolis_f=float(np.float32(0.1)))
            o.set_state(u, v, h)
            o.step(5)
            for k in got:
                if got[k].tobytes() != o.get_field(k).tobytes():
                    msgs.append(f"pyweather_sim on slabs: field {k} differs")
            o.close()
        out_q.put(("ok" if not msgs else "; ".join(msgs), rank))
    except Exception as e:  # pragma: no cover
        import traceback
        out_q.put((f"rank {rank}: {type(e).__name__}: {e}\n{traceback.format_exc()}", rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_capi.device_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_decomposition_bit_identical(world):
    if _capi.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    cases = [
        (300, 67, 2, "step_fused_tma", 5, "float32"),
        (300, 67, 2, "step_fused_reg", 5, "float32"),
        (300, 67, 2, "stage_direct", 5, "float32"),
        (130, 41, 1, "step_fused", 4, "float32"),
        (130, 41, 0, "step_fused", 4, "float32"),
        (64, 64, 1, "step_fused", 4, "float64"),
        (2048, 2048, 2, "step_fused", 3, "float32"),
    ]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for msg, _ in results), results
