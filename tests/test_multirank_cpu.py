"""CPU multi-process tests (gloo, world_size 2) of the host-side decomposition logic.

The GPU data path (ncclSend/ncclRecv of ghost rows inside libweather_b200.so) cannot run here, so these
tests pin what surrounds it: the partition arithmetic shared by Python and C, the rendezvous helper that
ships the NCCL id, gather of slabs, and -- with the CPU oracle standing in for the kernels -- the claim the
slab scheme rests on: a slab extended by `nstages` ghost rows per side, advanced one step with the global
clamp applied only at true domain edges, reproduces the single-domain result bit-for-bit on its own rows.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, H, W, integ, steps, out_q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "nvidia-jetson-workload_b200"), os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from oracle_py import Oracle
    from weather_sim import distributed as wd
    from weather_sim import synthetic as syn

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. rendezvous helper: the bytes rank 0 publishes arrive everywhere
        token = wd.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 0)
        assert token == bytes(range(128))

        # 2. slab stepping emulated with the oracle + gloo ghost-row exchange
        depth = {0: 1, 1: 2, 2: 4}[integ]
        u, v, h = syn.white_noise_state(W, H, seed=11)
        r0, n = wd.slab_rows(H, world, rank)
        loc = {k: a[r0:r0 + n].copy() for k, a in (("u", u), ("v", v), ("h", h))}
        for _ in range(steps):
            ext = {}
            for k in loc:
                top = bot = None
                reqs = []
                if rank > 0:
                    reqs.append(dist.isend(torch.from_numpy(loc[k][:depth].copy()), rank - 1))
                    top = torch.empty((depth, W), dtype=torch.float32)
                    reqs.append(dist.irecv(top, rank - 1))
                if rank < world - 1:
                    reqs.append(dist.isend(torch.from_numpy(loc[k][-depth:].copy()), rank + 1))
                    bot = torch.empty((depth, W), dtype=torch.float32)
                    reqs.append(dist.irecv(bot, rank + 1))
                for q in reqs:
                    q.wait()
                parts = ([top.numpy()] if top is not None else []) + [loc[k]] + ([bot.numpy()] if bot is not None else [])
                ext[k] = np.concatenate(parts, axis=0)
            o = Oracle(W, ext["u"].shape[0], 0, integ, coriolis_f=0.1)
            o.set_state(ext["u"], ext["v"], ext["h"])
            o.step(1, diagnostics=False)
            lo = depth if rank > 0 else 0
            loc = {k: o.get_field(k)[lo:lo + n].copy() for k in loc}
            o.close()
        full = {k: wd.gather_rows(loc[k], 0) for k in loc}
        if rank == 0:
            g = Oracle(W, H, 0, integ, coriolis_f=0.1)
            g.set_state(u, v, h)
            g.step(steps, diagnostics=False)
            ok = all(np.array_equal(full[k].view(np.uint32), g.get_field(k).view(np.uint32)) for k in full)
            out_q.put(("ok" if ok else "slab result differs from the single-domain oracle", rank))
        else:
            out_q.put(("ok", rank))
    except Exception as e:  # pragma: no cover
        out_q.put((f"rank {rank}: {type(e).__name__}: {e}", rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("integ", [0, 1, 2])
def test_slab_scheme_with_gloo_world_size_2(integ):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, 23, integ, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(msg == "ok" for msg, _ in results), results


def test_local_slab_and_partition_agree():
    from weather_sim import distributed as wd
    a = np.arange(11 * 5, dtype=np.float32).reshape(11, 5)
    got = np.concatenate([wd.local_slab(a, 3, r) for r in range(3)], axis=0)
    assert np.array_equal(got, a)
    lv = np.arange(2 * 11 * 5, dtype=np.float32).reshape(2, 11, 5)
    got = np.concatenate([wd.local_slab(lv, 4, r) for r in range(4)], axis=-2)
    assert np.array_equal(got, lv)
