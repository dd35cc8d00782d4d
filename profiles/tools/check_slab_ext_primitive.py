#!/usr/bin/env python
"""2-rank check of the extended Primitive model on row slabs (tracer ghost rows): runs the case of
tests/test_multigpu.py (_extended_primitive_slabs) alone, one process per GPU.
usage: python profiles/tools/check_slab_ext_primitive.py [world]"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")


def worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "nvidia-jetson-workload_b200"), os.path.join(ROOT, "oracle"),
                    os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    msgs = []
    try:
        import test_multigpu
        test_multigpu._extended_primitive_slabs(rank, world, msgs)
        q.put((rank, "ok" if not msgs else "; ".join(msgs)))
    except Exception as e:
        import traceback
        q.put((rank, f"{type(e).__name__}: {e}\n{traceback.format_exc()}"))
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    import socket
    import torch.multiprocessing as mp
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(30)
    for r, m in res:
        print(f"rank {r}: {m}")
    sys.exit(0 if all(m == "ok" for _, m in res) else 1)
