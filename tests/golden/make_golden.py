#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REAL reference (oracle/_ref/libws_ref.so).

Runs only in the dev container (needs /root/reference to build oracle/_ref via oracle/build_ref.sh).
The reference's own tests hold no golden values for the time-stepping path, so
these reference-generated vectors are what pins the oracle and the CUDA path.

    python tests/golden/make_golden.py

Files:
  small_matrix.npz   every (model, integrator) pair x f in {0, 0.1} on 17x13 (white noise, dx=0.75, dy=1.3)
                     and 64x48 (Gaussian bump, dx=dy=1), 5 steps; inputs + u,v,h,p,T,q,vorticity outputs.
  edge_shapes.npz    degenerate grids 1x1, 1x9, 9x1, 2x2, 3x2 (RK4 + Euler, 3 steps).
  c1_swe256_euler1000.npz   BASELINE config 1: SWE 256x256 fp32 Euler, 1000 steps, Gaussian bump,
                     f = 0 and f = 0.1; final u,v,h,vorticity + float64 mass/energy.
  rk4_swe128_200.npz SWE 128x128 RK4 (reference-parity combine) 200 steps, bump, f=0.1.
  bookkeeping.npz    time/step counters after run(n) incl. the max_time early break.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "nvidia-jetson-workload_b200"))
from oracle_py import Reference, build_oracle  # noqa: E402
from weather_sim import synthetic as syn  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
FIELDS = ("u", "v", "h", "p", "t", "q", "vorticity")


def sha(*arrays):
    m = hashlib.sha256()
    for a in arrays:
        m.update(np.ascontiguousarray(a).tobytes())
    return m.hexdigest()


def small_matrix():
    out = {}
    cases = []
    for (W, H, ic, dx, dy) in ((17, 13, "noise", 0.75, 1.3), (64, 48, "bump", 1.0, 1.0)):
        u, v, h = syn.white_noise_state(W, H) if ic == "noise" else syn.gaussian_bump(W, H)
        out[f"in_{W}x{H}_u"], out[f"in_{W}x{H}_v"], out[f"in_{W}x{H}_h"] = u, v, h
        for model in range(4):
            for integ in range(5):
                for f in (0.0, 0.1):
                    r = Reference(W, H, model, integ, dx=dx, dy=dy, coriolis_f=f)
                    r.set_state(u, v, h)
                    r.step(5)
                    key = f"{W}x{H}_m{model}_i{integ}_f{f}"
                    cases.append((key, W, H, model, integ, f, dx, dy, 5))
                    for name in FIELDS:
                        out[f"{key}_{name}"] = r.get_field(name)
                    r.close()
    out["cases"] = np.array([c[0] for c in cases])
    out["case_params"] = np.array([c[1:] for c in cases], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "small_matrix.npz"), **out)


def edge_shapes():
    out = {}
    cases = []
    for (W, H) in ((1, 1), (1, 9), (9, 1), (2, 2), (3, 2)):
        u, v, h = syn.white_noise_state(W, H, seed=W * 31 + H)
        for integ in (0, 2):
            r = Reference(W, H, 0, integ, coriolis_f=0.1)
            r.set_state(u, v, h)
            r.step(3)
            key = f"{W}x{H}_i{integ}"
            cases.append(key)
            out[f"{key}_in_u"], out[f"{key}_in_v"], out[f"{key}_in_h"] = u, v, h
            for name in ("u", "v", "h", "vorticity"):
                out[f"{key}_{name}"] = r.get_field(name)
            r.close()
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(OUT, "edge_shapes.npz"), **out)


def config1():
    out = {}
    W = H = 256
    u, v, h = syn.gaussian_bump(W, H)
    out["input_sha256"] = np.array(sha(u, v, h))
    for f in (0.0, 0.1):
        r = Reference(W, H, 0, 0, coriolis_f=f)
        r.set_state(u, v, h)
        r.step(1000)
        st = {n: r.get_field(n) for n in ("u", "v", "h", "vorticity")}
        for n, a in st.items():
            out[f"f{f}_{n}"] = a
        out[f"f{f}_mass"] = np.float64(syn.total_mass(st["h"]))
        out[f"f{f}_energy"] = np.float64(syn.total_energy(st["u"], st["v"], st["h"]))
        out[f"f{f}_time"] = np.float32(r.time)
        r.close()
    np.savez_compressed(os.path.join(OUT, "c1_swe256_euler1000.npz"), **out)


def rk4_128():
    out = {}
    W = H = 128
    u, v, h = syn.gaussian_bump(W, H)
    out["input_sha256"] = np.array(sha(u, v, h))
    r = Reference(W, H, 0, 2, coriolis_f=0.1)
    r.set_state(u, v, h)
    r.step(200)
    st = {n: r.get_field(n) for n in ("u", "v", "h", "vorticity")}
    for n, a in st.items():
        out[n] = a
    out["mass"] = np.float64(syn.total_mass(st["h"]))
    out["energy"] = np.float64(syn.total_energy(st["u"], st["v"], st["h"]))
    np.savez_compressed(os.path.join(OUT, "rk4_swe128_200.npz"), **out)


def bookkeeping():
    """run(n) semantics (weather_simulation.cpp:68-115): float time accumulation and the early break."""
    import ctypes
    from oracle_py import _load_ref
    lib = _load_ref()
    out = {}
    # max_time is fixed to 1e30 in the shim, so emulate the break on the float time track instead:
    r = Reference(8, 8, 0, 0)
    times = []
    for _ in range(1200):
        r.step(1)
        times.append(r.time)
    out["time_track_dt0.01"] = np.array(times, dtype=np.float32)
    r.close()
    np.savez_compressed(os.path.join(OUT, "bookkeeping.npz"), **out)


if __name__ == "__main__":
    build_oracle()
    small_matrix()
    edge_shapes()
    config1()
    rk4_128()
    bookkeeping()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
