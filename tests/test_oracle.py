"""CPU tests (-m "not gpu"): pin the oracle restatement (oracle/ws_oracle.c) to the real reference.

* against the committed golden vectors generated from the reference itself
  (tests/golden/make_golden.py -> oracle/_ref/libws_ref.so);
* against oracle/_ref live, whenever it has been built in this checkout (dev container);
* properties of the restated algorithm the domain offers (uniform state is a fixed point, x/y symmetry,
  the RK4 aliasing identity of SURVEY.md F5, float time accumulation).
Bit-exact everywhere: the path is fp32 arithmetic with a fixed operation order.
"""
import os

import numpy as np
import pytest

from oracle_py import Oracle, Reference, oracle_tendencies, reference_available
from weather_sim import synthetic as syn

FIELDS = ("u", "v", "h", "p", "t", "q", "vorticity")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32 if a.dtype == np.float32 else np.uint64)


def assert_bit_equal(a, b, what=""):
    assert a.shape == b.shape and a.dtype == b.dtype, what
    same = bits(a) == bits(b)
    if not same.all():
        bad = np.argwhere(~same)
        raise AssertionError(f"{what}: {len(bad)} cells differ, first at {bad[0]}: {a[tuple(bad[0])]!r} vs {b[tuple(bad[0])]!r}")


def test_small_matrix_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "small_matrix.npz"))
    for key, params in zip(g["cases"], g["case_params"]):
        W, H, model, integ, f, dx, dy, steps = params
        W, H, model, integ, steps = int(W), int(H), int(model), int(integ), int(steps)
        o = Oracle(W, H, model, integ, dx=dx, dy=dy, coriolis_f=f)
        o.set_state(g[f"in_{W}x{H}_u"], g[f"in_{W}x{H}_v"], g[f"in_{W}x{H}_h"])
        o.step(steps)
        for name in FIELDS:
            assert_bit_equal(o.get_field(name), g[f"{key}_{name}"], f"{key}/{name}")


def test_edge_shapes_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "edge_shapes.npz"))
    for key in g["cases"]:
        shape, integ = key.split("_i")
        W, H = (int(x) for x in shape.split("x"))
        o = Oracle(W, H, 0, int(integ), coriolis_f=0.1)
        o.set_state(g[f"{key}_in_u"], g[f"{key}_in_v"], g[f"{key}_in_h"])
        o.step(3)
        for name in ("u", "v", "h", "vorticity"):
            assert_bit_equal(o.get_field(name), g[f"{key}_{name}"], f"{key}/{name}")


@pytest.mark.parametrize("f", [0.0, 0.1])
def test_config1_swe256_euler1000_golden(golden_dir, f):
    """BASELINE config 1: SWE 256x256 fp32, Euler, 1000 steps (the reference's CPU-runnable case)."""
    g = np.load(os.path.join(golden_dir, "c1_swe256_euler1000.npz"))
    u, v, h = syn.gaussian_bump(256, 256)
    o = Oracle(256, 256, 0, 0, coriolis_f=f)
    o.set_state(u, v, h)
    o.step(1000)
    st = {n: o.get_field(n) for n in ("u", "v", "h", "vorticity")}
    for n, a in st.items():
        assert_bit_equal(a, g[f"f{f}_{n}"], f"c1 f={f} {n}")
    assert syn.total_mass(st["h"]) == float(g[f"f{f}_mass"])
    assert syn.total_energy(st["u"], st["v"], st["h"]) == float(g[f"f{f}_energy"])
    assert np.float32(o.time) == g[f"f{f}_time"]


def test_rk4_swe128_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "rk4_swe128_200.npz"))
    u, v, h = syn.gaussian_bump(128, 128)
    o = Oracle(128, 128, 0, 2, coriolis_f=0.1)
    o.set_state(u, v, h)
    o.step(200)
    for n in ("u", "v", "h", "vorticity"):
        assert_bit_equal(o.get_field(n), g[n], f"rk4 128 {n}")
    st = o.state()
    assert syn.total_mass(st["h"]) == float(g["mass"])
    assert syn.total_energy(st["u"], st["v"], st["h"]) == float(g["energy"])


def test_time_track_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "bookkeeping.npz"))
    track = g["time_track_dt0.01"]
    o = Oracle(4, 4, 0, 0)
    got = []
    for _ in range(len(track)):
        o.step(1, diagnostics=False)
        got.append(np.float32(o.time))
    assert_bit_equal(np.array(got, dtype=np.float32), track, "float time accumulation")
    # the reference's run() stops after the first step with time >= max_time (10.0): step 1000
    assert np.argmax(track >= np.float32(10.0)) + 1 == 1000


@pytest.mark.skipif(not reference_available(), reason="oracle/_ref not built here (needs /root/reference)")
@pytest.mark.parametrize("shape", [(33, 21), (5, 40)])
def test_oracle_equals_live_reference(shape):
    W, H = shape
    u, v, h = syn.white_noise_state(W, H, seed=W + H)
    for model in range(4):
        for integ in range(5):
            o = Oracle(W, H, model, integ, dx=0.5, dy=2.5, dt=0.02, coriolis_f=0.3, gravity=3.7)
            r = Reference(W, H, model, integ, dx=0.5, dy=2.5, dt=0.02, coriolis_f=0.3, gravity=3.7)
            o.set_state(u, v, h)
            r.set_state(u, v, h)
            o.step(4)
            r.step(4)
            for name in FIELDS:
                assert_bit_equal(o.get_field(name), r.get_field(name), f"m{model} i{integ} {name}")
            assert np.float32(o.time) == np.float32(r.time) and o.steps == r.steps


def test_uniform_state_is_fixed_point():
    """Centred differences of a constant are zero: u,v,h never move (and the reference's own
    `Step` gtest that expects otherwise cannot pass)."""
    W, H = 19, 11
    for integ in (0, 1, 2):
        o = Oracle(W, H, 0, integ, coriolis_f=0.0)
        o.set_state(np.full((H, W), 1.0, np.float32), np.zeros((H, W), np.float32), np.full((H, W), 10.0, np.float32))
        o.step(3)
        assert (o.get_field("u") == 1.0).all() and (o.get_field("v") == 0.0).all() and (o.get_field("h") == 10.0).all()


def test_transpose_symmetry():
    """Swapping x<->y and u<->v maps the scheme onto itself when f = 0 (exactly, in floating point,
    only for the h equation's symmetric part; u/v swap roles)."""
    W, H = 12, 9
    u, v, h = syn.white_noise_state(W, H, seed=3)
    du, dv, dh = oracle_tendencies(u, v, h, dx=1.0, dy=1.0, coriolis_f=0.0)
    du2, dv2, dh2 = oracle_tendencies(v.T.copy(), u.T.copy(), h.T.copy(), dx=1.0, dy=1.0, coriolis_f=0.0)
    # u-equation of the transposed problem is the v-equation of the original, term order (-u*vx - v*vy)
    # becomes (-v'*..): not the same association, so compare to rounding rather than bitwise
    np.testing.assert_allclose(du2.T, dv, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dv2.T, du, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dh2.T, dh, rtol=1e-5, atol=1e-5)


def test_rk4_reference_combine_is_aliased():
    """reference RK4 == y + dt/6*(((k4+2k2)+2k3)+k4); the classical opt-in differs."""
    W, H = 24, 16
    u, v, h = syn.gaussian_bump(W, H)
    a = Oracle(W, H, 0, 2, coriolis_f=0.1)
    c = Oracle(W, H, 0, 2, coriolis_f=0.1, rk4_classical=True)
    a.set_state(u, v, h)
    c.set_state(u, v, h)
    a.step(1)
    c.step(1)
    # rebuild the aliased combine from single tendency evaluations
    dt = np.float32(0.01)
    half = np.float32(0.5) * dt
    y = (u, v, h)
    k1 = oracle_tendencies(*y, coriolis_f=0.1)
    t = tuple(yy + half * kk for yy, kk in zip(y, k1))
    k2 = oracle_tendencies(*t, coriolis_f=0.1)
    t = tuple(yy + half * kk for yy, kk in zip(y, k2))
    k3 = oracle_tendencies(*t, coriolis_f=0.1)
    t = tuple(yy + dt * kk for yy, kk in zip(y, k3))
    k4 = oracle_tendencies(*t, coriolis_f=0.1)
    two = np.float32(2.0)
    dt6 = dt / np.float32(6.0)
    for name, yy, a2, a3, a4 in zip(("u", "v", "h"), y, k2, k3, k4):
        expect = yy + dt6 * (((a4 + two * a2) + two * a3) + a4)
        assert_bit_equal(a.get_field(name), expect.astype(np.float32), f"aliased {name}")
    assert not np.array_equal(a.get_field("h"), c.get_field("h"))


def test_barotropic_rk4_is_swe_rk2():
    """non-SWE models fall through to SWE tendencies, RK4 falls back to RK2."""
    W, H = 20, 14
    u, v, h = syn.random_vorticity(W, H)
    a = Oracle(W, H, 1, 2)
    b = Oracle(W, H, 0, 1)
    a.set_state(u, v, h)
    b.set_state(u, v, h)
    a.step(6)
    b.step(6)
    for n in ("u", "v", "h", "vorticity"):
        assert_bit_equal(a.get_field(n), b.get_field(n), n)


def test_primitive_constant_tp_drift():
    """T += dt*288.15f, p += dt*1013.25f per step; q alternates between two buffers."""
    o = Oracle(8, 8, 2, 0)
    o.step(100)
    T = o.get_field("t")
    expect = np.float32(288.15)
    for _ in range(100):
        expect = np.float32(expect + np.float32(0.01) * np.float32(288.15))
    assert (T == expect).all()


def test_fp64_instantiation_matches_numpy_float64():
    """The double instantiation is the same code with T=double: compare one tendency evaluation with
    a float64 numpy evaluation of the same association (<= 1e-12 relative is the north-star bound)."""
    W, H = 31, 17
    u, v, h = (a.astype(np.float64) for a in syn.white_noise_state(W, H, seed=5))
    du, dv, dh = oracle_tendencies(u, v, h, dx=0.7, dy=1.1, gravity=9.81, coriolis_f=0.2)

    def nb(a):
        p = np.pad(a, 1, mode="edge")
        return p[1:-1, :-2], p[1:-1, 2:], p[:-2, 1:-1], p[2:, 1:-1]

    (uL, uR, uU, uD), (vL, vR, vU, vD), (hL, hR, hU, hD) = nb(u), nb(v), nb(h)
    ux, uy = (uR - uL) / (2.0 * 0.7), (uD - uU) / (2.0 * 1.1)
    vx, vy = (vR - vL) / (2.0 * 0.7), (vD - vU) / (2.0 * 1.1)
    hx, hy = (hR - hL) / (2.0 * 0.7), (hD - hU) / (2.0 * 1.1)
    edu = -u * ux - v * uy - 9.81 * hx + 0.2 * v
    edv = -u * vx - v * vy - 9.81 * hy - 0.2 * u
    edh = -h * (ux + vy) - u * hx - v * hy
    assert_bit_equal(du, edu, "du")
    assert_bit_equal(dv, edv, "dv")
    assert_bit_equal(dh, edh, "dh")
