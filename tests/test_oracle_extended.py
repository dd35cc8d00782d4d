"""CPU tests (-m "not gpu") of the EXTENDED physics restatement (oracle/ws_oracle_body.inc, `ext` branches).

The extended physics has no counterpart in the reference's compute code, so the oracle is its
specification and cannot be pinned to the reference. What can be checked without it:
* with beta = viscosity = diffusivity = 0 the extended tendencies ARE the reference's (weather_simulation.cpp:535-537);
* the double instantiation equals a float64 numpy evaluation of the same association, bit for bit (SWE terms, the beta
  plane, the 5-point Laplacians, the tracer transport of the PrimitiveEquations model);
* properties of the equations: a uniform tracer stays uniform, a tracer blob moves with the flow, viscosity damps a
  checkerboard, the Coriolis parameter varies linearly in y about the mid row.
"""
import numpy as np
import pytest

from oracle_py import Oracle
from weather_sim import synthetic as syn

from test_oracle import assert_bit_equal

DX, DY, DT, G, F0 = 0.7, 1.1, 0.01, 9.81, 0.2
EXT = (0.03, 0.05, 0.02)  # beta, viscosity, diffusivity


def nb(a):
    """left, right, upper, lower neighbours with the reference's clamp (weather_simulation.cpp:510-513)"""
    p = np.pad(a, 1, mode="edge")
    return p[1:-1, :-2], p[1:-1, 2:], p[:-2, 1:-1], p[2:, 1:-1]


def ddx(a):
    l, r, _, _ = nb(a)
    return (r - l) / (2.0 * DX)


def ddy(a):
    _, _, t, b = nb(a)
    return (b - t) / (2.0 * DY)


def lap(a):
    l, r, t, b = nb(a)
    return ((r - 2.0 * a) + l) * (1.0 / (DX * DX)) + ((b - 2.0 * a) + t) * (1.0 / (DY * DY))


def swe_ext_tendencies(u, v, h, beta, nu, kappa):
    H = u.shape[0]
    fy = (F0 + (beta * DY) * (np.arange(H, dtype=np.float64) - (H - 1) * 0.5))[:, None]
    ux, uy, vx, vy, hx, hy = ddx(u), ddy(u), ddx(v), ddy(v), ddx(h), ddy(h)
    du = (-u * ux - v * uy - G * hx + fy * v) + nu * lap(u)
    dv = (-u * vx - v * vy - G * hy - fy * u) + nu * lap(v)
    dh = (-h * (ux + vy) - u * hx - v * hy) + kappa * lap(h)
    return du, dv, dh


def tracer_tendency(c, u, v, kappa):
    return (-u * ddx(c) - v * ddy(c)) + kappa * lap(c)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("model,integ", [(0, 0), (0, 1), (0, 2), (1, 2)])
def test_zero_coefficients_reduce_to_the_reference_tendencies(dtype, model, integ):
    W, H = 37, 23
    u, v, h = (a.astype(dtype) for a in syn.white_noise_state(W, H, seed=11))
    a = Oracle(W, H, model, integ, dx=DX, dy=DY, coriolis_f=F0, dtype=dtype)
    b = Oracle(W, H, model, integ, dx=DX, dy=DY, coriolis_f=F0, dtype=dtype, extended=(0.0, 0.0, 0.0))
    for o in (a, b):
        o.set_state(u, v, h)
        o.step(5)
    for f in ("u", "v", "h"):
        assert np.array_equal(a.get_field(f), b.get_field(f)), f  # x + 0*lap == x (only the sign of a zero may differ)


def test_fp64_swe_extended_euler_step_matches_numpy():
    W, H = 31, 17
    u, v, h = (a.astype(np.float64) for a in syn.white_noise_state(W, H, seed=5))
    o = Oracle(W, H, 0, 0, dx=DX, dy=DY, dt=DT, gravity=G, coriolis_f=F0, dtype=np.float64, extended=EXT)
    o.set_state(u, v, h)
    o.step(1)
    du, dv, dh = swe_ext_tendencies(u, v, h, *EXT)
    assert_bit_equal(o.get_field("u"), u + DT * du, "u")
    assert_bit_equal(o.get_field("v"), v + DT * dv, "v")
    assert_bit_equal(o.get_field("h"), h + DT * dh, "h")


def test_fp64_swe_extended_rk2_step_matches_numpy():
    """midpoint rule (weather_simulation.cpp:220-323) around the extended tendencies"""
    W, H = 19, 29
    u, v, h = (a.astype(np.float64) for a in syn.white_noise_state(W, H, seed=6))
    o = Oracle(W, H, 0, 1, dx=DX, dy=DY, dt=DT, gravity=G, coriolis_f=F0, dtype=np.float64, extended=EXT)
    o.set_state(u, v, h)
    o.step(1)
    k = swe_ext_tendencies(u, v, h, *EXT)
    mid = [y + 0.5 * DT * ky for y, ky in zip((u, v, h), k)]
    k = swe_ext_tendencies(*mid, *EXT)
    for name, y, ky in zip("uvh", (u, v, h), k):
        assert_bit_equal(o.get_field(name), y + DT * ky, name)


@pytest.mark.parametrize("integ", [0, 1])
def test_fp64_primitive_tracer_transport_matches_numpy(integ):
    W, H = 23, 21
    rng = np.random.default_rng(3)
    u, v, h = (a.astype(np.float64) for a in syn.white_noise_state(W, H, seed=8))
    p, t, q = (base + rng.standard_normal((H, W)) for base in (1013.25, 288.15, 0.5))
    o = Oracle(W, H, 2, integ, dx=DX, dy=DY, dt=DT, gravity=G, coriolis_f=F0, dtype=np.float64, extended=EXT)
    o.set_state(u, v, h, p=p, t=t, q=q)
    o.step(1)
    kappa = EXT[2]
    if integ == 0:
        want = [c + DT * tracer_tendency(c, u, v, kappa) for c in (p, t, q)]
    else:
        k = swe_ext_tendencies(u, v, h, *EXT)
        um, vm = u + 0.5 * DT * k[0], v + 0.5 * DT * k[1]
        mids = [c + 0.5 * DT * tracer_tendency(c, u, v, kappa) for c in (p, t, q)]
        want = [c + DT * tracer_tendency(cm, um, vm, kappa) for c, cm in zip((p, t, q), mids)]
    for name, w in zip(("p", "t", "q"), want):
        assert_bit_equal(o.get_field(name), w, name)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_uniform_tracer_stays_uniform_and_blob_moves_with_the_flow(dtype):
    W, H = 96, 32
    o = Oracle(W, H, 2, 1, dt=0.01, dtype=dtype, extended=(0.0, 0.0, 0.0))
    x = np.arange(W, dtype=np.float64)[None, :] + np.zeros((H, 1))
    blob = np.exp(-((x - 30.0) ** 2) / (2 * 5.0 ** 2))
    o.set_state(np.ones((H, W)), np.zeros((H, W)), np.full((H, W), 10.0), t=288.15 + blob, q=np.full((H, W), 0.25))
    o.step(200)  # t = 2 in a uniform flow u = 1: the blob moves two cells to the right
    q, t = o.get_field("q").astype(np.float64), o.get_field("t").astype(np.float64) - float(dtype(288.15))
    assert (q == q.dtype.type(0.25)).all()
    centre = (t[H // 2] * x[0]).sum() / t[H // 2].sum()
    assert abs(centre - 32.0) < 0.02, centre
    # transported, not created (fp32: one ulp of 288 is 3e-5 per cell and step)
    assert abs(t.sum() / blob.sum() - 1.0) < (1e-6 if dtype == np.float64 else 2e-3)


def test_viscosity_damps_a_checkerboard_and_beta_plane_is_linear_in_y():
    W, H = 16, 16
    yy, xx = np.mgrid[0:H, 0:W]
    cb = np.where((xx + yy) % 2 == 0, 0.01, -0.01)
    o = Oracle(W, H, 0, 0, dtype=np.float64, extended=(0.0, 0.1, 0.0))
    o.set_state(cb, np.zeros((H, W)), np.full((H, W), 10.0))
    o.step(1)
    assert np.abs(o.get_field("u")[2:-2, 2:-2]).max() < 0.01 * (1.0 - 0.5 * 8 * 0.1 * 0.01)  # lap(cb) = -8 cb inside
    # beta plane: with u = 0, uniform v and flat h one Euler step gives du = dt * f(y) * v, f(y) = f0 + beta*(y - (H-1)/2)
    beta, v0 = 0.05, 0.5
    o = Oracle(W, H, 0, 0, coriolis_f=F0, dtype=np.float64, extended=(beta, 0.0, 0.0))
    o.set_state(np.zeros((H, W)), np.full((H, W), v0), np.full((H, W), 10.0))
    o.step(1)
    fy = F0 + beta * (np.arange(H) - (H - 1) * 0.5)
    assert np.allclose(o.get_field("u"), (0.01 * fy * v0)[:, None] * np.ones((1, W)), rtol=1e-14, atol=0)
    assert np.isclose(fy[0] - F0, -(fy[-1] - F0), rtol=1e-12)  # antisymmetric about the mid row
