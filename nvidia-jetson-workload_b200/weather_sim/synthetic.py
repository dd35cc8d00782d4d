"""Synthetic initial conditions for parity tests and benchmarks (numpy, host side).

These are the two initial conditions BASELINE.json names ("Gaussian height bump, random vorticity
field"), defined in SURVEY.md section 8(d). They are generated in float64 and cast once, then pushed
through ``set_velocity_field`` / ``set_height_field`` into whichever implementation is under test, so
every implementation starts from bit-identical inputs.
"""
import numpy as np


def gaussian_bump(width, height, dtype=np.float32, amplitude=1.0, h_mean=10.0, sigma_frac=0.1):
    """IC-A: h = h_mean + A*exp(-r^2 / (2 sigma^2)), sigma = sigma_frac*min(W,H); u = v = 0.

    Stays finite for >= 1000 reference steps at dt = 0.01.
    """
    y = np.arange(height, dtype=np.float64)[:, None]
    x = np.arange(width, dtype=np.float64)[None, :]
    sigma = sigma_frac * min(width, height)
    r2 = (x - (width - 1) / 2.0) ** 2 + (y - (height - 1) / 2.0) ** 2
    h = h_mean + amplitude * np.exp(-r2 / (2.0 * sigma * sigma))
    u = np.zeros((height, width), dtype=dtype)
    v = np.zeros((height, width), dtype=dtype)
    return u, v, h.astype(dtype)


def _smooth_noise(rng, height, width, passes=6):
    """Cheap separable smoother (binomial passes with clamped edges); no scipy dependency."""
    a = rng.standard_normal((height, width))
    for _ in range(passes):
        ap = np.pad(a, 1, mode="edge")
        a = 0.25 * (ap[:-2, 1:-1] + ap[2:, 1:-1]) + 0.5 * ap[1:-1, 1:-1]
        ap = np.pad(a, 1, mode="edge")
        a = 0.25 * (ap[1:-1, :-2] + ap[1:-1, 2:]) + 0.5 * ap[1:-1, 1:-1]
    return a


def random_vorticity(width, height, dtype=np.float32, seed=1234, max_speed=0.5, h_mean=10.0):
    """IC-B: smooth random stream function psi; u = -dpsi/dy, v = dpsi/dx, scaled to max|u,v| = max_speed.

    Use for <= 50 steps: the reference scheme amplifies grid-scale noise.
    """
    rng = np.random.default_rng(seed)
    psi = _smooth_noise(rng, height, width)
    pp = np.pad(psi, 1, mode="edge")
    u = -(pp[2:, 1:-1] - pp[:-2, 1:-1]) * 0.5
    v = (pp[1:-1, 2:] - pp[1:-1, :-2]) * 0.5
    scale = max_speed / max(np.abs(u).max(), np.abs(v).max(), 1e-30)
    h = np.full((height, width), h_mean, dtype=np.float64)
    return (u * scale).astype(dtype), (v * scale).astype(dtype), h.astype(dtype)


def white_noise_state(width, height, dtype=np.float32, seed=7, amplitude=0.5, h_mean=10.0):
    """Rough field exercising every stencil term at every cell (short parity runs only)."""
    rng = np.random.default_rng(seed)
    u = amplitude * rng.uniform(-1.0, 1.0, (height, width))
    v = amplitude * rng.uniform(-1.0, 1.0, (height, width))
    h = h_mean + amplitude * rng.uniform(-1.0, 1.0, (height, width))
    return u.astype(dtype), v.astype(dtype), h.astype(dtype)


def total_mass(h):
    """Sum of h in float64."""
    return float(np.sum(h, dtype=np.float64))


def total_energy(u, v, h, gravity=9.81):
    """Sum of 0.5*h*(u^2+v^2) + 0.5*g*h^2 in float64."""
    u64, v64, h64 = (np.asarray(a, dtype=np.float64) for a in (u, v, h))
    return float(np.sum(0.5 * h64 * (u64 * u64 + v64 * v64) + 0.5 * gravity * h64 * h64))


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 in float64 (0 when both are identically zero)."""
    a64 = np.asarray(a, dtype=np.float64)
    b64 = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b64)
    num = np.linalg.norm(a64 - b64)
    return float(num / den) if den > 0 else float(num)
