"""CPU tests: the C-ABI library loads and exports every symbol include/weather_b200.h declares; host-only
entry points behave; the product fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

from weather_sim import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "weather_b200.h")).read()
    return sorted(set(re.findall(r"WSB_API\s+[^;()]*?\b(wsb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 40
    for must in ("wsb_sim_create", "wsb_sim_step", "wsb_sim_run", "wsb_grid_set_field", "wsb_grid_get_field",
                 "wsb_grid_calculate_diagnostics", "wsb_last_error", "wsb_nccl_get_unique_id"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_binding_covers_header():
    assert sorted(_capi.SIGNATURES) == declared_symbols()


def test_struct_sizes_match_c_layout():
    # offsets the C compiler produces for wsb_config (checked against a tiny C probe at build time would be
    # nicer; the library itself validates struct_size on every wsb_sim_create)
    assert ctypes.sizeof(_capi.wsb_config) % 8 == 0
    cfg = _capi.wsb_config()
    cfg.struct_size = ctypes.sizeof(_capi.wsb_config) - 4
    cfg.grid_width = cfg.grid_height = 8
    cfg.num_levels = 1
    h = ctypes.c_void_p()
    st = _capi.load_library().wsb_sim_create(ctypes.byref(cfg), ctypes.byref(h))
    assert st == _capi.WSB_ERR_INVALID_ARGUMENT and "struct_size" in _capi.last_error()


def test_partition_rows_is_balanced_and_contiguous():
    for H in (1, 7, 64, 8191, 32768):
        for G in (1, 2, 3, 4, 8):
            if G > H:
                with pytest.raises(ValueError):
                    _capi.partition_rows(H, G, 0)
                continue
            nxt = 0
            sizes = []
            for r in range(G):
                r0, n = _capi.partition_rows(H, G, r)
                assert r0 == nxt and n > 0
                nxt = r0 + n
                sizes.append(n)
            assert nxt == H and max(sizes) - min(sizes) <= 1


def test_argument_validation_needs_no_gpu():
    with pytest.raises(ValueError, match="Grid dimensions must be positive"):
        _capi.Simulation(0, 16)
    with pytest.raises(ValueError, match="Grid dimensions must be positive"):
        _capi.Simulation(16, 16, num_levels=0)
    with pytest.raises(ValueError):
        _capi.Simulation(16, 16, model=7)
    with pytest.raises(ValueError):
        _capi.Simulation(16, 4, nranks=8, nccl_id=b"\0" * 128)  # more ranks than rows


@pytest.mark.skipif(_capi.device_count() > 0, reason="a CUDA device is present")
def test_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without a CUDA device the product refuses to construct."""
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _capi.Simulation(16, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _capi.Grid(16, 16)


def test_product_does_not_link_or_import_the_oracle():
    """The shipped library and package never reference oracle/."""
    import subprocess
    out = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "ws_ref" not in out
    pkg = os.path.join(ROOT, "nvidia-jetson-workload_b200")
    for d, _, files in os.walk(pkg):
        if os.path.basename(d) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle_py" not in src and "libws_oracle" not in src and "ws_ref" not in src, f


def test_header_is_valid_c_and_the_c_example_links():
    """include/weather_b200.h must be consumable by a C compiler (no C++-isms); the plain-C example is the
    reference's benchmark loop on the C-ABI and has to build against the library."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "weather_b200.h")
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "nvidia-jetson-workload_b200"), "example"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exe = os.path.join(ROOT, "nvidia-jetson-workload_b200", "build", "swe_example")
    assert os.path.exists(exe)
    if _capi.device_count() == 0:  # fails loudly, no CPU fallback
        r = subprocess.run([exe, "64", "64", "2"], capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_exact_reciprocal_kernels_contain_no_contracted_fma():
    """Parity guard on the BUILT library: ptxas 12.9 was seen contracting packed mul + add into FFMA2
    even under --fmad=false; profiles/check_no_fma.sh disassembles every exact-reciprocal kernel and fails on any
    fused multiply-add that is not exact by construction. The Makefile runs the same script after linking."""
    import shutil
    import subprocess
    if not (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")):
        pytest.skip("cuobjdump not available")
    env = dict(os.environ, PATH=os.environ.get("PATH", "") + ":/usr/local/cuda/bin")
    r = subprocess.run([os.path.join(ROOT, "profiles", "check_no_fma.sh"), _capi.LIB_PATH], capture_output=True,
                       text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "check_no_fma: ok" in r.stdout


def test_exact_division_reciprocal_is_proven_on_the_host():
    """Non-power-of-two spacing: the kernels divide by 2dx, 2dy with q = x*r; e = x - q*d; q' = q + e*r. For fp32 the
    library proves per divisor, by exhaustion over all 2^23 significands, that the sequence returns the IEEE quotient
    (cached); the same loop is repeated here in numpy for two divisors, on a stride of the significands."""
    for d in (1.6, 3.4, 1.4, 3.8):
        r = _capi.exact_division_reciprocal(d)
        df = np.float32(d)
        assert r == float(np.float32(1.0) / df) and r != 0.0
        bits = (np.arange(0, 1 << 23, 37, dtype=np.uint32) | np.uint32(0x3F800000)).view(np.float32)
        q = bits * np.float32(r)
        # e = x - q*d exactly (what the FMA computes), in float64: 24-bit x 24-bit products are exact there
        e = (bits.astype(np.float64) - q.astype(np.float64) * np.float64(df)).astype(np.float32)
        q1 = (q.astype(np.float64) + e.astype(np.float64) * np.float64(np.float32(r))).astype(np.float32)
        assert np.array_equal(q1, bits / df)
    assert _capi.exact_division_reciprocal(3.0e7) == 0.0          # outside the proven window: IEEE division
    assert _capi.exact_division_reciprocal(1.6, np.float64) == 1.0 / 1.6
