"""CPU tests: the initial conditions (wsb_ic_fill_host, pure host code behind the C-ABI) against the
reference's own InitialCondition classes, driven through the reference's pybind11 module built into
oracle/_ref (oracle/build_ref.sh). Bit-exact, including the reference's std::to_string/std::stof parameter
round trip. Skipped when oracle/_ref is not present (it is built wherever /root/reference is mounted and
travels to the GPU box with the snapshot).
"""
import ctypes
import glob
import importlib.machinery
import importlib.util
import os

import numpy as np
import pytest

from weather_sim import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MODULES = glob.glob(os.path.join(ROOT, "oracle", "_ref", "pyweather_sim*.so"))

pytestmark = pytest.mark.skipif(not REF_MODULES, reason="oracle/_ref/pyweather_sim*.so not built")


@pytest.fixture(scope="module")
def ref():
    loader = importlib.machinery.ExtensionFileLoader("pyweather_sim", REF_MODULES[0])
    spec = importlib.util.spec_from_loader("pyweather_sim", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def ours(name, params, W, H, seed=0, profile=None):
    lib = _capi.load_library()
    out = {k: np.full((H, W), np.nan, np.float32) for k in "uvhptq"}
    arr = (ctypes.c_double * max(len(params), 1))(*params)
    st = lib.wsb_ic_fill_host(name.encode(), arr, len(params), seed, profile.encode() if profile else None, W, H,
                              1.0, 1.0, *[out[k].ctypes.data for k in "uvhptq"])
    assert st == 0, _capi.last_error()
    return out


def reference_fields(ref, ic, W, H):
    g = ref.WeatherGrid(W, H)
    ic.initialize(g)
    u, v = g.get_velocity_field()
    return dict(u=u, v=v, h=g.get_height_field(), p=g.get_pressure_field(), t=g.get_temperature_field(),
                q=g.get_humidity_field())


CASES = [
    ("uniform", "UniformInitialCondition", [(), (1.5, -2.25, 9.0, 990.0, 280.0, 0.5)]),
    ("zonal_flow", "ZonalFlowInitialCondition", [(), (7.3, 12.0, 0.05)]),
    ("vortex", "VortexInitialCondition", [(), (0.31, 0.62, 0.2, 4.0, 11.0), (0.123456789, 0.5, 0.15, 3.3333333, 10.0)]),
    ("jet_stream", "JetStreamInitialCondition", [(), (0.4, 0.07, 6.5, 9.5)]),
    ("breaking_wave", "BreakingWaveInitialCondition", [(), (0.6, 0.35, 10.5)]),
    ("front", "FrontInitialCondition", [(), (0.45, 0.08, 7.0, 3.0)]),
    ("mountain", "MountainInitialCondition", [(), (0.25, 0.55, 0.12, 0.8, 4.0)]),
]


@pytest.mark.parametrize("shape", [(64, 48), (33, 57)])
def test_parameterised_ics_match_reference(ref, shape):
    W, H = shape
    for name, cls, param_sets in CASES:
        for params in param_sets:
            want = reference_fields(ref, getattr(ref, cls)(*params), W, H)
            got = ours(name, params, W, H)
            base = reference_fields(ref, ref.UniformInitialCondition(0, 0, 10.0, 1013.25, 288.15, 0), W, H)
            for k in "uvhptq":
                if np.isnan(got[k]).all():  # field not written by this IC: the reference leaves reset() values
                    np.testing.assert_array_equal(want[k], base[k], err_msg=f"{name}{params} {k} (untouched)")
                else:
                    assert got[k].tobytes() == want[k].tobytes(), f"{name}{params} field {k}"


def test_random_ic_matches_reference_mt19937(ref):
    W, H = 40, 24
    for seed, amp in ((0, 1.0), (42, 0.25), (7, 3.0)):
        want = reference_fields(ref, ref.RandomInitialCondition(seed, amp), W, H)
        got = ours("random", (amp,), W, H, seed=seed)
        for k in "uvh":
            assert got[k].tobytes() == want[k].tobytes(), (seed, amp, k)


def test_atmospheric_profiles_match_reference(ref):
    W, H = 50, 30
    for prof in ("standard", "tropical", "polar", "something-else"):
        want = reference_fields(ref, ref.AtmosphericProfileInitialCondition(prof), W, H)
        got = ours("atmospheric_profile", (), W, H, profile=prof)
        for k in "uvptq":
            assert got[k].tobytes() == want[k].tobytes(), (prof, k)
    # the factory names map onto the same three profiles (initial_conditions.cpp:653-665)
    ref.register_all_initial_conditions()
    fac = ref.InitialConditionFactory.get_instance()
    for fname in ("standard_atmosphere", "tropical_atmosphere", "polar_atmosphere"):
        want = reference_fields(ref, fac.create_initial_condition(fname), W, H)
        got = ours(fname, (), W, H)
        for k in "uvptq":
            assert got[k].tobytes() == want[k].tobytes(), (fname, k)


def test_factory_names_match_reference(ref):
    import weather_sim as ws
    ref.register_all_initial_conditions()
    assert ws.get_available_initial_conditions() == ref.InitialConditionFactory.get_instance().get_available_initial_conditions()
    assert ws.InitialConditionFactory.get_instance().create_initial_condition("nope") is None
    for name in ws.get_available_initial_conditions():
        a = ws.InitialConditionFactory.get_instance().create_initial_condition(name)
        b = ref.InitialConditionFactory.get_instance().create_initial_condition(name)
        assert a.get_name() == b.get_name()


def test_unknown_ic_is_an_error():
    lib = _capi.load_library()
    buf = np.zeros((4, 4), np.float32)
    st = lib.wsb_ic_fill_host(b"bogus", None, 0, 0, None, 4, 4, 1.0, 1.0, buf.ctypes.data, None, None, None, None, None)
    assert st == _capi.WSB_ERR_INVALID_ARGUMENT and "bogus" in _capi.last_error()


@pytest.mark.parametrize("threads", [2, 5, 16])
def test_threaded_fill_is_split_independent(threads, monkeypatch):
    """Rows are evaluated on several host threads (wsb_ic.cpp: ic_dispatch_parallel); every split must give the
    bytes of the serial evaluation, including the mt19937 stream of "random"."""
    W, H = 97, 61
    names = [c[0] for c in CASES] + ["random", "standard_atmosphere", "tropical_atmosphere"]
    monkeypatch.setenv("WSB_IC_THREADS", "1")
    serial = {n: ours(n, (), W, H, seed=7) for n in names}
    monkeypatch.setenv("WSB_IC_THREADS", str(threads))
    for n in names:
        got = ours(n, (), W, H, seed=7)
        for k in "uvhptq":
            assert got[k].tobytes() == serial[n][k].tobytes(), (n, k)
