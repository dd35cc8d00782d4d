"""The reference's OWN Python wrapper (src/weather-sim/python/weather_simulation.py, unmodified, byte-compiled by
oracle/build_ref.sh into oracle/_ref/weather_simulation.pyc.bin) imported on top of the B200 `pyweather_sim` module --
the drop-in claim of SURVEY.md section 7 step 3 / section 8b: `from .pyweather_sim import (...)` at
weather_simulation.py:16-29 must find every name, and `WeatherSimulationWrapper` (:194-371) must drive the GPU path.

CPU part: the import resolves against our module (no mock fallback) and construction fails loudly without a device.
GPU part: the same user script (the calls of examples/shallow_water_example.py) runs over the reference's own pybind
module and over ours, in one fresh interpreter each; every number must agree bit for bit.
"""
import glob
import json
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYC = os.path.join(ROOT, "oracle", "_ref", "weather_simulation.pyc.bin")
OURS = glob.glob(os.path.join(ROOT, "nvidia-jetson-workload_b200", "weather_sim", "pyweather_sim*.so"))
REFS = glob.glob(os.path.join(ROOT, "oracle", "_ref", "pyweather_sim*.so"))

DRIVER = textwrap.dedent("""
    import hashlib, importlib, json, os, sys
    import numpy as np
    pkg_dir, backend, script = sys.argv[1], sys.argv[2], sys.argv[3]
    sys.path.insert(0, os.path.dirname(pkg_dir))
    if backend == "cuda":  # libweather_b200.so must be resident before the module that needs it is dlopen'ed from a copy
        sys.path.insert(0, os.path.join(%(root)r, "nvidia-jetson-workload_b200"))
        from weather_sim import _capi
        _capi.load_library()
    ws = importlib.import_module(os.path.basename(pkg_dir) + ".weather_simulation")
    out = {"mock": not hasattr(ws, "register_all_initial_conditions"),
           "module": ws.WeatherSimulation.__module__, "ics": sorted(ws.get_available_initial_conditions())}
    if script == "import":
        try:
            ws.WeatherSimulationWrapper(width=16, height=16)
            out["constructed"] = True
        except Exception as e:
            out["constructed"] = False
            out["error"] = type(e).__name__ + ": " + str(e)
        print(json.dumps(out)); sys.exit(0)
    def digest(a):
        a = np.ascontiguousarray(a)
        return [str(a.dtype), list(a.shape), hashlib.sha256(a.tobytes()).hexdigest()]
    runs = {}
    for ic, kw in (("vortex", dict(x_center=0.5, y_center=0.5, radius=0.2, strength=10.0)),
                   ("jet_stream", dict(y_center=0.5, width=0.1, strength=20.0)),
                   ("breaking_wave", dict(amplitude=1.0, wavelength=0.2)),
                   ("zonal_flow", dict(u_max=20.0, beta=0.2)), ("standard_atmosphere", {})):
        for model, method in (("shallow_water", "rk4"), ("barotropic", "rk4"), ("primitive", "euler")):
            sim = ws.WeatherSimulationWrapper(width=96, height=64, model=model, dt=0.01, integration_method=method,
                                              backend=backend, output_interval=5)
            sim.config.max_time = 1.0e9
            sim.set_initial_condition(ic, **kw)
            sim.initialize()
            g = sim.get_grid()
            u0, v0 = g.get_velocity_field()
            sim.run(10)
            for _ in range(7):
                sim.step()
            sim.run_until(0.2)
            g = sim.get_grid()
            u, v = g.get_velocity_field()
            snaps = sim.get_output_data()
            m = sim.get_metrics()
            runs[f"{ic}/{model}/{method}"] = {
                "u0": digest(u0), "u": digest(u), "v": digest(v), "h": digest(g.get_height_field()),
                "vort": digest(g.get_vorticity_field()), "t": digest(g.get_temperature_field()),
                "p": digest(g.get_pressure_field()), "time": sim.simulation.get_current_time(),
                "step": sim.simulation.get_current_step(), "num_steps": m.num_steps,
                "snap_steps": [s["step"] for s in snaps], "snap_h": [digest(s["height"]) for s in snaps],
                "snap_vort": [digest(s["vorticity"]) for s in snaps]}
    out["runs"] = runs
    out["cuda_available"] = bool(ws.is_cuda_available())
    out["device_info_keys"] = sorted(ws.get_device_info())
    print(json.dumps(out))
""") % {"root": ROOT}


def _run(tmp_path, module_so, backend, script):
    """A package holding the reference's wrapper (.pyc) next to `module_so`, imported in a fresh interpreter."""
    pkg = tmp_path / f"refwrap_{backend}"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    shutil.copy(PYC, pkg / "weather_simulation.pyc")
    shutil.copy(module_so, pkg / os.path.basename(module_so))
    driver = tmp_path / f"driver_{backend}.py"
    driver.write_text(DRIVER)
    env = dict(os.environ, WSB_QUIET="1", OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, str(driver), str(pkg), backend, script], capture_output=True, text=True, env=env,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


needs_artifacts = pytest.mark.skipif(not (os.path.exists(PYC) and OURS),
                                     reason="oracle/_ref/weather_simulation.pyc.bin or the pyweather_sim shim not built")


@needs_artifacts
def test_reference_wrapper_imports_over_the_shim_without_the_mock_fallback(tmp_path):
    from weather_sim import _capi
    out = _run(tmp_path, OURS[0], "cuda", "import")
    assert out["mock"] is False and out["module"].endswith("pyweather_sim")
    assert out["ics"] == sorted(["breaking_wave", "front", "jet_stream", "mountain", "polar_atmosphere", "random",
                                 "standard_atmosphere", "tropical_atmosphere", "uniform", "vortex", "zonal_flow"])
    if _capi.device_count() == 0:  # no CPU fallback behind the reference's wrapper either
        assert out["constructed"] is False and "no CPU fallback" in out["error"]
    else:
        assert out["constructed"] is True


@pytest.mark.gpu
@needs_artifacts
@pytest.mark.skipif(not REFS, reason="oracle/_ref/pyweather_sim*.so not built")
def test_reference_wrapper_same_script_same_numbers(tmp_path):
    ours = _run(tmp_path, OURS[0], "cuda", "run")
    ref = _run(tmp_path, REFS[0], "cpu", "run")
    assert ours["mock"] is False and ref["mock"] is False
    assert ours["ics"] == ref["ics"]
    assert ours["cuda_available"] is True and ours["device_info_keys"] == ref["device_info_keys"]
    assert sorted(ours["runs"]) == sorted(ref["runs"]) and len(ours["runs"]) == 15
    for key, r in ref["runs"].items():
        o = ours["runs"][key]
        for k in r:
            assert o[k] == r[k], f"{key}: {k} differs between the reference module and the B200 module"
