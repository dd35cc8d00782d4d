"""TEST INFRASTRUCTURE ONLY -- ctypes views of the CPU oracle and of the real reference build.

* ``Oracle``    : oracle/libws_oracle.so  (plain-C restatement, float and double; oracle/ws_oracle.c)
* ``Reference`` : oracle/_ref/libws_ref.so (the reference's own WeatherSimulation, patched to compile;
                  oracle/build_ref.sh + oracle/ref_shim.cpp). float only -- the reference has no fp64.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product path (libweather_b200.so / pyweather_sim) never does.

Enum values follow the reference (weather_sim.hpp:30-56):
  model:      0 ShallowWater, 1 Barotropic, 2 PrimitiveEquations, 3 General
  integrator: 0 ExplicitEuler, 1 RungeKutta2, 2 RungeKutta4, 3 AdamsBashforth, 4 SemiImplicit
Field ids: 0 u, 1 v, 2 h, 3 p, 4 T, 5 q, 6 vorticity, 7 divergence (oracle only).
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libws_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libws_ref.so")

FIELD_IDS = {"u": 0, "v": 1, "h": 2, "p": 3, "t": 4, "q": 5, "vorticity": 6, "divergence": 7}


def build_oracle(force=False):
    """Compile oracle/libws_oracle.so (and oracle/_ref when /root/reference is mounted)."""
    if force or not os.path.exists(ORACLE_SO) or any(
        os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(ORACLE_SO)
        for f in ("ws_oracle.c", "ws_oracle_body.inc", "Makefile")
    ):
        subprocess.check_call(["make", "-s", "-C", HERE, "libws_oracle.so"])
    subprocess.check_call([os.path.join(HERE, "build_ref.sh")])


def reference_available():
    return os.path.exists(REF_SO)


_oracle_lib = None
_ref_lib = None


def _load_oracle():
    global _oracle_lib
    if _oracle_lib is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = ctypes.CDLL(ORACLE_SO)
        for sfx, ct in (("_f32", ctypes.c_float), ("_f64", ctypes.c_double)):
            f = getattr(lib, "wso_create" + sfx)
            f.restype = ctypes.c_void_p
            f.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double] * 5 + [ctypes.c_int]
            getattr(lib, "wso_destroy" + sfx).argtypes = [ctypes.c_void_p]
            getattr(lib, "wso_destroy" + sfx).restype = None
            for name in ("wso_set_field", "wso_get_field"):
                f = getattr(lib, name + sfx)
                f.restype = ctypes.c_int
                f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
            f = getattr(lib, "wso_step" + sfx)
            f.restype = None
            f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
            getattr(lib, "wso_diagnostics" + sfx).argtypes = [ctypes.c_void_p]
            getattr(lib, "wso_diagnostics" + sfx).restype = None
            getattr(lib, "wso_time" + sfx).argtypes = [ctypes.c_void_p]
            getattr(lib, "wso_time" + sfx).restype = ctypes.c_double
            getattr(lib, "wso_steps" + sfx).argtypes = [ctypes.c_void_p]
            getattr(lib, "wso_steps" + sfx).restype = ctypes.c_int
            getattr(lib, "wso_set_dt" + sfx).argtypes = [ctypes.c_void_p, ctypes.c_double]
            getattr(lib, "wso_set_dt" + sfx).restype = None
            getattr(lib, "wso_set_extended" + sfx).argtypes = [ctypes.c_void_p] + [ctypes.c_double] * 3
            getattr(lib, "wso_set_extended" + sfx).restype = None
            f = getattr(lib, "wso_tendencies" + sfx)
            f.restype = None
            f.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_double] * 4 + [ctypes.c_void_p] * 6
        lib.wso_build_flags.restype = ctypes.c_char_p
        _oracle_lib = lib
    return _oracle_lib


def _load_ref():
    global _ref_lib
    if _ref_lib is None:
        lib = ctypes.CDLL(REF_SO)
        lib.wsref_create.restype = ctypes.c_void_p
        lib.wsref_create.argtypes = [ctypes.c_int] * 4 + [ctypes.c_float] * 5
        lib.wsref_destroy.argtypes = [ctypes.c_void_p]
        lib.wsref_destroy.restype = None
        for name in ("wsref_set_field", "wsref_get_field"):
            f = getattr(lib, name)
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.wsref_step.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.wsref_step.restype = None
        lib.wsref_run.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.wsref_run.restype = None
        lib.wsref_time.argtypes = [ctypes.c_void_p]
        lib.wsref_time.restype = ctypes.c_float
        lib.wsref_steps.argtypes = [ctypes.c_void_p]
        lib.wsref_steps.restype = ctypes.c_int
        lib.wsref_set_dt.argtypes = [ctypes.c_void_p, ctypes.c_float]
        lib.wsref_set_dt.restype = None
        lib.wsref_diagnostics.argtypes = [ctypes.c_void_p]
        lib.wsref_diagnostics.restype = None
        lib.wsref_apply_ic.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        lib.wsref_apply_ic.restype = ctypes.c_int
        lib.wsref_set_omp_threads.argtypes = [ctypes.c_int]
        lib.wsref_set_omp_threads.restype = None
        lib.wsref_omp_threads.argtypes = []
        lib.wsref_omp_threads.restype = ctypes.c_int
        _ref_lib = lib
    return _ref_lib


class _SimBase:
    dtype = np.float32

    def set_state(self, u=None, v=None, h=None, **others):
        for name, arr in dict(u=u, v=v, h=h, **others).items():
            if arr is not None:
                self.set_field(name, arr)

    def state(self, names=("u", "v", "h")):
        return {n: self.get_field(n) for n in names}


class Oracle(_SimBase):
    """The C restatement. dtype np.float32 (reference arithmetic) or np.float64."""

    def __init__(self, width, height, model=0, integrator=2, dx=1.0, dy=1.0, dt=0.01, gravity=9.81,
                 coriolis_f=0.0, dtype=np.float32, rk4_classical=False, extended=None):
        """extended=(beta, viscosity, diffusivity) switches on the non-reference beta-plane / viscous tendencies."""
        self.lib = _load_oracle()
        self.dtype = np.dtype(dtype)
        self.sfx = "_f32" if self.dtype == np.float32 else "_f64"
        self.W, self.H = int(width), int(height)
        # the reference's config fields are float (weather_sim.hpp:166-172): round through float32 for
        # the float instantiation so that e.g. dt == 0.01f exactly as in the reference.
        if self.dtype == np.float32:
            dx, dy, dt, gravity, coriolis_f = (float(np.float32(x)) for x in (dx, dy, dt, gravity, coriolis_f))
        self.h = getattr(self.lib, "wso_create" + self.sfx)(model, integrator, self.W, self.H, dx, dy, dt,
                                                             gravity, coriolis_f, int(rk4_classical))
        if not self.h:
            raise ValueError("Grid dimensions must be positive")
        if extended is not None:
            ext = [float(np.float32(x)) if self.dtype == np.float32 else float(x) for x in extended]
            getattr(self.lib, "wso_set_extended" + self.sfx)(self.h, *ext)

    def _f(self, name):
        return getattr(self.lib, name + self.sfx)

    def set_field(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=self.dtype)
        assert a.shape == (self.H, self.W), (a.shape, (self.H, self.W))
        rc = self._f("wso_set_field")(self.h, FIELD_IDS[name], a.ctypes.data)
        assert rc == 0

    def get_field(self, name):
        out = np.empty((self.H, self.W), dtype=self.dtype)
        rc = self._f("wso_get_field")(self.h, FIELD_IDS[name], out.ctypes.data)
        assert rc == 0
        return out

    def step(self, n=1, diagnostics=True):
        self._f("wso_step")(self.h, int(n), int(bool(diagnostics)))

    def diagnostics(self):
        self._f("wso_diagnostics")(self.h)

    def set_dt(self, dt):
        self._f("wso_set_dt")(self.h, float(np.float32(dt)) if self.dtype == np.float32 else float(dt))

    @property
    def time(self):
        return self._f("wso_time")(self.h)

    @property
    def steps(self):
        return self._f("wso_steps")(self.h)

    def close(self):
        if self.h:
            self._f("wso_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def oracle_tendencies(u, v, h, dx=1.0, dy=1.0, gravity=9.81, coriolis_f=0.0):
    """One evaluation of the stencil (weather_simulation.cpp:473-540) on numpy arrays."""
    lib = _load_oracle()
    dt = np.dtype(u.dtype)
    sfx = "_f32" if dt == np.float32 else "_f64"
    if dt == np.float32:
        dx, dy, gravity, coriolis_f = (float(np.float32(x)) for x in (dx, dy, gravity, coriolis_f))
    u, v, h = (np.ascontiguousarray(a, dtype=dt) for a in (u, v, h))
    H, W = u.shape
    du, dv, dh = (np.empty_like(u) for _ in range(3))
    getattr(lib, "wso_tendencies" + sfx)(W, H, dx, dy, gravity, coriolis_f, u.ctypes.data, v.ctypes.data,
                                         h.ctypes.data, du.ctypes.data, dv.ctypes.data, dh.ctypes.data)
    return du, dv, dh


def oracle_omp_threads(n=None):
    """Set (n > 0) and return the OpenMP team size of the C port."""
    lib = _load_oracle()
    lib.wso_set_omp_threads.argtypes = [ctypes.c_int]
    lib.wso_omp_threads.restype = ctypes.c_int
    if n:
        lib.wso_set_omp_threads(int(n))
    return int(lib.wso_omp_threads())


def reference_omp_threads(n=None):
    """Set (n > 0) and return the OpenMP team size of the reference build -- whatever OMP_NUM_THREADS said at load."""
    lib = _load_ref()
    if n:
        lib.wsref_set_omp_threads(int(n))
    return int(lib.wsref_omp_threads())


class Reference(_SimBase):
    """The reference's own WeatherSimulation (float only)."""

    def __init__(self, width, height, model=0, integrator=2, dx=1.0, dy=1.0, dt=0.01, gravity=9.81,
                 coriolis_f=0.0):
        if not reference_available():
            raise RuntimeError("oracle/_ref/libws_ref.so not built (needs /root/reference; run oracle/build_ref.sh)")
        self.lib = _load_ref()
        self.W, self.H = int(width), int(height)
        self.h = self.lib.wsref_create(model, integrator, self.W, self.H, dx, dy, dt, gravity, coriolis_f)
        if not self.h:
            raise ValueError("Grid dimensions must be positive")

    def set_field(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        assert a.shape == (self.H, self.W)
        assert self.lib.wsref_set_field(self.h, FIELD_IDS[name], a.ctypes.data) == 0

    def get_field(self, name):
        out = np.empty((self.H, self.W), dtype=np.float32)
        assert self.lib.wsref_get_field(self.h, FIELD_IDS[name], out.ctypes.data) == 0
        return out

    def step(self, n=1):
        self.lib.wsref_step(self.h, int(n))

    def run(self, n):
        self.lib.wsref_run(self.h, int(n))

    def diagnostics(self):
        self.lib.wsref_diagnostics(self.h)

    def set_dt(self, dt):
        self.lib.wsref_set_dt(self.h, dt)

    def apply_ic(self, name):
        if self.lib.wsref_apply_ic(self.h, name.encode()) != 0:
            raise KeyError(name)

    @property
    def time(self):
        return float(self.lib.wsref_time(self.h))

    @property
    def steps(self):
        return self.lib.wsref_steps(self.h)

    def close(self):
        if self.h:
            self.lib.wsref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
