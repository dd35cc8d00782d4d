"""ctypes binding of libweather_b200.so (include/weather_b200.h) -- numpy in, numpy out.

This is the thin C-ABI route named by the north star ("C++ host code reached from Python through a thin
C-ABI (ctypes/pybind11), NumPy buffers handed straight to device memory"). The pybind11 module
``pyweather_sim`` exposes the reference's class surface on top of the same C-ABI; this module is the
lower-level view used by the benchmark and the parity tests.

There is no CPU fallback: if the library or a CUDA device is missing, construction raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# WSB_LIBRARY points at an alternative build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get("WSB_LIBRARY") or os.path.join(os.path.dirname(_HERE), "lib", "libweather_b200.so")

WSB_OK = 0
WSB_ERR_INVALID_ARGUMENT = -1
WSB_F32, WSB_F64 = 0, 1
NCCL_UNIQUE_ID_BYTES = 128

MODEL = {"shallow_water": 0, "barotropic": 1, "primitive": 2, "general": 3}
INTEGRATOR = {"euler": 0, "rk2": 1, "rk4": 2, "adams_bashforth": 3, "semi_implicit": 4}
FIELD = {"u": 0, "v": 1, "h": 2, "height": 2, "p": 3, "pressure": 3, "t": 4, "temperature": 4, "q": 5,
         "humidity": 5, "vorticity": 6, "divergence": 7}
KERNEL = {"auto": 0, "stage_direct": 1, "step_fused_reg": 2, "step_fused_tma": 3, "step_fused": 3}
ARITH = {"strict": 0, "folded": 1}


class wsb_config(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("model", ctypes.c_int32),
        ("integration_method", ctypes.c_int32),
        ("grid_width", ctypes.c_int32),
        ("grid_height", ctypes.c_int32),
        ("num_levels", ctypes.c_int32),
        ("dx", ctypes.c_double),
        ("dy", ctypes.c_double),
        ("dt", ctypes.c_double),
        ("gravity", ctypes.c_double),
        ("coriolis_f", ctypes.c_double),
        ("max_time", ctypes.c_double),
        ("dtype", ctypes.c_int32),
        ("device_id", ctypes.c_int32),
        ("rk4_mode", ctypes.c_int32),
        ("kernel_variant", ctypes.c_int32),
        ("rank", ctypes.c_int32),
        ("nranks", ctypes.c_int32),
        ("nccl_unique_id", ctypes.c_void_p),
        ("arith_mode", ctypes.c_int32),
        ("physics_mode", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 6),
        ("beta", ctypes.c_double),
        ("viscosity", ctypes.c_double),
        ("diffusivity", ctypes.c_double),
    ]


class wsb_metrics(ctypes.Structure):
    _fields_ = [
        ("total_time_ms", ctypes.c_double),
        ("compute_time_ms", ctypes.c_double),
        ("memory_transfer_time_ms", ctypes.c_double),
        ("io_time_ms", ctypes.c_double),
        ("num_steps", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("halo_time_ms", ctypes.c_double),
        ("kernel_launches", ctypes.c_uint64),
    ]


class wsb_device_caps(ctypes.Structure):
    _fields_ = [
        ("device_type", ctypes.c_int32),
        ("compute_capability_major", ctypes.c_int32),
        ("compute_capability_minor", ctypes.c_int32),
        ("cuda_cores", ctypes.c_int32),
        ("multiprocessors", ctypes.c_int32),
        ("global_memory", ctypes.c_uint64),
        ("shared_memory_per_block", ctypes.c_uint64),
        ("max_threads_per_block", ctypes.c_int32),
        ("max_threads_per_multiprocessor", ctypes.c_int32),
        ("clock_rate_khz", ctypes.c_int32),
        ("memory_clock_rate_khz", ctypes.c_int32),
        ("memory_bus_width", ctypes.c_int32),
        ("compute_power_ratio", ctypes.c_float),
        ("device_name", ctypes.c_char * 256),
    ]


class wsb_grid_info(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("num_levels", ctypes.c_int32),
        ("dtype", ctypes.c_int32),
        ("dx", ctypes.c_double),
        ("dy", ctypes.c_double),
        ("device_id", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


# name -> (restype, argtypes); every symbol include/weather_b200.h declares
_vp, _i32, _i64, _dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
SIGNATURES = {
    "wsb_version": (ctypes.c_char_p, []),
    "wsb_last_error": (ctypes.c_char_p, []),
    "wsb_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "wsb_device_capabilities": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(wsb_device_caps)]),
    "wsb_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(_vp)]),
    "wsb_host_free": (ctypes.c_int, [_vp]),
    "wsb_nccl_get_unique_id": (ctypes.c_int, [_vp]),
    "wsb_partition_rows": (ctypes.c_int, [_i32, _i32, _i32, ctypes.POINTER(_i32), ctypes.POINTER(_i32)]),
    "wsb_grid_create": (ctypes.c_int, [_i32, _i32, _i32, _dbl, _dbl, _i32, _i32, ctypes.POINTER(_vp)]),
    "wsb_grid_destroy": (None, [_vp]),
    "wsb_grid_reset": (ctypes.c_int, [_vp]),
    "wsb_grid_get_info": (ctypes.c_int, [_vp, ctypes.POINTER(wsb_grid_info)]),
    "wsb_grid_set_spacing": (ctypes.c_int, [_vp, _dbl, _dbl]),
    "wsb_grid_set_field": (ctypes.c_int, [_vp, _i32, _vp, _i32, _i64, _i64, _i64]),
    "wsb_grid_get_field": (ctypes.c_int, [_vp, _i32, _vp, _i32, _i64, _i64, _i64]),
    "wsb_grid_calculate_diagnostics": (ctypes.c_int, [_vp]),
    "wsb_grid_swap": (ctypes.c_int, [_vp, _vp]),
    "wsb_grid_device_pointer": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(_vp), ctypes.POINTER(_i64)]),
    "wsb_sim_create": (ctypes.c_int, [ctypes.POINTER(wsb_config), ctypes.POINTER(_vp)]),
    "wsb_sim_destroy": (None, [_vp]),
    "wsb_sim_initialize": (ctypes.c_int, [_vp]),
    "wsb_sim_current_grid": (_vp, [_vp]),
    "wsb_sim_step": (ctypes.c_int, [_vp]),
    "wsb_sim_step_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wsb_sim_run": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(_i32)]),
    "wsb_sim_run_until": (ctypes.c_int, [_vp, _dbl, ctypes.POINTER(_i32)]),
    "wsb_sim_advance_async": (ctypes.c_int, [_vp, _i32]),
    "wsb_sim_synchronize": (ctypes.c_int, [_vp]),
    "wsb_sim_last_run_device_ms": (ctypes.c_int, [_vp, ctypes.POINTER(_dbl)]),
    "wsb_sim_get_time": (_dbl, [_vp]),
    "wsb_sim_get_step": (_i32, [_vp]),
    "wsb_sim_get_dt": (_dbl, [_vp]),
    "wsb_sim_set_dt": (ctypes.c_int, [_vp, _dbl]),
    "wsb_sim_get_config": (ctypes.c_int, [_vp, ctypes.POINTER(wsb_config)]),
    "wsb_sim_get_metrics": (ctypes.c_int, [_vp, ctypes.POINTER(wsb_metrics)]),
    "wsb_sim_reset_metrics": (ctypes.c_int, [_vp]),
    "wsb_sim_local_rows": (ctypes.c_int, [_vp, ctypes.POINTER(_i32), ctypes.POINTER(_i32)]),
    "wsb_sim_kernel_name": (ctypes.c_char_p, [_vp]),
    "wsb_sim_mass_energy": (ctypes.c_int, [_vp, ctypes.POINTER(_dbl), ctypes.POINTER(_dbl)]),
    "wsb_sim_time_halo_exchange": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(_dbl), ctypes.POINTER(_i64)]),
    "wsb_exact_division_reciprocal": (ctypes.c_int, [_dbl, _i32, ctypes.POINTER(_dbl)]),
    "wsb_ic_apply": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.POINTER(_dbl), _i32, ctypes.c_uint32, ctypes.c_char_p]),
    "wsb_ic_fill_host": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(_dbl), _i32, ctypes.c_uint32, ctypes.c_char_p,
                                        _i32, _i32, _dbl, _dbl, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def load_library(path=None):
    """dlopen libweather_b200.so and bind every C-ABI symbol. Raises if the library is missing."""
    global _lib
    if _lib is None:
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise ImportError(
                f"{p} is missing: build it with `make -C nvidia-jetson-workload_b200` "
                "(or python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback"
            )
        lib = ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load_library().wsb_last_error().decode()


def _check(status):
    if status == WSB_OK:
        return
    msg = last_error()
    if status == WSB_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    raise RuntimeError(msg)


def device_count():
    n = ctypes.c_int(0)
    _check(load_library().wsb_device_count(ctypes.byref(n)))
    return n.value


def device_capabilities(device_id=0):
    caps = wsb_device_caps()
    _check(load_library().wsb_device_capabilities(device_id, ctypes.byref(caps)))
    return caps


def partition_rows(grid_height, nranks, rank):
    """(row0, nrows) of the slab `rank` owns -- same arithmetic as wsb_sim_create."""
    r0, n = ctypes.c_int32(), ctypes.c_int32()
    _check(load_library().wsb_partition_rows(grid_height, nranks, rank, ctypes.byref(r0), ctypes.byref(n)))
    return r0.value, n.value


def nccl_unique_id():
    buf = ctypes.create_string_buffer(NCCL_UNIQUE_ID_BYTES)
    _check(load_library().wsb_nccl_get_unique_id(buf))
    return buf.raw


def exact_division_reciprocal(divisor, dtype=np.float32):
    """RN(1/divisor) if the three-operation division is proven exact for this divisor (fp32: by exhaustion), else 0."""
    r = ctypes.c_double()
    _check(load_library().wsb_exact_division_reciprocal(float(divisor), WSB_F64 if np.dtype(dtype) == np.float64 else WSB_F32,
                                                        ctypes.byref(r)))
    return r.value


def pinned_empty(shape, dtype=np.float32):
    """numpy array backed by page-locked host memory (full-rate H2D/D2H). Keep the array alive while in use."""
    lib = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    ptr = ctypes.c_void_p()
    _check(lib.wsb_host_alloc(max(n, 1) * dtype.itemsize, ctypes.byref(ptr)))
    buf = (ctypes.c_char * (max(n, 1) * dtype.itemsize)).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = ptr.value
    return arr


_PINNED = {}


def pinned_free(arr):
    addr = arr.__array_interface__["data"][0]
    ptr = _PINNED.pop(addr, None)
    if ptr is not None:
        _check(load_library().wsb_host_free(ctypes.c_void_p(ptr)))


def _np_dtype(code):
    return np.float64 if code == WSB_F64 else np.float32


class DeviceField:
    """Holder of a `__cuda_array_interface__` dict; keeps the owning grid alive."""

    def __init__(self, iface, owner):
        self.__cuda_array_interface__ = iface
        self._owner = owner


class Grid:
    """Device-resident WeatherGrid (weather_sim.hpp:254-412). Getters return fresh host copies."""

    def __init__(self, width=None, height=None, num_levels=1, dx=1.0, dy=1.0, dtype=np.float32, device_id=0,
                 _handle=None, _owner=None):
        self._lib = load_library()
        self._owner = _owner
        if _handle is not None:
            self._h = _handle
        else:
            h = ctypes.c_void_p()
            code = WSB_F64 if np.dtype(dtype) == np.float64 else WSB_F32
            _check(self._lib.wsb_grid_create(width, height, num_levels, dx, dy, code, device_id, ctypes.byref(h)))
            self._h = h.value

    @property
    def info(self):
        gi = wsb_grid_info()
        _check(self._lib.wsb_grid_get_info(self._h, ctypes.byref(gi)))
        return gi

    @property
    def shape(self):
        gi = self.info
        return (gi.num_levels, gi.height, gi.width)

    @property
    def dtype(self):
        return np.dtype(_np_dtype(self.info.dtype))

    def set_field(self, name, arr):
        gi = self.info
        a = np.asarray(arr)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(_np_dtype(gi.dtype))
        a = np.ascontiguousarray(a)
        if a.ndim == 2:
            lv, (r, c) = 1, a.shape
        elif a.ndim == 3:
            lv, r, c = a.shape
        else:
            raise RuntimeError("Number of dimensions must be 2")  # python_bindings.cpp:63-65
        code = WSB_F64 if a.dtype == np.float64 else WSB_F32
        _check(self._lib.wsb_grid_set_field(self._h, FIELD[name], a.ctypes.data, code, lv, r, c))

    def get_field(self, name, out=None):
        gi = self.info
        shape = (gi.height, gi.width) if gi.num_levels == 1 else (gi.num_levels, gi.height, gi.width)
        if out is None:
            out = np.empty(shape, dtype=_np_dtype(gi.dtype))
        assert out.flags.c_contiguous and out.shape == shape
        code = WSB_F64 if out.dtype == np.float64 else WSB_F32
        _check(self._lib.wsb_grid_get_field(self._h, FIELD[name], out.ctypes.data, code, gi.num_levels, gi.height,
                                            gi.width))
        return out

    def device_view(self, name):
        """Zero-copy view of a field in HBM: an object exposing
        `__cuda_array_interface__` (version 3) that cupy.asarray / torch.as_tensor wrap without a copy. Rows are
        pitched (128-byte aligned); the view is valid until the next step() (the u, v, h planes rotate)."""
        ptr, pitch = ctypes.c_void_p(), ctypes.c_int64()
        _check(self._lib.wsb_grid_device_pointer(self._h, FIELD[name], ctypes.byref(ptr), ctypes.byref(pitch)))
        gi = self.info
        es = 8 if gi.dtype == WSB_F64 else 4
        lead = 5  # rows between consecutive levels beyond H: 2 * (4 ghost + 1 guard), see DESIGN.md section 3
        shape = (gi.height, gi.width) if gi.num_levels == 1 else (gi.num_levels, gi.height, gi.width)
        strides = (pitch.value * es, es) if gi.num_levels == 1 else \
            ((gi.height + 2 * lead) * pitch.value * es, pitch.value * es, es)
        return DeviceField({"shape": shape, "typestr": "<f8" if es == 8 else "<f4", "data": (ptr.value, False),
                            "version": 3, "strides": strides}, self)

    def apply_ic(self, name, params=(), seed=0, profile=None):
        """InitialCondition::initialize(grid) of the reference (initial_conditions.cpp:59-535) by name; params in
        constructor order. On a row slab it evaluates this rank's rows of the global initial condition."""
        arr = (ctypes.c_double * max(len(params), 1))(*params)
        _check(self._lib.wsb_ic_apply(self._h, name.encode(), arr, len(params), int(seed),
                                      profile.encode() if profile else None))

    def reset(self):
        _check(self._lib.wsb_grid_reset(self._h))

    def set_spacing(self, dx, dy):
        _check(self._lib.wsb_grid_set_spacing(self._h, dx, dy))

    def calculate_diagnostics(self):
        _check(self._lib.wsb_grid_calculate_diagnostics(self._h))

    def swap(self, other):
        _check(self._lib.wsb_grid_swap(self._h, other._h))

    def close(self):
        if self._h and self._owner is None:
            self._lib.wsb_grid_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Simulation:
    """Device-resident WeatherSimulation (weather_sim.hpp:417-544) over the C-ABI."""

    def __init__(self, width, height, model="shallow_water", integrator="rk4", dx=1.0, dy=1.0, dt=0.01,
                 gravity=9.81, coriolis_f=0.0, max_time=10.0, dtype=np.float32, num_levels=1, device_id=0,
                 rk4_classical=False, kernel="auto", rank=0, nranks=1, nccl_id=None, arith="strict", extended=None):
        """extended=(beta, viscosity, diffusivity): the non-reference beta-plane / viscous tendencies
        (WSB_PHYSICS_EXTENDED, include/weather_b200.h)."""
        self._lib = load_library()
        cfg = wsb_config()
        cfg.struct_size = ctypes.sizeof(wsb_config)
        cfg.model = MODEL[model] if isinstance(model, str) else int(model)
        cfg.integration_method = INTEGRATOR[integrator] if isinstance(integrator, str) else int(integrator)
        cfg.grid_width, cfg.grid_height, cfg.num_levels = int(width), int(height), int(num_levels)
        cfg.dx, cfg.dy, cfg.dt = dx, dy, dt
        cfg.gravity, cfg.coriolis_f, cfg.max_time = gravity, coriolis_f, max_time
        cfg.dtype = WSB_F64 if np.dtype(dtype) == np.float64 else WSB_F32
        cfg.device_id = device_id
        cfg.rk4_mode = 1 if rk4_classical else 0
        cfg.kernel_variant = KERNEL[kernel] if isinstance(kernel, str) else int(kernel)
        cfg.rank, cfg.nranks = rank, nranks
        cfg.arith_mode = ARITH[arith] if isinstance(arith, str) else int(arith)
        if extended is not None:
            cfg.physics_mode = 1
            cfg.beta, cfg.viscosity, cfg.diffusivity = (float(x) for x in extended)
        self._id_buf = None
        if nranks > 1:
            self._id_buf = ctypes.create_string_buffer(bytes(nccl_id), NCCL_UNIQUE_ID_BYTES)
            cfg.nccl_unique_id = ctypes.cast(self._id_buf, ctypes.c_void_p)
        h = ctypes.c_void_p()
        _check(self._lib.wsb_sim_create(ctypes.byref(cfg), ctypes.byref(h)))
        self._h = h.value
        self.dtype = np.dtype(dtype)
        self.grid = Grid(_handle=self._lib.wsb_sim_current_grid(self._h), _owner=self)

    # -- state ----------------------------------------------------------------------------
    def set_state(self, u=None, v=None, h=None, **others):
        for name, arr in dict(u=u, v=v, h=h, **others).items():
            if arr is not None:
                self.grid.set_field(name, arr)

    def get_field(self, name, out=None):
        return self.grid.get_field(name, out)

    def state(self, names=("u", "v", "h")):
        return {n: self.grid.get_field(n) for n in names}

    @property
    def local_rows(self):
        r0, n = ctypes.c_int32(), ctypes.c_int32()
        _check(self._lib.wsb_sim_local_rows(self._h, ctypes.byref(r0), ctypes.byref(n)))
        return r0.value, n.value

    # -- stepping -------------------------------------------------------------------------
    def initialize(self):
        _check(self._lib.wsb_sim_initialize(self._h))

    def step(self, n=1):
        """n reference step() calls (no max_time check), synchronised at the end."""
        _check(self._lib.wsb_sim_advance_async(self._h, int(n)))
        _check(self._lib.wsb_sim_synchronize(self._h))

    def step_host(self, u, v, h, out_u=None, out_v=None, out_h=None):
        """One step with host-resident state: streams u, v, h through the GPU in row slabs (upload, step and
        download overlap) and returns the new (u, v, h). Arrays must have the simulation's dtype and shape;
        page-locked arrays (pinned_empty) give full overlap. Outputs default to new arrays."""
        ins = [np.ascontiguousarray(a, dtype=self.dtype) for a in (u, v, h)]
        outs = [np.empty_like(ins[0]) if o is None else o for o in (out_u, out_v, out_h)]
        exp = self.grid.shape if self.grid.shape[0] > 1 else self.grid.shape[1:]
        for a in ins + outs:
            if tuple(a.shape) != tuple(exp) or a.dtype != self.dtype or not a.flags.c_contiguous:
                raise RuntimeError("Array dimensions must match field dimensions")
        _check(self._lib.wsb_sim_step_host(self._h, *[a.ctypes.data for a in ins], *[a.ctypes.data for a in outs]))
        return tuple(outs)

    def advance_async(self, n):
        _check(self._lib.wsb_sim_advance_async(self._h, int(n)))

    def synchronize(self):
        _check(self._lib.wsb_sim_synchronize(self._h))

    def run(self, n):
        done = ctypes.c_int32(0)
        _check(self._lib.wsb_sim_run(self._h, int(n), ctypes.byref(done)))
        return done.value

    def run_until(self, max_time):
        done = ctypes.c_int32(0)
        _check(self._lib.wsb_sim_run_until(self._h, float(max_time), ctypes.byref(done)))
        return done.value

    @property
    def last_run_device_ms(self):
        ms = ctypes.c_double()
        _check(self._lib.wsb_sim_last_run_device_ms(self._h, ctypes.byref(ms)))
        return ms.value

    @property
    def time(self):
        return self._lib.wsb_sim_get_time(self._h)

    @property
    def steps(self):
        return self._lib.wsb_sim_get_step(self._h)

    @property
    def dt(self):
        return self._lib.wsb_sim_get_dt(self._h)

    def set_dt(self, dt):
        _check(self._lib.wsb_sim_set_dt(self._h, float(dt)))

    @property
    def metrics(self):
        m = wsb_metrics()
        _check(self._lib.wsb_sim_get_metrics(self._h, ctypes.byref(m)))
        return m

    def reset_metrics(self):
        _check(self._lib.wsb_sim_reset_metrics(self._h))

    @property
    def kernel_name(self):
        return self._lib.wsb_sim_kernel_name(self._h).decode()

    def mass_energy(self):
        m, e = ctypes.c_double(), ctypes.c_double()
        _check(self._lib.wsb_sim_mass_energy(self._h, ctypes.byref(m), ctypes.byref(e)))
        return m.value, e.value

    def time_halo_exchange(self, reps=100):
        """(microseconds per bare ghost-row exchange, bytes per neighbour per direction); (0, bytes) on one rank."""
        us, nbytes = ctypes.c_double(), ctypes.c_int64()
        _check(self._lib.wsb_sim_time_halo_exchange(self._h, int(reps), ctypes.byref(us), ctypes.byref(nbytes)))
        return us.value, nbytes.value

    def close(self):
        if getattr(self, "_h", None):
            self.grid._h = None
            self._lib.wsb_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
