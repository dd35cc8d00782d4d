"""High-level Python API of the Weather Simulation workload on top of the B200-native extension.

Mirrors the public names and call semantics of the reference's wrapper module
(src/weather-sim/python/weather_simulation.py): `WeatherSimulationWrapper` (:194-371),
`create_initial_condition` (:376-453), `get_available_initial_conditions` (:455-469), `is_cuda_available`
(:471-481), `get_device_info` (:483-520). The reference's pure-Python mock classes (:30-189) are NOT
reproduced: a missing extension is an ImportError here, never a silent no-op simulation.
"""
import time as _time

from .pyweather_sim import (  # noqa: F401
    AdaptiveKernelManager,
    AtmosphericProfileInitialCondition,
    BoundaryCondition,
    BreakingWaveInitialCondition,
    ComputeBackend,
    DeviceType,
    FrontInitialCondition,
    GridType,
    InitialConditionFactory,
    IntegrationMethod,
    JetStreamInitialCondition,
    MountainInitialCondition,
    OutputConfig,
    OutputFormat,
    OutputManager,
    PerformanceMetrics,
    RandomInitialCondition,
    SimulationConfig,
    SimulationModel,
    UniformInitialCondition,
    VortexInitialCondition,
    WeatherGrid,
    WeatherSimulation,
    ZonalFlowInitialCondition,
    register_all_initial_conditions,
)

register_all_initial_conditions()

_MODELS = {
    "shallow_water": SimulationModel.ShallowWater,
    "barotropic": SimulationModel.Barotropic,
    "primitive": SimulationModel.PrimitiveEquations,
    "general": SimulationModel.General,
}
_METHODS = {
    "euler": IntegrationMethod.ExplicitEuler,
    "rk2": IntegrationMethod.RungeKutta2,
    "rk4": IntegrationMethod.RungeKutta4,
    "adams_bashforth": IntegrationMethod.AdamsBashforth,
    "semi_implicit": IntegrationMethod.SemiImplicit,
}
_BACKENDS = {
    "cuda": ComputeBackend.CUDA,
    "cpu": ComputeBackend.CPU,
    "hybrid": ComputeBackend.Hybrid,
    "adaptive": ComputeBackend.AdaptiveHybrid,
}

# constructor keyword -> (class, ordered (name, default) pairs), as bound at python_bindings.cpp:291-329
_IC_TABLE = {
    "uniform": (UniformInitialCondition, (("u", 0.0), ("v", 0.0), ("h", 10.0), ("p", 1000.0), ("t", 300.0), ("q", 0.0))),
    "random": (RandomInitialCondition, (("seed", 0), ("amplitude", 1.0))),
    "zonal_flow": (ZonalFlowInitialCondition, (("u_max", 10.0), ("h_mean", 10.0), ("beta", 0.1))),
    "vortex": (VortexInitialCondition,
               (("x_center", 0.5), ("y_center", 0.5), ("radius", 0.1), ("strength", 10.0), ("h_mean", 10.0))),
    "jet_stream": (JetStreamInitialCondition, (("y_center", 0.5), ("width", 0.1), ("strength", 10.0), ("h_mean", 10.0))),
    "breaking_wave": (BreakingWaveInitialCondition, (("amplitude", 1.0), ("wavelength", 0.2), ("h_mean", 10.0))),
    "front": (FrontInitialCondition,
              (("y_position", 0.5), ("width", 0.05), ("temp_difference", 10.0), ("wind_shear", 5.0))),
    "mountain": (MountainInitialCondition,
                 (("x_center", 0.3), ("y_center", 0.5), ("radius", 0.1), ("height", 1.0), ("u_base", 5.0))),
    "atmospheric_profile": (AtmosphericProfileInitialCondition, (("profile_name", "standard"),)),
}


def _pick(table, value, default):
    if isinstance(value, str):
        return table.get(value.lower(), default)
    return value


class WeatherSimulationWrapper:
    """Convenience driver: builds a SimulationConfig from keywords, owns the simulation, keeps snapshots."""

    def __init__(self, width=256, height=256, model="shallow_water", dt=0.01, integration_method="rk4",
                 backend="adaptive", device_id=0, threads=0, output_interval=10, output_path="./output",
                 double_precision=False, num_levels=1):
        cfg = SimulationConfig()
        cfg.grid_width = width
        cfg.grid_height = height
        cfg.num_levels = num_levels
        cfg.dt = dt
        cfg.output_interval = output_interval
        cfg.output_path = output_path
        cfg.device_id = device_id
        cfg.num_threads = threads
        cfg.double_precision = double_precision
        cfg.model = _pick(_MODELS, model, SimulationModel.ShallowWater)
        cfg.integration_method = _pick(_METHODS, integration_method, IntegrationMethod.RungeKutta4)
        cfg.compute_backend = _pick(_BACKENDS, backend, ComputeBackend.AdaptiveHybrid)
        self.config = cfg
        self.simulation = WeatherSimulation(cfg)
        self.initialized = False
        self.output_data = []

    def set_initial_condition(self, condition_name, **kwargs):
        ic = create_initial_condition(condition_name, **kwargs)
        if ic:
            self.simulation.set_initial_condition(ic)

    def initialize(self):
        self.simulation.initialize()
        self.initialized = True

    def step(self):
        if not self.initialized:
            self.initialize()
        self.simulation.step()
        every = self.config.output_interval
        if every > 0 and self.simulation.get_current_step() % every == 0:
            self._store_output()

    def run(self, steps):
        if not self.initialized:
            self.initialize()
        t0 = _time.time()
        self.simulation.run(steps)
        elapsed = (_time.time() - t0) * 1000.0
        print(f"Completed {steps} steps in {elapsed:.2f} ms ({elapsed / steps:.2f} ms/step)")

    def run_until(self, max_time):
        if not self.initialized:
            self.initialize()
        t0 = _time.time()
        self.simulation.run_until(max_time)
        elapsed = (_time.time() - t0) * 1000.0
        steps = max(self.simulation.get_current_step(), 1)
        print(f"Reached time {max_time} in {elapsed:.2f} ms ({elapsed / steps:.2f} ms/step)")

    def get_grid(self):
        return self.simulation.get_current_grid()

    def get_metrics(self):
        return self.simulation.get_performance_metrics()

    def get_output_data(self):
        return self.output_data

    def _store_output(self):
        grid = self.simulation.get_current_grid()
        u, v = grid.get_velocity_field()  # getters return fresh host copies
        self.output_data.append({
            "time": self.simulation.get_current_time(),
            "step": self.simulation.get_current_step(),
            "u": u,
            "v": v,
            "height": grid.get_height_field(),
            "vorticity": grid.get_vorticity_field(),
        })


def create_initial_condition(name, **kwargs):
    """Initial condition by name with keyword parameters; unknown names go through the factory; None on failure."""
    try:
        if name in _IC_TABLE:
            cls, spec = _IC_TABLE[name]
            return cls(*[kwargs.get(key, default) for key, default in spec])
        return InitialConditionFactory.get_instance().create_initial_condition(name)
    except Exception as exc:
        print(f"Error creating initial condition '{name}': {exc}")
        return None


def get_available_initial_conditions():
    return InitialConditionFactory.get_instance().get_available_initial_conditions()


def is_cuda_available():
    return AdaptiveKernelManager.get_instance().is_cuda_available()


_DEVICE_TYPE_NAMES = {
    DeviceType.Unknown: "Unknown",
    DeviceType.CPU: "CPU",
    DeviceType.JetsonOrinNX: "Jetson Orin NX",
    DeviceType.T4: "NVIDIA T4",
    DeviceType.HighEndGPU: "High-End GPU",
    DeviceType.OtherGPU: "Other GPU",
}


def get_device_info():
    try:
        mgr = AdaptiveKernelManager.get_instance()
        mgr.initialize()
        caps = mgr.get_device_capabilities()
        return {
            "device_type": _DEVICE_TYPE_NAMES.get(caps.device_type, "Unknown"),
            "device_name": caps.device_name,
            "compute_capability": f"{caps.compute_capability_major}.{caps.compute_capability_minor}",
            "cuda_cores": caps.cuda_cores,
            "multiprocessors": caps.multiprocessors,
            "global_memory_mb": caps.global_memory / (1024 * 1024),
            "compute_power_ratio": caps.compute_power_ratio,
            "cuda_available": mgr.is_cuda_available(),
        }
    except Exception as exc:
        return {"device_type": "Unknown", "device_name": "Unknown", "error": str(exc), "cuda_available": False}
