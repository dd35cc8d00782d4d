"""weather_sim -- the Weather Simulation Python package, backed by the B200-native libweather_b200.

Same import surface as the reference package (src/weather-sim/python/__init__.py:12-24): the names below
come from the `pyweather_sim` extension module, which here is the pybind11 shim over the C-ABI
(csrc/pyweather_sim.cpp). The reference's plotting helpers (`visualization`, matplotlib) are out of scope.

Unlike the reference there is no mock fallback: if the extension (or a CUDA device) is missing, imports
and constructors raise instead of silently doing nothing.
"""
from .weather_simulation import (  # noqa: F401
    WeatherSimulation,
    WeatherGrid,
    SimulationConfig,
    PerformanceMetrics,
    OutputConfig,
    OutputManager,
    IntegrationMethod,
    SimulationModel,
    ComputeBackend,
    GridType,
    BoundaryCondition,
    DeviceType,
    OutputFormat,
    InitialConditionFactory,
    AdaptiveKernelManager,
    WeatherSimulationWrapper,
    create_initial_condition,
    get_available_initial_conditions,
    is_cuda_available,
    get_device_info,
)

__version__ = "0.1.0"
