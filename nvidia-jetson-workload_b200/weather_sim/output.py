"""Output managers for the `OutputManager` hook of the simulation (weather_sim.hpp:549-570).

The reference declares CSV / NetCDF / VTK managers (output_manager.hpp:51-269) but ships no implementation
(SURVEY.md section 2.1); its binding comment says "implementations would be in Python". This module is
that: a CSV writer following the declared `OutputConfig` fields and a writer for the `weather.proto` slice
stream (`src/proto/weather.proto:57-118`), both built on the Python-subclassable `OutputManager` of the B200
shim. They run on the host at `output_interval` boundaries only -- the stepping between two outputs stays one
asynchronous stream of kernels.
"""
import os

import numpy as np

from . import proto_stream
from .pyweather_sim import OutputConfig, OutputFormat, OutputManager

_GETTERS = {
    "velocity": lambda g: dict(zip(("u", "v"), g.get_velocity_field())),
    "height": lambda g: {"height": g.get_height_field()},
    "pressure": lambda g: {"pressure": g.get_pressure_field()},
    "temperature": lambda g: {"temperature": g.get_temperature_field()},
    "humidity": lambda g: {"humidity": g.get_humidity_field()},
    "vorticity": lambda g: {"vorticity": g.get_vorticity_field()},
    "divergence": lambda g: {"divergence": g.get_divergence_field()},
}
_DIAGNOSTICS = ("vorticity", "divergence")


class CSVOutputManager(OutputManager):
    """One CSV file per output: `x,y,<field columns>` rows in row-major order (output_manager.hpp:35-97)."""

    def __init__(self, config=None):
        super().__init__()
        self.config = config or OutputConfig()
        if self.config.format != OutputFormat.CSV:
            raise ValueError("CSVOutputManager writes OutputFormat.CSV only")
        self.output_count = 0
        self.files = []

    def get_config(self):
        return self.config

    def set_config(self, config):
        self.config = config

    def initialize(self, simulation):
        os.makedirs(self.config.output_dir, exist_ok=True)
        self.output_count = 0
        self.files = []

    def _columns(self, grid):
        cols = {}
        for name in self.config.fields:
            if name in _DIAGNOSTICS and not self.config.include_diagnostics:
                continue
            if name in _GETTERS:
                cols.update(_GETTERS[name](grid))
        return cols

    def write_output(self, simulation):
        grid = simulation.get_current_grid()
        cols = self._columns(grid)
        h, w = grid.get_height(), grid.get_width()
        yy, xx = np.mgrid[0:h, 0:w]
        table = np.column_stack([xx.ravel(), yy.ravel()] + [np.asarray(a).reshape(-1, h * w)[0] for a in cols.values()])
        suffix = ".csv.gz" if self.config.compress else ".csv"
        path = os.path.join(self.config.output_dir,
                            f"{self.config.prefix}_{simulation.get_current_step():06d}{suffix}")
        header = "x,y," + ",".join(cols)
        np.savetxt(path, table, delimiter=",", header=header, comments="",
                   fmt=["%d", "%d"] + ["%.9g"] * len(cols))
        self.files.append(path)
        self.output_count += 1

    def finalize(self, simulation):
        pass


class ProtoSliceOutputManager(OutputManager):
    """Appends one length-delimited `WeatherSimUpdate` (with the current `AtmosphericSlice` of every `stride`-th
    cell of level `z_level`) per output interval to `<output_dir>/<prefix>.pb` (weather.proto:69-74,104-118).
    The per-cell message format of the schema is meant for visualisation-sized slices: pick `stride` accordingly."""

    def __init__(self, config=None, run_id="run", stride=1, z_level=0, total_steps=None):
        super().__init__()
        self.config = config or OutputConfig()
        self.run_id, self.stride, self.z_level, self.total_steps = run_id, int(stride), int(z_level), total_steps
        self.path = None
        self.output_count = 0

    def get_config(self):
        return self.config

    def set_config(self, config):
        self.config = config

    def initialize(self, simulation):
        os.makedirs(self.config.output_dir, exist_ok=True)
        self.path = os.path.join(self.config.output_dir, f"{self.config.prefix}.pb")
        open(self.path, "wb").close()
        self.output_count = 0

    def write_output(self, simulation):
        grid = simulation.get_current_grid()
        k, z = self.stride, self.z_level

        def pick(a):
            a = np.asarray(a)
            return (a[z] if a.ndim == 3 else a)[::k, ::k]

        u, v = grid.get_velocity_field()
        sl = proto_stream.encode_slice(z, pick(u), pick(v), pick(grid.get_temperature_field()),
                                       pick(grid.get_pressure_field()), pick(grid.get_humidity_field()))
        done = 100.0 * simulation.get_current_step() / self.total_steps if self.total_steps else 0.0
        msg = proto_stream.encode_update(self.run_id, simulation.get_current_time(), done, sl)
        with open(self.path, "ab") as f:
            f.write(proto_stream.frame(msg))
        self.output_count += 1

    def finalize(self, simulation):
        pass
