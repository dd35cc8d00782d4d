// pyweather_sim -- pybind11 module with the reference's Python surface (python_bindings.cpp:116-372),
// implemented on top of the C-ABI of libweather_b200.so (include/weather_b200.h).
//
// The reference's Python package (`weather_simulation.py:16-29`) imports this module by name; with this
// file's build product next to it, every field lives in B200 HBM and every step is a CUDA kernel. There is
// no CPU path here: construction raises RuntimeError when no CUDA device is usable.
//
// Differences from the reference surface (all additive):
//   * SimulationConfig.double_precision is honoured (the reference ignores it): True selects
//     the fp64 kernels and float64 arrays.
//   * SimulationConfig.num_levels > 1 runs that many independent 2-D levels; arrays are (L, H, W).
//   * SimulationConfig.rk4_classical / kernel_variant / folded_arithmetic, WeatherSimulation.get_kernel_name(),
//     WeatherSimulation.mass_energy(): B200-specific knobs and diagnostics.
//   * SimulationConfig.rank / nranks / nccl_unique_id + pyweather_sim.nccl_unique_id(): row slabs over the GPUs of
//     one box, one process per GPU; get_current_grid() is then this rank's slab (get_local_rows()).
//   * OutputManager can be subclassed from Python (the reference binds the abstract base only).
//   * get_current_grid() always views the CURRENT state (the reference's handle goes stale on odd steps).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "weather_b200.h"

namespace py = pybind11;

namespace {

// ---- enums (weather_sim.hpp:30-76, gpu_adaptability.hpp:23-30, output_manager.hpp:24-30) -----------------
enum class SimulationModel { ShallowWater, Barotropic, PrimitiveEquations, General };
enum class BoundaryCondition { Periodic, Reflective, Outflow, Custom };
enum class IntegrationMethod { ExplicitEuler, RungeKutta2, RungeKutta4, AdamsBashforth, SemiImplicit };
enum class GridType { Cartesian, Staggered, Icosahedral, SphericalHarmonic };
enum class ComputeBackend { CUDA, CPU, Hybrid, AdaptiveHybrid };
enum class DeviceType { Unknown, CPU, JetsonOrinNX, T4, HighEndGPU, OtherGPU };
enum class OutputFormat { CSV, NetCDF, VTK, PNG, Custom };

[[noreturn]] void raise_status(int status) {
    const std::string msg = wsb_last_error();
    if (status == WSB_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);  // -> ValueError
    throw std::runtime_error(msg);                                             // -> RuntimeError
}

inline void check(int status) {
    if (status != WSB_OK) raise_status(status);
}

// ---- plain structs ------------------------------------------------------------------------------------
struct SimulationConfig {  // weather_sim.hpp:155-191, same defaults
    SimulationModel model = SimulationModel::ShallowWater;
    GridType grid_type = GridType::Staggered;
    IntegrationMethod integration_method = IntegrationMethod::RungeKutta4;
    BoundaryCondition boundary_condition = BoundaryCondition::Periodic;
    int32_t grid_width = 256;
    int32_t grid_height = 256;
    int32_t num_levels = 1;
    float dx = 1.0f;
    float dy = 1.0f;
    float dt = 0.01f;
    float gravity = 9.81f;
    float coriolis_f = 0.0f;
    float beta = 0.0f;
    float viscosity = 0.0f;
    float diffusivity = 0.0f;
    ComputeBackend compute_backend = ComputeBackend::CUDA;
    bool double_precision = false;
    int device_id = 0;
    int num_threads = 0;
    float max_time = 10.0f;
    int max_steps = 1000;
    int output_interval = 10;
    std::string output_path = "./output";
    unsigned int random_seed = std::random_device{}();
    // B200 additions
    bool rk4_classical = false;
    int kernel_variant = WSB_KERNEL_AUTO;
    bool folded_arithmetic = false;  // wsb_arith_mode: WSB_ARITH_FOLDED (opt-in, see include/weather_b200.h)
    bool extended_physics = false;   // wsb_physics_mode: WSB_PHYSICS_EXTENDED -- beta, viscosity, diffusivity are READ
    // row-slab decomposition over the GPUs of one box (one process per GPU): this process is `rank` of `nranks`,
    // grid_height is the GLOBAL height, nccl_unique_id the 128 bytes of pyweather_sim.nccl_unique_id() made on rank 0
    // and handed to every rank (e.g. torch.distributed.broadcast_object_list / mpi4py bcast)
    int rank = 0;
    int nranks = 1;
    std::string nccl_unique_id;
};

struct PerformanceMetrics {  // weather_sim.hpp:196-223
    double total_time_ms = 0.0;
    double compute_time_ms = 0.0;
    double memory_transfer_time_ms = 0.0;
    double io_time_ms = 0.0;
    int num_steps = 0;
    double halo_time_ms = 0.0;
    uint64_t kernel_launches = 0;
    void reset() { *this = PerformanceMetrics(); }
    void print() const {
        std::cout << "Performance Metrics:" << std::endl;
        std::cout << "  Total time: " << total_time_ms << " ms" << std::endl;
        std::cout << "  Compute time: " << compute_time_ms << " ms (" << (compute_time_ms / total_time_ms * 100.0)
                  << "%)" << std::endl;
        std::cout << "  Memory transfer time: " << memory_transfer_time_ms << " ms ("
                  << (memory_transfer_time_ms / total_time_ms * 100.0) << "%)" << std::endl;
        std::cout << "  I/O time: " << io_time_ms << " ms (" << (io_time_ms / total_time_ms * 100.0) << "%)" << std::endl;
        std::cout << "  Steps: " << num_steps << std::endl;
        std::cout << "  Time per step: " << (total_time_ms / num_steps) << " ms" << std::endl;
    }
};

struct OutputConfig {  // output_manager.hpp:35-46
    std::string output_dir = "./output";
    std::string prefix = "weather_sim";
    OutputFormat format = OutputFormat::CSV;
    int output_interval = 10;
    bool compress = false;
    bool include_diagnostics = true;
    std::vector<std::string> fields = {"velocity", "height", "pressure", "temperature", "humidity", "vorticity",
                                       "divergence"};
};

struct DeviceCapabilities {  // gpu_adaptability.hpp:35-88
    DeviceType device_type = DeviceType::Unknown;
    int compute_capability_major = 0;
    int compute_capability_minor = 0;
    int cuda_cores = 0;
    int multiprocessors = 0;
    size_t global_memory = 0;
    size_t shared_memory_per_block = 0;
    int max_threads_per_block = 0;
    int max_threads_per_multiprocessor = 0;
    int clock_rate_khz = 0;
    int memory_clock_rate_khz = 0;
    int memory_bus_width = 0;
    float compute_power_ratio = 0.0f;
    std::string device_name;

    std::string summary() const {  // same lines as gpu_adaptability.cpp:500-540
        static const char *names[] = {"Unknown", "CPU", "Jetson Orin NX", "NVIDIA T4", "High-End GPU", "Other GPU"};
        std::stringstream ss;
        ss << "Device: " << device_name << std::endl;
        ss << "Device Type: " << names[static_cast<int>(device_type)] << std::endl;
        if (device_type != DeviceType::CPU) {
            ss << "Compute Capability: " << compute_capability_major << "." << compute_capability_minor << std::endl;
            ss << "CUDA Cores: " << cuda_cores << std::endl;
            ss << "Multiprocessors: " << multiprocessors << std::endl;
            ss << "Global Memory: " << (global_memory / (1024 * 1024)) << " MB" << std::endl;
            ss << "Shared Memory Per Block: " << (shared_memory_per_block / 1024) << " KB" << std::endl;
            ss << "Max Threads Per Block: " << max_threads_per_block << std::endl;
            ss << "Clock Rate: " << (clock_rate_khz / 1000) << " MHz" << std::endl;
        }
        ss << "Compute Power Ratio: " << compute_power_ratio << "x" << std::endl;
        return ss.str();
    }
};

// ---- WeatherGrid ---------------------------------------------------------------------------------------
class WeatherGrid {
public:
    WeatherGrid(int32_t width, int32_t height, int32_t num_levels) {
        check(wsb_grid_create(width, height, num_levels, 1.0, 1.0, WSB_F32, 0, &g_));
        owned_ = true;
    }
    explicit WeatherGrid(const SimulationConfig &c) {
        check(wsb_grid_create(c.grid_width, c.grid_height, c.num_levels, c.dx, c.dy,
                              c.double_precision ? WSB_F64 : WSB_F32, c.device_id, &g_));
        owned_ = true;
    }
    explicit WeatherGrid(wsb_grid *borrowed) : g_(borrowed), owned_(false) {}
    WeatherGrid(const WeatherGrid &) = delete;
    WeatherGrid &operator=(const WeatherGrid &) = delete;
    ~WeatherGrid() {
        if (owned_ && g_) wsb_grid_destroy(g_);
    }

    wsb_grid *handle() const { return g_; }
    wsb_grid_info info() const {
        wsb_grid_info gi;
        check(wsb_grid_get_info(g_, &gi));
        return gi;
    }
    void reset() { check(wsb_grid_reset(g_)); }
    int32_t width() const { return info().width; }
    int32_t height() const { return info().height; }
    int32_t levels() const { return info().num_levels; }
    float dx() const { return static_cast<float>(info().dx); }
    float dy() const { return static_cast<float>(info().dy); }
    void set_spacing(double dx, double dy) { check(wsb_grid_set_spacing(g_, dx, dy)); }
    void calculate_diagnostics() { check(wsb_grid_calculate_diagnostics(g_)); }

    py::array get(int field) const {
        const wsb_grid_info gi = info();
        std::vector<py::ssize_t> shape;
        if (gi.num_levels > 1) shape.push_back(gi.num_levels);
        shape.push_back(gi.height);
        shape.push_back(gi.width);
        if (gi.dtype == WSB_F64) {
            py::array_t<double> out(shape);
            check(wsb_grid_get_field(g_, field, out.mutable_data(), WSB_F64, gi.num_levels, gi.height, gi.width));
            return std::move(out);
        }
        py::array_t<float> out(shape);
        check(wsb_grid_get_field(g_, field, out.mutable_data(), WSB_F32, gi.num_levels, gi.height, gi.width));
        return std::move(out);
    }

    void set(int field, const py::array &arr) {
        const wsb_grid_info gi = info();
        const int want_dims = gi.num_levels > 1 ? 3 : 2;
        // forcecast like py::array_t<scalar_t, c_style> (python_bindings.cpp:60): anything castable is accepted;
        // float64 input stays float64 so that fp64 grids lose nothing
        const bool f64 = py::isinstance<py::array_t<double>>(arr) && gi.dtype == WSB_F64;
        py::array a = f64 ? py::array(py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(arr))
                          : py::array(py::array_t<float, py::array::c_style | py::array::forcecast>::ensure(arr));
        if (!a) throw std::runtime_error("array cannot be converted to a floating-point C-contiguous array");
        if (a.ndim() != want_dims && !(a.ndim() == 2 && gi.num_levels == 1))
            throw std::runtime_error("Number of dimensions must be 2");  // python_bindings.cpp:63-65
        const int64_t lv = a.ndim() == 3 ? a.shape(0) : 1;
        const int64_t rows = a.shape(a.ndim() - 2), cols = a.shape(a.ndim() - 1);
        check(wsb_grid_set_field(g_, field, a.data(), f64 ? WSB_F64 : WSB_F32, lv, rows, cols));
    }

    void set_velocity(const py::array &u, const py::array &v) {
        // validate both before touching the device (python_bindings.cpp:92-104)
        const wsb_grid_info gi = info();
        for (const py::array *a : {&u, &v}) {
            if (a->ndim() != (gi.num_levels > 1 ? 3 : 2)) throw std::runtime_error("Number of dimensions must be 2");
            if (a->shape(a->ndim() - 1) != gi.width || a->shape(a->ndim() - 2) != gi.height)
                throw std::runtime_error("Array dimensions must match field dimensions");
        }
        set(WSB_FIELD_U, u);
        set(WSB_FIELD_V, v);
    }

private:
    wsb_grid *g_ = nullptr;
    bool owned_ = false;
};

// ---- initial conditions (python_bindings.cpp:287-329) ----------------------------------------------------
class InitialCondition {
public:
    virtual ~InitialCondition() = default;
    virtual std::string name() const = 0;
    virtual void initialize(WeatherGrid &grid) const {
        check(wsb_ic_apply(grid.handle(), ic_name().c_str(), params_.empty() ? nullptr : params_.data(),
                           static_cast<int32_t>(params_.size()), seed_, profile_.empty() ? nullptr : profile_.c_str()));
    }

protected:
    virtual std::string ic_name() const { return name(); }
    std::vector<double> params_;
    uint32_t seed_ = 0;
    std::string profile_;
};

#define WSB_IC_CLASS(Cls, pyname)                       \
    class Cls : public InitialCondition {               \
    public:                                             \
        std::string name() const override { return pyname; }

WSB_IC_CLASS(UniformInitialCondition, "uniform")
    UniformInitialCondition(float u, float v, float h, float p, float t, float q) { params_ = {u, v, h, p, t, q}; }
};
WSB_IC_CLASS(RandomInitialCondition, "random")
    RandomInitialCondition(unsigned int seed, float amplitude) {
        seed_ = seed;
        params_ = {amplitude};
    }
};
WSB_IC_CLASS(ZonalFlowInitialCondition, "zonal_flow")
    ZonalFlowInitialCondition(float u_max, float h_mean, float beta) { params_ = {u_max, h_mean, beta}; }
};
WSB_IC_CLASS(VortexInitialCondition, "vortex")
    VortexInitialCondition(float xc, float yc, float radius, float strength, float h_mean) {
        params_ = {xc, yc, radius, strength, h_mean};
    }
};
WSB_IC_CLASS(JetStreamInitialCondition, "jet_stream")
    JetStreamInitialCondition(float yc, float width, float strength, float h_mean) {
        params_ = {yc, width, strength, h_mean};
    }
};
WSB_IC_CLASS(BreakingWaveInitialCondition, "breaking_wave")
    BreakingWaveInitialCondition(float amplitude, float wavelength, float h_mean) {
        params_ = {amplitude, wavelength, h_mean};
    }
};
WSB_IC_CLASS(FrontInitialCondition, "front")
    FrontInitialCondition(float y_position, float width, float temp_difference, float wind_shear) {
        params_ = {y_position, width, temp_difference, wind_shear};
    }
};
WSB_IC_CLASS(MountainInitialCondition, "mountain")
    MountainInitialCondition(float xc, float yc, float radius, float height, float u_base) {
        params_ = {xc, yc, radius, height, u_base};
    }
};
WSB_IC_CLASS(AtmosphericProfileInitialCondition, "atmospheric_profile")
    explicit AtmosphericProfileInitialCondition(const std::string &profile_name) { profile_ = profile_name; }
};
#undef WSB_IC_CLASS

class InitialConditionFactory {  // initial_conditions.cpp:15-46, 611-666
public:
    static InitialConditionFactory &instance() {
        static InitialConditionFactory f;
        return f;
    }
    void add(const std::string &name, std::function<std::shared_ptr<InitialCondition>()> creator) {
        creators_[name] = std::move(creator);
    }
    std::shared_ptr<InitialCondition> create(const std::string &name) {
        auto it = creators_.find(name);
        return it == creators_.end() ? nullptr : it->second();
    }
    std::vector<std::string> available() const {
        std::vector<std::string> names;
        for (const auto &e : creators_) names.push_back(e.first);
        return names;
    }

private:
    std::map<std::string, std::function<std::shared_ptr<InitialCondition>()>> creators_;
};

void register_all_initial_conditions() {
    auto &f = InitialConditionFactory::instance();
    f.add("uniform", [] { return std::make_shared<UniformInitialCondition>(0.0f, 0.0f, 10.0f, 1000.0f, 300.0f, 0.0f); });
    f.add("random", [] { return std::make_shared<RandomInitialCondition>(0u, 1.0f); });
    f.add("zonal_flow", [] { return std::make_shared<ZonalFlowInitialCondition>(10.0f, 10.0f, 0.1f); });
    f.add("vortex", [] { return std::make_shared<VortexInitialCondition>(0.5f, 0.5f, 0.1f, 10.0f, 10.0f); });
    f.add("jet_stream", [] { return std::make_shared<JetStreamInitialCondition>(0.5f, 0.1f, 10.0f, 10.0f); });
    f.add("breaking_wave", [] { return std::make_shared<BreakingWaveInitialCondition>(1.0f, 0.2f, 10.0f); });
    f.add("front", [] { return std::make_shared<FrontInitialCondition>(0.5f, 0.05f, 10.0f, 5.0f); });
    f.add("mountain", [] { return std::make_shared<MountainInitialCondition>(0.3f, 0.5f, 0.1f, 1.0f, 5.0f); });
    f.add("standard_atmosphere", [] { return std::make_shared<AtmosphericProfileInitialCondition>("standard"); });
    f.add("tropical_atmosphere", [] { return std::make_shared<AtmosphericProfileInitialCondition>("tropical"); });
    f.add("polar_atmosphere", [] { return std::make_shared<AtmosphericProfileInitialCondition>("polar"); });
}

// ---- simulation ----------------------------------------------------------------------------------------
class WeatherSimulation;

class OutputManager {  // weather_sim.hpp:549-570
public:
    virtual ~OutputManager() = default;
    virtual void initialize(const WeatherSimulation &sim) = 0;
    virtual void write_output(const WeatherSimulation &sim) = 0;
    virtual void finalize(const WeatherSimulation &sim) = 0;
};

class PyOutputManager : public OutputManager {
public:
    using OutputManager::OutputManager;
    void initialize(const WeatherSimulation &sim) override { PYBIND11_OVERRIDE_PURE(void, OutputManager, initialize, sim); }
    void write_output(const WeatherSimulation &sim) override {
        PYBIND11_OVERRIDE_PURE(void, OutputManager, write_output, sim);
    }
    void finalize(const WeatherSimulation &sim) override { PYBIND11_OVERRIDE_PURE(void, OutputManager, finalize, sim); }
};

bool quiet() {
    static const bool q = std::getenv("WSB_QUIET") != nullptr;
    return q;
}

class WeatherSimulation {
public:
    explicit WeatherSimulation(const SimulationConfig &config) : config_(config) {
        wsb_config c;
        std::memset(&c, 0, sizeof(c));
        c.struct_size = sizeof(c);
        c.model = static_cast<int32_t>(config.model);
        c.integration_method = static_cast<int32_t>(config.integration_method);
        c.grid_width = config.grid_width;
        c.grid_height = config.grid_height;
        c.num_levels = config.num_levels;
        c.dx = config.dx;
        c.dy = config.dy;
        c.dt = config.dt;
        c.gravity = config.gravity;
        c.coriolis_f = config.coriolis_f;
        c.max_time = config.max_time;
        c.dtype = config.double_precision ? WSB_F64 : WSB_F32;
        c.device_id = config.device_id;
        c.rk4_mode = config.rk4_classical ? WSB_RK4_CLASSICAL : WSB_RK4_REFERENCE;
        c.kernel_variant = config.kernel_variant;
        c.arith_mode = config.folded_arithmetic ? WSB_ARITH_FOLDED : WSB_ARITH_STRICT;
        c.physics_mode = config.extended_physics ? WSB_PHYSICS_EXTENDED : WSB_PHYSICS_REFERENCE;
        c.beta = config.beta;
        c.viscosity = config.viscosity;
        c.diffusivity = config.diffusivity;
        c.rank = config.rank;
        c.nranks = config.nranks;
        if (config.nranks > 1) {
            if (config.nccl_unique_id.size() != WSB_NCCL_UNIQUE_ID_BYTES)
                throw std::invalid_argument("SimulationConfig.nccl_unique_id must hold the 128 bytes of "
                                            "pyweather_sim.nccl_unique_id() when nranks > 1");
            c.nccl_unique_id = config.nccl_unique_id.data();
        }
        check(wsb_sim_create(&c, &sim_));
        grid_.reset(new WeatherGrid(wsb_sim_current_grid(sim_)));
        // every backend of the reference maps onto the one B200 path (weather_simulation.cpp:562-591)
        config_.compute_backend = ComputeBackend::CUDA;
        if (!quiet()) std::cout << "Using compute backend: CUDA GPU (" << wsb_sim_kernel_name(sim_) << ", sm_100a)" << std::endl;
    }
    WeatherSimulation(const WeatherSimulation &) = delete;
    WeatherSimulation &operator=(const WeatherSimulation &) = delete;
    ~WeatherSimulation() {
        grid_.reset();
        if (sim_) wsb_sim_destroy(sim_);
    }

    void set_initial_condition(std::shared_ptr<InitialCondition> ic) { ic_ = std::move(ic); }
    void set_output_manager(std::shared_ptr<OutputManager> om) { om_ = std::move(om); }

    void initialize() {  // weather_simulation.cpp:46-66
        check(wsb_sim_initialize(sim_));
        if (ic_) ic_->initialize(*grid_);
        if (om_) om_->initialize(*this);
    }

    void run(int num_steps) {  // weather_simulation.cpp:68-103
        if (num_steps <= 0) return;
        wsb_metrics before;
        check(wsb_sim_get_metrics(sim_, &before));
        int remaining = num_steps;
        const int interval = config_.output_interval;
        bool stopped = false;
        while (remaining > 0 && !stopped) {
            int chunk = remaining;
            if (om_ && interval > 0) chunk = std::min(remaining, interval - (wsb_sim_get_step(sim_) % interval));
            int32_t done = 0;
            {
                py::gil_scoped_release nogil;
                check(wsb_sim_run(sim_, chunk, &done));
            }
            remaining -= done;
            // max_time reached (:87-89). `done < chunk` misses a stop on the LAST step of a chunk (e.g. the defaults:
            // dt=0.01, max_time=10, interval=10 reach t >= 10 at step 1000, a chunk boundary), so the library's own
            // clock is tested too -- against max_time as the library rounded it for the simulation's dtype
            wsb_config lc;
            check(wsb_sim_get_config(sim_, &lc));
            if (done < chunk || wsb_sim_get_time(sim_) >= lc.max_time) stopped = true;
            if (om_ && interval > 0 && wsb_sim_get_step(sim_) % interval == 0) om_->write_output(*this);
            if (done == 0) break;
        }
        wsb_metrics after;
        check(wsb_sim_get_metrics(sim_, &after));
        const double ms = after.total_time_ms - before.total_time_ms;
        if (!quiet())
            std::cout << "Completed " << num_steps << " steps in " << ms << " ms (" << (ms / num_steps) << " ms/step)"
                      << std::endl;
    }

    void run_until(float max_time) {  // weather_simulation.cpp:105-115, float arithmetic
        const float t = static_cast<float>(wsb_sim_get_time(sim_)), dt = static_cast<float>(wsb_sim_get_dt(sim_));
        if (max_time <= t) return;
        run(static_cast<int>((max_time - t) / dt) + 1);
    }

    void step() {
        py::gil_scoped_release nogil;
        check(wsb_sim_step(sim_));
    }

    float current_time() const { return static_cast<float>(wsb_sim_get_time(sim_)); }
    double current_time_f64() const { return wsb_sim_get_time(sim_); }
    int current_step() const { return wsb_sim_get_step(sim_); }
    float dt() const { return static_cast<float>(wsb_sim_get_dt(sim_)); }
    void set_dt(double dt) { check(wsb_sim_set_dt(sim_, dt)); }
    const SimulationConfig &config() const { return config_; }
    WeatherGrid &current_grid() { return *grid_; }
    const PerformanceMetrics &metrics() {
        wsb_metrics m;
        check(wsb_sim_get_metrics(sim_, &m));
        metrics_.total_time_ms = m.total_time_ms;
        metrics_.compute_time_ms = m.compute_time_ms;
        metrics_.memory_transfer_time_ms = m.memory_transfer_time_ms;
        metrics_.io_time_ms = m.io_time_ms;
        metrics_.num_steps = m.num_steps;
        metrics_.halo_time_ms = m.halo_time_ms;
        metrics_.kernel_launches = m.kernel_launches;
        return metrics_;
    }
    void reset_metrics() {
        check(wsb_sim_reset_metrics(sim_));
        metrics_.reset();
    }
    std::string kernel_name() const { return wsb_sim_kernel_name(sim_); }
    std::pair<int, int> local_rows() const {  // (row0, nrows) of the global grid this rank owns
        int32_t r0 = 0, n = 0;
        check(wsb_sim_local_rows(sim_, &r0, &n));
        return {r0, n};
    }
    std::pair<double, double> mass_energy() {
        double m = 0, e = 0;
        check(wsb_sim_mass_energy(sim_, &m, &e));
        return {m, e};
    }

private:
    SimulationConfig config_;
    wsb_sim *sim_ = nullptr;
    std::unique_ptr<WeatherGrid> grid_;
    std::shared_ptr<InitialCondition> ic_;
    std::shared_ptr<OutputManager> om_;
    PerformanceMetrics metrics_;
};

// ---- AdaptiveKernelManager (python_bindings.cpp:365-371) -------------------------------------------------
// The reference's kernel-selection layer is replaced by fixed hand-written sm_100a kernels; what stays is the
// Python-visible device query.
class AdaptiveKernelManager {
public:
    static AdaptiveKernelManager &instance() {
        static AdaptiveKernelManager m;
        return m;
    }
    bool initialize(int device_id) {
        int n = 0;
        wsb_device_count(&n);
        available_ = n > 0 && device_id >= 0 && device_id < n;
        caps_ = DeviceCapabilities();
        if (!available_) {
            caps_.device_type = DeviceType::CPU;
            caps_.device_name = "CPU";
            return false;
        }
        wsb_device_caps c;
        if (wsb_device_capabilities(device_id, &c) != WSB_OK) {
            available_ = false;
            return false;
        }
        caps_.device_type = static_cast<DeviceType>(c.device_type);
        caps_.compute_capability_major = c.compute_capability_major;
        caps_.compute_capability_minor = c.compute_capability_minor;
        caps_.cuda_cores = c.cuda_cores;
        caps_.multiprocessors = c.multiprocessors;
        caps_.global_memory = c.global_memory;
        caps_.shared_memory_per_block = c.shared_memory_per_block;
        caps_.max_threads_per_block = c.max_threads_per_block;
        caps_.max_threads_per_multiprocessor = c.max_threads_per_multiprocessor;
        caps_.clock_rate_khz = c.clock_rate_khz;
        caps_.memory_clock_rate_khz = c.memory_clock_rate_khz;
        caps_.memory_bus_width = c.memory_bus_width;
        caps_.compute_power_ratio = c.compute_power_ratio;
        caps_.device_name = c.device_name;
        initialized_ = true;
        return true;
    }
    bool is_cuda_available() {
        if (!initialized_) initialize(0);
        return available_;
    }
    const DeviceCapabilities &capabilities() {
        if (!initialized_) initialize(0);
        return caps_;
    }
    float gpu_workload_ratio(const std::string &) { return is_cuda_available() ? 1.0f : 0.0f; }
    ComputeBackend optimal_backend(int, int, const std::string &) {
        return is_cuda_available() ? ComputeBackend::CUDA : ComputeBackend::CPU;
    }

private:
    bool initialized_ = false, available_ = false;
    DeviceCapabilities caps_;
};

}  // namespace

PYBIND11_MODULE(pyweather_sim, m) {
    m.doc() = "B200-native drop-in for the Weather Simulation workload's pyweather_sim module";
    m.attr("__backend__") = "libweather_b200 (sm_100a)";
    m.def("library_version", [] { return std::string(wsb_version()); });
    m.def("nccl_unique_id", [] {
        char id[WSB_NCCL_UNIQUE_ID_BYTES];
        check(wsb_nccl_get_unique_id(id));
        return py::bytes(id, sizeof(id));
    }, "128-byte NCCL id for SimulationConfig.nccl_unique_id: make it on rank 0, hand the same bytes to every rank");

    py::enum_<SimulationModel>(m, "SimulationModel")
        .value("ShallowWater", SimulationModel::ShallowWater)
        .value("Barotropic", SimulationModel::Barotropic)
        .value("PrimitiveEquations", SimulationModel::PrimitiveEquations)
        .value("General", SimulationModel::General)
        .export_values();
    py::enum_<IntegrationMethod>(m, "IntegrationMethod")
        .value("ExplicitEuler", IntegrationMethod::ExplicitEuler)
        .value("RungeKutta2", IntegrationMethod::RungeKutta2)
        .value("RungeKutta4", IntegrationMethod::RungeKutta4)
        .value("AdamsBashforth", IntegrationMethod::AdamsBashforth)
        .value("SemiImplicit", IntegrationMethod::SemiImplicit)
        .export_values();
    py::enum_<GridType>(m, "GridType")
        .value("Cartesian", GridType::Cartesian)
        .value("Staggered", GridType::Staggered)
        .value("Icosahedral", GridType::Icosahedral)
        .value("SphericalHarmonic", GridType::SphericalHarmonic)
        .export_values();
    py::enum_<BoundaryCondition>(m, "BoundaryCondition")
        .value("Periodic", BoundaryCondition::Periodic)
        .value("Reflective", BoundaryCondition::Reflective)
        .value("Outflow", BoundaryCondition::Outflow)
        .value("Custom", BoundaryCondition::Custom)
        .export_values();
    py::enum_<ComputeBackend>(m, "ComputeBackend")
        .value("CUDA", ComputeBackend::CUDA)
        .value("CPU", ComputeBackend::CPU)
        .value("Hybrid", ComputeBackend::Hybrid)
        .value("AdaptiveHybrid", ComputeBackend::AdaptiveHybrid)
        .export_values();
    py::enum_<DeviceType>(m, "DeviceType")
        .value("Unknown", DeviceType::Unknown)
        .value("CPU", DeviceType::CPU)
        .value("JetsonOrinNX", DeviceType::JetsonOrinNX)
        .value("T4", DeviceType::T4)
        .value("HighEndGPU", DeviceType::HighEndGPU)
        .value("OtherGPU", DeviceType::OtherGPU)
        .export_values();
    py::enum_<OutputFormat>(m, "OutputFormat")
        .value("CSV", OutputFormat::CSV)
        .value("NetCDF", OutputFormat::NetCDF)
        .value("VTK", OutputFormat::VTK)
        .value("PNG", OutputFormat::PNG)
        .value("Custom", OutputFormat::Custom)
        .export_values();

    py::class_<SimulationConfig>(m, "SimulationConfig")
        .def(py::init<>())
        .def_readwrite("model", &SimulationConfig::model)
        .def_readwrite("grid_type", &SimulationConfig::grid_type)
        .def_readwrite("integration_method", &SimulationConfig::integration_method)
        .def_readwrite("boundary_condition", &SimulationConfig::boundary_condition)
        .def_readwrite("grid_width", &SimulationConfig::grid_width)
        .def_readwrite("grid_height", &SimulationConfig::grid_height)
        .def_readwrite("num_levels", &SimulationConfig::num_levels)
        .def_readwrite("dx", &SimulationConfig::dx)
        .def_readwrite("dy", &SimulationConfig::dy)
        .def_readwrite("dt", &SimulationConfig::dt)
        .def_readwrite("gravity", &SimulationConfig::gravity)
        .def_readwrite("coriolis_f", &SimulationConfig::coriolis_f)
        .def_readwrite("beta", &SimulationConfig::beta)
        .def_readwrite("viscosity", &SimulationConfig::viscosity)
        .def_readwrite("diffusivity", &SimulationConfig::diffusivity)
        .def_readwrite("compute_backend", &SimulationConfig::compute_backend)
        .def_readwrite("double_precision", &SimulationConfig::double_precision)
        .def_readwrite("device_id", &SimulationConfig::device_id)
        .def_readwrite("num_threads", &SimulationConfig::num_threads)
        .def_readwrite("max_time", &SimulationConfig::max_time)
        .def_readwrite("max_steps", &SimulationConfig::max_steps)
        .def_readwrite("output_interval", &SimulationConfig::output_interval)
        .def_readwrite("output_path", &SimulationConfig::output_path)
        .def_readwrite("random_seed", &SimulationConfig::random_seed)
        .def_readwrite("rk4_classical", &SimulationConfig::rk4_classical)
        .def_readwrite("kernel_variant", &SimulationConfig::kernel_variant)
        .def_readwrite("folded_arithmetic", &SimulationConfig::folded_arithmetic)
        .def_readwrite("extended_physics", &SimulationConfig::extended_physics)
        .def_readwrite("rank", &SimulationConfig::rank)
        .def_readwrite("nranks", &SimulationConfig::nranks)
        .def_property(
            "nccl_unique_id", [](const SimulationConfig &c) { return py::bytes(c.nccl_unique_id); },
            [](SimulationConfig &c, const py::bytes &b) { c.nccl_unique_id = static_cast<std::string>(b); });

    py::class_<PerformanceMetrics>(m, "PerformanceMetrics")
        .def(py::init<>())
        .def_readwrite("total_time_ms", &PerformanceMetrics::total_time_ms)
        .def_readwrite("compute_time_ms", &PerformanceMetrics::compute_time_ms)
        .def_readwrite("memory_transfer_time_ms", &PerformanceMetrics::memory_transfer_time_ms)
        .def_readwrite("io_time_ms", &PerformanceMetrics::io_time_ms)
        .def_readwrite("num_steps", &PerformanceMetrics::num_steps)
        .def_readwrite("halo_time_ms", &PerformanceMetrics::halo_time_ms)
        .def_readwrite("kernel_launches", &PerformanceMetrics::kernel_launches)
        .def("reset", &PerformanceMetrics::reset)
        .def("print", &PerformanceMetrics::print);

    py::class_<OutputConfig>(m, "OutputConfig")
        .def(py::init<>())
        .def_readwrite("output_dir", &OutputConfig::output_dir)
        .def_readwrite("prefix", &OutputConfig::prefix)
        .def_readwrite("format", &OutputConfig::format)
        .def_readwrite("output_interval", &OutputConfig::output_interval)
        .def_readwrite("compress", &OutputConfig::compress)
        .def_readwrite("include_diagnostics", &OutputConfig::include_diagnostics)
        .def_readwrite("fields", &OutputConfig::fields);

    py::class_<DeviceCapabilities>(m, "DeviceCapabilities")
        .def(py::init<>())
        .def_readonly("device_type", &DeviceCapabilities::device_type)
        .def_readonly("compute_capability_major", &DeviceCapabilities::compute_capability_major)
        .def_readonly("compute_capability_minor", &DeviceCapabilities::compute_capability_minor)
        .def_readonly("cuda_cores", &DeviceCapabilities::cuda_cores)
        .def_readonly("multiprocessors", &DeviceCapabilities::multiprocessors)
        .def_readonly("global_memory", &DeviceCapabilities::global_memory)
        .def_readonly("shared_memory_per_block", &DeviceCapabilities::shared_memory_per_block)
        .def_readonly("max_threads_per_block", &DeviceCapabilities::max_threads_per_block)
        .def_readonly("max_threads_per_multiprocessor", &DeviceCapabilities::max_threads_per_multiprocessor)
        .def_readonly("clock_rate_khz", &DeviceCapabilities::clock_rate_khz)
        .def_readonly("memory_clock_rate_khz", &DeviceCapabilities::memory_clock_rate_khz)
        .def_readonly("memory_bus_width", &DeviceCapabilities::memory_bus_width)
        .def_readonly("compute_power_ratio", &DeviceCapabilities::compute_power_ratio)
        .def_readonly("device_name", &DeviceCapabilities::device_name)
        .def("get_summary", &DeviceCapabilities::summary);

    py::class_<WeatherGrid>(m, "WeatherGrid")
        .def(py::init<int32_t, int32_t, int32_t>(), py::arg("width"), py::arg("height"), py::arg("num_levels") = 1)
        .def(py::init<const SimulationConfig &>())
        .def("reset", &WeatherGrid::reset)
        .def("get_width", &WeatherGrid::width)
        .def("get_height", &WeatherGrid::height)
        .def("get_num_levels", &WeatherGrid::levels)
        .def("get_dx", &WeatherGrid::dx)
        .def("get_dy", &WeatherGrid::dy)
        .def("set_spacing", &WeatherGrid::set_spacing)
        .def("calculate_diagnostics", &WeatherGrid::calculate_diagnostics)
        .def("get_velocity_field",
             [](WeatherGrid &g) { return py::make_tuple(g.get(WSB_FIELD_U), g.get(WSB_FIELD_V)); })
        .def("get_height_field", [](WeatherGrid &g) { return g.get(WSB_FIELD_HEIGHT); })
        .def("get_pressure_field", [](WeatherGrid &g) { return g.get(WSB_FIELD_PRESSURE); })
        .def("get_temperature_field", [](WeatherGrid &g) { return g.get(WSB_FIELD_TEMPERATURE); })
        .def("get_humidity_field", [](WeatherGrid &g) { return g.get(WSB_FIELD_HUMIDITY); })
        .def("get_vorticity_field", [](WeatherGrid &g) { return g.get(WSB_FIELD_VORTICITY); })
        .def("get_divergence_field", [](WeatherGrid &g) { return g.get(WSB_FIELD_DIVERGENCE); })
        .def("set_velocity_field", &WeatherGrid::set_velocity)
        .def("set_height_field", [](WeatherGrid &g, const py::array &a) { g.set(WSB_FIELD_HEIGHT, a); })
        .def("set_pressure_field", [](WeatherGrid &g, const py::array &a) { g.set(WSB_FIELD_PRESSURE, a); })
        .def("set_temperature_field", [](WeatherGrid &g, const py::array &a) { g.set(WSB_FIELD_TEMPERATURE, a); })
        .def("set_humidity_field", [](WeatherGrid &g, const py::array &a) { g.set(WSB_FIELD_HUMIDITY, a); });

    py::class_<InitialCondition, std::shared_ptr<InitialCondition>>(m, "InitialCondition")
        .def("initialize", &InitialCondition::initialize)
        .def("get_name", &InitialCondition::name);
    py::class_<UniformInitialCondition, InitialCondition, std::shared_ptr<UniformInitialCondition>>(
        m, "UniformInitialCondition")
        .def(py::init<float, float, float, float, float, float>(), py::arg("u") = 0.0f, py::arg("v") = 0.0f,
             py::arg("h") = 10.0f, py::arg("p") = 1000.0f, py::arg("t") = 300.0f, py::arg("q") = 0.0f);
    py::class_<RandomInitialCondition, InitialCondition, std::shared_ptr<RandomInitialCondition>>(
        m, "RandomInitialCondition")
        .def(py::init<unsigned int, float>(), py::arg("seed") = 0, py::arg("amplitude") = 1.0f);
    py::class_<ZonalFlowInitialCondition, InitialCondition, std::shared_ptr<ZonalFlowInitialCondition>>(
        m, "ZonalFlowInitialCondition")
        .def(py::init<float, float, float>(), py::arg("u_max") = 10.0f, py::arg("h_mean") = 10.0f,
             py::arg("beta") = 0.1f);
    py::class_<VortexInitialCondition, InitialCondition, std::shared_ptr<VortexInitialCondition>>(
        m, "VortexInitialCondition")
        .def(py::init<float, float, float, float, float>(), py::arg("x_center") = 0.5f, py::arg("y_center") = 0.5f,
             py::arg("radius") = 0.1f, py::arg("strength") = 10.0f, py::arg("h_mean") = 10.0f);
    py::class_<JetStreamInitialCondition, InitialCondition, std::shared_ptr<JetStreamInitialCondition>>(
        m, "JetStreamInitialCondition")
        .def(py::init<float, float, float, float>(), py::arg("y_center") = 0.5f, py::arg("width") = 0.1f,
             py::arg("strength") = 10.0f, py::arg("h_mean") = 10.0f);
    py::class_<BreakingWaveInitialCondition, InitialCondition, std::shared_ptr<BreakingWaveInitialCondition>>(
        m, "BreakingWaveInitialCondition")
        .def(py::init<float, float, float>(), py::arg("amplitude") = 1.0f, py::arg("wavelength") = 0.2f,
             py::arg("h_mean") = 10.0f);
    py::class_<FrontInitialCondition, InitialCondition, std::shared_ptr<FrontInitialCondition>>(
        m, "FrontInitialCondition")
        .def(py::init<float, float, float, float>(), py::arg("y_position") = 0.5f, py::arg("width") = 0.05f,
             py::arg("temp_difference") = 10.0f, py::arg("wind_shear") = 5.0f);
    py::class_<MountainInitialCondition, InitialCondition, std::shared_ptr<MountainInitialCondition>>(
        m, "MountainInitialCondition")
        .def(py::init<float, float, float, float, float>(), py::arg("x_center") = 0.3f, py::arg("y_center") = 0.5f,
             py::arg("radius") = 0.1f, py::arg("height") = 1.0f, py::arg("u_base") = 5.0f);
    py::class_<AtmosphericProfileInitialCondition, InitialCondition,
               std::shared_ptr<AtmosphericProfileInitialCondition>>(m, "AtmosphericProfileInitialCondition")
        .def(py::init<const std::string &>(), py::arg("profile_name") = "standard");

    py::class_<OutputManager, PyOutputManager, std::shared_ptr<OutputManager>>(m, "OutputManager")
        .def(py::init<>())
        .def("initialize", &OutputManager::initialize)
        .def("write_output", &OutputManager::write_output)
        .def("finalize", &OutputManager::finalize);

    py::class_<WeatherSimulation>(m, "WeatherSimulation")
        .def(py::init<const SimulationConfig &>())
        .def("set_initial_condition", &WeatherSimulation::set_initial_condition, py::keep_alive<1, 2>())
        .def("set_output_manager", &WeatherSimulation::set_output_manager, py::keep_alive<1, 2>())
        .def("initialize", &WeatherSimulation::initialize)
        .def("run", &WeatherSimulation::run)
        .def("run_until", &WeatherSimulation::run_until)
        .def("step", &WeatherSimulation::step)
        .def("get_current_time", &WeatherSimulation::current_time)
        .def("get_current_step", &WeatherSimulation::current_step)
        .def("get_dt", &WeatherSimulation::dt)
        .def("set_dt", &WeatherSimulation::set_dt)
        .def("get_config", &WeatherSimulation::config, py::return_value_policy::reference_internal)
        .def("get_current_grid", &WeatherSimulation::current_grid, py::return_value_policy::reference_internal)
        .def("get_performance_metrics", &WeatherSimulation::metrics, py::return_value_policy::reference_internal)
        .def("reset_performance_metrics", &WeatherSimulation::reset_metrics)
        .def("get_kernel_name", &WeatherSimulation::kernel_name)
        .def("get_local_rows", &WeatherSimulation::local_rows)
        .def("get_current_time_f64", &WeatherSimulation::current_time_f64)
        .def("mass_energy", &WeatherSimulation::mass_energy);

    m.def("register_all_initial_conditions", &register_all_initial_conditions);
    py::class_<InitialConditionFactory>(m, "InitialConditionFactory")
        .def_static("get_instance", &InitialConditionFactory::instance, py::return_value_policy::reference)
        .def("create_initial_condition", &InitialConditionFactory::create)
        .def("get_available_initial_conditions", &InitialConditionFactory::available);

    py::class_<AdaptiveKernelManager>(m, "AdaptiveKernelManager")
        .def_static("get_instance", &AdaptiveKernelManager::instance, py::return_value_policy::reference)
        .def("initialize", &AdaptiveKernelManager::initialize, py::arg("device_id") = 0)
        .def("is_cuda_available", &AdaptiveKernelManager::is_cuda_available)
        .def("get_device_capabilities", &AdaptiveKernelManager::capabilities,
             py::return_value_policy::reference_internal)
        .def("get_gpu_workload_ratio", &AdaptiveKernelManager::gpu_workload_ratio)
        .def("determine_optimal_backend", &AdaptiveKernelManager::optimal_backend);
}
