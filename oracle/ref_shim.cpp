// TEST INFRASTRUCTURE ONLY (oracle build). A ctypes-friendly C view of the *reference's own*
// WeatherSimulation (weather_sim.hpp:417-544), compiled against the patched reference sources by
// oracle/build_ref.sh into oracle/_ref/libws_ref.so. No arithmetic lives here: every number comes
// out of the reference's weather_simulation.cpp / weather_grid.cpp.
#include "weather_sim/weather_sim.hpp"
#include "weather_sim/initial_conditions.hpp"
#include <cstring>
#include <cstdio>
#include <unistd.h>
#include <fcntl.h>
#include <omp.h>

using namespace weather_sim;

namespace {
// The reference prints "Using compute backend: ..." from its constructor
// (weather_simulation.cpp:576-590); keep test logs quiet.
struct QuietStdout {
    int saved = -1;
    QuietStdout() {
        fflush(stdout);
        std::cout.flush();
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) { dup2(nul, 1); close(nul); }
    }
    ~QuietStdout() {
        fflush(stdout);
        std::cout.flush();
        if (saved >= 0) { dup2(saved, 1); close(saved); }
    }
};

ScalarField2D* scalar_field(WeatherGrid& g, int field) {
    switch (field) {
        case 2: return &g.getHeightField();
        case 3: return &g.getPressureField();
        case 4: return &g.getTemperatureField();
        case 5: return &g.getHumidityField();
        case 6: return &g.getVorticityField();
        default: return nullptr;
    }
}
}  // namespace

extern "C" {

void* wsref_create(int model, int integrator, int width, int height, float dx, float dy, float dt,
                   float gravity, float coriolis_f) {
    SimulationConfig cfg;
    cfg.model = static_cast<SimulationModel>(model);
    cfg.integration_method = static_cast<IntegrationMethod>(integrator);
    cfg.grid_width = width;
    cfg.grid_height = height;
    cfg.dx = dx;
    cfg.dy = dy;
    cfg.dt = dt;
    cfg.gravity = gravity;
    cfg.coriolis_f = coriolis_f;
    cfg.compute_backend = ComputeBackend::CPU;
    cfg.max_time = 1.0e30f;  // run() breaks at current_time_ >= max_time (weather_simulation.cpp:87-89)
    cfg.output_interval = 0;
    try {
        QuietStdout q;
        return new WeatherSimulation(cfg);
    } catch (...) {
        return nullptr;
    }
}

void wsref_destroy(void* s) { delete static_cast<WeatherSimulation*>(s); }

// field ids: 0 u, 1 v, 2 h, 3 p, 4 T, 5 q, 6 vorticity
int wsref_set_field(void* s, int field, const float* src) {
    WeatherGrid& g = static_cast<WeatherSimulation*>(s)->getCurrentGrid();
    size_t n = static_cast<size_t>(g.getWidth()) * static_cast<size_t>(g.getVelocityField().height);
    if (field == 0) { std::memcpy(g.getVelocityField().u.data(), src, n * sizeof(float)); return 0; }
    if (field == 1) { std::memcpy(g.getVelocityField().v.data(), src, n * sizeof(float)); return 0; }
    ScalarField2D* f = scalar_field(g, field);
    if (!f) return -1;
    std::memcpy(f->data.data(), src, n * sizeof(float));
    return 0;
}

int wsref_get_field(void* s, int field, float* dst) {
    WeatherGrid& g = static_cast<WeatherSimulation*>(s)->getCurrentGrid();
    size_t n = static_cast<size_t>(g.getWidth()) * static_cast<size_t>(g.getVelocityField().height);
    if (field == 0) { std::memcpy(dst, g.getVelocityField().u.data(), n * sizeof(float)); return 0; }
    if (field == 1) { std::memcpy(dst, g.getVelocityField().v.data(), n * sizeof(float)); return 0; }
    ScalarField2D* f = scalar_field(g, field);
    if (!f) return -1;
    std::memcpy(dst, f->data.data(), n * sizeof(float));
    return 0;
}

// n calls of the reference's step() (weather_simulation.cpp:117-158)
void wsref_step(void* s, int n) {
    WeatherSimulation* sim = static_cast<WeatherSimulation*>(s);
    for (int i = 0; i < n; ++i) sim->step();
}

// the reference's run() (weather_simulation.cpp:68-103), stdout silenced
void wsref_run(void* s, int n) {
    QuietStdout q;
    static_cast<WeatherSimulation*>(s)->run(n);
}

float wsref_time(void* s) { return static_cast<WeatherSimulation*>(s)->getCurrentTime(); }
int wsref_steps(void* s) { return static_cast<WeatherSimulation*>(s)->getCurrentStep(); }
void wsref_set_dt(void* s, float dt) { static_cast<WeatherSimulation*>(s)->setDt(dt); }
void wsref_diagnostics(void* s) { static_cast<WeatherSimulation*>(s)->getCurrentGrid().calculateDiagnostics(); }

// Apply one of the reference's registered initial conditions by factory name
// (initial_conditions.cpp:611-666) through WeatherSimulation::initialize() (:46-66).
int wsref_apply_ic(void* s, const char* name) {
    static bool registered = false;
    if (!registered) { registerAllInitialConditions(); registered = true; }
    auto ic = InitialConditionFactory::getInstance().createInitialCondition(name);
    if (!ic) return -1;
    WeatherSimulation* sim = static_cast<WeatherSimulation*>(s);
    sim->setInitialCondition(ic);
    sim->initialize();
    return 0;
}

// OpenMP team size the reference's `#pragma omp parallel for` (weather_simulation.cpp:503) really gets, and a
// runtime override (launchers such as torchrun export OMP_NUM_THREADS=1 before this library is loaded).
void wsref_set_omp_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int wsref_omp_threads(void) {
    int n = 1;
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
    return n;
}

}  // extern "C"
