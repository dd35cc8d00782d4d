/*
 * weather_b200.h -- C-ABI of libweather_b200.so, the B200-native (sm_100a) implementation of the
 * weather-sim time-stepping hot path of scttfrdmn/nvidia-jetson-workload.
 *
 * This header is the drop-in boundary: plain C, plain pointers and sizes, no C++/torch/pybind types.
 * Every entry point names the reference interface it replaces. Reference paths are relative to
 * /root/reference/src/weather-sim/cpp/ ("ws.cpp" = src/weather_simulation.cpp, "wg.cpp" =
 * src/weather_grid.cpp, "ws.hpp" = include/weather_sim/weather_sim.hpp, "pb.cpp" =
 * src/python_bindings.cpp, "ga.hpp" = include/weather_sim/gpu_adaptability.hpp).
 *
 * Conventions
 *   - every function returns a wsb_status (0 = ok, negative = error) unless documented otherwise;
 *     nothing throws across the ABI. wsb_last_error() gives the message of the calling thread's last
 *     failure. The Python shim maps WSB_ERR_INVALID_ARGUMENT -> ValueError (reference:
 *     std::invalid_argument) and every other error -> RuntimeError (reference: std::runtime_error).
 *   - host buffers are caller-owned, dense row-major (rows, cols) == (H, W) like the reference's numpy
 *     views (pb.cpp:22-114); for num_levels > 1 they are (L, H, W). Device memory is library-owned.
 *   - there is NO CPU fallback: without a usable CUDA device wsb_sim_create / wsb_grid_create fail with
 *     WSB_ERR_CUDA.
 *   - a handle is driven from one host thread at a time (as the reference: no locking, pb.cpp holds the GIL).
 */
#ifndef WEATHER_B200_H
#define WEATHER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSB_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------- enums -- */

typedef enum wsb_status {
    WSB_OK = 0,
    WSB_ERR_INVALID_ARGUMENT = -1, /* std::invalid_argument in the reference (wg.cpp:29-31,74-76,125-127) */
    WSB_ERR_RUNTIME = -2,          /* std::runtime_error in the reference (pb.cpp:63-72,92-104) */
    WSB_ERR_CUDA = -3,             /* CUDA runtime/driver failure, or no device */
    WSB_ERR_NCCL = -4,             /* NCCL failure / libnccl not loadable when nranks > 1 */
    WSB_ERR_OUT_OF_MEMORY = -5
} wsb_status;

/* ws.hpp:30-35 SimulationModel */
typedef enum wsb_model {
    WSB_MODEL_SHALLOW_WATER = 0,
    WSB_MODEL_BAROTROPIC = 1,          /* reference: SWE tendencies (ws.cpp:542-550) */
    WSB_MODEL_PRIMITIVE_EQUATIONS = 2, /* reference: SWE tendencies + constant T/p drift (ws.cpp:552-560,201-214) */
    WSB_MODEL_GENERAL = 3              /* reference: default branch == SWE (ws.cpp:172-175) */
} wsb_model;

/* ws.hpp:50-56 IntegrationMethod */
typedef enum wsb_integrator {
    WSB_INT_EXPLICIT_EULER = 0,
    WSB_INT_RUNGE_KUTTA_2 = 1,
    WSB_INT_RUNGE_KUTTA_4 = 2,   /* non-SWE models fall back to RK2 (ws.cpp:334-338) */
    WSB_INT_ADAMS_BASHFORTH = 3, /* reference: Euler (ws.cpp:457-463) */
    WSB_INT_SEMI_IMPLICIT = 4    /* reference: Euler (ws.cpp:465-471) */
} wsb_integrator;

typedef enum wsb_dtype { WSB_F32 = 0, WSB_F64 = 1 } wsb_dtype;

/* Fields of a WeatherGrid (ws.hpp:403-411). */
typedef enum wsb_field {
    WSB_FIELD_U = 0,
    WSB_FIELD_V = 1,
    WSB_FIELD_HEIGHT = 2,
    WSB_FIELD_PRESSURE = 3,
    WSB_FIELD_TEMPERATURE = 4,
    WSB_FIELD_HUMIDITY = 5,
    WSB_FIELD_VORTICITY = 6,
    WSB_FIELD_DIVERGENCE = 7,
    WSB_NUM_FIELDS = 8
} wsb_field;

/* How the RK4 combine treats k1 (ws.cpp:350-351,437-451). */
typedef enum wsb_rk4_mode {
    WSB_RK4_REFERENCE = 0, /* bit-for-bit the reference: the k1 alias reads k4 at the combine */
    WSB_RK4_CLASSICAL = 1  /* textbook RK4 (non-parity opt-in) */
} wsb_rk4_mode;

/* Floating-point evaluation of the packed fp32 whole-step kernels. */
typedef enum wsb_arith_mode {
    WSB_ARITH_STRICT = 0, /* every operation of the reference, in its order: bit-identical for ALL inputs */
    WSB_ARITH_FOLDED = 1  /* opt-in, fp32, dx == dy with 2dx a power of two: the exact scale 1/(2dx) is folded
                             into the stage coefficients (6 fewer multiplies per cell-stage). Bit-identical
                             to STRICT unless an intermediate of the reference is subnormal or within 2dx of
                             overflow; the sign of a zero tendency may differ. Ignored where it does not apply. */
} wsb_arith_mode;

/* Which tendencies the models evaluate. */
typedef enum wsb_physics_mode {
    WSB_PHYSICS_REFERENCE = 0, /* the reference's: every model is the shallow-water stencil (ws.cpp:473-560) */
    WSB_PHYSICS_EXTENDED = 1   /* NOT in the reference's compute code (non-parity opt-in; the oracle's own
                                  restatement is the specification, oracle/ws_oracle_body.inc): the barotropic model
                                  the reference advertises and stubs (README.md:7-13, ws.cpp:542-550), in
                                  primitive-variable form -- shallow-water tendencies on a beta plane
                                  f(y) = coriolis_f + beta*dy*(y - (H-1)/2), eddy viscosity on u, v and
                                  diffusivity on h (5-point Laplacian, same clamped neighbours). Applies to
                                  every model and integrator; power-of-two spacing on the whole-step kernels.
                                  On WSB_MODEL_PRIMITIVE_EQUATIONS it also replaces the constant T/p drift by
                                  transport of p, T, q with the level's flow (advection + diffusivity; per-stage
                                  kernels; on row slabs the tracers' ghost rows are exchanged every stage). */
} wsb_physics_mode;

/* Which hand-written sm_100a path advances the state. All variants produce bit-identical results. */
typedef enum wsb_kernel_variant {
    WSB_KERNEL_AUTO = 0,           /* best available for the configuration */
    WSB_KERNEL_STAGE_DIRECT = 1,   /* one fused tendency+update pass per RK stage, direct global loads */
    WSB_KERNEL_STEP_FUSED_REG = 2, /* all RK stages of a step in ONE pass: register-resident row sweep, warp
                                      shuffles, y rows prefetched into registers with 64-bit global loads */
    WSB_KERNEL_STEP_FUSED_TMA = 3  /* same sweep; y rows staged by TMA bulk copies (cp.async.bulk + mbarrier) into
                                      a per-warp shared-memory ring, k2/k3 windows in shared memory */
} wsb_kernel_variant;

/* ------------------------------------------------------------- structs -- */

/* Mirrors SimulationConfig (ws.hpp:155-191) for the fields the time-stepping path reads, plus the
 * B200-specific knobs. Zero-initialise, set struct_size = sizeof(wsb_config), then fill. */
typedef struct wsb_config {
    uint32_t struct_size;
    int32_t model;              /* wsb_model */
    int32_t integration_method; /* wsb_integrator */
    int32_t grid_width;         /* cells in x (global) */
    int32_t grid_height;        /* cells in y (GLOBAL height when nranks > 1) */
    int32_t num_levels;         /* independent 2-D levels; the reference stores but ignores it */
    double dx, dy, dt;          /* rounded to float for WSB_F32, exactly as the reference's scalar_t fields */
    double gravity, coriolis_f;
    double max_time;            /* run() stops once time >= max_time (ws.cpp:87-89) */
    int32_t dtype;              /* wsb_dtype: arithmetic AND storage type (reference: always float, F8) */
    int32_t device_id;          /* CUDA device ordinal */
    int32_t rk4_mode;           /* wsb_rk4_mode */
    int32_t kernel_variant;     /* wsb_kernel_variant */
    /* Row-slab domain decomposition. nranks <= 1: single GPU.
     * Each rank owns rows [row0, row0+nrows) of the global grid (wsb_sim_local_rows) and exchanges
     * ghost rows with rank-1 / rank+1 by ncclSend/ncclRecv over NVLink. */
    int32_t rank;
    int32_t nranks;
    const void *nccl_unique_id; /* WSB_NCCL_UNIQUE_ID_BYTES bytes from wsb_nccl_get_unique_id on rank 0 */
    int32_t arith_mode;         /* wsb_arith_mode (0 = strict, the default) */
    int32_t physics_mode;       /* wsb_physics_mode (0 = the reference's tendencies, the default) */
    int32_t reserved[6];
    /* SimulationConfig.beta / viscosity / diffusivity (ws.hpp:172-175): read by WSB_PHYSICS_EXTENDED only -- the
     * reference never reads them, and neither does the default mode. */
    double beta, viscosity, diffusivity;
} wsb_config;

#define WSB_NCCL_UNIQUE_ID_BYTES 128

/* PerformanceMetrics (ws.hpp:196-223). Times are measured with CUDA events on the library's streams. */
typedef struct wsb_metrics {
    double total_time_ms;           /* wall time spent inside wsb_sim_run / run_until */
    double compute_time_ms;         /* device time of the stepping kernels (CUDA events) */
    double memory_transfer_time_ms; /* device time of H2D/D2H field copies */
    double io_time_ms;              /* always 0 (the reference never writes it either) */
    int32_t num_steps;
    int32_t reserved;
    double halo_time_ms;            /* device time of the NCCL ghost-row exchange (comm stream) */
    uint64_t kernel_launches;       /* stepping-kernel launches issued so far */
} wsb_metrics;

/* DeviceCapabilities (ga.hpp:35-88), filled from cudaDeviceProp. */
typedef struct wsb_device_caps {
    int32_t device_type; /* DeviceType (ga.hpp:23-30): 0 Unknown, 1 CPU, 2 JetsonOrinNX, 3 T4, 4 HighEndGPU, 5 OtherGPU */
    int32_t compute_capability_major;
    int32_t compute_capability_minor;
    int32_t cuda_cores;
    int32_t multiprocessors;
    uint64_t global_memory;
    uint64_t shared_memory_per_block;
    int32_t max_threads_per_block;
    int32_t max_threads_per_multiprocessor;
    int32_t clock_rate_khz;
    int32_t memory_clock_rate_khz;
    int32_t memory_bus_width;
    float compute_power_ratio;
    char device_name[256];
} wsb_device_caps;

typedef struct wsb_grid_info {
    int32_t width, height, num_levels; /* local (this rank's) extent */
    int32_t dtype;
    double dx, dy;
    int32_t device_id;
    int32_t reserved;
} wsb_grid_info;

typedef struct wsb_sim wsb_sim;   /* replaces weather_sim::WeatherSimulation (ws.hpp:417-544) */
typedef struct wsb_grid wsb_grid; /* replaces weather_sim::WeatherGrid (ws.hpp:254-412), device resident */

/* ------------------------------------------------------- library-level -- */

WSB_API const char *wsb_version(void);
/* Message of the calling thread's most recent failing call ("" if none). */
WSB_API const char *wsb_last_error(void);
/* AdaptiveKernelManager::isCudaAvailable / getDeviceCapabilities (ga.hpp:128-190; pb.cpp:365-371). */
WSB_API int wsb_device_count(int *count);
WSB_API int wsb_device_capabilities(int device_id, wsb_device_caps *out);
/* Pinned host staging buffers so that field copies run at full PCIe rate (pb.cpp:22-114 copies
 * element-wise through pageable memory). */
WSB_API int wsb_host_alloc(size_t bytes, void **out);
WSB_API int wsb_host_free(void *ptr);
/* Balanced row-slab partition used by wsb_sim_create: the first (H % nranks) ranks own one extra row.
 * Pure host arithmetic (usable without a GPU). */
WSB_API int wsb_partition_rows(int32_t grid_height, int32_t nranks, int32_t rank, int32_t *row0, int32_t *nrows);
/* 128-byte NCCL unique id for wsb_config.nccl_unique_id (call on rank 0, broadcast out of band). */
WSB_API int wsb_nccl_get_unique_id(void *out128);

/* ---------------------------------------------------------------- grid -- */

/* WeatherGrid(width, height, num_levels) (wg.cpp:15-34); dims <= 0 -> WSB_ERR_INVALID_ARGUMENT
 * "Grid dimensions must be positive". Fields start at reset() defaults. */
WSB_API int wsb_grid_create(int32_t width, int32_t height, int32_t num_levels, double dx, double dy,
                            int32_t dtype, int32_t device_id, wsb_grid **out);
WSB_API void wsb_grid_destroy(wsb_grid *grid);
/* WeatherGrid::reset (wg.cpp:57-71): u=v=0, h=10, p=1013.25, T=288.15, q=0, vorticity=divergence=0. */
WSB_API int wsb_grid_reset(wsb_grid *grid);
WSB_API int wsb_grid_get_info(const wsb_grid *grid, wsb_grid_info *out);
/* WeatherGrid::setSpacing (wg.cpp:73-80); <= 0 -> WSB_ERR_INVALID_ARGUMENT "Grid spacing must be positive". */
WSB_API int wsb_grid_set_spacing(wsb_grid *grid, double dx, double dy);
/* numpyTo{Scalar,Vector}Field (pb.cpp:60-114): host -> device. host_dtype may differ from the grid's
 * dtype (converted on the way, like pybind's forcecast). Shape mismatch -> WSB_ERR_RUNTIME
 * "Array dimensions must match field dimensions". Vorticity/divergence cannot be set (no reference setter). */
WSB_API int wsb_grid_set_field(wsb_grid *grid, int32_t field, const void *host, int32_t host_dtype,
                               int64_t levels, int64_t rows, int64_t cols);
/* {scalar,vector}FieldToNumpy (pb.cpp:22-57): device -> host copy. */
WSB_API int wsb_grid_get_field(wsb_grid *grid, int32_t field, void *host, int32_t host_dtype,
                               int64_t levels, int64_t rows, int64_t cols);
/* WeatherGrid::calculateDiagnostics (wg.cpp:82-121): vorticity and divergence of the current u, v. */
WSB_API int wsb_grid_calculate_diagnostics(wsb_grid *grid);
/* WeatherGrid::swap (wg.cpp:123-142); dimension mismatch -> WSB_ERR_INVALID_ARGUMENT. */
WSB_API int wsb_grid_swap(wsb_grid *a, wsb_grid *b);
/* Raw device pointer + row pitch (elements) of a field, for zero-copy consumers. */
WSB_API int wsb_grid_device_pointer(wsb_grid *grid, int32_t field, void **dev_ptr, int64_t *pitch_elems);

/* ---------------------------------------------------------- simulation -- */

/* WeatherSimulation(config) (ws.cpp:17-36). */
WSB_API int wsb_sim_create(const wsb_config *config, wsb_sim **out);
WSB_API void wsb_sim_destroy(wsb_sim *sim);
/* WeatherSimulation::initialize (ws.cpp:46-66) minus the initial condition (applied by the caller through
 * wsb_grid_set_field / wsb_ic_apply on the current grid): time = 0, step = 0, metrics reset, grid reset. */
WSB_API int wsb_sim_initialize(wsb_sim *sim);
/* getCurrentGrid (ws.hpp:493): borrowed handle, always a live view of the CURRENT state (the
 * reference's handle goes stale on odd step counts because of its pointer swap, SURVEY.md hard part 10). */
WSB_API wsb_grid *wsb_sim_current_grid(wsb_sim *sim);
/* WeatherSimulation::step (ws.cpp:117-158): one time step; time += dt (in the sim's dtype); step++. */
WSB_API int wsb_sim_step(wsb_sim *sim);
/* One time step with the state STREAMED through the GPU (host buffers in, host buffers out): equivalent to
 * set_field(u,v,h) + wsb_sim_step + get_field(u,v,h), but pipelined in row slabs on three streams -- slab
 * i+1 uploads while slab i is stepped and slab i-1 downloads -- so a host-resident step costs one PCIe
 * direction, not two. What a user of the reference does per step when the state lives in numpy
 * (pb.cpp:60-114 in, ws.cpp:117-158, pb.cpp:22-57 out). Buffers: (H, W) (or (L, H, W)) of the simulation's
 * dtype, page-locked for full overlap (wsb_host_alloc). Outputs may alias the inputs. */
WSB_API int wsb_sim_step_host(wsb_sim *sim, const void *u, const void *v, const void *h, void *out_u, void *out_v,
                              void *out_h);
/* WeatherSimulation::run (ws.cpp:68-103) incl. the early stop at time >= max_time. steps_done may be NULL.
 * The whole run is enqueued asynchronously and synchronised once at the end. */
WSB_API int wsb_sim_run(wsb_sim *sim, int32_t num_steps, int32_t *steps_done);
/* WeatherSimulation::runUntil (ws.cpp:105-115). */
WSB_API int wsb_sim_run_until(wsb_sim *sim, double max_time, int32_t *steps_done);
/* Enqueue num_steps steps without synchronising and without the max_time check (benchmark loops);
 * pair with wsb_sim_synchronize. */
WSB_API int wsb_sim_advance_async(wsb_sim *sim, int32_t num_steps);
WSB_API int wsb_sim_synchronize(wsb_sim *sim);
/* Device time in ms (CUDA events on the stepping stream) of the most recent run/advance+synchronize. */
WSB_API int wsb_sim_last_run_device_ms(wsb_sim *sim, double *ms);
WSB_API double wsb_sim_get_time(const wsb_sim *sim);  /* getCurrentTime (ws.hpp:463) */
WSB_API int32_t wsb_sim_get_step(const wsb_sim *sim); /* getCurrentStep (ws.hpp:469) */
WSB_API double wsb_sim_get_dt(const wsb_sim *sim);    /* getDt (ws.hpp:475) */
WSB_API int wsb_sim_set_dt(wsb_sim *sim, double dt);  /* setDt (ws.hpp:481) */
WSB_API int wsb_sim_get_config(const wsb_sim *sim, wsb_config *out);    /* getConfig (ws.hpp:487) */
WSB_API int wsb_sim_get_metrics(wsb_sim *sim, wsb_metrics *out);        /* getPerformanceMetrics (ws.hpp:505) */
WSB_API int wsb_sim_reset_metrics(wsb_sim *sim);                        /* resetPerformanceMetrics (ws.hpp:510) */
/* Rows of the global grid this rank owns (row0 = 0, nrows = grid_height when nranks <= 1). */
WSB_API int wsb_sim_local_rows(const wsb_sim *sim, int32_t *row0, int32_t *nrows);
/* Name of the kernel path selected for this configuration (static string). */
WSB_API const char *wsb_sim_kernel_name(const wsb_sim *sim);
/* Sum over this rank's cells, accumulated in double on the device: mass = sum h,
 * energy = sum 0.5*h*(u^2+v^2) + 0.5*g*h^2 (SURVEY.md section 8d conservation checks). */
WSB_API int wsb_sim_mass_energy(wsb_sim *sim, double *mass, double *energy);

/* Measurement aid for the ghost-row phase (no counterpart in the reference, which has no
 * decomposition): `reps` bare exchanges of the current state's ghost rows -- the very ncclSend/ncclRecv group a step
 * issues, without any compute -- timed with CUDA events on the comm stream after a device-side rendezvous of the
 * ranks. bytes_per_neighbour = ghost depth x 3 fields x row pitch. Single-rank simulations report 0 us. */
WSB_API int wsb_sim_time_halo_exchange(wsb_sim *sim, int32_t reps, double *us_per_exchange,
                                       int64_t *bytes_per_neighbour);

/* Introspection (pure host arithmetic, usable without a GPU): the reciprocal the kernels use to divide by a
 * loop-invariant spacing `divisor` = 2dx or 2dy with the exact three-operation sequence q = x*r; e = x - q*d;
 * q' = q + e*r (ws.cpp:521-528 divides six times per cell and stage). *reciprocal is RN(1/divisor) if the sequence
 * is proven to return the correctly rounded quotient for this divisor -- fp32: by exhaustion over all 2^23
 * significands (about 70 ms, cached per divisor) -- and 0 if the kernels fall back to the IEEE division. */
WSB_API int wsb_exact_division_reciprocal(double divisor, int32_t dtype, double *reciprocal);

/* ------------------------------------------------- initial conditions -- */

/* The reference's InitialCondition::initialize(grid) family (initial_conditions.cpp:59-535), evaluated
 * on the host in the reference's own float arithmetic and uploaded. `name` is one of: uniform, random,
 * zonal_flow, vortex, jet_stream, breaking_wave, front, mountain, atmospheric_profile. params/nparams
 * follow the constructor argument order bound at pb.cpp:291-329; seed is used by "random" only;
 * profile is used by "atmospheric_profile" only (may be NULL = "standard"). */
WSB_API int wsb_ic_apply(wsb_grid *grid, const char *name, const double *params, int32_t nparams,
                         uint32_t seed, const char *profile);
/* Same arithmetic into caller-owned host arrays (each rows*cols floats; any pointer may be NULL).
 * Pure host code: usable without a GPU. */
WSB_API int wsb_ic_fill_host(const char *name, const double *params, int32_t nparams, uint32_t seed,
                             const char *profile, int32_t width, int32_t height, double dx, double dy,
                             float *u, float *v, float *h, float *p, float *t, float *q);

#ifdef __cplusplus
}
#endif
#endif /* WEATHER_B200_H */
