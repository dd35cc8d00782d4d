// Host side of libweather_b200.so: device-resident WeatherGrid / WeatherSimulation behind the C-ABI
// of include/weather_b200.h. Mirrors the reference's driver (weather_simulation.cpp:17-158) and grid
// (weather_grid.cpp:15-142), with the arithmetic in hand-written sm_100a kernels.
//
// There is NO CPU compute path in this file: every field lives in HBM and every update is a kernel.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "wsb_internal.h"

namespace wsb {

// ------------------------------------------------------------------------------------ errors --
static thread_local std::string g_last_error;

void set_last_error(const std::string &msg) { g_last_error = msg; }

int fail(int status, const std::string &msg) {
    g_last_error = msg;
    return status;
}

int cuda_fail(cudaError_t err, const char *what, const char *file, int line) {
    g_last_error = std::string("CUDA error: ") + cudaGetErrorString(err) + " in " + what + " (" + file + ":" +
                   std::to_string(line) + ")";
    // leave the sticky error readable for the caller but clear the "last error" slot
    cudaGetLastError();
    return err == cudaErrorMemoryAllocation ? WSB_ERR_OUT_OF_MEMORY : WSB_ERR_CUDA;
}

static inline size_t elem_size(int dtype) { return dtype == WSB_F64 ? sizeof(double) : sizeof(float); }

// rows x width_bytes between a dense host array and pitched device rows; one flat copy when the pitch is dense
static inline cudaError_t copy_rows(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width_bytes,
                                    size_t rows, cudaMemcpyKind kind, cudaStream_t st) {
    if (dpitch == width_bytes && spitch == width_bytes) return cudaMemcpyAsync(dst, src, width_bytes * rows, kind, st);
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, rows, kind, st);
}

// reset() defaults, weather_grid.cpp:57-71. The float literals are the reference's; for WSB_F64 they are
// the same float values widened (what `std::fill(..., 288.15f)` would store in a vector<double>).
static const float kResetValue[WSB_NUM_FIELDS] = {0.0f, 0.0f, 10.0f, 1013.25f, 288.15f, 0.0f, 0.0f, 0.0f};

static bool is_pow2(double x) {
    if (!(x > 0) || !std::isfinite(x)) return false;
    int e;
    return std::frexp(x, &e) == 0.5;
}

// Reciprocal for the three-operation division by a loop-invariant divisor (wsb_arith.cuh, div_by_invariant):
// RN(1/d) if the sequence is PROVEN to return the correctly rounded quotient for this d, else 0 (= use the IEEE
// division). fp32: proven by exhaustion -- inside the window the kernels apply it to, the sequence commutes with
// scaling x by powers of two, so all 2^23 significands of one binade cover every operand (about 30 ms per divisor,
// cached). fp64: by the theorem (Brisebarre, Muller, Raina 2004: correct for every x when r = RN(1/d), except
// possibly for the all-ones significand of d, which is excluded).
static double proven_reciprocal(double d, int dtype) {
    if (std::getenv("WSB_IEEE_DIV")) return 0.0;
    if (!(d >= 9.5367431640625e-07 && d <= 1048576.0)) return 0.0;  // 2^-20 .. 2^20: keeps q and the remainder normal
    if (dtype == WSB_F64) {
        unsigned long long bits;
        std::memcpy(&bits, &d, sizeof(bits));
        if ((bits & 0xFFFFFFFFFFFFFull) == 0xFFFFFFFFFFFFFull) return 0.0;
        return 1.0 / d;
    }
    static std::mutex mu;
    static std::map<float, float> cache;
    const float df = (float)d;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(df);
    if (it != cache.end()) return (double)it->second;
    const volatile float rv = 1.0f / df;  // IEEE division in float: correctly rounded
    const float r = rv;
    bool ok = true;
    for (uint32_t m = 0; m < (1u << 23) && ok; ++m) {
        uint32_t b = 0x3F800000u | m;
        float x;
        std::memcpy(&x, &b, sizeof(x));
        const volatile float q0 = x * r;  // volatile: one rounding each, nothing fused by the host compiler
        const float q = q0;
        const float e = std::fmaf(-q, df, x);
        const float q1 = std::fmaf(e, r, q);
        const volatile float want = x / df;
        ok = q1 == want;
    }
    cache[df] = ok ? r : 0.0f;
    return (double)cache[df];
}

}  // namespace wsb

using namespace wsb;

// -------------------------------------------------------------------------------------- grid --
// One plane set of a field: lazily allocated. While `ptr == nullptr` the field is uniformly `uniform`
// (e.g. p/T/q of a shallow-water run are never touched: they cost no HBM).
struct FieldBuf {
    void *base = nullptr;  // allocation start (includes ghost rows)
    double uniform = 0.0;
};

struct wsb_grid {
    int W = 0, H = 0, L = 1;     // local extent
    int dtype = WSB_F32;
    int device = 0;
    double dx = 1.0, dy = 1.0;   // already rounded to the grid dtype
    double rdx = 0.5, rdy = 0.5; // reciprocals of 2dx, 2dy as the kernels use them (set_spacing_derived)
    bool recip = true;
    int pitch = 0;               // elements
    long long level_stride = 0;  // elements
    int row0 = 0, Hglobal = 0;   // slab position (standalone grid: 0, H)
    FieldBuf f[WSB_NUM_FIELDS];
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    wsb_sim *owner = nullptr;    // non-null for a simulation's current grid
    double transfer_ms = 0.0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    size_t plane_bytes() const { return (size_t)L * (size_t)level_stride * elem_size(dtype); }
    // pointer to (level 0, local row 0, x 0)
    void *origin(int field) const {
        if (!f[field].base) return nullptr;
        return (char *)f[field].base + (size_t)kLeadRows * pitch * elem_size(dtype);
    }
    template <typename T>
    Geometry<T> geom() const {
        Geometry<T> g;
        g.W = W; g.H = H; g.L = L; g.pitch = pitch; g.level_stride = level_stride; g.row0 = row0; g.Hglobal = Hglobal;
        return g;
    }
    template <typename T>
    Physics<T> physics(double gravity, double coriolis) const {
        Physics<T> p;
        const T tdx = (T)dx, tdy = (T)dy;
        p.ddx = T(2.0f) * tdx;  // `2.0f * dx` (weather_simulation.cpp:521): exact in both types
        p.ddy = T(2.0f) * tdy;
        p.recip = recip;
        p.rdx = (T)rdx;
        p.rdy = (T)rdy;
        p.g = (T)gravity;
        p.f = (T)coriolis;
        p.ext = 0;
        p.bdy = p.yc = p.nu = p.kappa = p.idx2 = p.idy2 = T(0);
        return p;
    }
};

// what the kernels need besides dx, dy: whether 2dx and 2dy are powers of two (then (a-b)/(2dx) IS the multiplication
// by the exact reciprocal), else the proven reciprocals of the three-operation division (0 = IEEE division)
static void set_spacing_derived(wsb_grid *g) {
    const double ddx = 2.0 * g->dx, ddy = 2.0 * g->dy;  // exact: dx, dy are already values of the grid dtype
    g->recip = is_pow2(ddx) && is_pow2(ddy);
    if (g->recip) {
        g->rdx = 1.0 / ddx;
        g->rdy = 1.0 / ddy;
    } else {
        g->rdx = proven_reciprocal(ddx, g->dtype);
        g->rdy = proven_reciprocal(ddy, g->dtype);
        if (g->rdx == 0.0 || g->rdy == 0.0) g->rdx = g->rdy = 0.0;
    }
}

static int grid_alloc_plane(wsb_grid *g, void **out) {
    WSB_CUDA(cudaSetDevice(g->device));
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, g->plane_bytes());
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(field plane)", __FILE__, __LINE__);
    *out = p;
    return WSB_OK;
}

static int grid_fill(wsb_grid *g, void *base, double value) {
    void *origin = (char *)base + (size_t)kLeadRows * g->pitch * elem_size(g->dtype);
    if (g->dtype == WSB_F64) WSB_CUDA(launch_fill<double>(g->geom<double>(), (double *)origin, value, true, g->stream));
    else WSB_CUDA(launch_fill<float>(g->geom<float>(), (float *)origin, (float)value, true, g->stream));
    return WSB_OK;
}

// make sure the field has device storage (materialising its uniform value)
static int grid_materialize(wsb_grid *g, int field) {
    if (g->f[field].base) return WSB_OK;
    void *p = nullptr;
    WSB_TRY(grid_alloc_plane(g, &p));
    g->f[field].base = p;
    return grid_fill(g, p, g->f[field].uniform);
}

static void grid_release(wsb_grid *g, int field, double uniform) {
    if (g->f[field].base) {
        cudaFree(g->f[field].base);
        g->f[field].base = nullptr;
    }
    g->f[field].uniform = uniform;
}

static int grid_init(wsb_grid *g, int W, int H, int L, double dx, double dy, int dtype, int device, int row0,
                     int Hglobal, cudaStream_t stream) {
    if (W <= 0 || H <= 0 || L <= 0)
        return fail(WSB_ERR_INVALID_ARGUMENT, "Grid dimensions must be positive");  // weather_grid.cpp:29-31
    if (dtype != WSB_F32 && dtype != WSB_F64) return fail(WSB_ERR_INVALID_ARGUMENT, "unknown dtype");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(WSB_ERR_CUDA, "no CUDA device available: libweather_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(WSB_ERR_INVALID_ARGUMENT, "device_id out of range");
    g->W = W; g->H = H; g->L = L; g->dtype = dtype; g->device = device;
    g->dx = dtype == WSB_F32 ? (double)(float)dx : dx;
    g->dy = dtype == WSB_F32 ? (double)(float)dy : dy;
    set_spacing_derived(g);
    const int align = 128 / (int)elem_size(dtype);  // rows start on 128-byte boundaries
    g->pitch = (W + align - 1) / align * align;
    g->level_stride = (long long)(H + 2 * kLeadRows) * g->pitch;
    g->row0 = row0;
    g->Hglobal = Hglobal;
    WSB_CUDA(cudaSetDevice(device));
    if (stream) {
        g->stream = stream;
    } else {
        WSB_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
        g->owns_stream = true;
    }
    WSB_CUDA(cudaEventCreate(&g->ev0));
    WSB_CUDA(cudaEventCreate(&g->ev1));
    for (int k = 0; k < WSB_NUM_FIELDS; ++k) g->f[k].uniform = (double)kResetValue[k];
    // u, v, h always live on the device; the rest are materialised on first touch
    for (int k = 0; k <= WSB_FIELD_HEIGHT; ++k) WSB_TRY(grid_materialize(g, k));
    return WSB_OK;
}

static void grid_fini(wsb_grid *g) {
    cudaSetDevice(g->device);
    for (int k = 0; k < WSB_NUM_FIELDS; ++k)
        if (g->f[k].base) cudaFree(g->f[k].base);
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    if (g->owns_stream && g->stream) cudaStreamDestroy(g->stream);
}

// --------------------------------------------------------------------------------------- sim --
enum KernelPath { PATH_STAGE_DIRECT = 1, PATH_STEP_REG = 2, PATH_STEP_TMA = 3 };
static inline bool is_step_path(int p) { return p == PATH_STEP_REG || p == PATH_STEP_TMA; }

struct wsb_sim {
    wsb_config cfg{};
    int dtype = WSB_F32;
    double dt = 0.01;         // rounded to dtype
    double time = 0.0;        // accumulated in dtype (float add for WSB_F32, weather_sim.hpp:527)
    int step = 0;
    int nstages = 1;          // 1 Euler-like, 2 RK2, 4 RK4
    int path = PATH_STAGE_DIRECT;
    int row0 = 0, nrows = 0;
    wsb_grid cur;             // the live "current grid" (u,v,h pointers rotate with `next`)
    // the reference swaps whole grids (weather_simulation.cpp:217): p/T/q of the "next" grid become current
    FieldBuf alt[WSB_NUM_FIELDS];
    void *next[3] = {nullptr, nullptr, nullptr};   // y_{n+1} planes
    void *trA[3] = {nullptr, nullptr, nullptr};    // midpoint p, T, q of the extended Primitive model (RK2)
    void *tA[3] = {nullptr, nullptr, nullptr};     // stage scratch (per-stage paths)
    void *tB[3] = {nullptr, nullptr, nullptr};
    void *k1[3] = {nullptr, nullptr, nullptr};
    void *k2[3] = {nullptr, nullptr, nullptr};
    void *k3[3] = {nullptr, nullptr, nullptr};
    bool diag_dirty = false;  // vorticity/divergence need recomputing from the current u, v
    bool halo_valid = false;  // ghost rows of the current state are up to date (nranks > 1)
    static constexpr int kMaxSlabs = 128;  // row slabs of the streamed host step
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_up[kMaxSlabs] = {}, ev_done[kMaxSlabs] = {};
    cudaStream_t stream = nullptr, comm_stream = nullptr, edge_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_edge = nullptr, ev_halo = nullptr, ev_interior = nullptr;
    cudaEvent_t ev_h0 = nullptr, ev_h1 = nullptr, ev_align = nullptr, ev_tracer = nullptr;
    bool run_open = false;    // ev_start recorded, ev_stop pending
    double last_run_ms = 0.0;
    wsb_metrics metrics{};
    HaloComm *comm = nullptr;
    // fused ghost exchange over peer memory (PeerExchange, wsb_internal.h): all ranks or none
    bool peer_ok = false;
    PeerLink peer_up, peer_dn;
    unsigned *xflags = nullptr;     // [0] bumped by the upper neighbour, [1] by the lower one
    unsigned x_seq = 0;             // fused steps so far
    int band_ctas = 0;              // CTAs per band and step (strips x levels): what one step adds to a flag
    void *setA[3] = {}, *setB[3] = {};  // the two plane sets as they were when the neighbours mapped them
    bool ghosts_in_flight = false;  // the ghost rows of the current state are being pushed by the neighbours' last
                                    // step: kernels that wait on the flags may use them, host-ordered consumers not yet
    bool nccl_ghosts_pending = false;  // an NCCL exchange was enqueued that the main stream has not waited for
    double *d_partial = nullptr;
    int npartial = 0;
    // step overlap (StepArgs::ovl_*): one counter per chunk row and level, protocol steps issued so far
    unsigned *ovl_done = nullptr;
    unsigned ovl_seq = 0;
    int rpc = 64;  // rows per chunk of a full-height launch of the TMA whole-step kernel
    unsigned *ovl_err = nullptr;  // mapped host word raised by a CTA whose dependency timed out
    bool ovl_enabled = false;
};

static void *plane_origin(const wsb_sim *s, void *base) {
    return base ? (char *)base + (size_t)kLeadRows * s->cur.pitch * elem_size(s->dtype) : nullptr;
}

static int sim_alloc3(wsb_sim *s, void *dst[3]) {
    for (int k = 0; k < 3; ++k) WSB_TRY(grid_alloc_plane(&s->cur, &dst[k]));
    return WSB_OK;
}

static void sim_free3(void *p[3]) {
    for (int k = 0; k < 3; ++k) {
        if (p[k]) cudaFree(p[k]);
        p[k] = nullptr;
    }
}

template <typename T>
static Planes3<T> planes(const wsb_sim *s, void *const p[3]) {
    Planes3<T> r;
    r.u = (T *)plane_origin(s, p[0]);
    r.v = (T *)plane_origin(s, p[1]);
    r.h = (T *)plane_origin(s, p[2]);
    return r;
}

template <typename T>
static Planes3<T> null_planes() {
    Planes3<T> r;
    r.u = r.v = r.h = nullptr;
    return r;
}

static int effective_stages(const wsb_config &c) {
    switch (c.integration_method) {
        case WSB_INT_RUNGE_KUTTA_2: return 2;
        case WSB_INT_RUNGE_KUTTA_4:
            return c.model == WSB_MODEL_SHALLOW_WATER ? 4 : 2;  // weather_simulation.cpp:334-338
        default: return 1;  // Euler, AdamsBashforth, SemiImplicit (:457-471)
    }
}

// extended physics on the PrimitiveEquations model: p, T, q are transported by the flow (tracer_stage_kernel)
// instead of drifting by the reference's constants
static inline bool prim_ext(const wsb_sim *s) {
    return s->cfg.physics_mode == WSB_PHYSICS_EXTENDED && s->cfg.model == WSB_MODEL_PRIMITIVE_EQUATIONS;
}

// Physics of one tendency evaluation for this simulation: the grid's spacing + the configuration's constants
// (+ the extended-physics constants, each rounded once in T exactly as oracle/ws_oracle_body.inc computes them)
template <typename T>
static Physics<T> sim_physics(const wsb_sim *s) {
    Physics<T> p = s->cur.physics<T>(s->cfg.gravity, s->cfg.coriolis_f);
    p.ext = s->cfg.physics_mode == WSB_PHYSICS_EXTENDED;
    if (p.ext) {
        const T dx = (T)s->cur.dx, dy = (T)s->cur.dy;
        p.idx2 = T(1.0f) / (dx * dx);
        p.idy2 = T(1.0f) / (dy * dy);
        p.bdy = (T)s->cfg.beta * dy;
        p.yc = (T)(s->cfg.grid_height - 1) * T(0.5f);
        p.nu = (T)s->cfg.viscosity;
        p.kappa = (T)s->cfg.diffusivity;
    }
    return p;
}

// ghost-row exchange of 3 planes on the comm stream, ordered after `after` and signalling ev_halo
static int sim_exchange(wsb_sim *s, void *const p[3], int nrows_halo, cudaEvent_t after) {
    if (!s->comm) return WSB_OK;
    WSB_CUDA(cudaStreamWaitEvent(s->comm_stream, after, 0));
    WSB_CUDA(cudaEventRecord(s->ev_h0, s->comm_stream));
    void *origins[3] = {plane_origin(s, p[0]), plane_origin(s, p[1]), plane_origin(s, p[2])};
    WSB_TRY(halo_exchange(s->comm, origins, 3, elem_size(s->dtype), s->cur.pitch, s->cur.H, nrows_halo, s->cur.L,
                          s->cur.level_stride, s->comm_stream));
    WSB_CUDA(cudaEventRecord(s->ev_h1, s->comm_stream));
    WSB_CUDA(cudaEventRecord(s->ev_halo, s->comm_stream));
    s->nccl_ghosts_pending = true;
    return WSB_OK;
}

// Enqueue one fused RK stage. With a decomposition the edge rows go first so that their exchange
// overlaps the interior launch.
template <typename T>
static int enqueue_stage(wsb_sim *s, StageArgs<T> a, void *const out_planes[3], bool exchange_output) {
    const Geometry<T> g = s->cur.geom<T>();
    const Physics<T> ph = sim_physics<T>(s);
    const int H = s->cur.H;
    if (!s->comm) {
        a.y_begin = 0;
        a.y_end = H;
        WSB_CUDA(launch_stage_direct<T>(g, ph, a, s->stream));
        s->metrics.kernel_launches += 1;
        return WSB_OK;
    }
    // input ghosts must have landed
    WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
    const int e = std::min(1, H);
    a.y_begin = 0; a.y_end = e;
    WSB_CUDA(launch_stage_direct<T>(g, ph, a, s->stream));
    s->metrics.kernel_launches += 1;
    if (H > e) {
        a.y_begin = std::max(e, H - 1); a.y_end = H;
        WSB_CUDA(launch_stage_direct<T>(g, ph, a, s->stream));
        s->metrics.kernel_launches += 1;
    }
    WSB_CUDA(cudaEventRecord(s->ev_edge, s->stream));
    if (exchange_output) WSB_TRY(sim_exchange(s, out_planes, 1, s->ev_edge));
    if (H > 2) {
        a.y_begin = 1; a.y_end = H - 1;
        WSB_CUDA(launch_stage_direct<T>(g, ph, a, s->stream));
        s->metrics.kernel_launches += 1;
    }
    return WSB_OK;
}

template <typename T>
static int enqueue_step_stages(wsb_sim *s) {
    const T dt = (T)s->dt;
    const T half_dt = T(0.5f) * dt;  // `0.5f * dt_` (weather_simulation.cpp:249): one rounding
    const T dt6 = dt / T(6.0f);      // `dt_ / 6.0f` (:438)
    void *cur3[3] = {s->cur.f[0].base, s->cur.f[1].base, s->cur.f[2].base};
    const Planes3<const T> Y = planes<const T>(s, cur3);
    StageArgs<T> a{};
    a.Y = Y;
    a.KS = null_planes<T>();
    a.KA = a.KB = a.K1 = null_planes<const T>();
    a.dt6 = dt6;
    a.final_stage = 0;
    // extended Primitive model: one tracer stage (p, T, q) beside every (u, v, h) stage, carried by that stage's
    // input velocity; Primitive runs Euler or RK2 only (RK4 -> RK2, :334-338)
    const int tf[3] = {WSB_FIELD_PRESSURE, WSB_FIELD_TEMPERATURE, WSB_FIELD_HUMIDITY};
    auto tracer_stage = [&](const Planes3<const T> &S, void *const Cin[3], void *const Out[3], T c) -> int {
        TracerArgs<T> t{};
        t.u = S.u; t.v = S.v; t.c = c;
        for (int k = 0; k < 3; ++k) {
            t.C[k] = (const T *)plane_origin(s, Cin[k]);
            t.Y[k] = (const T *)s->cur.origin(tf[k]);
            t.O[k] = (T *)plane_origin(s, Out[k]);
        }
        WSB_CUDA(launch_tracer_stage<T>(s->cur.geom<T>(), sim_physics<T>(s), t, s->stream));
        s->metrics.kernel_launches += 1;
        return WSB_OK;
    };
    // Row slabs: a tracer stage reads one ghost row of its input tracers (the stage's velocity ghosts come with the
    // (u, v, h) exchange of enqueue_stage). Exchanged on the comm stream after everything the main stream has
    // enqueued so far (the planes' producer, and the last reader of the ghost rows being overwritten); the main
    // stream waits for it. No overlap is attempted: this model is an HBM-bound opt-in on the per-stage kernels.
    auto tracer_ghosts = [&](void *const P[3]) -> int {
        if (!s->comm) return WSB_OK;
        WSB_CUDA(cudaEventRecord(s->ev_tracer, s->stream));
        WSB_TRY(sim_exchange(s, P, 1, s->ev_tracer));
        WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
        return WSB_OK;
    };
    void *tr_cur[3] = {nullptr, nullptr, nullptr}, *tr_next[3] = {nullptr, nullptr, nullptr};
    if (prim_ext(s)) {
        for (int k = 0; k < 3; ++k) {
            WSB_TRY(grid_materialize(&s->cur, tf[k]));
            if (!s->alt[tf[k]].base) WSB_TRY(grid_alloc_plane(&s->cur, &s->alt[tf[k]].base));
            if (s->nstages == 2 && !s->trA[k]) WSB_TRY(grid_alloc_plane(&s->cur, &s->trA[k]));
            tr_cur[k] = s->cur.f[tf[k]].base;
            tr_next[k] = s->alt[tf[k]].base;
        }
        WSB_TRY(tracer_ghosts(tr_cur));  // every step: p, T, q may have been set by the user in between
    }
    if (s->nstages == 1) {  // weather_simulation.cpp:160-218
        a.S = Y; a.O = planes<T>(s, s->next); a.c = dt;
        WSB_TRY(enqueue_stage<T>(s, a, s->next, true));
        if (prim_ext(s)) WSB_TRY(tracer_stage(Y, tr_cur, tr_next, dt));
    } else if (s->nstages == 2) {  // :220-323
        a.S = Y; a.O = planes<T>(s, s->tA); a.c = half_dt;
        WSB_TRY(enqueue_stage<T>(s, a, s->tA, true));
        if (prim_ext(s)) {
            WSB_TRY(tracer_stage(Y, tr_cur, s->trA, half_dt));
            WSB_TRY(tracer_ghosts(s->trA));
        }
        a.S = planes<const T>(s, s->tA); a.O = planes<T>(s, s->next); a.c = dt;
        WSB_TRY(enqueue_stage<T>(s, a, s->next, true));
        if (prim_ext(s)) WSB_TRY(tracer_stage(planes<const T>(s, s->tA), s->trA, tr_next, dt));
    } else {  // :325-455
        const bool classical = s->cfg.rk4_mode == WSB_RK4_CLASSICAL;
        a.S = Y; a.O = planes<T>(s, s->tA); a.c = half_dt;
        if (classical) a.KS = planes<T>(s, s->k1);
        WSB_TRY(enqueue_stage<T>(s, a, s->tA, true));
        a.S = planes<const T>(s, s->tA); a.O = planes<T>(s, s->tB); a.c = half_dt; a.KS = planes<T>(s, s->k2);
        WSB_TRY(enqueue_stage<T>(s, a, s->tB, true));
        a.S = planes<const T>(s, s->tB); a.O = planes<T>(s, s->tA); a.c = dt; a.KS = planes<T>(s, s->k3);
        WSB_TRY(enqueue_stage<T>(s, a, s->tA, true));
        a.S = planes<const T>(s, s->tA); a.O = planes<T>(s, s->next); a.KS = null_planes<T>();
        a.KA = planes<const T>(s, s->k2); a.KB = planes<const T>(s, s->k3);
        a.K1 = classical ? planes<const T>(s, s->k1) : null_planes<const T>();
        a.final_stage = 1;
        WSB_TRY(enqueue_stage<T>(s, a, s->next, true));
    }
    return WSB_OK;
}

template <typename T>
static cudaError_t launch_step(const wsb_sim *s, const Geometry<T> &g, const Physics<T> &ph, const StepArgs<T> &a,
                               cudaStream_t st) {
    return s->path == PATH_STEP_TMA ? launch_step_tma<T>(g, ph, a, s->nstages, st)
                                    : launch_step_fused<T>(g, ph, a, s->nstages, st);
}

// Arguments of one whole-step launch from the current state to `next` (both callers: the device-resident step
// and the streamed host step); row ranges are filled in by the caller.
template <typename T>
static StepArgs<T> step_args(const wsb_sim *s) {
    void *cur3[3] = {s->cur.f[0].base, s->cur.f[1].base, s->cur.f[2].base};
    StepArgs<T> a{};
    a.Y = planes<const T>(s, cur3);
    a.O = planes<T>(s, s->next);
    a.dt = (T)s->dt;
    a.half_dt = T(0.5f) * a.dt;  // `0.5f * dt_` (weather_simulation.cpp:249): one rounding
    a.dt6 = a.dt / T(6.0f);      // `dt_ / 6.0f` (:438)
    a.classical = s->cfg.rk4_mode == WSB_RK4_CLASSICAL;
    a.fold = s->cfg.arith_mode == WSB_ARITH_FOLDED;
    a.rows_per_chunk = s->path == PATH_STEP_TMA ? s->rpc : 0;  // the register twin keeps its own default
    return a;
}

// Whole-step kernels. Single GPU: one launch per step. Row slabs: the `halo` rows next to each slab edge
// need the neighbours' ghost rows, the rest does not, so a step is
//   edge stream (high priority): wait ghosts + previous interior -> ONE launch over both edge bands -> ev_edge
//   comm stream                : wait ev_edge -> ncclSend/ncclRecv of the new edge rows -> ev_halo
//   main stream                : wait previous ev_edge -> interior launch -> ev_interior
// i.e. the edge bands and their exchange run beside the interior sweep instead of in front of it.
template <typename T>
static int enqueue_step_fused(wsb_sim *s, bool chain) {
    const Geometry<T> g = s->cur.geom<T>();
    const Physics<T> ph = sim_physics<T>(s);
    StepArgs<T> a = step_args<T>(s);
    const int H = s->cur.H;
    const int halo = s->nstages;  // one ghost row per fused stage
    if (!s->comm) {
        a.y_begin = 0; a.y_end = H;
        if (s->ovl_enabled) {
            a.ovl_done = s->ovl_done;
            a.ovl_err = s->ovl_err;
            a.ovl_target = s->ovl_seq;  // every chunk row has been bumped by all strips of ovl_seq earlier steps
            a.ovl_chain = chain;        // only directly behind another protocol step on the stream
            ++s->ovl_seq;               // (counters and targets wrap together: the kernel compares differences)
        }
        WSB_CUDA(launch_step<T>(s, g, ph, a, s->stream));
        s->metrics.kernel_launches += 1;
        return WSB_OK;
    }
    if (s->peer_ok) {
        // fused ghost exchange: ONE launch; the band CTAs wait for, push to and signal the neighbours themselves
        const int which = s->next[0] == s->setB[0] ? 1 : s->next[0] == s->setA[0] ? 0 : -1;
        if (which < 0) return fail(WSB_ERR_RUNTIME, "fused ghost exchange: the state planes are not the mapped ones");
        const size_t lead = (size_t)kLeadRows * s->cur.pitch * sizeof(T);
        auto peer_planes = [&](const PeerLink &l) {
            Planes3<T> p;
            p.u = l.present ? (T *)((char *)l.ptr[which * 3 + 0] + lead) : nullptr;
            p.v = l.present ? (T *)((char *)l.ptr[which * 3 + 1] + lead) : nullptr;
            p.h = l.present ? (T *)((char *)l.ptr[which * 3 + 2] + lead) : nullptr;
            return p;
        };
        a.y_begin = 0; a.y_end = H;
        a.px.band = halo;
        a.px.up = peer_planes(s->peer_up);
        a.px.dn = peer_planes(s->peer_dn);
        a.px.up_H = s->peer_up.H;
        a.px.dn_H = s->peer_dn.H;
        a.px.wait = s->xflags;
        a.px.sig_up = s->peer_up.present ? (unsigned *)s->peer_up.ptr[6] + 1 : nullptr;
        a.px.sig_dn = s->peer_dn.present ? (unsigned *)s->peer_dn.ptr[6] + 0 : nullptr;
        a.px.target = s->x_seq;
        a.ovl_err = s->ovl_err;
        if (s->ovl_enabled) {
            a.ovl_done = s->ovl_done;
            a.ovl_target = s->ovl_seq;
            a.ovl_chain = chain && !s->nccl_ghosts_pending;
            ++s->ovl_seq;
        }
        if (s->nccl_ghosts_pending) {  // ghost rows that came through NCCL (first step after an upload)
            WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
            s->nccl_ghosts_pending = false;
        }
        WSB_CUDA(launch_step<T>(s, g, ph, a, s->stream));
        s->metrics.kernel_launches += 1;
        ++s->x_seq;
        s->band_ctas = step_tma_strips(s->nstages, s->dtype, s->cur.W) * s->cur.L;
        s->ghosts_in_flight = true;
        return WSB_OK;
    }
    // dependencies captured BEFORE this step re-records the events
    WSB_CUDA(cudaStreamWaitEvent(s->edge_stream, s->ev_interior, 0));  // previous step's interior rows
    WSB_CUDA(cudaStreamWaitEvent(s->edge_stream, s->ev_halo, 0));      // ghost rows of the current state
    WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_edge, 0));           // previous step's edge rows
    // edge bands: rows [0, e) and [max(e, H-halo), H) in one launch, short chunks for parallelism
    const int e = std::min(halo, H);
    StepArgs<T> b = a;
    b.y_begin = 0; b.y_end = e;
    b.y_begin2 = std::max(e, H - halo); b.y_end2 = H;
    b.rows_per_chunk = 2;
    WSB_CUDA(launch_step<T>(s, g, ph, b, s->edge_stream));
    s->metrics.kernel_launches += 1;
    WSB_CUDA(cudaEventRecord(s->ev_edge, s->edge_stream));
    WSB_TRY(sim_exchange(s, s->next, halo, s->ev_edge));
    if (H > 2 * halo) {
        a.y_begin = halo; a.y_end = H - halo;
        WSB_CUDA(launch_step<T>(s, g, ph, a, s->stream));
        s->metrics.kernel_launches += 1;
    }
    WSB_CUDA(cudaEventRecord(s->ev_interior, s->stream));
    // keep the main stream a superset of the edge stream: whatever is enqueued on it next (the next interior
    // sweep, a read-back, diagnostics) sees this step's edge rows
    WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_edge, 0));
    return WSB_OK;
}

// the constant T/p drift of the PrimitiveEquations model (weather_simulation.cpp:201-214, 311-319; F7):
// next.T = cur.T + dt*288.15f, next.p = cur.p + dt*1013.25f (the tendency grid holds its reset() values)
template <typename T>
static int enqueue_primitive_tp(wsb_sim *s) {
    const Geometry<T> g = s->cur.geom<T>();
    const int fields[2] = {WSB_FIELD_TEMPERATURE, WSB_FIELD_PRESSURE};
    for (int k = 0; k < 2; ++k) {
        WSB_TRY(grid_materialize(&s->cur, fields[k]));
        if (!s->alt[fields[k]].base) WSB_TRY(grid_alloc_plane(&s->cur, &s->alt[fields[k]].base));
    }
    WSB_CUDA(launch_axpy_const2<T>(g, (const T *)s->cur.origin(WSB_FIELD_TEMPERATURE),
                                   (T *)plane_origin(s, s->alt[WSB_FIELD_TEMPERATURE].base),
                                   (const T *)s->cur.origin(WSB_FIELD_PRESSURE),
                                   (T *)plane_origin(s, s->alt[WSB_FIELD_PRESSURE].base), (T)s->dt,
                                   (T)kResetValue[WSB_FIELD_TEMPERATURE], (T)kResetValue[WSB_FIELD_PRESSURE],
                                   s->stream));
    s->metrics.kernel_launches += 1;
    return WSB_OK;
}

// `flag_aware`: the consumer is a fused-exchange step, whose band CTAs wait for the neighbours' flags themselves.
// Anything else (diagnostics, a rendezvous) needs the ghost rows to be complete in stream order: if the last steps
// delivered them by peer stores, they are exchanged once more through NCCL (same values; a late peer store into the
// same rows is harmless).
static int sim_ensure_halo(wsb_sim *s, bool flag_aware = false) {
    if (s->comm && s->ghosts_in_flight && !flag_aware) {
        s->ghosts_in_flight = false;
        s->halo_valid = false;
    }
    if (!s->comm || s->halo_valid) return WSB_OK;
    void *cur3[3] = {s->cur.f[0].base, s->cur.f[1].base, s->cur.f[2].base};
    WSB_CUDA(cudaEventRecord(s->ev_edge, s->stream));      // "the current state is complete" for the comm stream
    WSB_CUDA(cudaEventRecord(s->ev_interior, s->stream));  // ... and for the edge stream
    const int depth = is_step_path(s->path) ? s->nstages : 1;
    WSB_TRY(sim_exchange(s, cur3, depth, s->ev_edge));
    s->halo_valid = true;
    return WSB_OK;
}

// one step, enqueued (weather_simulation.cpp:117-158 without the host-side bookkeeping).
// `follows_step`: the caller enqueued the previous step right before this one and nothing else in between -- the
// only situation in which the step-overlap launch attribute is used.
static int sim_enqueue_step(wsb_sim *s, bool follows_step = false) {
    WSB_TRY(sim_ensure_halo(s, s->peer_ok && is_step_path(s->path)));
    if (is_step_path(s->path)) {
        const bool chain = follows_step && s->cfg.model != WSB_MODEL_PRIMITIVE_EQUATIONS;
        if (s->dtype == WSB_F64) WSB_TRY(enqueue_step_fused<double>(s, chain));
        else WSB_TRY(enqueue_step_fused<float>(s, chain));
    } else {
        if (s->dtype == WSB_F64) WSB_TRY(enqueue_step_stages<double>(s));
        else WSB_TRY(enqueue_step_stages<float>(s));
    }
    if (s->cfg.model == WSB_MODEL_PRIMITIVE_EQUATIONS && !prim_ext(s)) {
        if (s->dtype == WSB_F64) WSB_TRY(enqueue_primitive_tp<double>(s));
        else WSB_TRY(enqueue_primitive_tp<float>(s));
    }
    // current_grid_.swap(next_grid_) (:217, :322, :454): every field of the two grids trades places
    for (int k = 0; k < 3; ++k) std::swap(s->cur.f[k].base, s->next[k]);
    for (int k = WSB_FIELD_PRESSURE; k <= WSB_FIELD_HUMIDITY; ++k) std::swap(s->cur.f[k], s->alt[k]);
    s->diag_dirty = true;  // step() recomputes vorticity/divergence of the new state (:149); done lazily
    // time and step bookkeeping in the reference's scalar type (:145-146)
    if (s->dtype == WSB_F32) s->time = (double)((float)s->time + (float)s->dt);
    else s->time += s->dt;
    s->step += 1;
    s->metrics.num_steps += 1;
    return WSB_OK;
}

// `align`: with row slabs, rendezvous all ranks on the device first (one-float all-reduce on the comm stream, which
// the stepping streams wait for), so that the timed region of a multi-step run starts within microseconds on every
// GPU instead of carrying the host-side skew between the processes into the first ghost exchanges.
static int sim_begin_timing(wsb_sim *s, bool align = false) {
    if (!s->run_open) {
        if (align && s->comm) {
            WSB_TRY(sim_ensure_halo(s, s->peer_ok));
            WSB_CUDA(cudaEventRecord(s->ev_align, s->stream));
            WSB_CUDA(cudaStreamWaitEvent(s->comm_stream, s->ev_align, 0));
            WSB_TRY(halo_align(s->comm, s->comm_stream));
            WSB_CUDA(cudaEventRecord(s->ev_halo, s->comm_stream));  // also still "ghost rows are valid"
            WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
        }
        WSB_CUDA(cudaEventRecord(s->ev_start, s->stream));
        s->run_open = true;
    }
    return WSB_OK;
}

static int sim_sync(wsb_sim *s) {
    WSB_CUDA(cudaSetDevice(s->cur.device));
    if (s->comm) {  // the timed region ends when the edge bands and the ghost-row exchange have finished too
        WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_edge, 0));
        WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
    }
    if (s->run_open) WSB_CUDA(cudaEventRecord(s->ev_stop, s->stream));
    if (s->edge_stream) WSB_CUDA(cudaStreamSynchronize(s->edge_stream));
    if (s->comm_stream) WSB_CUDA(cudaStreamSynchronize(s->comm_stream));
    WSB_CUDA(cudaStreamSynchronize(s->stream));
    if (s->ovl_err && *s->ovl_err)  // a step-overlap dependency that never arrived (protocol bug)
        return fail(WSB_ERR_RUNTIME, "step overlap: a chunk-row dependency timed out; the state is invalid");
    if (s->run_open) {
        float ms = 0.f;
        WSB_CUDA(cudaEventElapsedTime(&ms, s->ev_start, s->ev_stop));
        s->last_run_ms = ms;
        s->metrics.compute_time_ms += ms;
        s->run_open = false;
        if (s->comm) {
            float hms = 0.f;
            if (cudaEventElapsedTime(&hms, s->ev_h0, s->ev_h1) == cudaSuccess) s->metrics.halo_time_ms += hms;
            else cudaGetLastError();
        }
    }
    return WSB_OK;
}

static int sim_materialize_diagnostics(wsb_sim *s) {
    if (!s->diag_dirty) return WSB_OK;
    wsb_grid *g = &s->cur;
    WSB_TRY(sim_ensure_halo(s));
    if (s->comm) WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
    s->diag_dirty = false;
    return wsb_grid_calculate_diagnostics(g);
}

namespace wsb {

void grid_slab_position(const wsb_grid *g, int *row0, int *global_height) {
    *row0 = g->row0;
    *global_height = g->Hglobal;
}

int grid_upload_rows(wsb_grid *g, int field, const float *host_rows, int y0, int nrows) {
    if (!g || !host_rows || field < 0 || field > WSB_FIELD_HUMIDITY || y0 < 0 || nrows <= 0 || y0 + nrows > g->H)
        return fail(WSB_ERR_INVALID_ARGUMENT, "bad row block");
    WSB_CUDA(cudaSetDevice(g->device));
    if (g->owner && field <= WSB_FIELD_V) WSB_TRY(sim_materialize_diagnostics(g->owner));
    WSB_TRY(grid_materialize(g, field));
    const size_t es = elem_size(g->dtype), n = (size_t)nrows * g->W;
    const void *src = host_rows;
    std::vector<double> wide;
    if (g->dtype == WSB_F64) {
        wide.resize(n);
        for (size_t i = 0; i < n; ++i) wide[i] = (double)host_rows[i];
        src = wide.data();
    }
    for (int l = 0; l < g->L; ++l) {
        char *dst = (char *)g->origin(field) + ((size_t)l * g->level_stride + (size_t)y0 * g->pitch) * es;
        WSB_CUDA(copy_rows(dst, (size_t)g->pitch * es, src, (size_t)g->W * es, (size_t)g->W * es, (size_t)nrows,
                           cudaMemcpyHostToDevice, g->stream));
    }
    WSB_CUDA(cudaStreamSynchronize(g->stream));  // the caller reuses its block buffer
    if (g->owner && field <= WSB_FIELD_HEIGHT) g->owner->halo_valid = false;
    return WSB_OK;
}

// the bookkeeping every write into a field of the grid shares
static int grid_prepare_write(wsb_grid *g, int field) {
    WSB_CUDA(cudaSetDevice(g->device));
    if (g->owner && field <= WSB_FIELD_V) WSB_TRY(sim_materialize_diagnostics(g->owner));
    if (!g->f[field].base) WSB_TRY(grid_alloc_plane(g, &g->f[field].base));
    if (g->owner && field <= WSB_FIELD_HEIGHT) {
        g->owner->halo_valid = false;
        g->owner->ghosts_in_flight = false;
    }
    return WSB_OK;
}

int grid_record_event(wsb_grid *g, cudaEvent_t ev) {
    WSB_CUDA(cudaEventRecord(ev, g->stream));
    return WSB_OK;
}

int grid_upload_rows_async(wsb_grid *g, int field, const float *host_rows, int y0, int nrows) {
    if (!g || !host_rows || field < 0 || field > WSB_FIELD_HUMIDITY || y0 < 0 || nrows <= 0 || y0 + nrows > g->H)
        return fail(WSB_ERR_INVALID_ARGUMENT, "bad row block");
    if (g->dtype != WSB_F32) return grid_upload_rows(g, field, host_rows, y0, nrows);  // widened through a host copy
    WSB_TRY(grid_materialize(g, field));
    WSB_TRY(grid_prepare_write(g, field));
    const size_t es = sizeof(float);
    for (int l = 0; l < g->L; ++l) {
        char *dst = (char *)g->origin(field) + ((size_t)l * g->level_stride + (size_t)y0 * g->pitch) * es;
        WSB_CUDA(copy_rows(dst, (size_t)g->pitch * es, host_rows, (size_t)g->W * es, (size_t)g->W * es, (size_t)nrows,
                           cudaMemcpyHostToDevice, g->stream));
    }
    return WSB_OK;
}

int grid_fill_uniform(wsb_grid *g, int field, float value) {
    if (!g || field < 0 || field > WSB_FIELD_HUMIDITY) return fail(WSB_ERR_INVALID_ARGUMENT, "bad field");
    WSB_TRY(grid_prepare_write(g, field));
    if (g->dtype == WSB_F64) WSB_CUDA(launch_fill<double>(g->geom<double>(), (double *)g->origin(field), (double)value, false, g->stream));
    else WSB_CUDA(launch_fill<float>(g->geom<float>(), (float *)g->origin(field), value, false, g->stream));
    return WSB_OK;
}

int grid_fill_separable(wsb_grid *g, int field, const float *rowv, const float *colv) {
    if (!g || !rowv || field < 0 || field > WSB_FIELD_HUMIDITY) return fail(WSB_ERR_INVALID_ARGUMENT, "bad field");
    WSB_TRY(grid_prepare_write(g, field));
    float *dev = nullptr;
    WSB_CUDA(cudaMalloc(&dev, sizeof(float) * ((size_t)g->H + (colv ? g->W : 0))));
    cudaError_t e = cudaMemcpyAsync(dev, rowv, sizeof(float) * g->H, cudaMemcpyHostToDevice, g->stream);
    if (e == cudaSuccess && colv)
        e = cudaMemcpyAsync(dev + g->H, colv, sizeof(float) * g->W, cudaMemcpyHostToDevice, g->stream);
    if (e == cudaSuccess) {
        if (g->dtype == WSB_F64)
            e = launch_expand_separable<double>(g->geom<double>(), (double *)g->origin(field), dev, colv ? dev + g->H : nullptr, g->stream);
        else
            e = launch_expand_separable<float>(g->geom<float>(), (float *)g->origin(field), dev, colv ? dev + g->H : nullptr, g->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->stream);  // the vectors are the caller's, the scratch is freed here
    cudaFree(dev);
    if (e != cudaSuccess) return cuda_fail(e, "separable initial condition", __FILE__, __LINE__);
    return WSB_OK;
}

}  // namespace wsb

// ------------------------------------------------------------------------------ C-ABI: misc --
extern "C" {

const char *wsb_version(void) { return "weather_b200 0.1.0 (sm_100a)"; }

const char *wsb_last_error(void) { return g_last_error.c_str(); }

int wsb_device_count(int *count) {
    if (!count) return fail(WSB_ERR_INVALID_ARGUMENT, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    *count = n;
    return WSB_OK;
}

int wsb_device_capabilities(int device_id, wsb_device_caps *out) {
    if (!out) return fail(WSB_ERR_INVALID_ARGUMENT, "out is NULL");
    std::memset(out, 0, sizeof(*out));
    cudaDeviceProp p;
    WSB_CUDA(cudaGetDeviceProperties(&p, device_id));
    out->compute_capability_major = p.major;
    out->compute_capability_minor = p.minor;
    out->multiprocessors = p.multiProcessorCount;
    out->cuda_cores = p.multiProcessorCount * 128;  // 128 FP32 lanes per SM on sm_80..sm_100
    out->global_memory = p.totalGlobalMem;
    out->shared_memory_per_block = p.sharedMemPerBlock;
    out->max_threads_per_block = p.maxThreadsPerBlock;
    out->max_threads_per_multiprocessor = p.maxThreadsPerMultiProcessor;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device_id) == cudaSuccess) out->clock_rate_khz = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMemoryClockRate, device_id) == cudaSuccess) out->memory_clock_rate_khz = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrGlobalMemoryBusWidth, device_id) == cudaSuccess) out->memory_bus_width = v;
    cudaGetLastError();
    // DeviceType (gpu_adaptability.hpp:23-30; detection rules gpu_adaptability.cpp:40-80)
    if (p.major == 8 && p.minor == 7) out->device_type = 2;
    else if (p.major == 7 && p.minor == 5) out->device_type = 3;
    else if (p.major >= 8) out->device_type = 4;
    else out->device_type = 5;
    out->compute_power_ratio = (float)out->cuda_cores * (float)out->clock_rate_khz / (8.0f * 3.0e6f * 16.0f);
    std::snprintf(out->device_name, sizeof(out->device_name), "%s", p.name);
    return WSB_OK;
}

int wsb_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(WSB_ERR_INVALID_ARGUMENT, "out is NULL");
    WSB_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return WSB_OK;
}

int wsb_host_free(void *ptr) {
    if (ptr) WSB_CUDA(cudaFreeHost(ptr));
    return WSB_OK;
}

int wsb_nccl_get_unique_id(void *out128) { return nccl_get_unique_id(out128); }

int wsb_partition_rows(int32_t grid_height, int32_t nranks, int32_t rank, int32_t *row0, int32_t *nrows) {
    if (grid_height <= 0 || nranks <= 0 || rank < 0 || rank >= nranks || nranks > grid_height)
        return fail(WSB_ERR_INVALID_ARGUMENT, "invalid partition request");
    const int base = grid_height / nranks, rem = grid_height % nranks;
    if (nrows) *nrows = base + (rank < rem ? 1 : 0);
    if (row0) *row0 = rank * base + std::min(rank, rem);
    return WSB_OK;
}

// ------------------------------------------------------------------------------ C-ABI: grid --
int wsb_grid_create(int32_t width, int32_t height, int32_t num_levels, double dx, double dy, int32_t dtype,
                    int32_t device_id, wsb_grid **out) {
    if (!out) return fail(WSB_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = nullptr;
    std::unique_ptr<wsb_grid> g(new wsb_grid());
    int st = grid_init(g.get(), width, height, num_levels, dx, dy, dtype, device_id, 0, height, nullptr);
    if (st != WSB_OK) {
        grid_fini(g.get());
        return st;
    }
    *out = g.release();
    return WSB_OK;
}

void wsb_grid_destroy(wsb_grid *grid) {
    if (!grid || grid->owner) return;  // a simulation's grid dies with the simulation
    cudaSetDevice(grid->device);
    cudaStreamSynchronize(grid->stream);
    grid_fini(grid);
    delete grid;
}

int wsb_grid_reset(wsb_grid *g) {
    if (!g) return fail(WSB_ERR_INVALID_ARGUMENT, "grid is NULL");
    WSB_CUDA(cudaSetDevice(g->device));
    for (int k = 0; k < WSB_NUM_FIELDS; ++k) {
        if (k <= WSB_FIELD_HEIGHT) WSB_TRY(grid_fill(g, g->f[k].base, (double)kResetValue[k]));
        else grid_release(g, k, (double)kResetValue[k]);
    }
    if (g->owner) {
        g->owner->diag_dirty = false;
        g->owner->halo_valid = false;
    }
    return WSB_OK;
}

int wsb_grid_get_info(const wsb_grid *g, wsb_grid_info *out) {
    if (!g || !out) return fail(WSB_ERR_INVALID_ARGUMENT, "NULL argument");
    out->width = g->W; out->height = g->H; out->num_levels = g->L; out->dtype = g->dtype;
    out->dx = g->dx; out->dy = g->dy; out->device_id = g->device; out->reserved = 0;
    return WSB_OK;
}

int wsb_grid_set_spacing(wsb_grid *g, double dx, double dy) {
    if (!g) return fail(WSB_ERR_INVALID_ARGUMENT, "grid is NULL");
    if (!(dx > 0.0) || !(dy > 0.0))
        return fail(WSB_ERR_INVALID_ARGUMENT, "Grid spacing must be positive");  // weather_grid.cpp:74-76
    if (g->owner && g->owner->cfg.physics_mode == WSB_PHYSICS_EXTENDED && g->owner->path == PATH_STEP_TMA) {
        const double fx = g->dtype == WSB_F32 ? (double)(float)dx : dx, fy = g->dtype == WSB_F32 ? (double)(float)dy : dy;
        if (!is_pow2(2.0 * fx) || !is_pow2(2.0 * fy))  // that kernel family exists for exact-reciprocal spacing only
            return fail(WSB_ERR_INVALID_ARGUMENT, "extended physics on the whole-step kernels needs power-of-two spacing; "
                                                  "create the simulation with kernel_variant STAGE_DIRECT for this spacing");
    }
    g->dx = g->dtype == WSB_F32 ? (double)(float)dx : dx;
    g->dy = g->dtype == WSB_F32 ? (double)(float)dy : dy;
    set_spacing_derived(g);
    return WSB_OK;
}

static int check_shape(const wsb_grid *g, int64_t levels, int64_t rows, int64_t cols) {
    if (rows != g->H || cols != g->W || levels != g->L)
        return fail(WSB_ERR_RUNTIME, "Array dimensions must match field dimensions");  // python_bindings.cpp:70-72
    return WSB_OK;
}

int wsb_grid_set_field(wsb_grid *g, int32_t field, const void *host, int32_t host_dtype, int64_t levels,
                       int64_t rows, int64_t cols) {
    if (!g || !host) return fail(WSB_ERR_INVALID_ARGUMENT, "NULL argument");
    if (field < 0 || field > WSB_FIELD_HUMIDITY) return fail(WSB_ERR_INVALID_ARGUMENT, "field cannot be set");
    if (host_dtype != WSB_F32 && host_dtype != WSB_F64) return fail(WSB_ERR_INVALID_ARGUMENT, "unknown host dtype");
    WSB_TRY(check_shape(g, levels, rows, cols));
    WSB_CUDA(cudaSetDevice(g->device));
    if (g->owner && field <= WSB_FIELD_V) WSB_TRY(sim_materialize_diagnostics(g->owner));
    if (!g->f[field].base) WSB_TRY(grid_alloc_plane(g, &g->f[field].base));
    const size_t es = elem_size(g->dtype);
    const size_t n = (size_t)levels * rows * cols;
    const void *src = host;
    std::vector<char> conv;
    if (host_dtype != g->dtype) {  // forcecast, like py::array_t<scalar_t> (python_bindings.cpp:60)
        conv.resize(n * es);
        if (g->dtype == WSB_F32) {
            const double *s = (const double *)host; float *d = (float *)conv.data();
            for (size_t i = 0; i < n; ++i) d[i] = (float)s[i];
        } else {
            const float *s = (const float *)host; double *d = (double *)conv.data();
            for (size_t i = 0; i < n; ++i) d[i] = (double)s[i];
        }
        src = conv.data();
    }
    WSB_CUDA(cudaEventRecord(g->ev0, g->stream));
    for (int64_t l = 0; l < levels; ++l) {
        char *dst = (char *)g->origin(field) + (size_t)l * g->level_stride * es;
        const char *s = (const char *)src + (size_t)l * rows * cols * es;
        WSB_CUDA(copy_rows(dst, (size_t)g->pitch * es, s, (size_t)cols * es, (size_t)cols * es, (size_t)rows,
                                   cudaMemcpyHostToDevice, g->stream));
    }
    WSB_CUDA(cudaEventRecord(g->ev1, g->stream));
    WSB_CUDA(cudaStreamSynchronize(g->stream));  // the host buffer is reusable on return
    float ms = 0.f;
    WSB_CUDA(cudaEventElapsedTime(&ms, g->ev0, g->ev1));
    g->transfer_ms += ms;
    if (g->owner) {
        g->owner->metrics.memory_transfer_time_ms += ms;
        if (field <= WSB_FIELD_HEIGHT) {
            g->owner->halo_valid = false;
            g->owner->ghosts_in_flight = false;
        }
    }
    return WSB_OK;
}

int wsb_grid_get_field(wsb_grid *g, int32_t field, void *host, int32_t host_dtype, int64_t levels, int64_t rows,
                       int64_t cols) {
    if (!g || !host) return fail(WSB_ERR_INVALID_ARGUMENT, "NULL argument");
    if (field < 0 || field >= WSB_NUM_FIELDS) return fail(WSB_ERR_INVALID_ARGUMENT, "unknown field");
    if (host_dtype != WSB_F32 && host_dtype != WSB_F64) return fail(WSB_ERR_INVALID_ARGUMENT, "unknown host dtype");
    WSB_TRY(check_shape(g, levels, rows, cols));
    WSB_CUDA(cudaSetDevice(g->device));
    if (g->owner && field >= WSB_FIELD_VORTICITY) WSB_TRY(sim_materialize_diagnostics(g->owner));
    const size_t es = elem_size(g->dtype);
    const size_t n = (size_t)levels * rows * cols;
    if (!g->f[field].base) {  // never touched: uniformly its reset()/swap value
        WSB_CUDA(cudaStreamSynchronize(g->stream));
        if (host_dtype == WSB_F32) std::fill((float *)host, (float *)host + n, (float)g->f[field].uniform);
        else std::fill((double *)host, (double *)host + n, g->f[field].uniform);
        return WSB_OK;
    }
    std::vector<char> conv;
    void *dst = host;
    if (host_dtype != g->dtype) {
        conv.resize(n * es);
        dst = conv.data();
    }
    WSB_CUDA(cudaEventRecord(g->ev0, g->stream));
    for (int64_t l = 0; l < levels; ++l) {
        const char *s = (const char *)g->origin(field) + (size_t)l * g->level_stride * es;
        char *d = (char *)dst + (size_t)l * rows * cols * es;
        WSB_CUDA(copy_rows(d, (size_t)cols * es, s, (size_t)g->pitch * es, (size_t)cols * es, (size_t)rows,
                                   cudaMemcpyDeviceToHost, g->stream));
    }
    WSB_CUDA(cudaEventRecord(g->ev1, g->stream));
    WSB_CUDA(cudaStreamSynchronize(g->stream));
    float ms = 0.f;
    WSB_CUDA(cudaEventElapsedTime(&ms, g->ev0, g->ev1));
    g->transfer_ms += ms;
    if (g->owner) g->owner->metrics.memory_transfer_time_ms += ms;
    if (host_dtype != g->dtype) {
        if (host_dtype == WSB_F32) {
            const double *s = (const double *)conv.data(); float *d = (float *)host;
            for (size_t i = 0; i < n; ++i) d[i] = (float)s[i];
        } else {
            const float *s = (const float *)conv.data(); double *d = (double *)host;
            for (size_t i = 0; i < n; ++i) d[i] = (double)s[i];
        }
    }
    return WSB_OK;
}

int wsb_grid_calculate_diagnostics(wsb_grid *g) {
    if (!g) return fail(WSB_ERR_INVALID_ARGUMENT, "grid is NULL");
    WSB_CUDA(cudaSetDevice(g->device));
    for (int k = WSB_FIELD_VORTICITY; k <= WSB_FIELD_DIVERGENCE; ++k)
        if (!g->f[k].base) WSB_TRY(grid_alloc_plane(g, &g->f[k].base));
    if (g->owner) {
        wsb_sim *s = g->owner;
        WSB_TRY(sim_ensure_halo(s));
        if (s->comm) WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
        s->diag_dirty = false;
    }
    if (g->dtype == WSB_F64) {
        WSB_CUDA(launch_diagnostics<double>(g->geom<double>(), g->physics<double>(0, 0), (const double *)g->origin(0),
                                            (const double *)g->origin(1), (double *)g->origin(WSB_FIELD_VORTICITY),
                                            (double *)g->origin(WSB_FIELD_DIVERGENCE), g->stream));
    } else {
        WSB_CUDA(launch_diagnostics<float>(g->geom<float>(), g->physics<float>(0, 0), (const float *)g->origin(0),
                                           (const float *)g->origin(1), (float *)g->origin(WSB_FIELD_VORTICITY),
                                           (float *)g->origin(WSB_FIELD_DIVERGENCE), g->stream));
    }
    return WSB_OK;
}

int wsb_grid_swap(wsb_grid *a, wsb_grid *b) {
    if (!a || !b) return fail(WSB_ERR_INVALID_ARGUMENT, "grid is NULL");
    if (a->W != b->W || a->H != b->H || a->L != b->L)
        return fail(WSB_ERR_INVALID_ARGUMENT, "Cannot swap grids of different dimensions");  // weather_grid.cpp:125-127
    if (a->dtype != b->dtype || a->device != b->device)
        return fail(WSB_ERR_INVALID_ARGUMENT, "Cannot swap grids of different dtype or device");
    if (a->owner) WSB_TRY(sim_materialize_diagnostics(a->owner));
    if (b->owner) WSB_TRY(sim_materialize_diagnostics(b->owner));
    WSB_CUDA(cudaStreamSynchronize(a->stream));
    WSB_CUDA(cudaStreamSynchronize(b->stream));
    for (int k = 0; k < WSB_NUM_FIELDS; ++k) std::swap(a->f[k], b->f[k]);
    for (wsb_grid *g : {a, b})
        if (g->owner) {
            g->owner->halo_valid = false;
            g->owner->ghosts_in_flight = false;
        }
    return WSB_OK;
}

int wsb_grid_device_pointer(wsb_grid *g, int32_t field, void **dev_ptr, int64_t *pitch_elems) {
    if (!g || !dev_ptr) return fail(WSB_ERR_INVALID_ARGUMENT, "NULL argument");
    if (field < 0 || field >= WSB_NUM_FIELDS) return fail(WSB_ERR_INVALID_ARGUMENT, "unknown field");
    WSB_CUDA(cudaSetDevice(g->device));
    if (g->owner && field >= WSB_FIELD_VORTICITY) WSB_TRY(sim_materialize_diagnostics(g->owner));
    WSB_TRY(grid_materialize(g, field));
    // the consumer reads on ITS stream: everything this library has enqueued for the grid (pending asynchronous
    // steps, the fill of a lazily materialised field, the diagnostics kernel) must have completed first
    if (g->owner) WSB_TRY(sim_sync(g->owner));
    else WSB_CUDA(cudaStreamSynchronize(g->stream));
    *dev_ptr = g->origin(field);
    if (pitch_elems) *pitch_elems = g->pitch;
    return WSB_OK;
}

// ------------------------------------------------------------------------------- C-ABI: sim --
static void sim_free(wsb_sim *s) {
    cudaSetDevice(s->cur.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->comm_stream) cudaStreamSynchronize(s->comm_stream);
    if (s->edge_stream) cudaStreamSynchronize(s->edge_stream);
    if (s->comm && s->peer_ok) {
        // The neighbours store into this rank's ghost rows. For every fused step this rank has run, each neighbour
        // runs the same step (SPMD) and bumps this rank's flag once per band CTA AFTER its stores: wait (bounded)
        // until both flags have seen all x_seq steps, i.e. no store into this memory is in flight or still to come.
        const unsigned want = s->x_seq * (unsigned)s->band_ctas;
        for (int tries = 0; tries < 5000; ++tries) {
            unsigned f[2] = {want, want};
            if (cudaMemcpy(f, s->xflags, sizeof(f), cudaMemcpyDeviceToHost) != cudaSuccess) break;
            const bool up_done = !s->peer_up.present || (int)(f[0] - want) >= 0;
            const bool dn_done = !s->peer_dn.present || (int)(f[1] - want) >= 0;
            if (up_done && dn_done) break;
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        peer_close(&s->peer_up);
        peer_close(&s->peer_dn);
    }
    if (s->xflags) cudaFree(s->xflags);
    if (s->comm) halo_comm_destroy(s->comm);
    sim_free3(s->trA);
    sim_free3(s->next); sim_free3(s->tA); sim_free3(s->tB); sim_free3(s->k1); sim_free3(s->k2); sim_free3(s->k3);
    for (int k = 0; k < WSB_NUM_FIELDS; ++k)
        if (s->alt[k].base) cudaFree(s->alt[k].base);
    if (s->d_partial) cudaFree(s->d_partial);
    if (s->ovl_done) cudaFree(s->ovl_done);
    if (s->ovl_err) cudaFreeHost(s->ovl_err);
    grid_fini(&s->cur);
    cudaEvent_t evs[] = {s->ev_start, s->ev_stop, s->ev_edge, s->ev_halo, s->ev_h0, s->ev_h1, s->ev_interior, s->ev_align,
                         s->ev_tracer};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
    if (s->edge_stream) cudaStreamDestroy(s->edge_stream);
    if (s->h2d_stream) {
        cudaStreamSynchronize(s->h2d_stream);
        cudaStreamSynchronize(s->d2h_stream);
        for (int i = 0; i < wsb_sim::kMaxSlabs; ++i) {
            if (s->ev_up[i]) cudaEventDestroy(s->ev_up[i]);
            if (s->ev_done[i]) cudaEventDestroy(s->ev_done[i]);
        }
        cudaStreamDestroy(s->h2d_stream);
        cudaStreamDestroy(s->d2h_stream);
    }
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int wsb_sim_create(const wsb_config *config, wsb_sim **out) {
    if (!config || !out) return fail(WSB_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = nullptr;
    if (config->struct_size != sizeof(wsb_config))
        return fail(WSB_ERR_INVALID_ARGUMENT, "wsb_config.struct_size does not match this library");
    wsb_config c = *config;
    if (c.grid_width <= 0 || c.grid_height <= 0 || c.num_levels <= 0)
        return fail(WSB_ERR_INVALID_ARGUMENT, "Grid dimensions must be positive");  // weather_grid.cpp:50-52
    if (c.dtype != WSB_F32 && c.dtype != WSB_F64) return fail(WSB_ERR_INVALID_ARGUMENT, "unknown dtype");
    if (c.model < 0 || c.model > WSB_MODEL_GENERAL) return fail(WSB_ERR_INVALID_ARGUMENT, "unknown model");
    if (c.integration_method < 0 || c.integration_method > WSB_INT_SEMI_IMPLICIT)
        return fail(WSB_ERR_INVALID_ARGUMENT, "unknown integration method");
    if (c.physics_mode != WSB_PHYSICS_REFERENCE && c.physics_mode != WSB_PHYSICS_EXTENDED)
        return fail(WSB_ERR_INVALID_ARGUMENT, "unknown physics_mode");
    if (c.arith_mode != WSB_ARITH_STRICT && c.arith_mode != WSB_ARITH_FOLDED)
        return fail(WSB_ERR_INVALID_ARGUMENT, "unknown arith_mode");
    if (const char *e = std::getenv("WSB_ARITH"))  // A/B runs: WSB_ARITH=folded|strict overrides the configuration
        c.arith_mode = std::strcmp(e, "folded") == 0 ? WSB_ARITH_FOLDED : WSB_ARITH_STRICT;
    if (c.nranks < 1) c.nranks = 1;
    if (c.rank < 0 || c.rank >= c.nranks) return fail(WSB_ERR_INVALID_ARGUMENT, "rank out of range");
    if (c.nranks > 1 && !c.nccl_unique_id) return fail(WSB_ERR_INVALID_ARGUMENT, "nccl_unique_id is required when nranks > 1");
    if (c.nranks > c.grid_height) return fail(WSB_ERR_INVALID_ARGUMENT, "more ranks than grid rows");

    wsb_sim *s = new wsb_sim();
    s->cfg = c;
    s->cfg.nccl_unique_id = nullptr;
    s->dtype = c.dtype;
    if (c.dtype == WSB_F32) {  // the reference's config fields are float (weather_sim.hpp:166-172)
        s->cfg.dt = (double)(float)c.dt; s->cfg.gravity = (double)(float)c.gravity;
        s->cfg.coriolis_f = (double)(float)c.coriolis_f; s->cfg.max_time = (double)(float)c.max_time;
        s->cfg.beta = (double)(float)c.beta; s->cfg.viscosity = (double)(float)c.viscosity;
        s->cfg.diffusivity = (double)(float)c.diffusivity;
    }
    s->dt = s->cfg.dt;
    s->nstages = effective_stages(c);
    wsb_partition_rows(c.grid_height, c.nranks, c.rank, &s->row0, &s->nrows);

    int st = WSB_OK;
    do {
        if (cudaSetDevice(c.device_id) != cudaSuccess) {
            cudaGetLastError();
            int ndev = 0;
            if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
                cudaGetLastError();
                st = fail(WSB_ERR_CUDA, "no CUDA device available: libweather_b200 has no CPU fallback");
            } else {
                st = fail(WSB_ERR_INVALID_ARGUMENT, "device_id out of range");
            }
            break;
        }
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            st = fail(WSB_ERR_CUDA, "no CUDA device available: libweather_b200 has no CPU fallback");
            break;
        }
        st = grid_init(&s->cur, c.grid_width, s->nrows, c.num_levels, c.dx, c.dy, c.dtype, c.device_id, s->row0,
                       c.grid_height, s->stream);
        if (st != WSB_OK) break;
        s->cur.owner = s;
        for (int k = 0; k < WSB_NUM_FIELDS; ++k) s->alt[k].uniform = (double)kResetValue[k];
        cudaEvent_t *evs[] = {&s->ev_start, &s->ev_stop, &s->ev_edge, &s->ev_halo, &s->ev_h0, &s->ev_h1, &s->ev_interior,
                              &s->ev_align, &s->ev_tracer};
        for (cudaEvent_t *e : evs)
            if (cudaEventCreate(e) != cudaSuccess) st = cuda_fail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__);
        if (st != WSB_OK) break;

        // kernel path. The whole-step kernels implement the reference (aliased) RK4 combine only.
        int want = c.kernel_variant;
        const bool classical4 = s->nstages == 4 && c.rk4_mode == WSB_RK4_CLASSICAL;
        // extended physics: whole-step TMA kernel on power-of-two spacing, else the per-stage kernel (any spacing)
        const bool ext = c.physics_mode == WSB_PHYSICS_EXTENDED;
        // (the extended Primitive model needs the midpoint velocity in HBM for its tracer stage: per-stage kernels)
        const bool pext = ext && c.model == WSB_MODEL_PRIMITIVE_EQUATIONS;
        const bool reg_ok = step_fused_supported(s->nstages, s->dtype) && !classical4 && !ext;
        const bool tma_ok = step_tma_supported(s->nstages, s->dtype) && !pext &&  // incl. the classical RK4 opt-in
                            (!ext || (is_pow2(2.0 * (double)(float)c.dx) && is_pow2(2.0 * (double)(float)c.dy)));
        if (want == WSB_KERNEL_AUTO)
            want = tma_ok ? WSB_KERNEL_STEP_FUSED_TMA : reg_ok ? WSB_KERNEL_STEP_FUSED_REG : WSB_KERNEL_STAGE_DIRECT;
        if (want == WSB_KERNEL_STEP_FUSED_REG || want == WSB_KERNEL_STEP_FUSED_TMA) {
            if (!(want == WSB_KERNEL_STEP_FUSED_REG ? reg_ok : tma_ok)) {
                st = fail(WSB_ERR_INVALID_ARGUMENT, "kernel_variant STEP_FUSED is not available for this configuration");
                break;
            }
            s->path = want == WSB_KERNEL_STEP_FUSED_REG ? PATH_STEP_REG : PATH_STEP_TMA;
        } else if (want == WSB_KERNEL_STAGE_DIRECT) {
            s->path = PATH_STAGE_DIRECT;
        } else {
            st = fail(WSB_ERR_INVALID_ARGUMENT, "unknown kernel_variant");
            break;
        }
        // every slab (this rank's AND its neighbours') must be at least as deep as the ghost band; tested on the
        // smallest slab of the balanced partition, which all ranks know, so that every rank accepts or rejects the
        // configuration identically BEFORE ncclCommInitRank (a rank failing alone would leave the others blocked)
        if (c.nranks > 1 && c.grid_height / c.nranks < (is_step_path(s->path) ? s->nstages : 1)) {
            st = fail(WSB_ERR_INVALID_ARGUMENT, "row slab thinner than the ghost depth");
            break;
        }

        if ((st = sim_alloc3(s, s->next)) != WSB_OK) break;
        if (!is_step_path(s->path)) {
            if (s->nstages >= 2 && (st = sim_alloc3(s, s->tA)) != WSB_OK) break;
            if (s->nstages == 4) {
                if ((st = sim_alloc3(s, s->tB)) != WSB_OK) break;
                if ((st = sim_alloc3(s, s->k2)) != WSB_OK) break;
                if ((st = sim_alloc3(s, s->k3)) != WSB_OK) break;
                if (c.rk4_mode == WSB_RK4_CLASSICAL && (st = sim_alloc3(s, s->k1)) != WSB_OK) break;
            }
        }
        // ghost rows of scratch planes are read at slab edges before being written on rank 0 / G-1: zero them
        void **sets[] = {s->next, s->tA, s->tB};
        for (void **set : sets)
            for (int k = 0; k < 3; ++k)
                if (set[k] && (st = grid_fill(&s->cur, set[k], 0.0)) != WSB_OK) break;
        if (st != WSB_OK) break;

        if (c.nranks > 1) {
            int prio_lo = 0, prio_hi = 0;
            cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
            if (cudaStreamCreateWithPriority(&s->edge_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
                cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
                st = cuda_fail(cudaGetLastError(), "cudaStreamCreate(comm)", __FILE__, __LINE__);
                break;
            }
            if ((st = halo_comm_create(c.rank, c.nranks, c.nccl_unique_id, &s->comm)) != WSB_OK) break;
        }
        // fused ghost exchange over peer memory: TMA whole-step kernel, slabs at least three ghost bands deep (tested
        // on the smallest slab of the partition so that every rank decides alike), every neighbour mappable
        if (c.nranks > 1 && s->path == PATH_STEP_TMA && c.grid_height / c.nranks >= 3 * s->nstages) {
            if (cudaMalloc(&s->xflags, 2 * sizeof(unsigned)) != cudaSuccess ||
                cudaMemsetAsync(s->xflags, 0, 2 * sizeof(unsigned), s->stream) != cudaSuccess) {
                st = cuda_fail(cudaGetLastError(), "cudaMalloc(exchange flags)", __FILE__, __LINE__);
                break;
            }
            for (int k = 0; k < 3; ++k) {
                s->setA[k] = s->cur.f[k].base;
                s->setB[k] = s->next[k];
            }
            void *local[kPeerPointers] = {s->setA[0], s->setA[1], s->setA[2], s->setB[0], s->setB[1], s->setB[2], s->xflags};
            if (cudaStreamSynchronize(s->stream) != cudaSuccess) {
                st = cuda_fail(cudaGetLastError(), "cudaStreamSynchronize", __FILE__, __LINE__);
                break;
            }
            if ((st = peer_setup(s->comm, local, s->nrows, &s->peer_up, &s->peer_dn, &s->peer_ok, s->comm_stream)) != WSB_OK) break;
        }
        // step overlap: TMA whole-step kernel, one launch per step (single GPU, or slabs with the fused exchange);
        // not for the Primitive model (its T/p pass sits between two steps). WSB_STEP_OVERLAP=0 switches it off.
        const bool one_launch = s->path == PATH_STEP_TMA && (c.nranks == 1 || s->peer_ok);
        {
            const char *e = std::getenv("WSB_STEP_OVERLAP");
            s->ovl_enabled = one_launch && c.model != WSB_MODEL_PRIMITIVE_EQUATIONS && !(e && std::atoi(e) == 0);
        }
        s->rpc = step_tma_rows_per_chunk(s->nstages, s->dtype, c.grid_width, s->nrows, c.num_levels, s->ovl_enabled);
        if (one_launch) {
            // one counter per chunk row and level; with the fused exchange the chunk rows are [bands | interior chunks]
            const int chunk_rows = s->peer_ok ? 2 + (s->nrows - 2 * s->nstages + s->rpc - 1) / s->rpc
                                              : (s->nrows + s->rpc - 1) / s->rpc;
            const size_t n = (size_t)c.num_levels * chunk_rows;
            // the error word lives in mapped host memory: the host reads it after a sync without any copy
            if ((s->ovl_enabled && (cudaMalloc(&s->ovl_done, n * sizeof(unsigned)) != cudaSuccess ||
                                    cudaMemsetAsync(s->ovl_done, 0, n * sizeof(unsigned), s->stream) != cudaSuccess)) ||
                cudaHostAlloc((void **)&s->ovl_err, sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess) {
                st = cuda_fail(cudaGetLastError(), "cudaMalloc(step overlap counters)", __FILE__, __LINE__);
                break;
            }
            *s->ovl_err = 0;
        }
        s->npartial = 1024;
        if (cudaMalloc(&s->d_partial, sizeof(double) * 2 * s->npartial) != cudaSuccess) {
            st = cuda_fail(cudaGetLastError(), "cudaMalloc(partials)", __FILE__, __LINE__);
            break;
        }
        if (cudaStreamSynchronize(s->stream) != cudaSuccess) {
            st = cuda_fail(cudaGetLastError(), "cudaStreamSynchronize", __FILE__, __LINE__);
            break;
        }
    } while (0);
    if (st != WSB_OK) {
        std::string keep = g_last_error;
        sim_free(s);
        g_last_error = keep;
        return st;
    }
    *out = s;
    return WSB_OK;
}

void wsb_sim_destroy(wsb_sim *sim) {
    if (sim) sim_free(sim);
}

int wsb_sim_initialize(wsb_sim *s) {
    if (!s) return fail(WSB_ERR_INVALID_ARGUMENT, "sim is NULL");
    WSB_TRY(sim_sync(s));
    s->time = 0.0;
    s->step = 0;
    std::memset(&s->metrics, 0, sizeof(s->metrics));
    return wsb_grid_reset(&s->cur);
}

wsb_grid *wsb_sim_current_grid(wsb_sim *s) { return s ? &s->cur : nullptr; }

int wsb_sim_advance_async(wsb_sim *s, int32_t num_steps) {
    if (!s) return fail(WSB_ERR_INVALID_ARGUMENT, "sim is NULL");
    WSB_CUDA(cudaSetDevice(s->cur.device));
    WSB_TRY(sim_begin_timing(s, num_steps > 1));
    for (int i = 0; i < num_steps; ++i) WSB_TRY(sim_enqueue_step(s, i > 0));
    return WSB_OK;
}

int wsb_sim_synchronize(wsb_sim *s) {
    if (!s) return fail(WSB_ERR_INVALID_ARGUMENT, "sim is NULL");
    return sim_sync(s);
}

int wsb_sim_step(wsb_sim *s) {
    WSB_TRY(wsb_sim_advance_async(s, 1));
    return sim_sync(s);
}

}  // extern "C"

// Streamed host step: see include/weather_b200.h. Row slabs flow H2D -> whole-step kernel -> D2H on three
// streams; slab i can be stepped up to `halo` rows short of its end as soon as it has landed.
template <typename T>
static int step_host_pipelined(wsb_sim *s, const void *const in[3], void *const out[3]) {
    wsb_grid *g = &s->cur;
    const int H = g->H, W = g->W, halo = s->nstages;
    const size_t es = sizeof(T), row_bytes = (size_t)W * es, pitch_bytes = (size_t)g->pitch * es;
    if (!s->h2d_stream) {
        WSB_CUDA(cudaStreamCreateWithFlags(&s->h2d_stream, cudaStreamNonBlocking));
        WSB_CUDA(cudaStreamCreateWithFlags(&s->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < wsb_sim::kMaxSlabs; ++i) {
            WSB_CUDA(cudaEventCreateWithFlags(&s->ev_up[i], cudaEventDisableTiming));
            WSB_CUDA(cudaEventCreateWithFlags(&s->ev_done[i], cudaEventDisableTiming));
        }
    }
    const Geometry<T> geo = g->geom<T>();
    const Physics<T> ph = sim_physics<T>(s);
    void *cur3[3] = {g->f[0].base, g->f[1].base, g->f[2].base};
    StepArgs<T> a = step_args<T>(s);
    // 16 slabs: the call ends one slab after the last upload ((1 + 1/n) x the transfer) and every slab costs about
    // 50-100 us of event / launch latency in the upload -> step -> download chain (profiles/r1/e2e_sweep*.txt)
    int want_slabs = 16;
    if (const char *e = std::getenv("WSB_HOST_SLABS")) want_slabs = std::atoi(e);
    // short row chunks: a slab's launch sits between its upload and its download, so its LATENCY counts, not its
    // efficiency (the whole step is 0.6 ms of an 18 ms transfer)
    a.rows_per_chunk = 16;
    if (const char *e = std::getenv("WSB_HOST_RPC")) a.rows_per_chunk = std::atoi(e);
    const int nslabs = std::max(1, std::min(std::min((int)wsb_sim::kMaxSlabs, want_slabs), H / 64));
    // WSB_HOST_TRACE=1: device timeline of the three streams of one call (tuning aid, prints to stderr)
    const bool trace = std::getenv("WSB_HOST_TRACE") != nullptr;
    cudaEvent_t tr[4] = {};
    if (trace) {
        for (auto &e : tr) WSB_CUDA(cudaEventCreate(&e));
    }
    const int rows_per_slab = (H + nslabs - 1) / nslabs;
    // everything already enqueued on the main stream (previous steps) precedes the first upload
    WSB_CUDA(cudaEventRecord(s->ev_edge, s->stream));
    WSB_CUDA(cudaStreamWaitEvent(s->h2d_stream, s->ev_edge, 0));
    WSB_CUDA(cudaStreamWaitEvent(s->d2h_stream, s->ev_edge, 0));
    if (trace) WSB_CUDA(cudaEventRecord(tr[0], s->h2d_stream));
    if (s->comm) {
        // row slabs: the neighbours need this rank's first and last `halo` input rows before their edge rows can
        // step, so those go up first and are exchanged (NCCL, comm stream) while the slabs stream behind them
        const int bands[2][2] = {{0, halo}, {H - halo, H}};
        for (const auto &b : bands)
            for (int l = 0; l < g->L; ++l)
                for (int k = 0; k < 3; ++k) {
                    char *dst = (char *)g->origin(k) + ((size_t)l * g->level_stride + (size_t)b[0] * g->pitch) * es;
                    const char *src = (const char *)in[k] + ((size_t)l * H + b[0]) * row_bytes;
                    WSB_CUDA(copy_rows(dst, pitch_bytes, src, row_bytes, row_bytes, (size_t)(b[1] - b[0]),
                                       cudaMemcpyHostToDevice, s->h2d_stream));
                }
        WSB_CUDA(cudaEventRecord(s->ev_interior, s->h2d_stream));
        WSB_TRY(sim_exchange(s, cur3, halo, s->ev_interior));
        WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));  // ghost rows precede the first launch
    }
    int done_rows = 0;
    for (int i = 0; i < nslabs; ++i) {
        const int r0 = i * rows_per_slab, r1 = std::min(H, r0 + rows_per_slab);
        if (r0 >= r1) break;
        for (int l = 0; l < g->L; ++l)
            for (int k = 0; k < 3; ++k) {
                char *dst = (char *)g->origin(k) + ((size_t)l * g->level_stride + (size_t)r0 * g->pitch) * es;
                const char *src = (const char *)in[k] + ((size_t)l * H + r0) * row_bytes;
                WSB_CUDA(copy_rows(dst, pitch_bytes, src, row_bytes, row_bytes, (size_t)(r1 - r0),
                                           cudaMemcpyHostToDevice, s->h2d_stream));
            }
        WSB_CUDA(cudaEventRecord(s->ev_up[i], s->h2d_stream));
        const int end = (r1 == H) ? H : r1 - halo;  // rows whose stencil inputs have all landed
        if (end <= done_rows) continue;
        WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_up[i], 0));
        a.y_begin = done_rows;
        a.y_end = end;
        WSB_CUDA(launch_step<T>(s, geo, ph, a, s->stream));
        s->metrics.kernel_launches += 1;
        WSB_CUDA(cudaEventRecord(s->ev_done[i], s->stream));
        WSB_CUDA(cudaStreamWaitEvent(s->d2h_stream, s->ev_done[i], 0));
        if (trace && done_rows == 0) WSB_CUDA(cudaEventRecord(tr[2], s->d2h_stream));
        for (int l = 0; l < g->L; ++l)
            for (int k = 0; k < 3; ++k) {
                const char *src = (const char *)plane_origin(s, s->next[k]) +
                                  ((size_t)l * g->level_stride + (size_t)done_rows * g->pitch) * es;
                char *dst = (char *)out[k] + ((size_t)l * H + done_rows) * row_bytes;
                WSB_CUDA(copy_rows(dst, row_bytes, src, pitch_bytes, row_bytes, (size_t)(end - done_rows),
                                           cudaMemcpyDeviceToHost, s->d2h_stream));
            }
        done_rows = end;
    }
    if (trace) {
        WSB_CUDA(cudaEventRecord(tr[1], s->h2d_stream));
        WSB_CUDA(cudaEventRecord(tr[3], s->d2h_stream));
        WSB_CUDA(cudaEventSynchronize(tr[3]));
        WSB_CUDA(cudaEventSynchronize(tr[1]));
        float up = 0, d0 = 0, d1 = 0;
        cudaEventElapsedTime(&up, tr[0], tr[1]);
        cudaEventElapsedTime(&d0, tr[0], tr[2]);
        cudaEventElapsedTime(&d1, tr[0], tr[3]);
        std::fprintf(stderr, "[wsb] step_host: %d slabs; uploads 0..%.3f ms, downloads %.3f..%.3f ms\n", nslabs, up, d0, d1);
        for (auto &e : tr) cudaEventDestroy(e);
    }
    // the main stream rejoins the copy streams: later work sees a complete state
    WSB_CUDA(cudaEventRecord(s->ev_halo, s->d2h_stream));
    WSB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
    return WSB_OK;
}

extern "C" {

int wsb_sim_step_host(wsb_sim *s, const void *u, const void *v, const void *h, void *out_u, void *out_v,
                      void *out_h) {
    if (!s || !u || !v || !h || !out_u || !out_v || !out_h) return fail(WSB_ERR_INVALID_ARGUMENT, "NULL argument");
    WSB_CUDA(cudaSetDevice(s->cur.device));
    wsb_grid *g = &s->cur;
    // with row slabs every rank must take the same branch (the two paths exchange ghost rows differently), so the
    // test uses the smallest slab of the balanced partition, which all ranks know
    const bool pipelined = is_step_path(s->path) && s->cfg.model != WSB_MODEL_PRIMITIVE_EQUATIONS &&
                           (!s->comm || s->cfg.grid_height / std::max(1, s->cfg.nranks) >= 2 * s->nstages);
    if (!pipelined) {  // same result, unpipelined: upload, step, download
        const void *in[3] = {u, v, h};
        void *out[3] = {out_u, out_v, out_h};
        for (int k = 0; k < 3; ++k) WSB_TRY(wsb_grid_set_field(g, k, in[k], s->dtype, g->L, g->H, g->W));
        WSB_TRY(wsb_sim_step(s));
        for (int k = 0; k < 3; ++k) WSB_TRY(wsb_grid_get_field(g, k, out[k], s->dtype, g->L, g->H, g->W));
        return WSB_OK;
    }
    const auto t0 = std::chrono::steady_clock::now();
    if (s->comm) {  // asynchronous earlier steps may still be exchanging ghost rows
        WSB_CUDA(cudaStreamSynchronize(s->edge_stream));
        WSB_CUDA(cudaStreamSynchronize(s->comm_stream));
    }
    WSB_TRY(sim_begin_timing(s));
    const void *in[3] = {u, v, h};
    void *out[3] = {out_u, out_v, out_h};
    if (s->dtype == WSB_F64) WSB_TRY(step_host_pipelined<double>(s, in, out));
    else WSB_TRY(step_host_pipelined<float>(s, in, out));
    // bookkeeping of one step (sim_enqueue_step without the launch)
    for (int k = 0; k < 3; ++k) std::swap(g->f[k].base, s->next[k]);
    for (int k = WSB_FIELD_PRESSURE; k <= WSB_FIELD_HUMIDITY; ++k) std::swap(g->f[k], s->alt[k]);
    s->diag_dirty = true;
    s->halo_valid = false;  // the new state's ghost rows have not been exchanged
    s->ghosts_in_flight = false;
    if (s->dtype == WSB_F32) s->time = (double)((float)s->time + (float)s->dt);
    else s->time += s->dt;
    s->step += 1;
    s->metrics.num_steps += 1;
    WSB_TRY(sim_sync(s));
    WSB_CUDA(cudaStreamSynchronize(s->h2d_stream));
    WSB_CUDA(cudaStreamSynchronize(s->d2h_stream));
    s->metrics.total_time_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return WSB_OK;
}

int wsb_sim_run(wsb_sim *s, int32_t num_steps, int32_t *steps_done) {
    if (!s) return fail(WSB_ERR_INVALID_ARGUMENT, "sim is NULL");
    if (steps_done) *steps_done = 0;
    if (num_steps <= 0) return WSB_OK;  // weather_simulation.cpp:69-71
    WSB_CUDA(cudaSetDevice(s->cur.device));
    const auto t0 = std::chrono::steady_clock::now();
    WSB_TRY(sim_begin_timing(s, num_steps > 1));
    int done = 0;
    for (int i = 0; i < num_steps; ++i) {
        WSB_TRY(sim_enqueue_step(s, i > 0));
        ++done;
        if (s->time >= s->cfg.max_time) break;  // :87-89, checked after the step
    }
    WSB_TRY(sim_sync(s));
    const auto t1 = std::chrono::steady_clock::now();
    s->metrics.total_time_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (steps_done) *steps_done = done;
    return WSB_OK;
}

int wsb_sim_run_until(wsb_sim *s, double max_time, int32_t *steps_done) {
    if (!s) return fail(WSB_ERR_INVALID_ARGUMENT, "sim is NULL");
    if (steps_done) *steps_done = 0;
    int n;
    if (s->dtype == WSB_F32) {  // weather_simulation.cpp:105-115 in float
        const float mt = (float)max_time, t = (float)s->time, dt = (float)s->dt;
        if (mt <= t) return WSB_OK;
        n = (int)((mt - t) / dt) + 1;
    } else {
        if (max_time <= s->time) return WSB_OK;
        n = (int)((max_time - s->time) / s->dt) + 1;
    }
    return wsb_sim_run(s, n, steps_done);
}

int wsb_sim_last_run_device_ms(wsb_sim *s, double *ms) {
    if (!s || !ms) return fail(WSB_ERR_INVALID_ARGUMENT, "NULL argument");
    *ms = s->last_run_ms;
    return WSB_OK;
}

double wsb_sim_get_time(const wsb_sim *s) { return s ? s->time : 0.0; }
int32_t wsb_sim_get_step(const wsb_sim *s) { return s ? s->step : 0; }
double wsb_sim_get_dt(const wsb_sim *s) { return s ? s->dt : 0.0; }

int wsb_sim_set_dt(wsb_sim *s, double dt) {
    if (!s) return fail(WSB_ERR_INVALID_ARGUMENT, "sim is NULL");
    s->dt = s->dtype == WSB_F32 ? (double)(float)dt : dt;
    return WSB_OK;
}

int wsb_sim_get_config(const wsb_sim *s, wsb_config *out) {
    if (!s || !out) return fail(WSB_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = s->cfg;
    return WSB_OK;
}

int wsb_sim_get_metrics(wsb_sim *s, wsb_metrics *out) {
    if (!s || !out) return fail(WSB_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = s->metrics;
    return WSB_OK;
}

int wsb_sim_reset_metrics(wsb_sim *s) {
    if (!s) return fail(WSB_ERR_INVALID_ARGUMENT, "sim is NULL");
    std::memset(&s->metrics, 0, sizeof(s->metrics));
    return WSB_OK;
}

int wsb_sim_local_rows(const wsb_sim *s, int32_t *row0, int32_t *nrows) {
    if (!s) return fail(WSB_ERR_INVALID_ARGUMENT, "sim is NULL");
    if (row0) *row0 = s->row0;
    if (nrows) *nrows = s->nrows;
    return WSB_OK;
}

const char *wsb_sim_kernel_name(const wsb_sim *s) {
    if (!s) return "";
    switch (s->path) {
        case PATH_STEP_REG: return "step_fused_reg";
        case PATH_STEP_TMA: return "step_fused_tma";
        default: return "stage_direct";
    }
}

int wsb_exact_division_reciprocal(double divisor, int32_t dtype, double *reciprocal) {
    if (!reciprocal || (dtype != WSB_F32 && dtype != WSB_F64)) return fail(WSB_ERR_INVALID_ARGUMENT, "bad argument");
    *reciprocal = proven_reciprocal(dtype == WSB_F32 ? (double)(float)divisor : divisor, dtype);
    return WSB_OK;
}

int wsb_sim_time_halo_exchange(wsb_sim *s, int32_t reps, double *us_per_exchange, int64_t *bytes_per_neighbour) {
    if (!s || !us_per_exchange) return fail(WSB_ERR_INVALID_ARGUMENT, "NULL argument");
    *us_per_exchange = 0.0;
    const int depth = is_step_path(s->path) ? s->nstages : 1;
    if (bytes_per_neighbour)
        *bytes_per_neighbour = (int64_t)depth * 3 * s->cur.L * s->cur.pitch * (int64_t)elem_size(s->dtype);
    if (!s->comm || reps <= 0) return WSB_OK;
    WSB_TRY(sim_sync(s));
    void *cur3[3] = {s->cur.f[0].base, s->cur.f[1].base, s->cur.f[2].base};
    void *origins[3] = {plane_origin(s, cur3[0]), plane_origin(s, cur3[1]), plane_origin(s, cur3[2])};
    const size_t es = elem_size(s->dtype);
    // the first exchanges absorb the skew between the ranks; the all-reduce pins the start of the timed ones
    for (int i = 0; i < 3; ++i)
        WSB_TRY(halo_exchange(s->comm, origins, 3, es, s->cur.pitch, s->cur.H, depth, s->cur.L, s->cur.level_stride,
                              s->comm_stream));
    WSB_TRY(halo_align(s->comm, s->comm_stream));
    WSB_CUDA(cudaEventRecord(s->ev_h0, s->comm_stream));
    for (int i = 0; i < reps; ++i)
        WSB_TRY(halo_exchange(s->comm, origins, 3, es, s->cur.pitch, s->cur.H, depth, s->cur.L, s->cur.level_stride,
                              s->comm_stream));
    WSB_CUDA(cudaEventRecord(s->ev_h1, s->comm_stream));
    WSB_CUDA(cudaStreamSynchronize(s->comm_stream));
    float ms = 0.f;
    WSB_CUDA(cudaEventElapsedTime(&ms, s->ev_h0, s->ev_h1));
    *us_per_exchange = 1.0e3 * ms / reps;
    s->halo_valid = false;  // (the next step re-establishes the ghost rows in its own stream order)
    s->ghosts_in_flight = false;
    return WSB_OK;
}

int wsb_sim_mass_energy(wsb_sim *s, double *mass, double *energy) {
    if (!s) return fail(WSB_ERR_INVALID_ARGUMENT, "sim is NULL");
    WSB_CUDA(cudaSetDevice(s->cur.device));
    const wsb_grid *g = &s->cur;
    if (s->dtype == WSB_F64) {
        WSB_CUDA(launch_mass_energy<double>(g->geom<double>(), (const double *)g->origin(0), (const double *)g->origin(1),
                                            (const double *)g->origin(2), s->cfg.gravity, s->d_partial, s->npartial,
                                            s->stream));
    } else {
        WSB_CUDA(launch_mass_energy<float>(g->geom<float>(), (const float *)g->origin(0), (const float *)g->origin(1),
                                           (const float *)g->origin(2), s->cfg.gravity, s->d_partial, s->npartial,
                                           s->stream));
    }
    std::vector<double> hp(2 * s->npartial);
    WSB_CUDA(cudaMemcpyAsync(hp.data(), s->d_partial, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    WSB_CUDA(cudaStreamSynchronize(s->stream));
    double m = 0.0, e = 0.0;
    for (int b = 0; b < s->npartial; ++b) {
        m += hp[2 * b];
        e += hp[2 * b + 1];
    }
    if (mass) *mass = m;
    if (energy) *energy = e;
    return WSB_OK;
}

}  // extern "C"
