// Initial conditions of the reference (initial_conditions.cpp:59-535), evaluated on the host in the
// reference's own mixed float/double arithmetic and uploaded to the device-resident grid. They run once per
// simulation and sit next to, not on, the time-stepping hot path.
//
// Quirk kept on purpose: the reference stores every constructor argument as std::to_string(value) and
// reads it back with std::stof (initial_conditions.hpp:74-76,95-117), i.e. parameters are rounded to six
// decimals. roundtrip() reproduces that so that results match bit-for-bit.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "wsb_internal.h"

namespace {

float roundtrip(double v) { return std::stof(std::to_string(static_cast<float>(v))); }

struct Params {
    const double *p;
    int n;
    // i-th constructor argument or its default, through the reference's to_string/stof round trip
    float get(int i, float dflt) const { return roundtrip(i < n ? p[i] : (double)dflt); }
};

struct HostFields {
    int W, H;          // full grid extent (the formulas normalise by it)
    int ys, ye;        // rows [ys, ye) are evaluated; the buffers hold exactly those rows
    float *u, *v, *h, *p, *t, *q;  // any may be null: that field is left untouched
    void set_uv(int x, int y, float uu, float vv) const {
        const size_t i = (size_t)(y - ys) * W + x;
        if (u) u[i] = uu;
        if (v) v[i] = vv;
    }
    void set(float *f, int x, int y, float val) const {
        if (f) f[(size_t)(y - ys) * W + x] = val;
    }
};

// bit mask of the fields an initial condition writes (the others keep their current values)
enum { F_U = 1, F_V = 2, F_H = 4, F_P = 8, F_T = 16, F_Q = 32 };

int ic_uniform(const Params &a, const HostFields &g) {  // initial_conditions.cpp:59-90
    const float u = a.get(0, 0.0f), v = a.get(1, 0.0f), h = a.get(2, 10.0f), p = a.get(3, 1000.0f),
                t = a.get(4, 300.0f), q = a.get(5, 0.0f);
    for (int y = g.ys; y < g.ye; ++y)
        for (int x = 0; x < g.W; ++x) {
            g.set_uv(x, y, u, v);
            g.set(g.h, x, y, h);
            g.set(g.p, x, y, p);
            g.set(g.t, x, y, t);
            g.set(g.q, x, y, q);
        }
    return F_U | F_V | F_H | F_P | F_T | F_Q;
}

int ic_random(const Params &a, std::mt19937 &rng, const HostFields &g) {  // :98-125
    // the reference seeds std::mt19937 with the int seed and draws u, v, h per cell in row-major order; the
    // generator is carried across row blocks so that blockwise evaluation draws the same sequence
    const float amplitude = a.get(0, 1.0f);
    std::uniform_real_distribution<float> dist(-amplitude, amplitude);
    for (int y = g.ys; y < g.ye; ++y)
        for (int x = 0; x < g.W; ++x) {
            const float u = dist(rng);
            const float v = dist(rng);
            const float h = 10.0f + dist(rng);
            g.set_uv(x, y, u, v);
            g.set(g.h, x, y, h);
        }
    return F_U | F_V | F_H;
}

int ic_zonal_flow(const Params &a, const HostFields &g) {  // :136-177
    const float u_max = a.get(0, 10.0f), h_mean = a.get(1, 10.0f), beta = a.get(2, 0.1f);
    for (int y = g.ys; y < g.ye; ++y) {
        const float y_norm = static_cast<float>(y) / (g.H - 1);
        const float u = u_max * std::sin(M_PI * y_norm);  // double product, rounded once on assignment
        for (int x = 0; x < g.W; ++x) {
            g.set_uv(x, y, u, 0.0f);
            const float f = 1.0e-4f + beta * (y_norm - 0.5f);
            const float h = h_mean - 0.5f * f * u * u / 9.81f;
            g.set(g.h, x, y, h);
        }
    }
    return F_U | F_V | F_H;
}

int ic_vortex(const Params &a, const HostFields &g) {  // :190-240
    const float x_center = a.get(0, 0.5f), y_center = a.get(1, 0.5f), radius = a.get(2, 0.1f),
                strength = a.get(3, 10.0f), h_mean = a.get(4, 10.0f);
    const float xc = x_center * (g.W - 1), yc = y_center * (g.H - 1);
    const float rg = radius * std::min(g.W, g.H);
    for (int y = g.ys; y < g.ye; ++y)
        for (int x = 0; x < g.W; ++x) {
            const float dx = x - xc, dy = y - yc;
            const float r = std::sqrt(dx * dx + dy * dy);
            float av = 0.0f, h = h_mean;
            if (r > 0.0f && r <= rg) {
                const float rn = r / rg;
                av = strength * rn * std::exp(1.0f - rn * rn);
                h = h_mean - 0.5f * av * av / 9.81f;
            }
            const float u = -av * dy / std::max(r, 1.0e-6f);
            const float v = av * dx / std::max(r, 1.0e-6f);
            g.set_uv(x, y, u, v);
            g.set(g.h, x, y, h);
        }
    return F_U | F_V | F_H;
}

int ic_jet_stream(const Params &a, const HostFields &g) {  // :252-290
    const float y_center = a.get(0, 0.5f), width = a.get(1, 0.1f), strength = a.get(2, 10.0f),
                h_mean = a.get(3, 10.0f);
    const float yc = y_center * (g.H - 1);
    const float wg = width * g.H;
    for (int y = g.ys; y < g.ye; ++y) {
        const float dy = y - yc;
        const float u = strength * std::exp(-(dy * dy) / (2.0f * wg * wg));
        const float dh_dy = -1.0e-4f * u / 9.81f;
        for (int x = 0; x < g.W; ++x) {
            g.set_uv(x, y, u, 0.0f);
            g.set(g.h, x, y, h_mean + dh_dy * dy);
        }
    }
    return F_U | F_V | F_H;
}

int ic_breaking_wave(const Params &a, const HostFields &g) {  // :301-343
    const float amplitude = a.get(0, 1.0f), wavelength = a.get(1, 0.2f), h_mean = a.get(2, 10.0f);
    const float wave_k = 2.0f * M_PI / (wavelength * g.W);
    for (int y = g.ys; y < g.ye; ++y) {
        const float y_norm = static_cast<float>(y) / (g.H - 1);
        const float u_base = 5.0f * std::sin(M_PI * y_norm);
        for (int x = 0; x < g.W; ++x) {
            const float wave_phase = wave_k * x - 0.1f * y_norm;
            // std::pow(float, int) and the exp around it are evaluated in double
            const float wave_amp = amplitude * std::exp(-std::pow(y_norm - 0.5f, 2) / 0.05f);
            const float u = u_base + wave_amp * std::sin(wave_phase);
            const float v = wave_amp * std::cos(wave_phase);
            const float h = h_mean + wave_amp * std::cos(wave_phase);
            g.set_uv(x, y, u, v);
            g.set(g.h, x, y, h);
        }
    }
    return F_U | F_V | F_H;
}

int ic_front(const Params &a, const HostFields &g) {  // :356-395
    const float y_position = a.get(0, 0.5f), width = a.get(1, 0.05f), temp_difference = a.get(2, 10.0f),
                wind_shear = a.get(3, 5.0f);
    const float yp = y_position * (g.H - 1);
    const float wg = width * g.H;
    for (int y = g.ys; y < g.ye; ++y) {
        const float dy = y - yp;
        const float tt = std::tanh(dy / wg);
        const float temperature = 288.15f + 0.5f * temp_difference * tt;
        const float u = 0.5f * wind_shear * tt;
        for (int x = 0; x < g.W; ++x) {
            g.set_uv(x, y, u, 0.0f);
            g.set(g.t, x, y, temperature);
            g.set(g.p, x, y, 1013.25f - 0.1f * temp_difference * tt);
        }
    }
    return F_U | F_V | F_T | F_P;
}

int ic_mountain(const Params &a, const HostFields &g) {  // :408-467
    const float x_center = a.get(0, 0.3f), y_center = a.get(1, 0.5f), radius = a.get(2, 0.1f),
                mountain_height = a.get(3, 1.0f), u_base = a.get(4, 5.0f);
    const float xc = x_center * (g.W - 1), yc = y_center * (g.H - 1);
    const float rg = radius * std::min(g.W, g.H);
    for (int y = g.ys; y < g.ye; ++y)
        for (int x = 0; x < g.W; ++x) {
            const float dx = x - xc, dy = y - yc;
            const float r = std::sqrt(dx * dx + dy * dy);
            float profile = 0.0f;
            if (r <= 2.0f * rg) profile = mountain_height * std::exp(-(r * r) / (rg * rg));
            const float h = 10.0f + profile;
            float u = u_base, v = 0.0f;
            if (r <= 3.0f * rg) {
                const float flow_reduction = 0.7f * profile / mountain_height;
                u *= (1.0f - flow_reduction);
                if (r > 0.0f) v = -0.5f * flow_reduction * u_base * dy / r;
            }
            g.set_uv(x, y, u, v);
            g.set(g.h, x, y, h);
        }
    return F_U | F_V | F_H;
}

struct Profile {
    float p[10], t[10], q[10], u[10], v[10];
};

// initial_conditions.cpp:540-609
const Profile kStandard = {{1013.0f, 1011.0f, 1009.0f, 1005.0f, 1000.0f, 995.0f, 990.0f, 985.0f, 980.0f, 975.0f},
                           {298.0f, 295.0f, 292.0f, 288.0f, 285.0f, 282.0f, 278.0f, 275.0f, 272.0f, 268.0f},
                           {0.8f, 0.75f, 0.7f, 0.65f, 0.6f, 0.55f, 0.5f, 0.45f, 0.4f, 0.35f},
                           {2.0f, 4.0f, 6.0f, 8.0f, 10.0f, 12.0f, 10.0f, 8.0f, 6.0f, 4.0f},
                           {0.0f, 1.0f, 2.0f, 1.0f, 0.0f, -1.0f, -2.0f, -1.0f, 0.0f, 1.0f}};
const Profile kTropical = {{1010.0f, 1009.0f, 1008.0f, 1007.0f, 1006.0f, 1005.0f, 1004.0f, 1003.0f, 1002.0f, 1001.0f},
                           {303.0f, 302.0f, 301.0f, 300.0f, 299.0f, 298.0f, 297.0f, 296.0f, 295.0f, 294.0f},
                           {0.9f, 0.89f, 0.88f, 0.87f, 0.86f, 0.85f, 0.84f, 0.83f, 0.82f, 0.81f},
                           {-5.0f, -6.0f, -7.0f, -8.0f, -7.0f, -6.0f, -5.0f, -4.0f, -3.0f, -2.0f},
                           {-1.0f, -0.5f, 0.0f, 0.5f, 1.0f, 1.0f, 0.5f, 0.0f, -0.5f, -1.0f}};
const Profile kPolar = {{1020.0f, 1018.0f, 1016.0f, 1014.0f, 1012.0f, 1010.0f, 1008.0f, 1006.0f, 1004.0f, 1002.0f},
                        {260.0f, 258.0f, 256.0f, 254.0f, 252.0f, 250.0f, 248.0f, 246.0f, 244.0f, 242.0f},
                        {0.3f, 0.29f, 0.28f, 0.27f, 0.26f, 0.25f, 0.24f, 0.23f, 0.22f, 0.21f},
                        {10.0f, 12.0f, 14.0f, 16.0f, 18.0f, 20.0f, 18.0f, 16.0f, 14.0f, 12.0f},
                        {0.0f, -1.0f, -2.0f, -3.0f, -4.0f, -3.0f, -2.0f, -1.0f, 0.0f, 1.0f}};

const Profile &profile_by_name(const char *profile) {
    const std::string name = profile ? profile : "standard";
    return name == "tropical" ? kTropical : name == "polar" ? kPolar : kStandard;
}
// the table row a grid row reads, and the three x-dependent perturbations (initial_conditions.cpp:499-528); shared by
// the cell-by-cell host evaluation below and the separable device fill (apply_separable)
size_t profile_index(int y, int H) {
    const size_t n = 10;
    const float y_norm = static_cast<float>(y) / (H - 1);
    return std::min(static_cast<size_t>(y_norm * (n - 1)), n - 1);
}
void profile_column(int x, int W, float &t_var, float &p_var, float &q_var) {
    const float x_norm = static_cast<float>(x) / (W - 1);
    t_var = 2.0f * std::sin(2.0f * M_PI * x_norm);  // double product, rounded once on assignment
    p_var = 2.0f * std::cos(2.0f * M_PI * x_norm);
    q_var = 0.02f * std::sin(4.0f * M_PI * x_norm);
}

int ic_atmospheric_profile(const char *profile, const HostFields &g) {  // :474-537
    const Profile &pr = profile_by_name(profile);
    for (int y = g.ys; y < g.ye; ++y) {
        const size_t idx = profile_index(y, g.H);
        const float t_base = pr.t[idx], p_base = pr.p[idx], q_base = pr.q[idx], u_base = pr.u[idx], v_base = pr.v[idx];
        for (int x = 0; x < g.W; ++x) {
            float t_var, p_var, q_var;
            profile_column(x, g.W, t_var, p_var, q_var);
            g.set(g.t, x, y, t_base + t_var);
            g.set(g.p, x, y, p_base + p_var);
            g.set(g.q, x, y, q_base + q_var);
            g.set_uv(x, y, u_base, v_base);
        }
    }
    return F_U | F_V | F_T | F_P | F_Q;
}

// returns the written-field mask, or -1 for an unknown name
int ic_dispatch(const char *name, const Params &a, std::mt19937 &rng, const char *profile, const HostFields &g) {
    const std::string n = name ? name : "";
    if (n == "uniform") return ic_uniform(a, g);
    if (n == "random") return ic_random(a, rng, g);
    if (n == "zonal_flow") return ic_zonal_flow(a, g);
    if (n == "vortex") return ic_vortex(a, g);
    if (n == "jet_stream") return ic_jet_stream(a, g);
    if (n == "breaking_wave") return ic_breaking_wave(a, g);
    if (n == "front") return ic_front(a, g);
    if (n == "mountain") return ic_mountain(a, g);
    if (n == "atmospheric_profile") return ic_atmospheric_profile(profile, g);
    if (n == "standard_atmosphere") return ic_atmospheric_profile("standard", g);  // factory names, :653-665
    if (n == "tropical_atmosphere") return ic_atmospheric_profile("tropical", g);
    if (n == "polar_atmosphere") return ic_atmospheric_profile("polar", g);
    return -1;
}

// Rows [g.ys, g.ye) split over host threads. Every cell is a pure function of (x, y) -- and, for "random", of the
// generator position, which each thread reaches with discard() -- so the result does not depend on the split.
int ic_dispatch_parallel(const char *name, const Params &a, std::mt19937 &rng, const char *profile,
                         const HostFields &g) {
    const int rows = g.ye - g.ys;
    const long long cells = (long long)rows * g.W;
    int nthreads = (int)std::min<long long>(std::max(1u, std::min(32u, std::thread::hardware_concurrency())),
                                            cells / (1 << 16));
    if (const char *e = std::getenv("WSB_IC_THREADS")) nthreads = std::atoi(e);
    nthreads = std::max(1, std::min(nthreads, rows));
    const bool is_random = name && std::strcmp(name, "random") == 0;
    int mask = 0;
    if (nthreads == 1) {
        mask = ic_dispatch(name, a, rng, profile, g);
        return mask;
    }
    std::vector<int> masks(nthreads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) {
        const int y0 = g.ys + (int)((long long)rows * t / nthreads), y1 = g.ys + (int)((long long)rows * (t + 1) / nthreads);
        pool.emplace_back([&, t, y0, y1] {
            const size_t off = (size_t)(y0 - g.ys) * g.W;
            auto at = [off](float *f) { return f ? f + off : f; };
            const HostFields sub{g.W, g.H, y0, y1, at(g.u), at(g.v), at(g.h), at(g.p), at(g.t), at(g.q)};
            std::mt19937 local = rng;
            if (is_random) local.discard(3ULL * off);
            masks[t] = ic_dispatch(name, a, local, profile, sub);
        });
    }
    for (auto &th : pool) th.join();
    if (is_random) rng.discard(3ULL * (unsigned long long)cells);
    for (int m : masks) mask = (m < 0 || mask < 0) ? -1 : (mask | m);
    return mask;
}

// ---- device-side path ------------------------------------------------------------------
// Six of the nine initial conditions are separable: a constant (uniform), a function of the row alone (zonal_flow,
// jet_stream, front -- their one libm call sits outside the x loop of the reference too), or table(y) + f(x) with one
// fp32 addition per cell (the three atmospheric profiles). For those the host evaluates O(W + H) values with the
// reference's own expressions (libm is part of the reference's results) and a kernel expands them in HBM: a 32768^2
// grid never exists on the host and costs one sweep per written field at memory bandwidth. Returns 1 if handled.
int apply_separable(wsb_grid *grid, const std::string &n, const Params &a, const char *profile, int W, int H, int row0,
                    int Hglobal) {
    const int fields[6] = {WSB_FIELD_U, WSB_FIELD_V, WSB_FIELD_HEIGHT, WSB_FIELD_PRESSURE, WSB_FIELD_TEMPERATURE,
                           WSB_FIELD_HUMIDITY};
    if (n == "uniform") {
        const float val[6] = {a.get(0, 0.0f), a.get(1, 0.0f), a.get(2, 10.0f), a.get(3, 1000.0f), a.get(4, 300.0f),
                              a.get(5, 0.0f)};
        for (int k = 0; k < 6; ++k)
            if (wsb::grid_fill_uniform(grid, fields[k], val[k]) != WSB_OK) return -1;
        return 1;
    }
    const bool rows_only = n == "zonal_flow" || n == "jet_stream" || n == "front";
    const bool prof = n == "atmospheric_profile" || n == "standard_atmosphere" || n == "tropical_atmosphere" ||
                      n == "polar_atmosphere";
    if (!rows_only && !prof) return 0;
    std::vector<float> row[6];
    for (auto &r : row) r.resize((size_t)H);
    if (rows_only) {
        // the reference's own row loop on a one-column grid: none of the three reads the width
        const HostFields g{1, Hglobal, row0, row0 + H, row[0].data(), row[1].data(), row[2].data(), row[3].data(),
                           row[4].data(), row[5].data()};
        std::mt19937 unused;
        const int mask = ic_dispatch(n.c_str(), a, unused, profile, g);
        if (mask < 0) return -1;
        for (int k = 0; k < 6; ++k)
            if ((mask & (1 << k)) && wsb::grid_fill_separable(grid, fields[k], row[k].data(), nullptr) != WSB_OK) return -1;
        return 1;
    }
    const Profile &pr = profile_by_name(n == "atmospheric_profile" ? profile
                                        : n == "tropical_atmosphere" ? "tropical"
                                        : n == "polar_atmosphere"    ? "polar"
                                                                     : "standard");
    std::vector<float> col[3];
    for (auto &c : col) c.resize((size_t)W);
    for (int x = 0; x < W; ++x) profile_column(x, W, col[0][x], col[1][x], col[2][x]);  // t, p, q perturbations
    for (int y = 0; y < H; ++y) {
        const size_t idx = profile_index(row0 + y, Hglobal);
        row[0][y] = pr.u[idx]; row[1][y] = pr.v[idx]; row[3][y] = pr.p[idx]; row[4][y] = pr.t[idx]; row[5][y] = pr.q[idx];
    }
    if (wsb::grid_fill_separable(grid, WSB_FIELD_U, row[0].data(), nullptr) != WSB_OK ||
        wsb::grid_fill_separable(grid, WSB_FIELD_V, row[1].data(), nullptr) != WSB_OK ||
        wsb::grid_fill_separable(grid, WSB_FIELD_TEMPERATURE, row[4].data(), col[0].data()) != WSB_OK ||
        wsb::grid_fill_separable(grid, WSB_FIELD_PRESSURE, row[3].data(), col[1].data()) != WSB_OK ||
        wsb::grid_fill_separable(grid, WSB_FIELD_HUMIDITY, row[5].data(), col[2].data()) != WSB_OK)
        return -1;
    return 1;
}

}  // namespace

extern "C" {

int wsb_ic_fill_host(const char *name, const double *params, int32_t nparams, uint32_t seed, const char *profile,
                     int32_t width, int32_t height, double dx, double dy, float *u, float *v, float *h, float *p,
                     float *t, float *q) {
    (void)dx;
    (void)dy;  // no reference initial condition reads the spacing
    if (width <= 0 || height <= 0) return wsb::fail(WSB_ERR_INVALID_ARGUMENT, "Grid dimensions must be positive");
    if (nparams < 0 || (nparams > 0 && !params)) return wsb::fail(WSB_ERR_INVALID_ARGUMENT, "bad parameter list");
    const Params a{params, nparams};
    const HostFields g{width, height, 0, height, u, v, h, p, t, q};
    std::mt19937 rng(static_cast<int>(seed));
    if (ic_dispatch_parallel(name, a, rng, profile, g) < 0)
        return wsb::fail(WSB_ERR_INVALID_ARGUMENT, std::string("unknown initial condition '") + (name ? name : "") + "'");
    return WSB_OK;
}

int wsb_ic_apply(wsb_grid *grid, const char *name, const double *params, int32_t nparams, uint32_t seed,
                 const char *profile) {
    if (!grid) return wsb::fail(WSB_ERR_INVALID_ARGUMENT, "grid is NULL");
    if (nparams < 0 || (nparams > 0 && !params)) return wsb::fail(WSB_ERR_INVALID_ARGUMENT, "bad parameter list");
    wsb_grid_info gi;
    WSB_TRY(wsb_grid_get_info(grid, &gi));
    const Params a{params, nparams};
    // a row slab of a decomposed simulation evaluates ITS rows of the global initial condition
    int row0 = 0, Hglobal = gi.height;
    wsb::grid_slab_position(grid, &row0, &Hglobal);
    std::mt19937 rng(static_cast<int>(seed));
    rng.discard(3ULL * (unsigned long long)row0 * (unsigned long long)gi.width);  // "random": 3 draws per cell
    // separable initial conditions are expanded on the device from O(W + H) host values
    if (!std::getenv("WSB_IC_HOST")) {
        const int handled = apply_separable(grid, name ? name : "", a, profile, gi.width, gi.height, row0, Hglobal);
        if (handled < 0) return WSB_ERR_CUDA;  // the message is the failing call's
        if (handled > 0) return wsb_grid_calculate_diagnostics(grid);
    }
    // The others (vortex, mountain, breaking_wave: libm per cell; random: the mt19937 stream) are evaluated by host
    // threads in row blocks -- host memory stays bounded (a 32768^2 grid is 4 GiB per field) and fields an initial
    // condition does not write are neither allocated nor touched on the device -- through two page-locked buffer sets:
    // block k uploads while block k+1 is evaluated.
    const int block = std::max(1, std::min(gi.height, (int)(((size_t)32 << 20) / sizeof(float) / (size_t)gi.width)));
    const int fields[6] = {WSB_FIELD_U, WSB_FIELD_V, WSB_FIELD_HEIGHT, WSB_FIELD_PRESSURE, WSB_FIELD_TEMPERATURE,
                           WSB_FIELD_HUMIDITY};
    const size_t block_floats = (size_t)block * gi.width;
    const int nsets = gi.height > block ? 2 : 1;
    float *pin = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr};
    if (cudaHostAlloc((void **)&pin, sizeof(float) * block_floats * 6 * nsets, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return wsb::fail(WSB_ERR_OUT_OF_MEMORY, "initial condition: cannot allocate the page-locked staging buffers");
    }
    for (int k = 0; k < nsets; ++k) cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming);
    int status = WSB_OK, nblk = 0;
    for (int ys = 0; ys < gi.height && status == WSB_OK; ys += block, ++nblk) {
        const int ye = std::min(gi.height, ys + block), set = nblk % nsets;
        float *b = pin + (size_t)set * 6 * block_floats;
        if (nblk >= nsets) cudaEventSynchronize(done[set]);  // the upload that last read this buffer set
        const HostFields g{gi.width, Hglobal, row0 + ys, row0 + ye, b, b + block_floats, b + 2 * block_floats,
                           b + 3 * block_floats, b + 4 * block_floats, b + 5 * block_floats};
        const int mask = ic_dispatch_parallel(name, a, rng, profile, g);
        if (mask < 0) {
            status = wsb::fail(WSB_ERR_INVALID_ARGUMENT, std::string("unknown initial condition '") + (name ? name : "") + "'");
            break;
        }
        for (int k = 0; k < 6 && status == WSB_OK; ++k)
            if (mask & (1 << k))  // every level receives the same 2-D initial condition (the reference has one level)
                status = wsb::grid_upload_rows_async(grid, fields[k], b + (size_t)k * block_floats, ys, ye - ys);
        if (status == WSB_OK) status = wsb::grid_record_event(grid, done[set]);
    }
    for (int k = 0; k < nsets; ++k) {
        if (done[k]) {
            cudaEventSynchronize(done[k]);
            cudaEventDestroy(done[k]);
        }
    }
    cudaFreeHost(pin);
    if (status != WSB_OK) return status;
    // every reference initial condition ends with grid.calculateDiagnostics()
    return wsb_grid_calculate_diagnostics(grid);
}

}  // extern "C"
