// Per-stage kernels and small utility kernels of libweather_b200.so (sm_100a).
//
// stage_direct_kernel: one RK stage fused into one pass -- tendency (the 5-point clamped stencil of
// weather_simulation.cpp:473-540) + stage update (:186-198 / :248-260 / :380-451) -- one cell per
// thread, coalesced direct global loads; vertical/horizontal neighbour re-reads are served by L1/L2.
// It is the simplest correct path and the A/B baseline for the TMA-tiled and whole-step fused variants.
#include "wsb_arith.cuh"
#include "wsb_internal.h"

namespace wsb {

namespace {

constexpr int kBX = 64;
constexpr int kBY = 4;

template <typename T, bool RECIP, bool FINAL, bool STORE_K>
__global__ void __launch_bounds__(kBX *kBY)
    stage_direct_kernel(const Geometry<T> g, const Physics<T> ph, const StageArgs<T> a) {
    const int x = blockIdx.x * kBX + threadIdx.x;
    const int y = a.y_begin + blockIdx.y * kBY + threadIdx.y;
    if (x >= g.W || y >= a.y_end) return;
    const long long base = (long long)blockIdx.z * g.level_stride;
    const long long i = base + (long long)y * g.pitch + x;
    const int gy = g.row0 + y;
    // clamp-to-self neighbours (weather_simulation.cpp:510-513); slab edges read ghost rows
    const long long iL = (x > 0) ? i - 1 : i;
    const long long iR = (x < g.W - 1) ? i + 1 : i;
    const long long iU = (gy > 0) ? i - g.pitch : i;
    const long long iD = (gy < g.Hglobal - 1) ? i + g.pitch : i;

    const T u = a.S.u[i], v = a.S.v[i], h = a.S.h[i];
    T du, dv, dh;
    if (ph.ext)  // extended physics (uniform): beta plane + viscosity / diffusivity on the same clamped neighbours
        tendency_cell_ext<T, RECIP>(ph, ext_coriolis<T>(ph, gy), u, v, h, a.S.u[iL], a.S.u[iR], a.S.u[iU], a.S.u[iD],
                                    a.S.v[iL], a.S.v[iR], a.S.v[iU], a.S.v[iD], a.S.h[iL], a.S.h[iR], a.S.h[iU],
                                    a.S.h[iD], du, dv, dh);
    else
    tendency_cell<T, RECIP>(ph, u, v, h, a.S.u[iL], a.S.u[iR], a.S.u[iU], a.S.u[iD], a.S.v[iL], a.S.v[iR], a.S.v[iU],
                            a.S.v[iD], a.S.h[iL], a.S.h[iR], a.S.h[iU], a.S.h[iD], du, dv, dh);
    if (STORE_K) {
        a.KS.u[i] = du;
        a.KS.v[i] = dv;
        a.KS.h[i] = dh;
    }
    const T yu = a.Y.u[i], yv = a.Y.v[i], yh = a.Y.h[i];
    if (FINAL) {
        const T k1u = a.K1.u ? a.K1.u[i] : du;
        const T k1v = a.K1.v ? a.K1.v[i] : dv;
        const T k1h = a.K1.h ? a.K1.h[i] : dh;
        a.O.u[i] = rk4_combine<T>(yu, a.dt6, k1u, a.KA.u[i], a.KB.u[i], du);
        a.O.v[i] = rk4_combine<T>(yv, a.dt6, k1v, a.KA.v[i], a.KB.v[i], dv);
        a.O.h[i] = rk4_combine<T>(yh, a.dt6, k1h, a.KA.h[i], a.KB.h[i], dh);
    } else {
        a.O.u[i] = axpy<T>(yu, a.c, du);
        a.O.v[i] = axpy<T>(yv, a.c, dv);
        a.O.h[i] = axpy<T>(yh, a.c, dh);
    }
}

// Tracer transport of the extended Primitive model: one cell per thread, the three tracers share the cell's u, v and
// neighbour indices. Operation order: oracle/ws_oracle_body.inc (tracer_tendencies) -- (-u*c_x - v*c_y) + kappa*lap.
template <typename T, bool RECIP>
__global__ void __launch_bounds__(kBX *kBY)
    tracer_stage_kernel(const Geometry<T> g, const Physics<T> ph, const TracerArgs<T> a) {
    using A = Ar<T>;
    const int x = blockIdx.x * kBX + threadIdx.x;
    const int y = blockIdx.y * kBY + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const long long i = (long long)blockIdx.z * g.level_stride + (long long)y * g.pitch + x;
    const int gy = g.row0 + y;
    const long long iL = (x > 0) ? i - 1 : i, iR = (x < g.W - 1) ? i + 1 : i;
    const long long iU = (gy > 0) ? i - g.pitch : i, iD = (gy < g.Hglobal - 1) ? i + g.pitch : i;
    const T u = a.u[i], v = a.v[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const T *c = a.C[k];
        const T cc = c[i], cL = c[iL], cR = c[iR], cU = c[iU], cD = c[iD];
        const T cx = cdiff<T, RECIP>(cR, cL, ph.ddx, ph.rdx);
        const T cy = cdiff<T, RECIP>(cD, cU, ph.ddy, ph.rdy);
        const T adv = A::sub(A::mul(-u, cx), A::mul(v, cy));
        const T dc = A::add(adv, A::mul(ph.kappa, ext_laplacian<T>(ph, cc, cL, cR, cU, cD)));
        a.O[k][i] = axpy<T>(a.Y[k][i], a.c, dc);
    }
}

// Vorticity and divergence (weather_grid.cpp:96-100 and :114-118) as a row sweep: a thread owns 16 bytes of a row
// (4 floats / 2 doubles), walks down kDiagRows rows with u and v in 3-row register windows (every element is read
// once per chunk: 16 B/cell fp32 of traffic instead of eight scattered loads), takes horizontal neighbours from
// its own vector or the warp shuffle network (lane 0 / 31: one scalar load), prefetches one row ahead.
constexpr int kDiagRows = 32;
constexpr int kDiagThreads = 128;

template <typename T, bool RECIP>
__global__ void __launch_bounds__(kDiagThreads)
    diagnostics_kernel(const Geometry<T> g, const Physics<T> ph, const T *__restrict__ u, const T *__restrict__ v,
                       T *__restrict__ vort, T *__restrict__ dvg) {
    using A = Ar<T>;
    constexpr int V = 16 / (int)sizeof(T);
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int x0 = (blockIdx.x * kDiagThreads + threadIdx.x) * V;
    const int xl = min(x0, g.pitch - V);  // trailing lanes stay inside the row (their values are never stored)
    const int y0 = blockIdx.y * kDiagRows, y1 = min(y0 + kDiagRows, g.H);
    if ((blockIdx.x * kDiagThreads + (threadIdx.x & ~31)) * V >= g.W) return;  // whole warp past the right edge
    const long long lvl = (long long)blockIdx.z * g.level_stride;
    const T *ub = u + lvl, *vb = v + lvl;

    struct Vec {
        T e[V];
    };
    auto load = [&](const T *base, int y) {  // row y of this rank (ghost rows are addressable), 128-bit access
        alignas(16) Vec r;
        *reinterpret_cast<int4 *>(r.e) = *reinterpret_cast<const int4 *>(base + (long long)y * g.pitch + xl);
        return r;
    };
    // clamp-to-self at the GLOBAL top and bottom (interior slab edges read the neighbour's ghost row)
    auto up_row = [&](int y) { return (g.row0 + y > 0) ? y - 1 : y; };
    auto down_row = [&](int y) { return (g.row0 + y < g.Hglobal - 1) ? y + 1 : y; };

    Vec uU = load(ub, up_row(y0)), vU = load(vb, up_row(y0));
    Vec uC = load(ub, y0), vC = load(vb, y0);
    Vec uD = load(ub, down_row(y0)), vD = load(vb, down_row(y0));
    for (int y = y0; y < y1; ++y) {
        // prefetch the "down" row of the next iteration
        const int yn = min(y + 1, y1 - 1);
        const Vec uN = load(ub, down_row(yn)), vN = load(vb, down_row(yn));
        // horizontal neighbours across threads
        T uLn = __shfl_up_sync(kFull, uC.e[V - 1], 1), vLn = __shfl_up_sync(kFull, vC.e[V - 1], 1);
        T uRn = __shfl_down_sync(kFull, uC.e[0], 1), vRn = __shfl_down_sync(kFull, vC.e[0], 1);
        const long long row = (long long)y * g.pitch;
        if (lane == 0 && x0 > 0 && x0 < g.W) {
            uLn = ub[row + x0 - 1];
            vLn = vb[row + x0 - 1];
        }
        if (lane == 31 && x0 + V < g.W) {
            uRn = ub[row + x0 + V];
            vRn = vb[row + x0 + V];
        }
        alignas(16) Vec zo, dv;
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int x = x0 + e;
            // weather_grid.cpp:93-94 / 111-112: l = max(x-1, 0), r = min(x+1, W-1)
            const T uL = (x == 0) ? uC.e[e] : (e == 0 ? uLn : uC.e[e > 0 ? e - 1 : 0]);
            const T vL = (x == 0) ? vC.e[e] : (e == 0 ? vLn : vC.e[e > 0 ? e - 1 : 0]);
            const T uR = (x >= g.W - 1) ? uC.e[e] : (e == V - 1 ? uRn : uC.e[e < V - 1 ? e + 1 : e]);
            const T vR = (x >= g.W - 1) ? vC.e[e] : (e == V - 1 ? vRn : vC.e[e < V - 1 ? e + 1 : e]);
            const T dv_dx = cdiff<T, RECIP>(vR, vL, ph.ddx, ph.rdx);
            const T du_dy = cdiff<T, RECIP>(uD.e[e], uU.e[e], ph.ddy, ph.rdy);
            const T du_dx = cdiff<T, RECIP>(uR, uL, ph.ddx, ph.rdx);
            const T dv_dy = cdiff<T, RECIP>(vD.e[e], vU.e[e], ph.ddy, ph.rdy);
            zo.e[e] = A::sub(dv_dx, du_dy);
            dv.e[e] = A::add(du_dx, dv_dy);
        }
        const long long o = lvl + row + x0;
        if (x0 + V <= g.W) {
            *reinterpret_cast<int4 *>(vort + o) = *reinterpret_cast<const int4 *>(zo.e);
            *reinterpret_cast<int4 *>(dvg + o) = *reinterpret_cast<const int4 *>(dv.e);
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e)
                if (x0 + e < g.W) {
                    vort[o + e] = zo.e[e];
                    dvg[o + e] = dv.e[e];
                }
        }
        // next row: at the last row of the domain "down" stays the row itself, which uN/vN already hold
        uU = uC; vU = vC;
        uC = uD; vC = vD;
        uD = uN; vD = vN;
    }
}

template <typename T>
__global__ void __launch_bounds__(kBX *kBY)
    axpy_const_kernel(const Geometry<T> g, const T *__restrict__ yv, T *__restrict__ o, T c, T k) {
    const int x = blockIdx.x * kBX + threadIdx.x;
    const int y = blockIdx.y * kBY + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const long long i = (long long)blockIdx.z * g.level_stride + (long long)y * g.pitch + x;
    o[i] = axpy<T>(yv[i], c, k);
}

// Both Primitive-equations drifts in one streaming pass: T' = T + c*kT and p' = p + c*kp over the
// contiguous rows [0, H) x pitch of one level (padding columns included: they are never read back).
// 128-bit accesses; 16 B/cell of traffic (weather_simulation.cpp:201-214, 311-319).
template <typename T>
__global__ void __launch_bounds__(256)
    axpy_const2_kernel(const Geometry<T> g, const T *__restrict__ ta, T *__restrict__ to, const T *__restrict__ pa,
                       T *__restrict__ po, T c, T kt, T kp) {
    constexpr int VEC = 16 / (int)sizeof(T);
    const long long n = (long long)g.H * g.pitch / VEC;  // pitch is a multiple of 128 bytes
    const long long base = (long long)blockIdx.y * g.level_stride;
    const T ckt = Ar<T>::mul(c, kt), ckp = Ar<T>::mul(c, kp);  // c*k once: the same product for every cell
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        alignas(16) T a[VEC];
        alignas(16) T b[VEC];
        *reinterpret_cast<int4 *>(a) = *reinterpret_cast<const int4 *>(ta + base + i * VEC);
        *reinterpret_cast<int4 *>(b) = *reinterpret_cast<const int4 *>(pa + base + i * VEC);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            a[e] = Ar<T>::add(a[e], ckt);
            b[e] = Ar<T>::add(b[e], ckp);
        }
        *reinterpret_cast<int4 *>(to + base + i * VEC) = *reinterpret_cast<const int4 *>(a);
        *reinterpret_cast<int4 *>(po + base + i * VEC) = *reinterpret_cast<const int4 *>(b);
    }
}

template <typename T>
__global__ void __launch_bounds__(kBX *kBY) fill_kernel(const Geometry<T> g, T *__restrict__ p, T value, int halo) {
    const int x = blockIdx.x * kBX + threadIdx.x;
    const int y = blockIdx.y * kBY + threadIdx.y - halo;
    if (x >= g.pitch || y >= g.H + halo) return;
    p[(long long)blockIdx.z * g.level_stride + (long long)y * g.pitch + x] = value;
}

template <typename T>
__global__ void __launch_bounds__(256)
    mass_energy_kernel(const Geometry<T> g, const T *__restrict__ u, const T *__restrict__ v, const T *__restrict__ h,
                       double gravity, double *__restrict__ partial) {
    double m = 0.0, e = 0.0;
    const long long cells_per_level = (long long)g.W * g.H;
    const long long total = cells_per_level * g.L;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total;
         c += (long long)gridDim.x * blockDim.x) {
        const long long lvl = c / cells_per_level;
        const long long r = c - lvl * cells_per_level;
        const long long y = r / g.W, x = r - y * g.W;
        const long long i = lvl * g.level_stride + y * g.pitch + x;
        const double hu = (double)u[i], hv = (double)v[i], hh = (double)h[i];
        m += hh;
        e += 0.5 * hh * (hu * hu + hv * hv) + 0.5 * gravity * hh * hh;
    }
    __shared__ double sm[256], se[256];
    sm[threadIdx.x] = m;
    se[threadIdx.x] = e;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            sm[threadIdx.x] += sm[threadIdx.x + s];
            se[threadIdx.x] += se[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = sm[0];
        partial[2 * blockIdx.x + 1] = se[0];
    }
}

template <typename T>
dim3 grid_for(const Geometry<T> &g, int rows, int cols) {
    return dim3((cols + kBX - 1) / kBX, (rows + kBY - 1) / kBY, g.L);
}

}  // namespace

template <typename T>
cudaError_t launch_stage_direct(const Geometry<T> &g, const Physics<T> &ph, const StageArgs<T> &a, cudaStream_t st) {
    const int rows = a.y_end - a.y_begin;
    if (rows <= 0) return cudaSuccess;
    const dim3 grid = grid_for(g, rows, g.W), block(kBX, kBY);
    const bool store = a.KS.u != nullptr;
#define WSB_LAUNCH(R, F, S) stage_direct_kernel<T, R, F, S><<<grid, block, 0, st>>>(g, ph, a)
    if (ph.recip) {
        if (a.final_stage) WSB_LAUNCH(true, true, false);
        else if (store) WSB_LAUNCH(true, false, true);
        else WSB_LAUNCH(true, false, false);
    } else {
        if (a.final_stage) WSB_LAUNCH(false, true, false);
        else if (store) WSB_LAUNCH(false, false, true);
        else WSB_LAUNCH(false, false, false);
    }
#undef WSB_LAUNCH
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_diagnostics(const Geometry<T> &g, const Physics<T> &ph, const T *u, const T *v, T *vort, T *dvg,
                               cudaStream_t st) {
    constexpr int V = 16 / (int)sizeof(T);
    const dim3 grid((g.W + kDiagThreads * V - 1) / (kDiagThreads * V), (g.H + kDiagRows - 1) / kDiagRows, g.L);
    if (ph.recip) diagnostics_kernel<T, true><<<grid, kDiagThreads, 0, st>>>(g, ph, u, v, vort, dvg);
    else diagnostics_kernel<T, false><<<grid, kDiagThreads, 0, st>>>(g, ph, u, v, vort, dvg);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_axpy_const(const Geometry<T> &g, const T *y, T *o, T c, T k, cudaStream_t st) {
    axpy_const_kernel<T><<<grid_for(g, g.H, g.W), dim3(kBX, kBY), 0, st>>>(g, y, o, c, k);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_axpy_const2(const Geometry<T> &g, const T *ta, T *to, const T *pa, T *po, T c, T kt, T kp,
                               cudaStream_t st) {
    axpy_const2_kernel<T><<<dim3(148 * 8, g.L), 256, 0, st>>>(g, ta, to, pa, po, c, kt, kp);
    return cudaGetLastError();
}

// Separable initial conditions (wsb_ic.cpp): field(x, y) = rowv[y] (+ colv[x]). The vectors hold the reference's own
// float values (evaluated on the host with libm, O(W + H) work); the sum is ONE fp32 addition per cell, exactly the
// reference's `base + var` (initial_conditions.cpp:520-528), then widened for fp64 grids. Every level gets the field.
template <typename T>
__global__ void __launch_bounds__(kBX *kBY)
    expand_separable_kernel(const Geometry<T> g, T *__restrict__ p, const float *__restrict__ rowv,
                            const float *__restrict__ colv) {
    const int x = blockIdx.x * kBX + threadIdx.x;
    const int y = blockIdx.y * kBY + threadIdx.y;
    if (x >= g.W || y >= g.H) return;
    const float v = colv ? __fadd_rn(rowv[y], colv[x]) : rowv[y];
    p[(long long)blockIdx.z * g.level_stride + (long long)y * g.pitch + x] = (T)v;
}

template <typename T>
cudaError_t launch_tracer_stage(const Geometry<T> &g, const Physics<T> &ph, const TracerArgs<T> &a, cudaStream_t st) {
    if (ph.recip) tracer_stage_kernel<T, true><<<grid_for(g, g.H, g.W), dim3(kBX, kBY), 0, st>>>(g, ph, a);
    else tracer_stage_kernel<T, false><<<grid_for(g, g.H, g.W), dim3(kBX, kBY), 0, st>>>(g, ph, a);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_expand_separable(const Geometry<T> &g, T *p, const float *rowv, const float *colv, cudaStream_t st) {
    expand_separable_kernel<T><<<grid_for(g, g.H, g.W), dim3(kBX, kBY), 0, st>>>(g, p, rowv, colv);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_fill(const Geometry<T> &g, T *p, T value, bool with_halo, cudaStream_t st) {
    const int halo = with_halo ? kHaloRows : 0;
    fill_kernel<T><<<grid_for(g, g.H + 2 * halo, g.pitch), dim3(kBX, kBY), 0, st>>>(g, p, value, halo);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_mass_energy(const Geometry<T> &g, const T *u, const T *v, const T *h, double gravity,
                               double *partial, int nblocks, cudaStream_t st) {
    mass_energy_kernel<T><<<nblocks, 256, 0, st>>>(g, u, v, h, gravity, partial);
    return cudaGetLastError();
}

#define WSB_INSTANTIATE(T)                                                                                          \
    template cudaError_t launch_stage_direct<T>(const Geometry<T> &, const Physics<T> &, const StageArgs<T> &,       \
                                                cudaStream_t);                                                      \
    template cudaError_t launch_diagnostics<T>(const Geometry<T> &, const Physics<T> &, const T *, const T *, T *,   \
                                               T *, cudaStream_t);                                                  \
    template cudaError_t launch_axpy_const<T>(const Geometry<T> &, const T *, T *, T, T, cudaStream_t);              \
    template cudaError_t launch_axpy_const2<T>(const Geometry<T> &, const T *, T *, const T *, T *, T, T, T,         \
                                               cudaStream_t);                                                       \
    template cudaError_t launch_fill<T>(const Geometry<T> &, T *, T, bool, cudaStream_t);                            \
    template cudaError_t launch_tracer_stage<T>(const Geometry<T> &, const Physics<T> &, const TracerArgs<T> &,       \
                                                cudaStream_t);                                                       \
    template cudaError_t launch_expand_separable<T>(const Geometry<T> &, T *, const float *, const float *,          \
                                                    cudaStream_t);                                                   \
    template cudaError_t launch_mass_energy<T>(const Geometry<T> &, const T *, const T *, const T *, double,         \
                                               double *, int, cudaStream_t);
WSB_INSTANTIATE(float)
WSB_INSTANTIATE(double)
#undef WSB_INSTANTIATE

}  // namespace wsb
