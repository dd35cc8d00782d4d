// Whole-step fused kernel: ALL Runge-Kutta stages of one time step in ONE pass over the grid (sm_100a).
//
// Per step the per-stage path moves 42-45 field elements per cell through HBM;
// this kernel reads y_n (3 elements) and writes y_{n+1} (3 elements). Everything in between lives in
// registers:
//
//   * a warp owns a strip of 32*V columns (V = elements per lane, 128-bit/64-bit vector loads) and
//     sweeps down the rows ("register-blocked sweep along the slow dimension");
//   * the sweep is time-skewed: when row n of y arrives, stage 1 is evaluated on row n-1, stage 2 on
//     row n-2, ... stage S on row n-S, each from a 3-row register window of the previous stage's output;
//   * horizontal neighbours come from warp shuffles; a strip carries S halo columns per side that are
//     recomputed (not exchanged), so warps never synchronise with each other: no shared memory, no
//     barriers;
//   * the clamp-to-self boundary of the reference (weather_simulation.cpp:510-513) is reproduced by
//     materialising the clamped copy of an edge row/column in the (otherwise dead) window slot just
//     outside the domain, so interior cells pay nothing for it.
//
// The arithmetic per cell is exactly wsb_arith.cuh (one IEEE operation per reference operation, no
// contraction), so results are bit-identical to the per-stage path and to the CPU oracle.
//
// Register windows (indices are iteration numbers m; slot = m mod period; the row loop is unrolled by
// 6 = lcm of all periods so every slot index is a compile-time constant):
//   Y  : y rows,           period 6  (rows m-S..m live, the rest is load prefetch)
//   Ls : stage-s output,   period 3  (s = 1..S-1)
//   K2 : k2 rows (RK4),    period 3     K3 : k3 rows (RK4), period 2
#include "wsb_arith.cuh"
#include "wsb_internal.h"

#include <cstdlib>

namespace wsb {

namespace {

constexpr int kWarpsPerCta = 1;  // one warp per CTA: every branch condition is provably warp-uniform
constexpr unsigned kFull = 0xffffffffu;

template <typename T, int V>
struct VecIO;

template <>
struct VecIO<float, 2> {
    static __device__ __forceinline__ void load(const float *p, float (&r)[2]) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
        r[0] = t.x; r[1] = t.y;
    }
    static __device__ __forceinline__ void store(float *p, const float (&r)[2]) {
        *reinterpret_cast<float2 *>(p) = make_float2(r[0], r[1]);
    }
};

template <>
struct VecIO<float, 1> {
    static __device__ __forceinline__ void load(const float *p, float (&r)[1]) { r[0] = __ldg(p); }
    static __device__ __forceinline__ void store(float *p, const float (&r)[1]) { *p = r[0]; }
};

template <>
struct VecIO<double, 1> {
    static __device__ __forceinline__ void load(const double *p, double (&r)[1]) { r[0] = __ldg(p); }
    static __device__ __forceinline__ void store(double *p, const double (&r)[1]) { *p = r[0]; }
};

template <>
struct VecIO<double, 2> {
    static __device__ __forceinline__ void load(const double *p, double (&r)[2]) {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
        r[0] = t.x; r[1] = t.y;
    }
    static __device__ __forceinline__ void store(double *p, const double (&r)[2]) {
        *reinterpret_cast<double2 *>(p) = make_double2(r[0], r[1]);
    }
};

template <typename T, int V>
struct Row {
    T f[3][V];  // [field u,v,h][element]
};

// Everything a warp keeps in registers while sweeping.
template <typename T, int NST, int V>
struct Windows {
    Row<T, V> Y[6];
    Row<T, V> Lv[(NST > 1 ? NST - 1 : 1)][3];
    Row<T, V> K2[3];
    Row<T, V> K3[2];
};

template <typename T, int V>
struct LaneCtx {
    int c0;          // first column held by this lane (may be out of the domain)
    int W;           // domain width
    bool edge_strip; // strip touches x = 0 or x = W-1 (warp uniform)
};

// Materialise the clamped copies just outside the domain: column -1 := column 0, column W := column W-1.
template <typename T, int V>
__device__ __forceinline__ void fix_columns(Row<T, V> &r, const LaneCtx<T, V> &lc) {
#pragma unroll
    for (int f = 0; f < 3; ++f) {
        if (V == 2) {
            const T from_right = __shfl_down_sync(kFull, r.f[f][0], 1);    // lane+1's first column
            const T from_left = __shfl_up_sync(kFull, r.f[f][V - 1], 1);   // lane-1's last column
            if (lc.c0 == -2) r.f[f][V - 1] = from_right;                   // column -1 := column 0
            if (lc.c0 == lc.W) r.f[f][0] = from_left;                      // column W := column W-1 (W even)
            if (lc.c0 + 1 == lc.W) r.f[f][V - 1] = r.f[f][0];              // column W := column W-1 (W odd)
        } else {
            const T from_right = __shfl_down_sync(kFull, r.f[f][0], 1);
            const T from_left = __shfl_up_sync(kFull, r.f[f][0], 1);
            if (lc.c0 == -1) r.f[f][0] = from_right;
            if (lc.c0 == lc.W) r.f[f][0] = from_left;
        }
    }
}

// k = tend(U, C, D) for the V cells of this lane; horizontal neighbours by shuffle.
template <typename T, int V, bool RECIP>
__device__ __forceinline__ void tendency_row(const Physics<T> &ph, const Row<T, V> &U, const Row<T, V> &C,
                                             const Row<T, V> &D, Row<T, V> &k) {
    T Lft[3], Rgt[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
        Lft[f] = __shfl_up_sync(kFull, C.f[f][V - 1], 1);   // left neighbour of my first cell
        Rgt[f] = __shfl_down_sync(kFull, C.f[f][0], 1);     // right neighbour of my last cell
    }
#pragma unroll
    for (int e = 0; e < V; ++e) {
        const T uL = (e == 0) ? Lft[0] : C.f[0][e - 1], uR = (e == V - 1) ? Rgt[0] : C.f[0][e + 1];
        const T vL = (e == 0) ? Lft[1] : C.f[1][e - 1], vR = (e == V - 1) ? Rgt[1] : C.f[1][e + 1];
        const T hL = (e == 0) ? Lft[2] : C.f[2][e - 1], hR = (e == V - 1) ? Rgt[2] : C.f[2][e + 1];
        tendency_cell<T, RECIP>(ph, C.f[0][e], C.f[1][e], C.f[2][e], uL, uR, U.f[0][e], D.f[0][e], vL, vR, U.f[1][e],
                                D.f[1][e], hL, hR, U.f[2][e], D.f[2][e], k.f[0][e], k.f[1][e], k.f[2][e]);
    }
}

template <typename T, int NST, int V, bool RECIP>
struct Sweep {
    using Win = Windows<T, NST, V>;

    const Geometry<T> &g;
    const Physics<T> &ph;
    const StepArgs<T> &a;
    LaneCtx<T, V> lc;
    int y0;          // first output row of this chunk (local)
    int niter;       // (y1 - y0) + 2*NST
    int gmin, gmax;  // local indices of global rows 0 and Hglobal
    int out_lo, out_hi;  // output column range of this strip
    long long lvl_off;
    int cl;          // clamped column for loads

    __device__ __forceinline__ Sweep(const Geometry<T> &g_, const Physics<T> &ph_, const StepArgs<T> &a_)
        : g(g_), ph(ph_), a(a_) {}

    // y row with iteration index m -> slot m % 6 (row index clamped into the global domain)
    __device__ __forceinline__ void load_y(Row<T, V> &dst, int m) const {
        int r = y0 - NST + m;
        r = max(r, gmin);
        r = min(r, gmax - 1);
        const long long off = lvl_off + (long long)r * g.pitch + cl;
        VecIO<T, V>::load(a.Y.u + off, dst.f[0]);
        VecIO<T, V>::load(a.Y.v + off, dst.f[1]);
        VecIO<T, V>::load(a.Y.h + off, dst.f[2]);
    }

    __device__ __forceinline__ void store_out(const Row<T, V> &o, int r) const {
        const long long off = lvl_off + (long long)r * g.pitch + lc.c0;
        const bool in0 = lc.c0 >= out_lo && lc.c0 < out_hi;
        const bool inl = lc.c0 + V - 1 >= out_lo && lc.c0 + V - 1 < out_hi;
        if (in0 && inl) {
            VecIO<T, V>::store(a.O.u + off, o.f[0]);
            VecIO<T, V>::store(a.O.v + off, o.f[1]);
            VecIO<T, V>::store(a.O.h + off, o.f[2]);
        } else if (V > 1) {
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const int c = lc.c0 + e;
                if (c >= out_lo && c < out_hi) {
                    a.O.u[off + e] = o.f[0][e];
                    a.O.v[off + e] = o.f[1][e];
                    a.O.h[off + e] = o.f[2][e];
                }
            }
        }
    }

    // Stage S1 (1-based) at iteration n, PH = n % 6. Center row index m = n - S1.
    template <int PH, int S1>
    __device__ __forceinline__ void stage(Win &w, int n) const {
        if (n < 2 * S1) return;                     // inputs not produced yet (pipeline fill)
        const int r = y0 - NST + n - S1;            // local row of the center
        constexpr int M = (PH - S1 + 12) % 6;       // m mod 6
        constexpr int m3 = M % 3, m3m = (M + 2) % 3, m3p = (M + 1) % 3;
        constexpr int m2 = M % 2;
        if (r < gmin) return;
        if (r >= gmax) {
            // one past the bottom edge: the next stage reads this slot as "down" of the last row
            if constexpr (S1 < NST) {
                if (r == gmax) {
                    asm volatile("");
                    w.Lv[S1 - 1][m3] = w.Lv[S1 - 1][m3m];
                }
            }
            return;
        }
        Row<T, V> ktmp;
        Row<T, V> &k = (NST == 4 && S1 == 2) ? w.K2[m3] : (NST == 4 && S1 == 3) ? w.K3[m2] : ktmp;
        if constexpr (S1 == 1) {
            tendency_row<T, V, RECIP>(ph, w.Y[(M + 5) % 6], w.Y[M], w.Y[(M + 1) % 6], k);
        } else {
            tendency_row<T, V, RECIP>(ph, w.Lv[S1 - 2][m3m], w.Lv[S1 - 2][m3], w.Lv[S1 - 2][m3p], k);
        }
        const Row<T, V> &yb = w.Y[M];
        if constexpr (S1 < NST) {
            // t_s = y + c*k   (c = 0.5f*dt for the half stages of RK2/RK4, dt for stage 3 of RK4)
            const T c = (NST == 4 && S1 == 3) ? a.dt : a.half_dt;
            Row<T, V> &t = w.Lv[S1 - 1][m3];
#pragma unroll
            for (int f = 0; f < 3; ++f)
#pragma unroll
                for (int e = 0; e < V; ++e) t.f[f][e] = axpy<T>(yb.f[f][e], c, k.f[f][e]);
            if (lc.edge_strip) fix_columns<T, V>(t, lc);
            if (r == gmin) {                         // row -1 := row 0 ("up" of the first row)
                asm volatile("");                    // keep this rare copy a branch, not 6 selects per row
                w.Lv[S1 - 1][m3m] = t;
            }
        } else {
            Row<T, V> o;
#pragma unroll
            for (int f = 0; f < 3; ++f)
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    if constexpr (NST == 4) {
                        // reference aliasing: "k1" reads k4 at the combine (weather_simulation.cpp:350-351, F5)
                        o.f[f][e] = rk4_combine<T>(yb.f[f][e], a.dt6, k.f[f][e], w.K2[m3].f[f][e], w.K3[m2].f[f][e],
                                                   k.f[f][e]);
                    } else {
                        o.f[f][e] = axpy<T>(yb.f[f][e], a.dt, k.f[f][e]);
                    }
                }
            store_out(o, r);
        }
    }

    template <int PH>
    __device__ __forceinline__ void iteration(Win &w, int n) const {
        constexpr int PF = 6 - (NST + 1 > 3 ? NST + 1 : 3);  // prefetch distance in rows
        if (n + PF < niter) {
            Row<T, V> &dst = w.Y[(PH + PF) % 6];
            load_y(dst, n + PF);
            if (lc.edge_strip) fix_columns<T, V>(dst, lc);
        }
        stage<PH, 1>(w, n);
        if constexpr (NST >= 2) stage<PH, 2>(w, n);
        if constexpr (NST >= 4) {
            stage<PH, 3>(w, n);
            stage<PH, 4>(w, n);
        }
    }

    __device__ __forceinline__ void run() const {
        constexpr int PF = 6 - (NST + 1 > 3 ? NST + 1 : 3);
        Win w;
        // prologue: rows 0..PF-1
#pragma unroll
        for (int m = 0; m < PF; ++m) {
            if (m < niter) {
                load_y(w.Y[m], m);
                if (lc.edge_strip) fix_columns<T, V>(w.Y[m], lc);
            }
        }
        int n = 0;
        for (; n + 6 <= niter; n += 6) {
            iteration<0>(w, n);
            iteration<1>(w, n + 1);
            iteration<2>(w, n + 2);
            iteration<3>(w, n + 3);
            iteration<4>(w, n + 4);
            iteration<5>(w, n + 5);
        }
        if (n < niter) iteration<0>(w, n);
        if (n + 1 < niter) iteration<1>(w, n + 1);
        if (n + 2 < niter) iteration<2>(w, n + 2);
        if (n + 3 < niter) iteration<3>(w, n + 3);
        if (n + 4 < niter) iteration<4>(w, n + 4);
    }
};

// MINB = minimum resident CTAs (= warps) per SM the register allocator must allow: 65536 / (32 * MINB) registers.
template <typename T, int NST, int V, bool RECIP, int MINB>
__global__ void __launch_bounds__(kWarpsPerCta * 32, MINB)
    step_fused_kernel(const Geometry<T> g, const Physics<T> ph, const StepArgs<T> a, const int rows_per_chunk) {
    constexpr int SW = 32 * V;
    constexpr int HX = (NST + V - 1) / V * V;  // halo columns per side, a multiple of V so vector loads stay aligned
    constexpr int OUTW = SW - 2 * HX;
    const int lane = threadIdx.x & 31;
    const int strip = blockIdx.x;
    if (strip * OUTW >= g.W) return;  // warp-uniform
    // blockIdx.y enumerates the chunks of the first row range, then those of the optional second one
    const int nchunks1 = (a.y_end - a.y_begin + rows_per_chunk - 1) / rows_per_chunk;
    const bool second = (int)blockIdx.y >= nchunks1;
    const int cy = second ? (int)blockIdx.y - nchunks1 : (int)blockIdx.y;
    const int y0 = (second ? a.y_begin2 : a.y_begin) + cy * rows_per_chunk;
    const int y1 = min(y0 + rows_per_chunk, second ? a.y_end2 : a.y_end);
    if (y0 >= y1) return;

    Sweep<T, NST, V, RECIP> sw(g, ph, a);
    const int xs = strip * OUTW - HX;
    sw.lc.c0 = xs + lane * V;
    sw.lc.W = g.W;
    sw.lc.edge_strip = (xs < 0) || (xs + SW > g.W);
    sw.y0 = y0;
    sw.niter = (y1 - y0) + 2 * NST;
    sw.gmin = -g.row0;
    sw.gmax = g.Hglobal - g.row0;
    sw.out_lo = strip * OUTW;
    sw.out_hi = min(sw.out_lo + OUTW, g.W);
    sw.lvl_off = (long long)blockIdx.z * g.level_stride;
    sw.cl = min(max(sw.lc.c0, 0), g.pitch - V);
    sw.run();
}

int env_int(const char *name, int dflt) {
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

// tuning knobs (read once): rows each warp sweeps, and the occupancy variant of the RK4 kernel
int rows_per_chunk_default() {
    static const int v = env_int("WSB_FUSED_ROWS_PER_CHUNK", 128);
    return v > 0 ? v : 128;
}

int minb_variant() {
    static const int v = env_int("WSB_FUSED_MINB", 0);
    return v;
}

template <typename T, int NST, int V, int MINB>
cudaError_t launch_impl(const Geometry<T> &g, const Physics<T> &ph, const StepArgs<T> &a, cudaStream_t st) {
    const int rows = a.y_end - a.y_begin, rows2 = a.y_end2 - a.y_begin2;
    if (rows <= 0) return cudaSuccess;
    constexpr int OUTW = 32 * V - 2 * ((NST + V - 1) / V * V);
    const int strips = (g.W + OUTW - 1) / OUTW;
    const int rpc = a.rows_per_chunk > 0 ? a.rows_per_chunk : rows_per_chunk_default();
    const int chunks = (rows + rpc - 1) / rpc + (rows2 > 0 ? (rows2 + rpc - 1) / rpc : 0);
    const dim3 grid((strips + kWarpsPerCta - 1) / kWarpsPerCta, chunks, g.L);
    const dim3 block(kWarpsPerCta * 32);
    if (ph.recip) step_fused_kernel<T, NST, V, true, MINB><<<grid, block, 0, st>>>(g, ph, a, rpc);
    else step_fused_kernel<T, NST, V, false, MINB><<<grid, block, 0, st>>>(g, ph, a, rpc);
    return cudaGetLastError();
}

}  // namespace

bool step_fused_supported(int nstages, int dtype) {
    // fp64 RK4 would need ~240 registers of window state per lane: it stays on the per-stage path
    if (dtype == WSB_F64) return nstages == 1 || nstages == 2;
    return nstages == 1 || nstages == 2 || nstages == 4;
}

template <>
cudaError_t launch_step_fused<float>(const Geometry<float> &g, const Physics<float> &ph, const StepArgs<float> &a,
                                     int nstages, cudaStream_t st) {
    if (a.classical && nstages == 4) return cudaErrorNotSupported;
    switch (nstages) {
        case 1: return launch_impl<float, 1, 2, 16>(g, ph, a, st);
        case 2: return launch_impl<float, 2, 2, 12>(g, ph, a, st);
        case 4:
            switch (minb_variant()) {
                case 12: return launch_impl<float, 4, 2, 12>(g, ph, a, st);
                case 16: return launch_impl<float, 4, 2, 16>(g, ph, a, st);
                default: return launch_impl<float, 4, 2, 8>(g, ph, a, st);
            }
        default: return cudaErrorNotSupported;
    }
}

template <>
cudaError_t launch_step_fused<double>(const Geometry<double> &g, const Physics<double> &ph, const StepArgs<double> &a,
                                      int nstages, cudaStream_t st) {
    if (a.classical && nstages == 4) return cudaErrorNotSupported;
    switch (nstages) {
        case 1: return launch_impl<double, 1, 1, 16>(g, ph, a, st);
        case 2: return launch_impl<double, 2, 1, 12>(g, ph, a, st);
        default: return cudaErrorNotSupported;
    }
}

}  // namespace wsb
