// Exactly-rounded scalar arithmetic for the stencil: every operation is ONE IEEE round-to-nearest
// operation in the reference's association. The __f*_rn / __d*_rn intrinsics are never contracted into
// FMAs by nvcc, whatever -fmad says (contraction alone moves the reference against
// itself by 5.6e-3 after 1000 steps).
#pragma once

#include <cuda_runtime.h>

#include "wsb_internal.h"

namespace wsb {

template <typename T>
struct Ar;

template <>
struct Ar<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};

template <>
struct Ar<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

// x / d for a loop-invariant divisor with r = RN(1/d) (correctly rounded, computed once on the host), as three
// operations instead of the ~10-instruction IEEE division sequence (weather_simulation.cpp:521-528 divides six times
// per cell and stage):  q = RN(x*r);  e = x - q*d (exact, one FMA);  q' = RN(q + e*r).
// Markstein's theorem: with r the correctly rounded reciprocal and q within one ulp of x/d, q' IS the correctly
// rounded quotient -- provided nothing under- or overflows on the way (the remainder e must be representable) and x
// is not -0 (the sequence returns +0). x is therefore tested against a window of ordinary magnitudes (or +0); any
// other operand takes the IEEE division. The FMAs here are the algorithm, not a contraction.
template <typename T>
struct FastDiv;
template <>
struct FastDiv<float> {
    static __device__ __forceinline__ bool ordinary(float x) {  // +0, or 2^-100 <= |x| < 2^100
        const unsigned b = __float_as_uint(x);
        return b == 0u || ((b << 1) - (27u << 24)) < (200u << 24);
    }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
};
template <>
struct FastDiv<double> {
    static __device__ __forceinline__ bool ordinary(double x) {  // +0, or 2^-900 <= |x| < 2^900
        const unsigned hi = (unsigned)__double2hiint(x);
        return __double_as_longlong(x) == 0ll || ((hi << 1) - (123u << 21)) < (1800u << 21);
    }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
};
template <typename T>
__device__ __forceinline__ T div_fast(T x, T d, T r) {
    const T q = Ar<T>::mul(x, r);
    const T e = FastDiv<T>::fma(-q, d, x);
    return FastDiv<T>::fma(e, r, q);
}
template <typename T>
__device__ __forceinline__ T div_by_invariant(T x, T d, T r) {
    // r == 0: the host found d outside the window where the sequence is proven
    return (r != T(0) && FastDiv<T>::ordinary(x)) ? div_fast<T>(x, d, r) : Ar<T>::div(x, d);
}
// The six centred differences of a cell divided at once. The IEEE divisions live in ONE out-of-line function per
// type: inlined at every site (six per cell and stage, hundreds per unrolled whole-step kernel) their ~30
// instructions plus slow-path call made the true-division kernels 17 k instructions long and instruction-fetch
// bound (3.1 ms per 8192^2 RK4 step); a cell takes the call only if one of its six operands is not ordinary.
template <typename T>
struct Six {
    T q[6];
};
template <typename T>
__device__ __noinline__ Six<T> ieee_div6(T a0, T a1, T a2, T a3, T a4, T a5, T ddx, T ddy) {
    Six<T> r;
    r.q[0] = Ar<T>::div(a0, ddx); r.q[1] = Ar<T>::div(a1, ddy); r.q[2] = Ar<T>::div(a2, ddx);
    r.q[3] = Ar<T>::div(a3, ddy); r.q[4] = Ar<T>::div(a4, ddx); r.q[5] = Ar<T>::div(a5, ddy);
    return r;
}
// x: the differences (ux, uy, vx, vy, hx, hy order: dx, dy, dx, dy, dx, dy)
template <typename T>
__device__ __forceinline__ Six<T> div6_by_invariant(const T (&x)[6], const Physics<T> &ph) {
    Six<T> r;
    bool ok = ph.rdx != T(0);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        ok = ok && FastDiv<T>::ordinary(x[k]);
        r.q[k] = div_fast<T>(x[k], (k & 1) ? ph.ddy : ph.ddx, (k & 1) ? ph.rdy : ph.rdx);
    }
    if (!ok) r = ieee_div6<T>(x[0], x[1], x[2], x[3], x[4], x[5], ph.ddx, ph.ddy);
    return r;
}

// (hi - lo) / (2*d): true IEEE division (by the exact three-operation sequence above), or the bit-identical multiply
// by an exact power-of-two reciprocal.
template <typename T, bool RECIP>
__device__ __forceinline__ T cdiff(T hi, T lo, T dd, T rd) {
    T d = Ar<T>::sub(hi, lo);
    if constexpr (RECIP) return Ar<T>::mul(d, rd);
    else return div_by_invariant<T>(d, dd, rd);
}

// the six centred differences of weather_simulation.cpp:521-528
template <typename T, bool RECIP>
__device__ __forceinline__ void centred_differences(const Physics<T> &ph, T uL, T uR, T uU, T uD, T vL, T vR, T vU, T vD,
                                                    T hL, T hR, T hU, T hD, T &ux, T &uy, T &vx, T &vy, T &hx, T &hy) {
    using A = Ar<T>;
    if constexpr (RECIP) {
        ux = A::mul(A::sub(uR, uL), ph.rdx); uy = A::mul(A::sub(uD, uU), ph.rdy);
        vx = A::mul(A::sub(vR, vL), ph.rdx); vy = A::mul(A::sub(vD, vU), ph.rdy);
        hx = A::mul(A::sub(hR, hL), ph.rdx); hy = A::mul(A::sub(hD, hU), ph.rdy);
    } else {
        const T d[6] = {A::sub(uR, uL), A::sub(uD, uU), A::sub(vR, vL), A::sub(vD, vU), A::sub(hR, hL), A::sub(hD, hU)};
        const Six<T> q = div6_by_invariant<T>(d, ph);
        ux = q.q[0]; uy = q.q[1]; vx = q.q[2]; vy = q.q[3]; hx = q.q[4]; hy = q.q[5];
    }
}

// weather_simulation.cpp:516-537 for one cell. L/R/U/D are the clamped neighbours (:510-513).
//   du = ((((-u)*ux) - (v*uy)) - (g*hx)) + (f*v)
//   dv = ((((-u)*vx) - (v*vy)) - (g*hy)) - (f*u)
//   dh = (((-h)*(ux+vy)) - (u*hx)) - (v*hy)
template <typename T, bool RECIP>
__device__ __forceinline__ void tendency_cell(const Physics<T> &ph, T u, T v, T h, T uL, T uR, T uU, T uD, T vL,
                                              T vR, T vU, T vD, T hL, T hR, T hU, T hD, T &du, T &dv, T &dh) {
    using A = Ar<T>;
    T ux, uy, vx, vy, hx, hy;
    centred_differences<T, RECIP>(ph, uL, uR, uU, uD, vL, vR, vU, vD, hL, hR, hU, hD, ux, uy, vx, vy, hx, hy);
    du = A::add(A::sub(A::sub(A::mul(-u, ux), A::mul(v, uy)), A::mul(ph.g, hx)), A::mul(ph.f, v));
    dv = A::sub(A::sub(A::sub(A::mul(-u, vx), A::mul(v, vy)), A::mul(ph.g, hy)), A::mul(ph.f, u));
    dh = A::sub(A::sub(A::mul(-h, A::add(ux, vy)), A::mul(u, hx)), A::mul(v, hy));
}

// Extended physics for one cell: the tendencies above on a beta plane plus viscosity / diffusivity. fy is the row's
// Coriolis parameter (ext_coriolis). Operation order: oracle/ws_oracle_body.inc (the specification of this mode).
template <typename T>
__device__ __forceinline__ T ext_coriolis(const Physics<T> &ph, int global_row) {
    using A = Ar<T>;
    return A::add(ph.f, A::mul(ph.bdy, A::sub((T)global_row, ph.yc)));
}
template <typename T>
__device__ __forceinline__ T ext_laplacian(const Physics<T> &ph, T c, T l, T r, T u, T d) {
    using A = Ar<T>;
    const T two_c = A::mul(T(2), c);
    const T tx = A::add(A::sub(r, two_c), l), ty = A::add(A::sub(d, two_c), u);
    return A::add(A::mul(tx, ph.idx2), A::mul(ty, ph.idy2));
}
template <typename T, bool RECIP>
__device__ __forceinline__ void tendency_cell_ext(const Physics<T> &ph, T fy, T u, T v, T h, T uL, T uR, T uU, T uD,
                                                  T vL, T vR, T vU, T vD, T hL, T hR, T hU, T hD, T &du, T &dv, T &dh) {
    using A = Ar<T>;
    T ux, uy, vx, vy, hx, hy;
    centred_differences<T, RECIP>(ph, uL, uR, uU, uD, vL, vR, vU, vD, hL, hR, hU, hD, ux, uy, vx, vy, hx, hy);
    const T a = A::add(A::sub(A::sub(A::mul(-u, ux), A::mul(v, uy)), A::mul(ph.g, hx)), A::mul(fy, v));
    const T b = A::sub(A::sub(A::sub(A::mul(-u, vx), A::mul(v, vy)), A::mul(ph.g, hy)), A::mul(fy, u));
    const T c = A::sub(A::sub(A::mul(-h, A::add(ux, vy)), A::mul(u, hx)), A::mul(v, hy));
    du = A::add(a, A::mul(ph.nu, ext_laplacian<T>(ph, u, uL, uR, uU, uD)));
    dv = A::add(b, A::mul(ph.nu, ext_laplacian<T>(ph, v, vL, vR, vU, vD)));
    dh = A::add(c, A::mul(ph.kappa, ext_laplacian<T>(ph, h, hL, hR, hU, hD)));
}

// y + c*k  (weather_simulation.cpp:187, 249, 381 ...: `0.5f * dt_ * k` is (0.5f*dt_)*k, c is pre-rounded)
template <typename T>
__device__ __forceinline__ T axpy(T y, T c, T k) {
    return Ar<T>::add(y, Ar<T>::mul(c, k));
}

// y + (dt/6.0f) * (((k1 + 2.0f*k2) + 2.0f*k3) + k4)   (weather_simulation.cpp:438-440)
template <typename T>
__device__ __forceinline__ T rk4_combine(T y, T dt6, T k1, T k2, T k3, T k4) {
    using A = Ar<T>;
    const T two = T(2);
    T s = A::add(A::add(A::add(k1, A::mul(two, k2)), A::mul(two, k3)), k4);
    return A::add(y, A::mul(dt6, s));
}

}  // namespace wsb

// ---- packed fp32x2 arithmetic (sm_100: FMUL2 / FADD2 / FFMA2, two IEEE operations per issue slot) -----
// ptxas 12.9 contracts a packed multiply feeding a packed add/sub into FFMA2 even for mul.rn/add.rn and
// even under --fmad=false, which would break bit parity. What it does NOT touch is an explicit
// fma.rn.f32x2 whose multiplier is the constant -1: `c - p` for a product p is therefore issued as
// fma(p, -1, c) = round((-p) + c), which IS the IEEE subtraction c - p, bit for bit (signed zeros included).
// Sums `c + a*b` become c - ((-a)*b) with the sign carried by a constant factor ((-a)*b == -(a*b) exactly), so
// every operation of the stencil is packed and none is contracted: the RECIP kernels contain FFMA2 only with
// the immediate multiplier -1 (or 2, folded mode), checked on the built library by profiles/check_no_fma.sh.
namespace wsb {

struct F2 {
    float x, y;
};

__device__ __forceinline__ unsigned long long f2_pack(F2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ F2 f2_unpack(unsigned long long v) {
    F2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ F2 f2_splat(float v) { return F2{v, v}; }
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
}
// sums/differences whose operands are NOT products (safe from contraction)
__device__ __forceinline__ F2 f2_sub_packed(F2 a, F2 b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
}
__device__ __forceinline__ F2 f2_add_packed(F2 a, F2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
}
// c - p for a product p: fma(p, -1, c), exact and never contracted further
__device__ __forceinline__ F2 f2_sub_prod(F2 c, F2 p) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(p)), "l"(0xBF800000BF800000ULL), "l"(f2_pack(c)));
    return f2_unpack(r);
}
// c + 2*k: one FFMA2 with the immediate 2. Identical to c + round(2*k) unless 2*k overflows (folded mode only)
__device__ __forceinline__ F2 f2_add_twice(F2 c, F2 k) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(k)), "l"(0x4000000040000000ULL), "l"(f2_pack(c)));
    return f2_unpack(r);
}

// Constants of the packed path, splatted once per kernel. The negated copies carry the signs (see above).
struct PhysicsF2 {
    F2 nrdx, rdy;  // -1/(2dx), 1/(2dy)
    F2 g, ng, f, nf;
};

__device__ __forceinline__ PhysicsF2 physics_f2(const Physics<float> &ph) {
    PhysicsF2 p;
    p.nrdx = f2_splat(-ph.rdx);
    p.rdy = f2_splat(ph.rdy);
    p.g = f2_splat(ph.g);
    p.ng = f2_splat(-ph.g);
    p.f = f2_splat(ph.f);
    p.nf = f2_splat(-ph.f);
    return p;
}

// tendency_cell for a pair of horizontally adjacent cells (x, y) = (cell 0, cell 1), exact-reciprocal spacing
// only. uLft/uRgt... are the cells left of cell 0 and right of cell 1 (from the neighbouring lanes). Every
// operation is the reference's (weather_simulation.cpp:516-537), in its order; with nqx := -(q_x):
//   (-u)*ux == u*nux,  g*hx == (-g)*nhx,  -(u*hx) == u*nhx,  a + f*v == a - (-f)*v      (all exact, signs included)
// so the results are bit-identical to tendency_cell<float, true>.
__device__ __forceinline__ void tendency_pair(const PhysicsF2 &ph, F2 u, F2 v, F2 h, float uLft, float uRgt, F2 uU, F2 uD,
                                              float vLft, float vRgt, F2 vU, F2 vD, float hLft, float hRgt, F2 hU,
                                              F2 hD, F2 &du, F2 &dv, F2 &dh) {
    // horizontal differences as scalar subtractions (no register pairing needed): R - L per cell
    const F2 dux = F2{__fsub_rn(u.y, uLft), __fsub_rn(uRgt, u.x)};
    const F2 dvx = F2{__fsub_rn(v.y, vLft), __fsub_rn(vRgt, v.x)};
    const F2 dhx = F2{__fsub_rn(h.y, hLft), __fsub_rn(hRgt, h.x)};
    const F2 nux = f2_mul(dux, ph.nrdx);
    const F2 uy = f2_mul(f2_sub_packed(uD, uU), ph.rdy);
    const F2 nvx = f2_mul(dvx, ph.nrdx);
    const F2 vy = f2_mul(f2_sub_packed(vD, vU), ph.rdy);
    const F2 nhx = f2_mul(dhx, ph.nrdx);
    const F2 hy = f2_mul(f2_sub_packed(hD, hU), ph.rdy);
    // du = ((((-u)*ux) - (v*uy)) - (g*hx)) + (f*v)
    du = f2_sub_prod(f2_sub_prod(f2_sub_prod(f2_mul(u, nux), f2_mul(v, uy)), f2_mul(ph.ng, nhx)), f2_mul(ph.nf, v));
    // dv = ((((-u)*vx) - (v*vy)) - (g*hy)) - (f*u)
    dv = f2_sub_prod(f2_sub_prod(f2_sub_prod(f2_mul(u, nvx), f2_mul(v, vy)), f2_mul(ph.g, hy)), f2_mul(ph.f, u));
    // dh = (((-h)*(ux+vy)) - (u*hx)) - (v*hy);  ux + vy == vy - nux;  (-P) - Q == (-Q) - P
    const F2 s = f2_sub_prod(vy, nux);
    dh = f2_sub_prod(f2_sub_prod(f2_mul(u, nhx), f2_mul(h, s)), f2_mul(v, hy));
}

// ---- packed path for a spacing whose 2dx / 2dy are NOT powers of two ---------------------------------------------
// tendency_pair with every centred difference divided by the exact three-operation sequence of div_by_invariant,
// packed: q = D*r, e = D - q*d (one FFMA2 -- an intended fused operation, not a contraction), q' = q + e*r. The x
// quotients are produced negated (r -> -r: every rounding is sign-symmetric), as tendency_pair wants them. Valid
// only where every difference is +0 or of ordinary magnitude: the CALLER guarantees that (wsb_step_tma.cu keeps
// a warp-uniform flag per input row: all cells +0 or 2^-70 <= |x| < 2^99, hence differences +0 or 2^-93 <= |D| <
// 2^100) and takes the scalar path with the IEEE fallback for any other row.
struct PhysicsDivF2 {
    F2 nrx, ry;    // -RN(1/(2dx)), RN(1/(2dy))
    F2 ddx, nddy;  // 2dx, -(2dy)
    F2 g, ng, f, nf;
};
__device__ __forceinline__ PhysicsDivF2 physics_div_f2(const Physics<float> &ph) {
    PhysicsDivF2 p;
    p.nrx = f2_splat(-ph.rdx);
    p.ry = f2_splat(ph.rdy);
    p.ddx = f2_splat(ph.ddx);
    p.nddy = f2_splat(-ph.ddy);
    p.g = f2_splat(ph.g);
    p.ng = f2_splat(-ph.g);
    p.f = f2_splat(ph.f);
    p.nf = f2_splat(-ph.f);
    return p;
}
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
    return f2_unpack(r);
}
// -(D / d) with nr = -RN(1/d):  nq = D*nr;  e = D + nq*d;  nq' = nq + e*nr
__device__ __forceinline__ F2 f2_neg_div(F2 D, F2 nr, F2 d) {
    const F2 nq = f2_mul(D, nr);
    return f2_fma(f2_fma(nq, d, D), nr, nq);
}
// D / d with r = RN(1/d), nd = -d:  q = D*r;  e = D + q*nd;  q' = q + e*r
__device__ __forceinline__ F2 f2_div(F2 D, F2 r, F2 nd) {
    const F2 q = f2_mul(D, r);
    return f2_fma(f2_fma(q, nd, D), r, q);
}
__device__ __forceinline__ bool ordinary_input(float x) {  // +0, or 2^-70 <= |x| < 2^99
    const unsigned b = __float_as_uint(x);
    return b == 0u || ((b << 1) - (57u << 24)) < (169u << 24);
}
__device__ __forceinline__ void tendency_pair_div(const PhysicsDivF2 &ph, F2 u, F2 v, F2 h, float uLft, float uRgt, F2 uU,
                                                  F2 uD, float vLft, float vRgt, F2 vU, F2 vD, float hLft, float hRgt,
                                                  F2 hU, F2 hD, F2 &du, F2 &dv, F2 &dh) {
    const F2 dux = F2{__fsub_rn(u.y, uLft), __fsub_rn(uRgt, u.x)};
    const F2 dvx = F2{__fsub_rn(v.y, vLft), __fsub_rn(vRgt, v.x)};
    const F2 dhx = F2{__fsub_rn(h.y, hLft), __fsub_rn(hRgt, h.x)};
    const F2 nux = f2_neg_div(dux, ph.nrx, ph.ddx);
    const F2 uy = f2_div(f2_sub_packed(uD, uU), ph.ry, ph.nddy);
    const F2 nvx = f2_neg_div(dvx, ph.nrx, ph.ddx);
    const F2 vy = f2_div(f2_sub_packed(vD, vU), ph.ry, ph.nddy);
    const F2 nhx = f2_neg_div(dhx, ph.nrx, ph.ddx);
    const F2 hy = f2_div(f2_sub_packed(hD, hU), ph.ry, ph.nddy);
    // from here on: tendency_pair, operation for operation
    du = f2_sub_prod(f2_sub_prod(f2_sub_prod(f2_mul(u, nux), f2_mul(v, uy)), f2_mul(ph.ng, nhx)), f2_mul(ph.nf, v));
    dv = f2_sub_prod(f2_sub_prod(f2_sub_prod(f2_mul(u, nvx), f2_mul(v, vy)), f2_mul(ph.g, hy)), f2_mul(ph.f, u));
    const F2 s = f2_sub_prod(vy, nux);
    dh = f2_sub_prod(f2_sub_prod(f2_mul(u, nhx), f2_mul(h, s)), f2_mul(v, hy));
}

// Extended physics, packed (exact-reciprocal spacing): tendency_pair with the row's Coriolis parameter instead of the
// constant f, plus nu*lap(u), nu*lap(v), kappa*lap(h). Same operations as tendency_cell_ext<float, true>, bit for bit:
// (a - 2c) is c2 := (-2)*c subtracted as a product, p*idx2 + q*idy2 is p*idx2 - q*(-idy2), a + nu*L is a - (-nu)*L.
struct PhysicsExtF2 {
    F2 idx2, nidy2, nnu, nkappa;
};
__device__ __forceinline__ PhysicsExtF2 physics_ext_f2(const Physics<float> &ph) {
    PhysicsExtF2 p;
    p.idx2 = f2_splat(ph.idx2);
    p.nidy2 = f2_splat(-ph.idy2);
    p.nnu = f2_splat(-ph.nu);
    p.nkappa = f2_splat(-ph.kappa);
    return p;
}
__device__ __forceinline__ F2 ext_laplacian_pair(const PhysicsExtF2 &pe, F2 c, float lft, float rgt, F2 up, F2 dn) {
    const F2 two_c = f2_mul(f2_splat(2.0f), c);
    // horizontal: (R - 2c) + L per cell, scalar (the neighbours sit in different registers of the pair)
    const F2 tx = F2{__fadd_rn(__fsub_rn(c.y, two_c.x), lft), __fadd_rn(__fsub_rn(rgt, two_c.y), c.x)};
    const F2 ty = f2_add_packed(f2_sub_prod(dn, two_c), up);
    return f2_sub_prod(f2_mul(tx, pe.idx2), f2_mul(ty, pe.nidy2));
}
__device__ __forceinline__ void tendency_pair_ext(const PhysicsF2 &ph, const PhysicsExtF2 &pe, float fy, F2 u, F2 v, F2 h,
                                                  float uLft, float uRgt, F2 uU, F2 uD, float vLft, float vRgt, F2 vU,
                                                  F2 vD, float hLft, float hRgt, F2 hU, F2 hD, F2 &du, F2 &dv, F2 &dh) {
    PhysicsF2 row = ph;  // the row's Coriolis parameter replaces the constant one
    row.f = f2_splat(fy);
    row.nf = f2_splat(-fy);
    F2 a, b, c;
    tendency_pair(row, u, v, h, uLft, uRgt, uU, uD, vLft, vRgt, vU, vD, hLft, hRgt, hU, hD, a, b, c);
    du = f2_sub_prod(a, f2_mul(pe.nnu, ext_laplacian_pair(pe, u, uLft, uRgt, uU, uD)));
    dv = f2_sub_prod(b, f2_mul(pe.nnu, ext_laplacian_pair(pe, v, vLft, vRgt, vU, vD)));
    dh = f2_sub_prod(c, f2_mul(pe.nkappa, ext_laplacian_pair(pe, h, hLft, hRgt, hU, hD)));
}

// ---- folded mode (opt-in, WSB_ARITH_FOLDED): dx == dy and 2dx a power of two, r = 1/(2dx) ---------------------
// Scaling by a power of two commutes with every rounding as long as nothing under- or overflows, so the six
// multiplications by r are dropped: the tendencies are carried as K' = K/r (differences unscaled, f' = f/r) and r
// is folded into the stage coefficients (c*r, dt6*r: exact). 32 instead of 38 operations per cell-stage.
// Results are bit-identical to the strict path unless an intermediate of the reference is subnormal (the
// reference then rounds (a-b)*r or a product at subnormal granularity, the folded form does not) or within a
// factor 1/r of overflow; the sign of an exact zero tendency can differ (nD = L - R is +0, not -0, when L == R),
// which is invisible unless a field holds -0. tests/test_parity_gpu.py states both bounds.
struct PhysicsFold {
    F2 g, ng, fs, nfs;  // fs = f/r
};

__device__ __forceinline__ PhysicsFold physics_fold(const Physics<float> &ph) {
    PhysicsFold p;
    p.g = f2_splat(ph.g);
    p.ng = f2_splat(-ph.g);
    p.fs = f2_splat(ph.f * ph.ddx);  // f / r, exact (ddx is a power of two)
    p.nfs = f2_splat(-(ph.f * ph.ddx));
    return p;
}

__device__ __forceinline__ void tendency_pair_folded(const PhysicsFold &ph, F2 u, F2 v, F2 h, float uLft, float uRgt,
                                                     F2 uU, F2 uD, float vLft, float vRgt, F2 vU, F2 vD, float hLft,
                                                     float hRgt, F2 hU, F2 hD, F2 &du, F2 &dv, F2 &dh) {
    // negated horizontal differences L - R per cell (scalar), vertical differences D - U (packed)
    const F2 ndux = F2{__fsub_rn(uLft, u.y), __fsub_rn(u.x, uRgt)};
    const F2 ndvx = F2{__fsub_rn(vLft, v.y), __fsub_rn(v.x, vRgt)};
    const F2 ndhx = F2{__fsub_rn(hLft, h.y), __fsub_rn(h.x, hRgt)};
    const F2 duy = f2_sub_packed(uD, uU), dvy = f2_sub_packed(vD, vU), dhy = f2_sub_packed(hD, hU);
    du = f2_sub_prod(f2_sub_prod(f2_sub_prod(f2_mul(u, ndux), f2_mul(v, duy)), f2_mul(ph.ng, ndhx)), f2_mul(ph.nfs, v));
    dv = f2_sub_prod(f2_sub_prod(f2_sub_prod(f2_mul(u, ndvx), f2_mul(v, dvy)), f2_mul(ph.g, dhy)), f2_mul(ph.fs, u));
    const F2 s = f2_sub_packed(dvy, ndux);
    dh = f2_sub_prod(f2_sub_prod(f2_mul(u, ndhx), f2_mul(h, s)), f2_mul(v, dhy));
}

// y + dt6r * (((k1 + 2*k2) + 2*k3) + k4) on folded tendencies, ndt6r = -(dt6*r); 2*k folded into FFMA2
__device__ __forceinline__ F2 rk4_combine_pair_folded(F2 y, F2 ndt6r, F2 k1, F2 k2, F2 k3, F2 k4) {
    const F2 s = f2_add_packed(f2_add_twice(f2_add_twice(k1, k2), k3), k4);
    return f2_sub_prod(y, f2_mul(ndt6r, s));
}

// y + c*k as y - (nc*k), nc = -c
__device__ __forceinline__ F2 axpy_pair(F2 y, F2 nc, F2 k) { return f2_sub_prod(y, f2_mul(nc, k)); }

// y + dt6 * (((k1 + 2*k2) + 2*k3) + k4), ndt6 = -dt6. 2*k as a product of its own (overflows like the reference's).
__device__ __forceinline__ F2 rk4_combine_pair(F2 y, F2 ndt6, F2 k1, F2 k2, F2 k3, F2 k4) {
    const F2 ntwo = f2_splat(-2.0f);
    const F2 s = f2_add_packed(f2_sub_prod(f2_sub_prod(k1, f2_mul(ntwo, k2)), f2_mul(ntwo, k3)), k4);
    return f2_sub_prod(y, f2_mul(ndt6, s));
}

}  // namespace wsb
