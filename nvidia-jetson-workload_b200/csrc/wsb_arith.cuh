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

// (hi - lo) / (2*d): true IEEE division, or the bit-identical multiply by an exact power-of-two reciprocal.
template <typename T, bool RECIP>
__device__ __forceinline__ T cdiff(T hi, T lo, T dd, T rd) {
    T d = Ar<T>::sub(hi, lo);
    return RECIP ? Ar<T>::mul(d, rd) : Ar<T>::div(d, dd);
}

// weather_simulation.cpp:516-537 for one cell. L/R/U/D are the clamped neighbours (:510-513).
//   du = ((((-u)*ux) - (v*uy)) - (g*hx)) + (f*v)
//   dv = ((((-u)*vx) - (v*vy)) - (g*hy)) - (f*u)
//   dh = (((-h)*(ux+vy)) - (u*hx)) - (v*hy)
template <typename T, bool RECIP>
__device__ __forceinline__ void tendency_cell(const Physics<T> &ph, T u, T v, T h, T uL, T uR, T uU, T uD, T vL,
                                              T vR, T vU, T vD, T hL, T hR, T hU, T hD, T &du, T &dv, T &dh) {
    using A = Ar<T>;
    const T ux = cdiff<T, RECIP>(uR, uL, ph.ddx, ph.rdx);
    const T uy = cdiff<T, RECIP>(uD, uU, ph.ddy, ph.rdy);
    const T vx = cdiff<T, RECIP>(vR, vL, ph.ddx, ph.rdx);
    const T vy = cdiff<T, RECIP>(vD, vU, ph.ddy, ph.rdy);
    const T hx = cdiff<T, RECIP>(hR, hL, ph.ddx, ph.rdx);
    const T hy = cdiff<T, RECIP>(hD, hU, ph.ddy, ph.rdy);
    du = A::add(A::sub(A::sub(A::mul(-u, ux), A::mul(v, uy)), A::mul(ph.g, hx)), A::mul(ph.f, v));
    dv = A::sub(A::sub(A::sub(A::mul(-u, vx), A::mul(v, vy)), A::mul(ph.g, hy)), A::mul(ph.f, u));
    dh = A::sub(A::sub(A::mul(-h, A::add(ux, vy)), A::mul(u, hx)), A::mul(v, hy));
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

// ---- packed fp32x2 arithmetic (sm_100: FMUL2 / FADD2, two IEEE operations per issue slot) -----------
// ptxas 12.9 contracts a packed multiply feeding a packed add/sub into FFMA2 even for mul.rn/add.rn and
// even under --fmad=false, which would break bit parity. The rule used here therefore is:
// every add/sub that consumes a product is issued as two SCALAR add.rn.f32 (never contracted), while
// multiplies and the adds/subs of non-products are packed. The build checks that the RECIP kernels
// contain no FFMA/FFMA2 at all (profiles/check_no_fma.sh).
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
// packed: products, and sums/differences whose operands are not products
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
    return f2_unpack(r);
}
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
// scalar pairs: for sums/differences that consume a product
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) { return F2{__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)}; }
__device__ __forceinline__ F2 f2_sub(F2 a, F2 b) { return F2{__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)}; }
__device__ __forceinline__ F2 f2_neg(F2 a) { return F2{-a.x, -a.y}; }

// Constants of the packed path, splatted once per kernel.
struct PhysicsF2 {
    F2 rdx, rdy, g, f;
};

// tendency_cell for a pair of horizontally adjacent cells (x, y) = (cell 0, cell 1), exact-reciprocal
// spacing only. uLft/uRgt... are the cells left of cell 0 and right of cell 1 (from the neighbouring lanes).
// Same operation order as tendency_cell<float, true>: results are bit-identical.
__device__ __forceinline__ void tendency_pair(const PhysicsF2 &ph, F2 u, F2 v, F2 h, float uLft, float uRgt, F2 uU, F2 uD,
                                              float vLft, float vRgt, F2 vU, F2 vD, float hLft, float hRgt, F2 hU,
                                              F2 hD, F2 &du, F2 &dv, F2 &dh) {
    // horizontal differences as scalar subtractions (no register pairing needed): R - L per cell
    const F2 dux = F2{__fsub_rn(u.y, uLft), __fsub_rn(uRgt, u.x)};
    const F2 dvx = F2{__fsub_rn(v.y, vLft), __fsub_rn(vRgt, v.x)};
    const F2 dhx = F2{__fsub_rn(h.y, hLft), __fsub_rn(hRgt, h.x)};
    const F2 ux = f2_mul(dux, ph.rdx);
    const F2 uy = f2_mul(f2_sub_packed(uD, uU), ph.rdy);
    const F2 vx = f2_mul(dvx, ph.rdx);
    const F2 vy = f2_mul(f2_sub_packed(vD, vU), ph.rdy);
    const F2 hx = f2_mul(dhx, ph.rdx);
    const F2 hy = f2_mul(f2_sub_packed(hD, hU), ph.rdy);
    // ((((-u)*ux) - (v*uy)) - (g*hx)) + (f*v); (-u)*ux == -(u*ux) exactly
    du = f2_add(f2_sub(f2_sub(f2_neg(f2_mul(u, ux)), f2_mul(v, uy)), f2_mul(ph.g, hx)), f2_mul(ph.f, v));
    dv = f2_sub(f2_sub(f2_sub(f2_neg(f2_mul(u, vx)), f2_mul(v, vy)), f2_mul(ph.g, hy)), f2_mul(ph.f, u));
    dh = f2_sub(f2_sub(f2_neg(f2_mul(h, f2_add(ux, vy))), f2_mul(u, hx)), f2_mul(v, hy));
}

__device__ __forceinline__ F2 axpy_pair(F2 y, F2 c, F2 k) { return f2_add(y, f2_mul(c, k)); }

// y + dt6 * (((k1 + 2*k2) + 2*k3) + k4)
__device__ __forceinline__ F2 rk4_combine_pair(F2 y, F2 dt6, F2 k1, F2 k2, F2 k3, F2 k4) {
    const F2 two = f2_splat(2.0f);
    const F2 s = f2_add_packed(f2_add(f2_add(k1, f2_mul(two, k2)), f2_mul(two, k3)), k4);
    return f2_add(y, f2_mul(dt6, s));
}

}  // namespace wsb
