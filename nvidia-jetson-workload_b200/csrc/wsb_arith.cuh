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
