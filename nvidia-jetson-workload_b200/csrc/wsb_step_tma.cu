// Whole-step fused kernel, TMA-staged variant (sm_100a): ALL Runge-Kutta stages of one time step in ONE
// pass over the grid, like wsb_step_fused.cu, but the rows of y_n are brought on chip by the TMA engine.
//
//   * each warp owns a strip of 32*V columns and sweeps down the rows with the same time skew as the
//     register variant (stage s works on row n-s when row n of y arrives);
//   * y rows land in a per-warp shared-memory ring (3 groups of 3 rows) through the TMA engine, one mbarrier
//     per group: in the common case ONE tiled copy per field (cp.async.bulk.tensor.3d over a CUtensorMap of
//     the plane, box = strip width x 3 rows, SASS UTMALDG; out-of-range columns are zero-filled by the
//     hardware), at the clamped domain top/bottom row-wise 1-D bulk copies (SASS UBLKCP). One elected lane is
//     the producer: as soon as the rows of triple q-2 are dead it re-arms their group with the rows of
//     triple q+1, so HBM latency is hidden without spending registers on prefetch;
//   * y is re-read from the ring wherever a stage needs it (stencil rows of stage 1, base rows y + c*k),
//     the k2/k3 rows the RK4 combine needs later are parked in shared memory too (private per lane, no
//     synchronisation), and only the 3-row windows of the intermediate stage states stay in registers:
//     92 registers instead of ~170, i.e. 18 instead of 12 resident warps per SM (shared-memory limited), and
//     a 3x instead of 6x unrolled loop body;
//   * a steady-state fast path (interior strip, pipeline full, no domain edge in reach) runs without a
//     single boundary test; the general path is the register variant's logic.
//
// Arithmetic per cell is wsb_arith.cuh: bit-identical to every other path and to the CPU oracle.
#include "wsb_arith.cuh"
#include "wsb_internal.h"

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

namespace wsb {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kGroups = 3;          // the y ring is 3 groups of 3 rows; the row loop is unrolled by 3,
constexpr int kRing = 3 * kGroups;  // so every shared-memory offset is a compile-time constant

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WSB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra WSB_DONE;\n"
        "bra WSB_WAIT;\n"
        "WSB_DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// TMA tiled copy: a box of the 3-D tensor (x, row, level) -> shared memory; out-of-range elements arrive as zeros
// and count towards complete_tx like the others (SASS: UTMALDG)
__device__ __forceinline__ void tensor_g2s(uint32_t dst, const CUtensorMap *map, int x, int y, int z, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// generic <-> async proxy ordering for GLOBAL memory: rows stored by other CTAs (generic proxy), acquired through a
// flag, are about to be read by the TMA engine (async proxy)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_relaxed_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

template <typename T, int V>
struct Row {
    T f[3][V];
};

template <typename T, int V>
struct SmemIO;
template <>
struct SmemIO<float, 2> {
    static __device__ __forceinline__ void ld(const float *p, float (&r)[2]) {
        const float2 t = *reinterpret_cast<const float2 *>(p);
        r[0] = t.x; r[1] = t.y;
    }
    static __device__ __forceinline__ void st(float *p, const float (&r)[2]) {
        *reinterpret_cast<float2 *>(p) = make_float2(r[0], r[1]);
    }
};
template <>
struct SmemIO<float, 4> {
    static __device__ __forceinline__ void ld(const float *p, float (&r)[4]) {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
    }
    static __device__ __forceinline__ void st(float *p, const float (&r)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(r[0], r[1], r[2], r[3]);
    }
};
template <>
struct SmemIO<double, 2> {
    static __device__ __forceinline__ void ld(const double *p, double (&r)[2]) {
        const double2 t = *reinterpret_cast<const double2 *>(p);
        r[0] = t.x; r[1] = t.y;
    }
    static __device__ __forceinline__ void st(double *p, const double (&r)[2]) {
        *reinterpret_cast<double2 *>(p) = make_double2(r[0], r[1]);
    }
};
template <>
struct SmemIO<double, 1> {
    static __device__ __forceinline__ void ld(const double *p, double (&r)[1]) { r[0] = *p; }
    static __device__ __forceinline__ void st(double *p, const double (&r)[1]) { *p = r[0]; }
};

template <typename T, int NST, int V>
struct Layout {
    static constexpr int SW = 32 * V;                                   // strip width in columns
    static constexpr int ALIGN = 16 / (int)sizeof(T);                   // TMA needs 16-byte aligned rows
    static constexpr int HX = (NST + ALIGN - 1) / ALIGN * ALIGN;        // halo columns per side
    static constexpr int OUTW = SW - 2 * HX;                            // output columns per strip
    static constexpr int FIELD_BYTES = SW * (int)sizeof(T);
    static constexpr int ROW_ELEMS = 3 * SW;
    static constexpr int ROW_BYTES = 3 * FIELD_BYTES;
    // a ring group holds 3 rows as [field][row][SW]: one TMA box (SW columns x 3 rows) per field
    static constexpr int GROUP_ELEMS = 3 * ROW_ELEMS;
    static constexpr int GROUP_BYTES = 3 * ROW_BYTES;
    static __host__ __device__ constexpr int y_elem(int row_in_group, int field) { return (field * 3 + row_in_group) * SW; }
    // RK4 parks two 3-row rings in shared memory (k2 and k3; classical opt-in: k1, k1+2k2, k1+2k2+2k3)
    static __host__ __device__ constexpr int k_rows(bool classical) { return NST == 4 ? (classical ? 9 : 6) : 0; }
    // rings + one mbarrier per group + the three peer-push offsets (PeerExchange)
    static __host__ __device__ constexpr int smem_bytes(bool classical) { return (kRing + k_rows(classical)) * ROW_BYTES + kGroups * 8 + 3 * 8; }
};

// One lane of the (converged) warp. With elect.sync the compiler knows exactly one lane runs the producer code, so
// the uniform-datapath UBLKCPs need no per-copy ELECT / vote / branch bookkeeping (0.622 -> 0.603 ms/step).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <typename T, int NST, int V, bool RECIP, bool CL, bool FOLD, bool EXT>
struct SweepT {
    using L = Layout<T, NST, V>;
    using Windows = Row<T, V>[(NST > 1 ? NST - 1 : 1)][3];

    // lane-offset shared-memory bases of the ring groups holding rows of triple q, q-1 and q-2
    struct Groups {
        const T *a, *b, *c;
    };

    const Geometry<T> &g;
    const Physics<T> &ph;
    const StepArgs<T> &a;
    PhysicsF2 ph2;    // splatted constants of the packed path
    PhysicsFold phf;  // ... of its folded variant (wsb_arith.cuh)
    PhysicsExtF2 phe; // ... of the extended physics (EXT)
    PhysicsDivF2 phd; // ... of the packed exact-division path (kPackedDiv)
    bool div_unproven; // the host has no proven reciprocal for this spacing (ph.rdx == 0): IEEE path everywhere
    float rs;         // FOLD: r = 1/(2dx), folded into every stage coefficient; else unused
    const CUtensorMap *tm_u, *tm_v, *tm_h;  // tiled-TMA descriptors of the y_n planes (kernel parameters)
    int lane, c0, xs;
    bool edge_strip, fix_left, fix_right;
    int y0, niter, gmin, gmax, out_lo, out_hi;
    long long lvl_off;
    T *ring;        // [kGroups][3 fields][3 rows][SW]
    T *k2s;         // [3][3][SW], lane-offset
    T *k3s;         // [3][3][SW], lane-offset
    T *k1s;         // [3][3][SW], lane-offset (classical RK4 only)
    static constexpr bool classical = NST == 4 && CL;  // textbook RK4 opt-in: its own kernel instantiation
    uint32_t ring_u32, bar_u32;
    bool st_vec, st_e[V], ragged;
    bool push;              // band CTA with a neighbour: every stored row also goes into the neighbour's ghost rows
    const long long *pdel;  // shared memory: byte offsets from this rank's output addresses to the peer's (u, v, h)

    // packed fp32x2 math (FMUL2/FADD2) is used for the float, 2-cells-per-lane, exact-reciprocal kernels
    static constexpr bool kPacked = std::is_same<T, float>::value && V % 2 == 0 && RECIP;
    // non-power-of-two spacing: the same packed arithmetic with the exact three-operation division, for rows whose
    // input cells are all ordinary (tracked per input row in `hz`, see stage()); the scalar path with the IEEE
    // fallback otherwise
#ifndef WSB_PACKED_DIV
#define WSB_PACKED_DIV 1
#endif
    static constexpr bool kPackedDiv = WSB_PACKED_DIV != 0 && std::is_same<T, float>::value && V % 2 == 0 && !RECIP && !EXT;
#ifndef WSB_STEADY_BODY
#define WSB_STEADY_BODY 1
#endif
    static constexpr bool kSteadyBody = WSB_STEADY_BODY != 0;

    __device__ __forceinline__ SweepT(const Geometry<T> &g_, const Physics<T> &ph_, const StepArgs<T> &a_)
        : g(g_), ph(ph_), a(a_) {}

    // ---- producer side (one elected lane): one group = 3 consecutive rows, one mbarrier ----------------
    __device__ __forceinline__ void issue_group(int q, int grp) const {
        const int m0 = 3 * q;
        const int nr = min(3, niter - m0);
        if (nr <= 0) return;
        const uint32_t bar = bar_u32 + 8u * grp;
        mbar_expect_tx(bar, (uint32_t)nr * L::ROW_BYTES);
        const int r0 = y0 - NST + m0;
        if (nr == 3 && r0 >= gmin && r0 + 2 < gmax) {
            // common case, three consecutive rows of the slab: one tiled-TMA box (SW columns x 3 rows) per field.
            // Columns outside [0, W) arrive as zeros (never used: clamp columns are patched in wait_group).
            const uint32_t dst = ring_u32 + (uint32_t)grp * L::GROUP_BYTES;
            const int ty = r0 + kLeadRows, tz = (int)blockIdx.z;  // rows count from the start of the allocation
            tensor_g2s(dst, tm_u, xs, ty, tz, bar);
            tensor_g2s(dst + 3 * L::FIELD_BYTES, tm_v, xs, ty, tz, bar);
            tensor_g2s(dst + 6 * L::FIELD_BYTES, tm_h, xs, ty, tz, bar);
            return;
        }
        for (int i = 0; i < nr; ++i) {
            int r = y0 - NST + m0 + i;
            r = max(r, gmin);
            r = min(r, gmax - 1);  // clamp-to-self rows of the reference (weather_simulation.cpp:512-513)
            const long long off = lvl_off + (long long)r * g.pitch + xs;
            const uint32_t dst = ring_u32 + (uint32_t)grp * L::GROUP_BYTES;
            bulk_g2s(dst + L::y_elem(i, 0) * sizeof(T), a.Y.u + off, L::FIELD_BYTES, bar);
            bulk_g2s(dst + L::y_elem(i, 1) * sizeof(T), a.Y.v + off, L::FIELD_BYTES, bar);
            bulk_g2s(dst + L::y_elem(i, 2) * sizeof(T), a.Y.h + off, L::FIELD_BYTES, bar);
        }
    }

    // ---- consumer side ------------------------------------------------------------------------------
    __device__ __forceinline__ void wait_group(int q, int grp, uint32_t parity) const {
        mbar_wait(bar_u32 + 8u * grp, parity);
        if (edge_strip) {
            // clamp-to-self columns (weather_simulation.cpp:510-511): column -1 := column 0, column W := column W-1
            if (lane == 0) {
                const int nr = min(3, niter - 3 * q);
                for (int i = 0; i < nr; ++i) {
#pragma unroll
                    for (int f = 0; f < 3; ++f) {
                        T *row = ring + grp * L::GROUP_ELEMS + L::y_elem(i, f);
                        if (fix_left) row[L::HX - 1] = row[L::HX];
                        if (fix_right) row[g.W - xs] = row[g.W - xs - 1];
                    }
                }
            }
            __syncwarp();
        }
    }

    // y row (n - J) at phase PH = n % 3: a compile-time offset from one of the three group bases
    template <int PH, int J>
    __device__ __forceinline__ void ld_y(Row<T, V> &dst, const Groups &G) const {
        constexpr int i = PH - J;
        const T *grp = (i >= 0) ? G.a : (i >= -3) ? G.b : G.c;
        constexpr int ri = (i >= 0) ? i : (i >= -3) ? 3 + i : 6 + i;
#pragma unroll
        for (int f = 0; f < 3; ++f) SmemIO<T, V>::ld(grp + L::y_elem(ri, f), dst.f[f]);
    }
    __device__ __forceinline__ void ld_k(Row<T, V> &dst, const T *base, int slot) const {
        const T *row = base + slot * L::ROW_ELEMS;
#pragma unroll
        for (int f = 0; f < 3; ++f) SmemIO<T, V>::ld(row + f * L::SW, dst.f[f]);
    }
    __device__ __forceinline__ void st_k(const Row<T, V> &src, T *base, int slot) const {
        T *row = base + slot * L::ROW_ELEMS;
#pragma unroll
        for (int f = 0; f < 3; ++f) SmemIO<T, V>::st(row + f * L::SW, src.f[f]);
    }

    __device__ __forceinline__ void fix_columns(Row<T, V> &r) const {
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const T from_right = __shfl_down_sync(kFull, r.f[f][0], 1);
            const T from_left = __shfl_up_sync(kFull, r.f[f][V - 1], 1);
            if (c0 + V - 1 == -1) r.f[f][V - 1] = from_right;           // column -1 := column 0
            if (c0 == g.W) r.f[f][0] = from_left;                       // column W := column W-1
#pragma unroll
            for (int e = 1; e < V; ++e)
                if (c0 + e == g.W) r.f[f][e] = r.f[f][e - 1];           // ... when W-1 and W share a lane
        }
    }

    // r: local index of the centre row (EXT: the beta plane needs the global one)
    // slow (kPackedDiv only, warp-uniform): one of the three input rows holds a cell that is not ordinary
    __device__ __forceinline__ void tendency_row(const Row<T, V> &U, const Row<T, V> &C, const Row<T, V> &D,
                                                 Row<T, V> &k, int r, bool slow = false) const {
        T fy = T(0);
        if constexpr (EXT) fy = ext_coriolis<T>(ph, r + g.row0);
        T Lft[3], Rgt[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            Lft[f] = __shfl_up_sync(kFull, C.f[f][V - 1], 1);
            Rgt[f] = __shfl_down_sync(kFull, C.f[f][0], 1);
        }
        if constexpr (kPacked || kPackedDiv) {
            if (!(kPackedDiv && slow)) {
            // pairs of adjacent cells (2p, 2p+1): outer neighbours come from the lane's other cells or the shuffles
#pragma unroll
            for (int p = 0; p < V / 2; ++p) {
                auto P = [p](const T(&r)[V]) { return F2{(float)r[2 * p], (float)r[2 * p + 1]}; };
                float lf[3], rg[3];
#pragma unroll
                for (int f = 0; f < 3; ++f) {
                    lf[f] = (p == 0) ? (float)Lft[f] : (float)C.f[f][2 * p - 1];
                    rg[f] = (p == V / 2 - 1) ? (float)Rgt[f] : (float)C.f[f][2 * p + 2];
                }
                F2 du, dv, dh;
                if constexpr (kPackedDiv)
                    tendency_pair_div(phd, P(C.f[0]), P(C.f[1]), P(C.f[2]), lf[0], rg[0], P(U.f[0]), P(D.f[0]), lf[1], rg[1],
                                      P(U.f[1]), P(D.f[1]), lf[2], rg[2], P(U.f[2]), P(D.f[2]), du, dv, dh);
                else if constexpr (EXT)
                    tendency_pair_ext(ph2, phe, (float)fy, P(C.f[0]), P(C.f[1]), P(C.f[2]), lf[0], rg[0], P(U.f[0]), P(D.f[0]),
                                      lf[1], rg[1], P(U.f[1]), P(D.f[1]), lf[2], rg[2], P(U.f[2]), P(D.f[2]), du, dv, dh);
                else if constexpr (FOLD)
                    tendency_pair_folded(phf, P(C.f[0]), P(C.f[1]), P(C.f[2]), lf[0], rg[0], P(U.f[0]), P(D.f[0]), lf[1],
                                         rg[1], P(U.f[1]), P(D.f[1]), lf[2], rg[2], P(U.f[2]), P(D.f[2]), du, dv, dh);
                else
                    tendency_pair(ph2, P(C.f[0]), P(C.f[1]), P(C.f[2]), lf[0], rg[0], P(U.f[0]), P(D.f[0]), lf[1], rg[1],
                                  P(U.f[1]), P(D.f[1]), lf[2], rg[2], P(U.f[2]), P(D.f[2]), du, dv, dh);
                k.f[0][2 * p] = du.x; k.f[0][2 * p + 1] = du.y;
                k.f[1][2 * p] = dv.x; k.f[1][2 * p + 1] = dv.y;
                k.f[2][2 * p] = dh.x; k.f[2][2 * p + 1] = dh.y;
            }
            return;
            }
        }
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const T uL = (e == 0) ? Lft[0] : C.f[0][e - 1], uR = (e == V - 1) ? Rgt[0] : C.f[0][e + 1];
            const T vL = (e == 0) ? Lft[1] : C.f[1][e - 1], vR = (e == V - 1) ? Rgt[1] : C.f[1][e + 1];
            const T hL = (e == 0) ? Lft[2] : C.f[2][e - 1], hR = (e == V - 1) ? Rgt[2] : C.f[2][e + 1];
            if constexpr (EXT)
                tendency_cell_ext<T, RECIP>(ph, fy, C.f[0][e], C.f[1][e], C.f[2][e], uL, uR, U.f[0][e], D.f[0][e], vL, vR,
                                            U.f[1][e], D.f[1][e], hL, hR, U.f[2][e], D.f[2][e], k.f[0][e], k.f[1][e],
                                            k.f[2][e]);
            else
            tendency_cell<T, RECIP>(ph, C.f[0][e], C.f[1][e], C.f[2][e], uL, uR, U.f[0][e], D.f[0][e], vL, vR,
                                    U.f[1][e], D.f[1][e], hL, hR, U.f[2][e], D.f[2][e], k.f[0][e], k.f[1][e],
                                    k.f[2][e]);
        }
    }

    // po: this lane's output addresses (u, v, h) of the row being stored, advanced one row per iteration by run()
    __device__ __forceinline__ void store_rows(const Row<T, V> &o, T *const (&po)[3]) const {
        if (!ragged) {  // block-uniform: every lane is entirely inside or entirely outside the output range
            if (st_vec) {
#pragma unroll
                for (int f = 0; f < 3; ++f) SmemIO<T, V>::st(po[f], o.f[f]);  // plain vector store (generic address)
            }
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e)
                if (st_e[e]) {
#pragma unroll
                    for (int f = 0; f < 3; ++f) po[f][e] = o.f[f][e];
                }
        }
    }
    __device__ __forceinline__ void store_out(const Row<T, V> &o, T *const (&po)[3]) const {
        store_rows(o, po);
        if (push) {  // block-uniform; the same row into the neighbour's ghost rows (peer mapping over NVLink)
            T *pp[3];
#pragma unroll
            for (int f = 0; f < 3; ++f) pp[f] = reinterpret_cast<T *>(reinterpret_cast<char *>(po[f]) + pdel[f]);
            store_rows(o, pp);
        }
    }

    // kPackedDiv: one warp-uniform bit per input row of every stage (group 0: the y rows, group s: the rows stage s
    // produced; slot = row mod 3) -- set when any lane holds a cell of that row that is not ordinary. A stage takes the
    // packed division only while the three rows of its input group are clean.
    __device__ __forceinline__ void mark_row(unsigned &hz, int bit, const Row<T, V> &row) const {
        if constexpr (kPackedDiv) {
            bool bad = false;
#pragma unroll
            for (int f = 0; f < 3; ++f)
#pragma unroll
                for (int e = 0; e < V; ++e) bad = bad || !ordinary_input((float)row.f[f][e]);
            const unsigned any = __any_sync(kFull, bad) ? 1u : 0u;
            hz = (hz & ~(1u << bit)) | (any << bit);
        }
    }
    static __device__ __forceinline__ void copy_mark(unsigned &hz, int from, int to) {
        if constexpr (kPackedDiv) hz = (hz & ~(1u << to)) | (((hz >> from) & 1u) << to);
    }

    // Steady state of the boundary-free body: every lane is entirely inside or outside the output range (interior
    // strips are never ragged), so the row goes out as three PREDICATED vector stores -- no branch, the loop body
    // stays one basic block.
    __device__ __forceinline__ void store_out_predicated(const Row<T, V> &o, T *const (&po)[3]) const {
        if constexpr (std::is_same<T, float>::value && V == 2) {
#pragma unroll
            for (int f = 0; f < 3; ++f)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.global.v2.f32 [%0], {%1, %2};\n\t}" ::"l"(po[f]),
                             "f"((float)o.f[f][0]), "f"((float)o.f[f][1]), "r"((unsigned)st_vec)
                             : "memory");
        } else {
            store_out(o, po);
        }
    }

    // Stage S1 (1-based) at iteration n; PH = n % 3. Center row index m = n - S1.
    // PURE: the CTA is nowhere near a domain edge (interior strip, every row it loads is a real row): no boundary
    // handling is compiled in at all, pipeline-fill iterations just compute on garbage that is never stored.
    // STEADY (PURE only): the pipeline is full and the triple is complete -- no test of any kind in the body.
    template <int PH, int S1, bool PURE, bool STEADY>
    __device__ __forceinline__ void stage(Windows &Lv, Row<T, V> (&Yw)[3], unsigned &hz, int n, const Groups &G,
                                          const bool FAST_, T *const (&po)[3]) const {
        const bool FAST = PURE || FAST_;
        const bool slow = kPackedDiv && (div_unproven || ((hz >> (3 * (S1 - 1))) & 7u) != 0u);
        constexpr int m3 = (PH - S1 + 12) % 3, m3m = (m3 + 2) % 3, m3p = (m3 + 1) % 3;
        const int r = y0 - NST + n - S1;
        if constexpr (PURE && S1 > 1 && !STEADY) {
            // pipeline fill: the kernel is fp32-pipe bound, the 20 useless stage-rows per 64-row chunk are worth a
            // uniform branch per stage (stage 1 always runs: it also loads the y row into the carried window)
            if (n < 2 * S1) return;
        }
        if (!FAST) {
            if (n < 2 * S1) return;  // pipeline fill
            if (r < gmin) return;
            if (r >= gmax) {
                if constexpr (S1 < NST) {
                    if (r == gmax) {  // one past the bottom edge: "down" of the last row is the last row
                        asm volatile("");
                        Lv[S1 - 1][m3] = Lv[S1 - 1][m3m];
                        copy_mark(hz, 3 * S1 + m3m, 3 * S1 + m3);
                    }
                }
                return;
            }
        }
        // boundary-free body: the last three y rows are carried in registers (slot = row mod 3), so stage 1 loads
        // one row per iteration instead of three and stage 2 finds its base row there
        constexpr bool kYwin = PURE;
        Row<T, V> k, yb;
        if constexpr (S1 == 1) {
            if constexpr (kYwin) {
                ld_y<PH, 0>(Yw[PH], G);
                mark_row(hz, PH, Yw[PH]);
                yb = Yw[(PH + 2) % 3];
                tendency_row(Yw[(PH + 1) % 3], yb, Yw[PH], k, r, kPackedDiv && (div_unproven || (hz & 7u) != 0u));
            } else {
                Row<T, V> U, D;
                ld_y<PH, 2>(U, G);
                ld_y<PH, 1>(yb, G);
                ld_y<PH, 0>(D, G);
                tendency_row(U, yb, D, k, r, slow);  // (the y rows are marked in iteration(), also while this stage idles)
            }
        } else {
            tendency_row(Lv[S1 - 2][m3m], Lv[S1 - 2][m3], Lv[S1 - 2][m3p], k, r, slow);
            if constexpr (kYwin && S1 == 2) yb = Yw[(PH + 1) % 3];
            else ld_y<PH, S1>(yb, G);
        }
        if constexpr (S1 < NST) {
            const T c = (NST == 4 && S1 == 3) ? a.dt : a.half_dt;
            Row<T, V> &t = Lv[S1 - 1][m3];
            if constexpr (kPacked) {
                const F2 c2 = f2_splat(FOLD ? -((float)c * rs) : -(float)c);  // axpy_pair takes the negated coefficient
#pragma unroll
                for (int f = 0; f < 3; ++f)
#pragma unroll
                    for (int p = 0; p < V; p += 2) {
                        const F2 rr = axpy_pair(F2{(float)yb.f[f][p], (float)yb.f[f][p + 1]}, c2,
                                                F2{(float)k.f[f][p], (float)k.f[f][p + 1]});
                        t.f[f][p] = rr.x; t.f[f][p + 1] = rr.y;
                    }
            } else {
#pragma unroll
                for (int f = 0; f < 3; ++f)
#pragma unroll
                    for (int e = 0; e < V; ++e) t.f[f][e] = axpy<T>(yb.f[f][e], c, k.f[f][e]);
            }
            if (!FAST && edge_strip) fix_columns(t);
            mark_row(hz, 3 * S1 + m3, t);
            if (!FAST) {
                if (r == gmin) {  // row -1 := row 0 ("up" of the first row)
                    asm volatile("");
                    Lv[S1 - 1][m3m] = t;
                    copy_mark(hz, 3 * S1 + m3, 3 * S1 + m3m);
                }
            }
            if constexpr (NST == 4) {
                if constexpr (!classical) {
                    if constexpr (S1 == 2) st_k(k, k2s, m3);
                    if constexpr (S1 == 3) st_k(k, k3s, m3);
                } else {
                    // textbook RK4 (opt-in): carry the running sum ((k1 + 2*k2) + 2*k3) instead of k2 and k3
                    if constexpr (S1 == 1) st_k(k, k1s, m3);
                    if constexpr (S1 == 2 || S1 == 3) {
                        Row<T, V> acc;
                        ld_k(acc, S1 == 2 ? k1s : k2s, m3);
#pragma unroll
                        for (int f = 0; f < 3; ++f)
#pragma unroll
                            for (int e = 0; e < V; ++e) acc.f[f][e] = Ar<T>::add(acc.f[f][e], Ar<T>::mul(T(2), k.f[f][e]));
                        st_k(acc, S1 == 2 ? k2s : k3s, m3);
                    }
                }
            }
        } else {
            Row<T, V> o;
            if constexpr (NST == 4) {
                Row<T, V> k2, k3;
                ld_k(k3, k3s, m3);
                if constexpr (classical) {
                    // y + dt6 * ((((k1 + 2*k2) + 2*k3)) + k4), the sum so far is in the k3 ring
#pragma unroll
                    for (int f = 0; f < 3; ++f)
#pragma unroll
                        for (int e = 0; e < V; ++e)
                            o.f[f][e] = Ar<T>::add(yb.f[f][e], Ar<T>::mul(a.dt6, Ar<T>::add(k3.f[f][e], k.f[f][e])));
                } else {
                ld_k(k2, k2s, m3);
                // reference aliasing: "k1" reads k4 at the combine (weather_simulation.cpp:350-351, F5)
                if constexpr (kPacked) {
                    const F2 dt6 = f2_splat(FOLD ? -((float)a.dt6 * rs) : -(float)a.dt6);  // negated, see rk4_combine_pair
#pragma unroll
                    for (int f = 0; f < 3; ++f)
#pragma unroll
                        for (int p = 0; p < V; p += 2) {
                            const F2 k4 = F2{(float)k.f[f][p], (float)k.f[f][p + 1]};
                            const F2 yy = F2{(float)yb.f[f][p], (float)yb.f[f][p + 1]};
                            const F2 kk2 = F2{(float)k2.f[f][p], (float)k2.f[f][p + 1]};
                            const F2 kk3 = F2{(float)k3.f[f][p], (float)k3.f[f][p + 1]};
                            const F2 rr = FOLD ? rk4_combine_pair_folded(yy, dt6, k4, kk2, kk3, k4)
                                               : rk4_combine_pair(yy, dt6, k4, kk2, kk3, k4);
                            o.f[f][p] = rr.x; o.f[f][p + 1] = rr.y;
                        }
                } else {
#pragma unroll
                    for (int f = 0; f < 3; ++f)
#pragma unroll
                        for (int e = 0; e < V; ++e)
                            o.f[f][e] = rk4_combine<T>(yb.f[f][e], a.dt6, k.f[f][e], k2.f[f][e], k3.f[f][e], k.f[f][e]);
                }
                }
            } else if constexpr (kPacked) {
                const F2 dt2 = f2_splat(FOLD ? -((float)a.dt * rs) : -(float)a.dt);
#pragma unroll
                for (int f = 0; f < 3; ++f)
#pragma unroll
                    for (int p = 0; p < V; p += 2) {
                        const F2 rr = axpy_pair(F2{(float)yb.f[f][p], (float)yb.f[f][p + 1]}, dt2,
                                                F2{(float)k.f[f][p], (float)k.f[f][p + 1]});
                        o.f[f][p] = rr.x; o.f[f][p + 1] = rr.y;
                    }
            } else {
#pragma unroll
                for (int f = 0; f < 3; ++f)
#pragma unroll
                    for (int e = 0; e < V; ++e) o.f[f][e] = axpy<T>(yb.f[f][e], a.dt, k.f[f][e]);
            }
            if constexpr (STEADY) store_out_predicated(o, po);
            else if (!PURE || n >= 2 * NST) store_out(o, po);
        }
    }

    template <int PH, bool PURE, bool STEADY = false>
    __device__ __forceinline__ void iteration(Windows &Lv, Row<T, V> (&Yw)[3], unsigned &hz, int n, const Groups &G,
                                              const bool fast, T *(&po)[3]) const {
        if constexpr (kPackedDiv && !PURE) {  // (the boundary-free body marks the row where stage 1 loads it)
            Row<T, V> yn;
            ld_y<PH, 0>(yn, G);
            mark_row(hz, PH, yn);
        }
        stage<PH, 1, PURE, STEADY>(Lv, Yw, hz, n, G, fast, po);
        if constexpr (NST >= 2) stage<PH, 2, PURE, STEADY>(Lv, Yw, hz, n, G, fast, po);
        if constexpr (NST >= 4) {
            stage<PH, 3, PURE, STEADY>(Lv, Yw, hz, n, G, fast, po);
            stage<PH, 4, PURE, STEADY>(Lv, Yw, hz, n, G, fast, po);
        }
#pragma unroll
        for (int f = 0; f < 3; ++f) po[f] += g.pitch;
    }

    // After iteration 3q of triple q the rows of triple q-2 are dead: their group takes the rows of triple q+1.
    __device__ __forceinline__ void refill(int q, int grp_next) const {
        if (edge_strip) fence_proxy_async();  // lane 0 patched clamp columns in that group with generic stores
        __syncwarp();  // every lane has consumed its reads of the dead group
        if (elect_one()) issue_group(q + 1, grp_next);  // elect.sync: no divergence bookkeeping around the UBLKCPs
    }

    template <bool PURE>
    __device__ __forceinline__ void run() const {
        Windows Lv;
        Row<T, V> Yw[3];  // y rows n, n-1, n-2 of the boundary-free body
        // kPackedDiv: every row counts as suspect until it has been looked at
        unsigned hz = 0xfffu;
        Groups G;
        const T *base = ring + lane * V;
        G.a = base;                            // group 0: rows of triple 0
        G.b = base + 2 * 3 * L::ROW_ELEMS;     // group 2 (triple -1: never read)
        G.c = base + 1 * 3 * L::ROW_ELEMS;     // group 1 (triple -2: never read; next to be filled)
        const int ntriples = (niter + 2) / 3;
        int grp = 0;          // ring group of triple q (q mod 3) and the parity of its mbarrier (q / 3 mod 2)
        uint32_t parity = 0;
        // output row of iteration n is y0 - 2*NST + n: the lane's three store addresses advance one row per iteration
        const long long o0 = lvl_off + (long long)(y0 - 2 * NST) * g.pitch + c0;
        T *po[3] = {a.O.u + o0, a.O.v + o0, a.O.h + o0};
        // steady state (pipeline full, no domain edge within reach of any stage, interior strip) for n in [n_lo, n_hi):
        //   n >= 2*NST, n + 2 < niter, y0 - 2*NST + n > gmin, y0 - NST + n + 1 < gmax
        const int n_lo = max(2 * NST, gmin - y0 + 2 * NST + 1);
        const int n_hi = edge_strip ? 0 : min(niter - 2, gmax - y0 + NST - 1);
        const unsigned n_span = (unsigned)max(0, n_hi - n_lo);
        for (int q = 0; q < ntriples; ++q) {
            const int n = 3 * q;
            const int grp_next = grp == kGroups - 1 ? 0 : grp + 1;
            wait_group(q, grp, parity);
            const bool fast = PURE || (unsigned)(n - n_lo) < n_span;
            if (PURE && kSteadyBody && n > 2 * NST && n + 2 < niter) {
                // boundary-free body, pipeline full, complete triple: three straight-line iterations (the only
                // branches left per triple are the mbarrier wait, the producer election and the loop itself)
                iteration<0, PURE, PURE>(Lv, Yw, hz, n, G, true, po);
                refill(q, grp_next);
                iteration<1, PURE, PURE>(Lv, Yw, hz, n + 1, G, true, po);
                iteration<2, PURE, PURE>(Lv, Yw, hz, n + 2, G, true, po);
            } else {
            // general body: one code path for steady state and boundaries, the boundary tests are skipped by
            // uniform branches in the steady state
            iteration<0, PURE>(Lv, Yw, hz, n, G, fast, po);
            refill(q, grp_next);
            if ((!PURE && fast) || n + 1 < niter) iteration<1, PURE>(Lv, Yw, hz, n + 1, G, fast, po);
            else { po[0] += g.pitch; po[1] += g.pitch; po[2] += g.pitch; }
            if ((!PURE && fast) || n + 2 < niter) iteration<2, PURE>(Lv, Yw, hz, n + 2, G, fast, po);
            }
            if (grp_next == 0) parity ^= 1u;
            grp = grp_next;
            const T *t = G.c;  // rotate: triple q+1 lives where triple q-2 lived
            G.c = G.b;
            G.b = G.a;
            G.a = t;
        }
    }
};

// Step overlap (StepArgs::ovl_*). Written so that the warp provably never diverges: every lane polls the same
// counters and the loop condition is a warp vote -- with a lane-0-only spin (or a call) in front of the sweep the
// compiler can no longer prove convergence at the shuffles and emits a WARPSYNC.COLLECTIVE fallback for each of them
// (twice the code, 6 % slower kernel).
// enter: let the next step's grid start as soon as every CTA of this one is resident, then wait for the chunk rows of
// the previous step this CTA reads (c-1, c, c+1) and overwrites (c). A dependency that has not arrived after ~4 s is
// a protocol bug: the CTA gives up waiting and raises the error word (mapped host memory), which the host turns into
// an error at the next synchronisation -- wrong numbers reported loudly instead of a hung device.
__device__ __forceinline__ void overlap_enter(unsigned *done, unsigned steps_so_far, unsigned *err, int li, int nli) {
    if (!done) return;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!steps_so_far) return;
    const unsigned *slot = done + (size_t)blockIdx.z * nli + li;  // li: position of this chunk row in ROW order
    const unsigned target = steps_so_far * gridDim.x;  // every strip of a chunk row, every step so far
    const unsigned *lo = li > 0 ? slot - 1 : slot, *hi = li + 1 < nli ? slot + 1 : slot;
    unsigned tries = 0;
    while (!__all_sync(kFull, (int)(ld_relaxed_gpu(lo) - target) >= 0 && (int)(ld_relaxed_gpu(slot) - target) >= 0 &&
                                  (int)(ld_relaxed_gpu(hi) - target) >= 0)) {
        __nanosleep(100);
        if (++tries > 40000000u) {
            if (err) *(volatile unsigned *)err = 1u;
            break;
        }
    }
    fence_acq_rel_gpu();         // acquire: the rows those CTAs stored before bumping their counters ...
    fence_proxy_async_global();  // ... are read by the TMA engine (async proxy)
}
// leave: this CTA's rows are stored
__device__ __forceinline__ void overlap_leave(unsigned *done, int li, int nli) {
    if (!done) return;
    __syncwarp();
    fence_acq_rel_gpu();  // release
    if (threadIdx.x == 0) atomicAdd(done + (size_t)blockIdx.z * nli + li, 1u);
}
// Fused ghost exchange (PeerExchange): a band CTA waits until the neighbour's band of the previous step has landed in
// this rank's ghost rows (flag in local memory, bumped over NVLink; same convergent polling as above) ...
__device__ __forceinline__ void peer_wait(const unsigned *flag, unsigned target, unsigned *err) {
    unsigned tries = 0;
    while (!__all_sync(kFull, (int)(ld_relaxed_sys(flag) - target) >= 0)) {
        __nanosleep(200);
        if (++tries > 20000000u) {
            if (err) *(volatile unsigned *)err = 2u;
            break;
        }
    }
    fence_acq_rel_sys();
    fence_proxy_async_global();
}
// ... and tells the neighbour when its own rows have been stored over there
__device__ __forceinline__ void peer_signal(unsigned *flag) {
    __syncwarp();
    fence_acq_rel_sys();  // release at system scope: the peer stores of every lane (ordered by the warp barrier)
    if (threadIdx.x == 0) atomicAdd_system(flag, 1u);
}

template <typename T, int NST, int V, bool RECIP, int MINB, bool CL, bool FOLD, bool EXT>
__global__ void __launch_bounds__(32, MINB)
    step_tma_kernel(const Geometry<T> g, const Physics<T> ph, const StepArgs<T> a, const int rows_per_chunk,
                    const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_v,
                    const __grid_constant__ CUtensorMap tm_h) {
    using L = Layout<T, NST, V>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int strip = blockIdx.x;
    int y0, y1, li = (int)blockIdx.y, role = 0;  // role: 0 interior, 1 top band, 2 bottom band (PeerExchange)
    const int nli = (int)gridDim.y;
    if (a.px.band > 0) {
        // fused ghost exchange: chunk rows are [top band | bottom band | interior chunks]; li = position in row order
        if (blockIdx.y == 0) { y0 = 0; y1 = a.px.band; role = 1; li = 0; }
        else if (blockIdx.y == 1) { y0 = g.H - a.px.band; y1 = g.H; role = 2; li = nli - 1; }
        else { y0 = a.px.band + ((int)blockIdx.y - 2) * rows_per_chunk; y1 = min(y0 + rows_per_chunk, g.H - a.px.band); li = (int)blockIdx.y - 1; }
    } else {
        // blockIdx.y enumerates the chunks of the first row range, then those of the optional second one
        const int nchunks1 = (a.y_end - a.y_begin + rows_per_chunk - 1) / rows_per_chunk;
        const bool second = (int)blockIdx.y >= nchunks1;
        const int cy = second ? (int)blockIdx.y - nchunks1 : (int)blockIdx.y;
        y0 = (second ? a.y_begin2 : a.y_begin) + cy * rows_per_chunk;
        y1 = min(y0 + rows_per_chunk, second ? a.y_end2 : a.y_end);
    }
    if (strip * L::OUTW >= g.W || y0 >= y1) return;  // block-uniform (never taken when a.ovl_done is set: exact grid)
    overlap_enter(a.ovl_done, a.ovl_target, a.ovl_err, li, nli);
    const bool peer_up = role == 1 && a.px.up.u != nullptr, peer_dn = role == 2 && a.px.dn.u != nullptr;
    if ((peer_up || peer_dn) && a.px.target)
        peer_wait(a.px.wait + (peer_dn ? 1 : 0), a.px.target * gridDim.x * gridDim.z, a.ovl_err);

    SweepT<T, NST, V, RECIP, CL, FOLD, EXT> sw(g, ph, a);
    if constexpr (std::is_same<T, float>::value) {
        sw.ph2 = physics_f2(ph);
        sw.phf = physics_fold(ph);
        sw.phe = physics_ext_f2(ph);
        sw.phd = physics_div_f2(ph);
        sw.div_unproven = ph.rdx == 0.0f;
        sw.rs = ph.rdx;
    }
    sw.tm_u = &tm_u;
    sw.tm_v = &tm_v;
    sw.tm_h = &tm_h;
    sw.lane = lane;
    sw.xs = strip * L::OUTW - L::HX;
    sw.c0 = sw.xs + lane * V;
    sw.fix_left = sw.xs < 0;
    sw.fix_right = sw.xs + L::SW > g.W;
    sw.edge_strip = sw.fix_left || sw.fix_right;
    sw.y0 = y0;
    sw.niter = (y1 - y0) + 2 * NST;
    sw.gmin = -g.row0;
    sw.gmax = g.Hglobal - g.row0;
    sw.out_lo = strip * L::OUTW;
    sw.out_hi = min(sw.out_lo + L::OUTW, g.W);
    sw.lvl_off = (long long)blockIdx.z * g.level_stride;
    sw.ring = reinterpret_cast<T *>(smem_raw);
    sw.k2s = sw.ring + kRing * L::ROW_ELEMS + lane * V;
    sw.k3s = sw.k2s + 3 * L::ROW_ELEMS;
    sw.k1s = sw.k3s + 3 * L::ROW_ELEMS;
    sw.ring_u32 = smem_u32(sw.ring);
    sw.bar_u32 = smem_u32(smem_raw + (kRing + L::k_rows(NST == 4 && CL)) * L::ROW_BYTES);
    bool all_in = true;
#pragma unroll
    for (int e = 0; e < V; ++e) {
        sw.st_e[e] = (sw.c0 + e >= sw.out_lo) && (sw.c0 + e < sw.out_hi);
        all_in = all_in && sw.st_e[e];
    }
    sw.st_vec = all_in;
    sw.ragged = V > 1 && ((sw.out_hi - sw.out_lo) % V != 0);
    // peer push: byte offsets from this rank's output addresses to the same cell of the neighbour's ghost rows
    // (rows [0, band) -> its rows [H_up, H_up + band); rows [H - band, H) -> its rows [-band, 0)), parked in shared memory
    sw.push = peer_up || peer_dn;
    long long *pdel = reinterpret_cast<long long *>(smem_raw + (kRing + L::k_rows(NST == 4 && CL)) * L::ROW_BYTES + kGroups * 8);
    sw.pdel = pdel;
    if (sw.push && lane < 3) {
        const Planes3<T> &P = peer_up ? a.px.up : a.px.dn;
        const int Hp = peer_up ? a.px.up_H : a.px.dn_H;
        const T *mine = lane == 0 ? a.O.u : lane == 1 ? a.O.v : a.O.h;
        const T *theirs = lane == 0 ? P.u : lane == 1 ? P.v : P.h;
        const long long lstride_p = (long long)(Hp + 2 * kLeadRows) * g.pitch;
        const long long shift = (long long)blockIdx.z * (lstride_p - g.level_stride) +
                                (long long)(peer_up ? Hp : -g.H) * g.pitch;
        pdel[lane] = (reinterpret_cast<const char *>(theirs) - reinterpret_cast<const char *>(mine)) +
                     shift * (long long)sizeof(T);
    }

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kGroups; ++s) mbar_init(sw.bar_u32 + 8u * s, 1);
        fence_mbar_init();
        fence_proxy_async();
        sw.issue_group(0, 0);
    }
    __syncwarp();
    // RK4: CTAs away from every domain edge (the vast majority) run a loop body without any boundary handling
    // (0.578 -> 0.564 ms/step at 8192^2; no gain for the 1- and 2-stage kernels, which keep the single body)
    if constexpr (NST == 4) {
        const bool pure = !sw.edge_strip && y0 - NST >= sw.gmin && y1 + NST <= sw.gmax && !sw.push;
        if (pure) {
            sw.template run<true>();
            overlap_leave(a.ovl_done, li, nli);  // this CTA's rows are stored: publish them to the next step's CTAs
            return;
        }
    }
    sw.template run<false>();
    if (sw.push) peer_signal(peer_up ? a.px.sig_up : a.px.sig_dn);
    overlap_leave(a.ovl_done, li, nli);
}

int env_int(const char *name, int dflt) {
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

// Rows each warp sweeps. Every chunk pays 2*NST pipeline-fill iterations, so tall chunks win wherever the step
// overlap hides the drain of the last CTAs (B200, 8192^2, profiles/r2/README.md): fp32 RK4 0.479 ms at 64 rows,
// 0.456 at 160-192, 0.453 at 256; fp64 RK2 (16384^2) 2.23 / 2.09 / 2.11 ms at 64 / 128 / 192. The HBM-bound fp32
// Euler kernel wants short chunks (0.257 / 0.260 / 0.282 / 0.313 ms at 64 / 88 / 128 / 192). Small grids get
// shorter chunks so that there are at least about two CTAs per warp slot of the GPU.
int rows_per_chunk_for(int nstages, bool f32, int W, int H, int L, bool overlapped) {
    static const int forced = env_int("WSB_FUSED_ROWS_PER_CHUNK", 0);
    if (forced > 0) return forced;
    // without the step overlap the drain of tall chunks is exposed at every step (0.60 vs 0.51 ms at 256 vs 64 rows)
    const int cap = !overlapped ? 64 : f32 ? (nstages == 4 ? 256 : 64) : (nstages == 2 ? 128 : 64);
    const int cols = f32 ? (nstages == 2 ? 120 : 56) : (nstages == 4 ? 24 : 60);  // output columns per strip
    const long long strips = (W + cols - 1) / cols;
    long long rpc = strips * (long long)H * L / (2 * 148 * 16);
    rpc = (rpc + 15) / 16 * 16;
    return (int)std::max<long long>(32, std::min<long long>(cap, rpc));
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link dependency on libcuda)
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// One y_n plane as a 3-D tensor (x, row incl. the lead rows, level); box = one strip width x 3 rows.
template <typename T>
cudaError_t make_plane_map(CUtensorMap *map, const Geometry<T> &g, const T *origin, int strip_width) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    void *base = const_cast<T *>(origin) - (size_t)kLeadRows * g.pitch;
    const cuuint64_t dims[3] = {(cuuint64_t)g.W, (cuuint64_t)(g.H + 2 * kLeadRows), (cuuint64_t)g.L};
    const cuuint64_t strides[2] = {(cuuint64_t)g.pitch * sizeof(T), (cuuint64_t)g.level_stride * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)strip_width, 3u, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(map, sizeof(T) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims,
                           strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);  // NONE / 256B: same within noise
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// the extended-physics kernels exist for the default cells-per-lane choice of each (type, stage count)
template <typename T, int NST, int V>
constexpr bool kExtInstantiated = std::is_same<T, float>::value ? (NST == 2 ? V == 4 : V == 2) : (NST == 4 ? V == 1 : V == 2);

template <typename T, int NST, int V, int MINB>
cudaError_t launch_impl(const Geometry<T> &g, const Physics<T> &ph, const StepArgs<T> &a, cudaStream_t st) {
    using L = Layout<T, NST, V>;
    const int rows = a.y_end - a.y_begin, rows2 = a.y_end2 - a.y_begin2;
    if (rows <= 0) return cudaSuccess;
    const int strips = (g.W + L::OUTW - 1) / L::OUTW;
    const int rpc = a.rows_per_chunk > 0 ? a.rows_per_chunk
                                         : rows_per_chunk_for(NST, std::is_same<T, float>::value, g.W, g.H, g.L, false);
    int chunks = (rows + rpc - 1) / rpc + (rows2 > 0 ? (rows2 + rpc - 1) / rpc : 0);
    if (a.px.band > 0)  // fused ghost exchange: [top band | bottom band | interior chunks] over the whole slab
        chunks = 2 + (g.H - 2 * a.px.band + rpc - 1) / rpc;
    const dim3 grid(strips, chunks, g.L);
    CUtensorMap tu, tv, th;
    if (cudaError_t e = make_plane_map<T>(&tu, g, a.Y.u, L::SW)) return e;
    if (cudaError_t e = make_plane_map<T>(&tv, g, a.Y.v, L::SW)) return e;
    if (cudaError_t e = make_plane_map<T>(&th, g, a.Y.h, L::SW)) return e;
    // step overlap needs ONE row range (the counters are indexed by blockIdx.y); ovl_chain adds the programmatic
    // stream serialization attribute: this grid may start while the previous step's last CTAs still run
    StepArgs<T> aa = a;
    if (rows2 > 0) aa.ovl_done = nullptr;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(32, 1, 1);
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = (aa.ovl_done && aa.ovl_chain) ? 1 : 0;
    auto launch = [&](auto kernel, size_t smem) {
        cfg.dynamicSmemBytes = smem;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, g, ph, aa, rpc, tu, tv, th);
        if (e != cudaSuccess && cfg.numAttrs) {
            // a context that refuses the programmatic attribute: the same launch, stream-serialised (the chunk-row
            // counters are satisfied trivially then)
            cudaGetLastError();
            cfg.numAttrs = 0;
            e = cudaLaunchKernelEx(&cfg, kernel, g, ph, aa, rpc, tu, tv, th);
        }
        return e;
    };
    if (ph.ext) {
        // extended physics (beta plane, viscosity, diffusivity): power-of-two spacing only on this path (the per-stage
        // path takes any spacing); instantiated for the default strip widths
        if constexpr (kExtInstantiated<T, NST, V>) {
            if (!ph.recip) return cudaErrorNotSupported;
            if constexpr (NST == 4) {
                if (a.classical) return launch(step_tma_kernel<T, NST, V, true, MINB, true, false, true>, L::smem_bytes(true));
            }
            return launch(step_tma_kernel<T, NST, V, true, MINB, false, false, true>, L::smem_bytes(false));
        } else {
            return cudaErrorNotSupported;
        }
    }
    if constexpr (NST == 4) {
        if (a.classical)  // textbook RK4 opt-in: separate instantiation, one more 3-row ring in shared memory
            return ph.recip ? launch(step_tma_kernel<T, NST, V, true, MINB, true, false, false>, L::smem_bytes(true))
                            : launch(step_tma_kernel<T, NST, V, false, MINB, true, false, false>, L::smem_bytes(true));
    }
    if constexpr (std::is_same<T, float>::value && V % 2 == 0) {
        // folded arithmetic (opt-in): one spacing, exact reciprocal
        if (a.fold && ph.recip && ph.rdx == ph.rdy)
            return launch(step_tma_kernel<T, NST, V, true, MINB, false, true, false>, L::smem_bytes(false));
    }
    return ph.recip ? launch(step_tma_kernel<T, NST, V, true, MINB, false, false, false>, L::smem_bytes(false))
                    : launch(step_tma_kernel<T, NST, V, false, MINB, false, false, false>, L::smem_bytes(false));
}

}  // namespace

// strips (CTAs per chunk row and level) of a launch: what one fused step adds to a neighbour's flag per level
int step_tma_strips(int nstages, int dtype, int W) {
    const bool f32 = dtype == WSB_F32;
    int cols;
    if (f32) cols = nstages == 1 ? (env_int("WSB_CELLS_PER_LANE", 2) == 4 ? Layout<float, 1, 4>::OUTW : Layout<float, 1, 2>::OUTW)
                  : nstages == 2 ? (env_int("WSB_CELLS_PER_LANE", 4) == 2 ? Layout<float, 2, 2>::OUTW : Layout<float, 2, 4>::OUTW)
                                 : Layout<float, 4, 2>::OUTW;
    else cols = nstages == 1 ? (env_int("WSB_CELLS_PER_LANE", 2) == 1 ? Layout<double, 1, 1>::OUTW : Layout<double, 1, 2>::OUTW)
                : nstages == 2 ? (env_int("WSB_CELLS_PER_LANE", 2) == 1 ? Layout<double, 2, 1>::OUTW : Layout<double, 2, 2>::OUTW)
                               : Layout<double, 4, 1>::OUTW;
    return (W + cols - 1) / cols;
}

// rows per chunk of a full-height launch: the host sizes the step-overlap counters with it (one per chunk row)
int step_tma_rows_per_chunk(int nstages, int dtype, int W, int H, int L, bool overlapped) {
    return rows_per_chunk_for(nstages, dtype == WSB_F32, W, H, L, overlapped);
}

namespace {

}  // namespace

bool step_tma_supported(int nstages, int dtype) {
    (void)dtype;  // every stage count has an fp32 and an fp64 instantiation (cells per lane differ, launch_step_tma)
    return nstages == 1 || nstages == 2 || nstages == 4;
}

template <>
cudaError_t launch_step_tma<float>(const Geometry<float> &g, const Physics<float> &ph, const StepArgs<float> &a,
                                   int nstages, cudaStream_t st) {
    switch (nstages) {
        // cells per lane (strip = 32 x cells columns), measured on B200 (profiles/r1/README.md): Euler is best at 2
        // (0.271 vs 0.276 ms at 8192^2), RK2 at 4 (1.79 vs 1.91 ms at 2048^2 x 64); WSB_CELLS_PER_LANE overrides for A/B
        case 1: return env_int("WSB_CELLS_PER_LANE", 2) == 4 ? launch_impl<float, 1, 4, 16>(g, ph, a, st)
                                                              : launch_impl<float, 1, 2, 20>(g, ph, a, st);
        case 2: return env_int("WSB_CELLS_PER_LANE", 4) == 2 ? launch_impl<float, 2, 2, 16>(g, ph, a, st)
                                                              : launch_impl<float, 2, 4, 12>(g, ph, a, st);
        // RK4 at 4 cells per lane: 168 registers, 9 warps per SM, 0.77 ms vs 0.623 ms -> not instantiated
        // 16 warps per SM = 128 registers: the boundary-free body carries three y rows in registers (120 used);
        // at 18 warps (96 registers) that spills (0.70 ms), without the carried rows 18 and 16 are equal
        case 4: return launch_impl<float, 4, 2, 16>(g, ph, a, st);
        default: return cudaErrorNotSupported;
    }
}

template <>
cudaError_t launch_step_tma<double>(const Geometry<double> &g, const Physics<double> &ph, const StepArgs<double> &a,
                                    int nstages, cudaStream_t st) {
    switch (nstages) {
        // two cells per lane: Barotropic 16384^2 Euler 2.11 ms (0.93 of HBM peak) vs 2.54 ms with one
        case 1: return env_int("WSB_CELLS_PER_LANE", 2) == 1 ? launch_impl<double, 1, 1, 20>(g, ph, a, st)
                                                              : launch_impl<double, 1, 2, 16>(g, ph, a, st);
        case 2: return env_int("WSB_CELLS_PER_LANE", 2) == 1 ? launch_impl<double, 2, 1, 16>(g, ph, a, st)
                                                              : launch_impl<double, 2, 2, 12>(g, ph, a, st);
        case 4: return launch_impl<double, 4, 1, 12>(g, ph, a, st);
        default: return cudaErrorNotSupported;
    }
}

}  // namespace wsb
