// Internal declarations shared by the translation units of libweather_b200.so.
// Nothing here is part of the C-ABI (include/weather_b200.h is).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>

#include "weather_b200.h"

namespace wsb {

// Ghost rows kept above and below every field so that a row slab can hold its neighbours' rows:
// 1 per RK stage, 4 for a whole fused RK4 step.
constexpr int kHaloRows = 4;
// One more never-written guard row on each side: the TMA-staged kernel copies whole 256-byte strip rows
// that may start up to 16 bytes before a row and end up to 240 bytes past it (columns outside the
// domain, whose values are never used), so the first/last ghost row needs slack inside the allocation.
constexpr int kGuardRows = 1;
constexpr int kLeadRows = kHaloRows + kGuardRows;  // rows between the allocation start and local row 0

// ---- error plumbing -------------------------------------------------------------------------
void set_last_error(const std::string &msg);
int fail(int status, const std::string &msg);
int cuda_fail(cudaError_t err, const char *what, const char *file, int line);

#define WSB_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) return ::wsb::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define WSB_TRY(call)                \
    do {                             \
        int _s = (call);             \
        if (_s != WSB_OK) return _s; \
    } while (0)

// ---- device-side view of a set of (u, v, h) planes -------------------------------------------
// All planes of a grid share one geometry: `pitch` elements per row, (rows + 2*kHaloRows) rows per
// level, pointer at (level 0, local row 0, x 0).
template <typename T>
struct Planes3 {
    T *u;
    T *v;
    T *h;
};

template <typename T>
struct Geometry {
    int W;                   // cells per row
    int H;                   // local rows (this rank)
    int L;                   // levels
    int pitch;               // elements between consecutive rows
    long long level_stride;  // elements between consecutive levels
    int row0;                // global index of local row 0
    int Hglobal;             // global rows (clamp boundary applies at 0 and Hglobal-1 only)
};

// Constants of one tendency evaluation (weather_simulation.cpp:486-489).
template <typename T>
struct Physics {
    T ddx, ddy;    // 2.0f*dx, 2.0f*dy                         (used when !recip)
    T rdx, rdy;    // recip: exact reciprocals (both spacings powers of two); else RN(1/ddx), RN(1/ddy) for the exact
                   // three-operation division, or 0 = plain IEEE division (wsb_arith.cuh)
    T g, f;
    int recip;     // 1: (a-b)*rdx is bit-identical to (a-b)/ddx for every input
    // Extended physics (WSB_PHYSICS_EXTENDED; NOT in the reference's compute code, see include/weather_b200.h): beta
    // plane f(y) = f + bdy*(y - yc), eddy viscosity nu on u, v, diffusivity kappa on h, 5-point Laplacian scaled by
    // idx2 = 1/dx^2, idy2 = 1/dy^2. All constants are rounded once on the host, in T, as oracle/ws_oracle_body.inc does.
    int ext;
    T bdy, yc, nu, kappa, idx2, idy2;
};

// One fused RK stage: k = tend(S); then either
//   UPDATE: O = Y + c*k              (optionally also store k into KS)
//   FINAL : O = Y + dt6*(((K1 + 2*KA) + 2*KB) + k)   with K1 := k when k1 planes are null
// (weather_simulation.cpp:186-198, 248-260, 380-451).
template <typename T>
struct StageArgs {
    Planes3<const T> S;   // stencil input state
    Planes3<const T> Y;   // base state y_n
    Planes3<T> O;         // output state
    Planes3<T> KS;        // where to store k (null = don't)
    Planes3<const T> KA;  // k2 (FINAL)
    Planes3<const T> KB;  // k3 (FINAL)
    Planes3<const T> K1;  // k1 (FINAL, classical mode only; null = reference aliasing, K1 := k4)
    T c;                  // stage coefficient, already rounded: dt or (0.5f*dt)
    T dt6;                // dt/6.0f
    int final_stage;
    int y_begin, y_end;   // local row range this launch covers
};

// Fused ghost exchange over peer memory (row slabs, TMA whole-step kernel): ONE launch per step and rank. The chunk
// rows are laid out as [top band | bottom band | interior chunks]; the CTAs of a band wait for the neighbour's flag
// before they read the ghost rows, store their output rows BOTH locally and -- through an NVLink peer mapping of the
// neighbour's planes (cudaIpc) -- into the neighbour's ghost rows, and bump the neighbour's flag. No NCCL kernel, no
// edge launch and no stream event is left in a step; the transfer rides under the interior sweep.
template <typename T>
struct PeerExchange {
    int band;                   // ghost depth (rows per band); 0 = exchange not fused into this launch
    Planes3<T> up, dn;          // the neighbours' y_{n+1} planes (origin pointers in THIS process), null at domain ends
    int up_H, dn_H;             // their local row counts (level stride and ghost-row position over there)
    unsigned *wait;             // local flags: [0] bumped by the upper neighbour's bottom band, [1] by the lower one's top
    unsigned *sig_up, *sig_dn;  // the flags to bump over there: up's [1], dn's [0]
    unsigned target;            // fused steps so far (flags count strips x levels per step)
};

// One RK stage of the tracer transport of the extended Primitive model (p, T, q carried by the stage's flow):
// O = Y + c * (-u*c_x - v*c_y + kappa*lap(c)) for each of the three tracers; u, v and the tracers C of the stage's
// input state (oracle/ws_oracle_body.inc, tracer_tendencies).
template <typename T>
struct TracerArgs {
    const T *u, *v;   // velocity of the stage's input state
    const T *C[3];    // tracers of the stage's input state (p, T, q)
    const T *Y[3];    // base state
    T *O[3];          // output state
    T c;              // stage coefficient, already rounded: dt or (0.5f*dt)
};

// Whole-step fused kernel arguments (all RK stages in one pass over the grid).
template <typename T>
struct StepArgs {
    Planes3<const T> Y;   // y_n
    Planes3<T> O;         // y_{n+1}
    T dt, half_dt, dt6;
    int classical;
    int fold;              // WSB_ARITH_FOLDED requested (used where the kernel and the spacing allow it)
    int y_begin, y_end;    // local OUTPUT row range this launch covers
    int y_begin2, y_end2;  // optional second range in the same launch (slab top + bottom edges); empty if equal
    int rows_per_chunk;    // rows each warp sweeps; 0 = library default
    // Step overlap (single range, TMA kernel): consecutive steps are launched with programmatic stream serialization,
    // so the CTAs of step n+1 fill the SMs while the last CTAs of step n drain. Data dependencies are tracked per
    // chunk row instead of per kernel: every CTA bumps ovl_done[level * nchunks + chunk] when its rows are stored,
    // and a CTA starts once the chunk rows c-1, c, c+1 of the PREVIOUS step have reached ovl_target x strips
    // (ovl_target = protocol steps so far). ovl_done == nullptr: plain launch. ovl_chain: launch with the
    // programmatic attribute.
    PeerExchange<T> px;
    unsigned *ovl_done;
    unsigned *ovl_err;     // raised (mapped host memory) by a CTA whose dependency did not arrive in ~4 s
    unsigned ovl_target;
    int ovl_chain;
};

// ---- kernel launchers (wsb_kernels.cu) -------------------------------------------------------
template <typename T>
cudaError_t launch_stage_direct(const Geometry<T> &g, const Physics<T> &ph, const StageArgs<T> &a, cudaStream_t st);

template <typename T>
cudaError_t launch_diagnostics(const Geometry<T> &g, const Physics<T> &ph, const T *u, const T *v, T *vort,
                               T *div, cudaStream_t st);

template <typename T>
cudaError_t launch_tracer_stage(const Geometry<T> &g, const Physics<T> &ph, const TracerArgs<T> &a, cudaStream_t st);

// O = Y + c*k for a constant k (the Primitive-equations T/p "tendencies", weather_simulation.cpp:201-214)
template <typename T>
cudaError_t launch_axpy_const(const Geometry<T> &g, const T *y, T *o, T c, T k, cudaStream_t st);

// both Primitive drifts (T and p) in one vectorised streaming pass
template <typename T>
cudaError_t launch_axpy_const2(const Geometry<T> &g, const T *ta, T *to, const T *pa, T *po, T c, T kt, T kp,
                               cudaStream_t st);

template <typename T>
cudaError_t launch_fill(const Geometry<T> &g, T *p, T value, bool with_halo, cudaStream_t st);

// field(x, y) = rowv[y] (+ colv[x], one fp32 addition) on every level: separable initial conditions (wsb_ic.cpp)
template <typename T>
cudaError_t launch_expand_separable(const Geometry<T> &g, T *p, const float *rowv, const float *colv, cudaStream_t st);

// per-block partial sums of mass and energy in double; partial[2*b], partial[2*b+1]
template <typename T>
cudaError_t launch_mass_energy(const Geometry<T> &g, const T *u, const T *v, const T *h, double gravity,
                               double *partial, int nblocks, cudaStream_t st);

// whole-step fused path (wsb_step_fused.cu). NSTAGES in {1, 2, 4}.
template <typename T>
cudaError_t launch_step_fused(const Geometry<T> &g, const Physics<T> &ph, const StepArgs<T> &a, int nstages,
                              cudaStream_t st);
bool step_fused_supported(int nstages, int dtype);

// whole-step fused path with TMA-staged y rows (wsb_step_tma.cu)
template <typename T>
cudaError_t launch_step_tma(const Geometry<T> &g, const Physics<T> &ph, const StepArgs<T> &a, int nstages,
                            cudaStream_t st);
bool step_tma_supported(int nstages, int dtype);
int step_tma_rows_per_chunk(int nstages, int dtype, int W, int H, int L, bool overlapped);
int step_tma_strips(int nstages, int dtype, int W);

// rows [y0, y0+nrows) of a field from a dense float host block (nrows x W), replicated to every level and
// converted to the grid's dtype (wsb_sim.cu; used by the blockwise initial conditions)
int grid_upload_rows(wsb_grid *grid, int field, const float *host_rows, int y0, int nrows);
// same, asynchronous: host_rows must be page-locked and stay untouched until *done_event (recorded on the grid's
// stream behind the copies) has completed
int grid_upload_rows_async(wsb_grid *grid, int field, const float *host_rows, int y0, int nrows);
int grid_record_event(wsb_grid *grid, cudaEvent_t ev);
// a whole field from host vectors of the reference's float values: field(x, y) = rowv[y] (+ colv[x]); rowv covers this
// grid's rows, colv (may be null) its columns
int grid_fill_separable(wsb_grid *grid, int field, const float *rowv, const float *colv);
// a whole field set to one value (all levels)
int grid_fill_uniform(wsb_grid *grid, int field, float value);
// where this grid sits in the global domain (row slab of a decomposed simulation; else 0 and its own height)
void grid_slab_position(const wsb_grid *grid, int *row0, int *global_height);

// ---- NCCL, loaded lazily with dlopen (wsb_nccl.cpp) -------------------------------------------
struct NcclApi;
int nccl_load(const NcclApi **api);
int nccl_get_unique_id(void *out128);
struct HaloComm;  // opaque: communicator + neighbour ranks
int halo_comm_create(int rank, int nranks, const void *unique_id, HaloComm **out);
void halo_comm_destroy(HaloComm *c);
int halo_align(HaloComm *c, cudaStream_t st);
// Peer mappings of one neighbour's memory for the fused ghost exchange: [0..2] u, v, h of plane set A, [3..5] of plane
// set B, [6] its two flag words.
constexpr int kPeerPointers = 7;
struct PeerLink {
    void *mapped[kPeerPointers] = {};  // allocation bases as mapped here
    bool owned[kPeerPointers] = {};    // this entry opened the mapping (several pointers may share an allocation)
    void *ptr[kPeerPointers] = {};     // the neighbour's pointers, usable in this process
    int H = 0;                         // its local row count
    bool present = false;
};
// Collective over all ranks of the communicator. *all_ok: every rank mapped every neighbour (else nothing is mapped).
int peer_setup(HaloComm *c, void *const local[kPeerPointers], int H, PeerLink *up, PeerLink *dn, bool *all_ok,
               cudaStream_t st);
void peer_close(PeerLink *link);
// Exchange `nrows` boundary rows of each of the `nplanes` planes with the up/down neighbours:
// sends local rows [0,nrows) up and [H-nrows,H) down; receives into ghost rows [-nrows,0) and [H,H+nrows),
// for each of the `levels` independent 2-D levels of every plane (one NCCL group for all of them).
int halo_exchange(HaloComm *c, void *const *planes, int nplanes, size_t elem_size, int pitch, int H, int nrows,
                  int levels, long long level_stride, cudaStream_t st);

}  // namespace wsb
