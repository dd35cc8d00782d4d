/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the weather-sim time-stepping path: a plain-C restatement of the reference's
 * algorithm (weather_simulation.cpp:117-560, weather_grid.cpp:57-121), instantiated for float (the
 * reference's scalar_t, weather_sim.hpp:24) and double (the reference has no fp64 path, SURVEY.md F8;
 * the double instantiation is the same code with T=double).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker or the timed CPU baseline. The product (libweather_b200.so and
 * the pyweather_sim shim) never links, loads or falls back to it.
 *
 * Parity pin: tests/test_oracle.py checks the float instantiation bit-for-bit against (a) the golden
 * vectors under tests/golden/ generated from the *real* reference (oracle/_ref/libws_ref.so, built by
 * oracle/build_ref.sh from /root/reference with six compile-only patches) and (b) that library itself
 * whenever it is present. The reference's own tests hold no golden values for this path (SURVEY.md
 * section 4), so reference-generated vectors are the pin.
 *
 * Build: oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp; contraction MUST stay off).
 */
#include <stdlib.h>
#include <string.h>
#include <stddef.h>

#if defined(__FP_FAST_FMA) && !defined(WSO_ALLOW_FMA_TARGET)
/* fine: the target has FMA, but -ffp-contract=off keeps the compiler from using it */
#endif

#define WSO_CAT_(a, b) a##b
#define WSO_CAT(a, b) WSO_CAT_(a, b)

#define T float
#define SFX(name) WSO_CAT(name, _f32)
#include "ws_oracle_body.inc"
#undef T
#undef SFX

#define T double
#define SFX(name) WSO_CAT(name, _f64)
#include "ws_oracle_body.inc"
#undef T
#undef SFX

const char *wso_build_flags(void) {
#ifdef WSO_BUILD_FLAGS
    return WSO_BUILD_FLAGS;
#else
    return "unknown";
#endif
}

/* OpenMP team size of the port's one parallel loop, and a runtime override (launchers such as torchrun
 * export OMP_NUM_THREADS=1 before this library is loaded). */
#ifdef _OPENMP
#include <omp.h>
void wso_set_omp_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int wso_omp_threads(void) {
    int n = 1;
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
    return n;
}
#else
void wso_set_omp_threads(int n) { (void)n; }
int wso_omp_threads(void) { return 1; }
#endif
