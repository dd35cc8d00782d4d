// placeholder translation unit; initial conditions are added below in a later commit
#include "wsb_internal.h"
extern "C" {
int wsb_ic_apply(wsb_grid *, const char *, const double *, int32_t, uint32_t, const char *) {
    return wsb::fail(WSB_ERR_RUNTIME, "initial conditions not built yet");
}
int wsb_ic_fill_host(const char *, const double *, int32_t, uint32_t, const char *, int32_t, int32_t, double, double,
                     float *, float *, float *, float *, float *, float *) {
    return wsb::fail(WSB_ERR_RUNTIME, "initial conditions not built yet");
}
}
