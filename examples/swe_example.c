/* Plain-C client of libweather_b200.so: the reference's benchmark loop (examples/weather_sim_example.cpp:
 * configure, initialise, run, report MCUPS = W*H*steps / (compute_time_ms * 1000), :113-116) through the
 * C-ABI only. Build: make -C nvidia-jetson-workload_b200 example   Run: ./nvidia-jetson-workload_b200/build/swe_example [W H steps] */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "weather_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int st_ = (call);                                                        \
        if (st_ != WSB_OK) {                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, st_, wsb_last_error()); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char **argv) {
    int W = argc > 1 ? atoi(argv[1]) : 2048, H = argc > 2 ? atoi(argv[2]) : 2048, steps = argc > 3 ? atoi(argv[3]) : 200;
    wsb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.struct_size = sizeof(cfg);
    cfg.model = WSB_MODEL_SHALLOW_WATER;
    cfg.integration_method = WSB_INT_RUNGE_KUTTA_4;
    cfg.grid_width = W;
    cfg.grid_height = H;
    cfg.num_levels = 1;
    cfg.dx = cfg.dy = 1.0;
    cfg.dt = 0.01;
    cfg.gravity = 9.81;
    cfg.max_time = 1.0e9;
    cfg.dtype = WSB_F32;
    cfg.nranks = 1;

    wsb_sim *sim = NULL;
    CHECK(wsb_sim_create(&cfg, &sim));
    wsb_grid *grid = wsb_sim_current_grid(sim);
    /* the reference's jet_stream initial condition, evaluated by the library */
    CHECK(wsb_ic_apply(grid, "jet_stream", NULL, 0, 0, NULL));
    CHECK(wsb_sim_run(sim, steps, NULL));

    wsb_metrics m;
    CHECK(wsb_sim_get_metrics(sim, &m));
    double mass = 0, energy = 0;
    CHECK(wsb_sim_mass_energy(sim, &mass, &energy));
    float *h = (float *)malloc((size_t)W * H * sizeof(float));
    CHECK(wsb_grid_get_field(grid, WSB_FIELD_HEIGHT, h, WSB_F32, 1, H, W));
    printf("kernel path      : %s\n", wsb_sim_kernel_name(sim));
    printf("grid             : %d x %d, %d RK4 steps, t = %.4f\n", W, H, m.num_steps, wsb_sim_get_time(sim));
    printf("compute time     : %.3f ms (%.4f ms/step)\n", m.compute_time_ms, m.compute_time_ms / m.num_steps);
    printf("MCUPS            : %.1f\n", (double)W * H * m.num_steps / (m.compute_time_ms * 1000.0));
    printf("mass, energy     : %.9e %.9e   h[H/2][W/2] = %.6f\n", mass, energy, h[(size_t)(H / 2) * W + W / 2]);
    free(h);
    wsb_sim_destroy(sim);
    return 0;
}
