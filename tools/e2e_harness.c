/* Developer probe: host wall-clock time of mpcb_mppi_compute called from plain C (no Python) — the cost a Rust
 * caller of the C ABI would see.  gcc -O2 -o tools/e2e_harness tools/e2e_harness.c -Iinclude -ldl
 *   tools/e2e_harness mpc_rs_b200/libmpc_b200.so [steps] */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mpc_b200.h"

static double now_us(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "mpc_rs_b200/libmpc_b200.so";
    int steps = argc > 2 ? atoi(argv[2]) : 2000;
    void* lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 1; }
    mpcb_status (*default_cfg)(int32_t, mpcb_mppi_cfg*) = dlsym(lib, "mpcb_mppi_default_cfg");
    mpcb_status (*create)(mpcb_mppi**, const mpcb_mppi_cfg*) = dlsym(lib, "mpcb_mppi_create");
    mpcb_status (*compute)(mpcb_mppi*, const double*, const double*, double*, mpcb_mppi_info*) = dlsym(lib, "mpcb_mppi_compute");
    void (*destroy)(mpcb_mppi*) = dlsym(lib, "mpcb_mppi_destroy");
    const char* (*errstr)(void) = dlsym(lib, "mpcb_last_error_string");
    mpcb_mppi_cfg cfg;
    default_cfg(MPCB_MODEL_NL, &cfg);
    cfg.horizon = 100; cfg.samples = 65536; cfg.model.dt = 0.008; cfg.precision = MPCB_F32;
    mpcb_mppi* h = NULL;
    if (create(&h, &cfg) != MPCB_OK) { fprintf(stderr, "create: %s\n", errstr()); return 1; }
    double x[4] = {0.5, 0.0, 0.1, 0.0}, u[100], out[100];
    memset(u, 0, sizeof(u));
    mpcb_mppi_info info;
    for (int i = 0; i < 50; ++i) { compute(h, x, u, out, &info); memcpy(u, out, sizeof(u)); }
    double t0 = now_us();
    for (int i = 0; i < steps; ++i) { compute(h, x, u, out, &info); memcpy(u, out, sizeof(u)); }
    double dt = (now_us() - t0) / steps;
    printf("C harness: mpcb_mppi_compute K=65536 H=100 f32: %.2f us/call (%.3e rollout-steps/s), status %d\n", dt, 6553600.0 / (dt * 1e-6), info.status);
    destroy(h);
    return 0;
}
