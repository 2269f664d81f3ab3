/* tools/e2e_loop.c — the e2e leg of bench.py as a compiled caller sees it: a closed loop around the C-ABI entry point
 * mpcb_mppi_compute with HOST buffers (x[4], u_in[H] -> u_out[H] + info; u_out is the next u_in), timed on the host.
 * The reference's users are compiled Rust programs calling Mppi::compute in exactly such a loop
 * (examples/mppi4.rs:41-68); a Python for-loop around the same call adds ~4 us of interpreter/ctypes time per step.
 * Built by __graft_entry__.build() into tools/libmpcb_e2e.so and loaded by bench.py with ctypes. */
#include <string.h>
#include <time.h>

#include "mpc_b200.h"

/* Runs n steps; returns the elapsed seconds, or -1.0 with *status set if a step fails. */
double mpcb_e2e_loop(mpcb_mppi* h, const double* x, double* u, double* out, int horizon, int n, int* status) {
    struct timespec a, b;
    mpcb_mppi_info info;
    *status = 0;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int i = 0; i < n; ++i) {
        const mpcb_status st = mpcb_mppi_compute(h, x, u, out, &info);
        if (st != MPCB_OK) {
            *status = (int)st;
            return -1.0;
        }
        memcpy(u, out, (size_t)horizon * sizeof(double));
    }
    clock_gettime(CLOCK_MONOTONIC, &b);
    return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
