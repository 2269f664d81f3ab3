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

/* Device-resident closed loop: n back-to-back control steps enqueued from compiled code (ping-pong u buffers on the
 * device: u_out of step i is u_in of step i + 1).  bench.py brackets this call with CUDA events on the handle's stream
 * to get the kernel's launch-to-launch duration: a Python loop over the same call spends ~30 us per step in the
 * interpreter, which is as long as the kernel itself.  Returns the first failing status, or MPCB_OK. */
int mpcb_device_loop(mpcb_mppi* h, const double* d_x, double* d_u0, double* d_u1, int n) {
    for (int i = 0; i < n; ++i) {
        const mpcb_status st = mpcb_mppi_compute_device(h, d_x, (i & 1) ? d_u1 : d_u0, NULL, MPCB_DT_F32, (i & 1) ? d_u0 : d_u1);
        if (st != MPCB_OK) return (int)st;
    }
    return 0;
}

/* The same loop with a CUDA event between consecutive launches (on the handle's stream): ms[i] = time from event i to
 * event i + 1, i.e. the launch-to-launch interval of step i as the GPU saw it.  Returns 0 or the failing status; -1 if
 * CUDA events fail.  n <= 4096. */
#include <cuda_runtime_api.h>
int mpcb_device_loop_events(mpcb_mppi* h, void* stream, const double* d_x, double* d_u0, double* d_u1, int n, float* ms) {
    static cudaEvent_t ev[4097];
    static int made = 0;
    if (n > 4096) n = 4096;
    for (; made <= 4096; ++made)
        if (cudaEventCreate(&ev[made]) != cudaSuccess) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    for (int i = 0; i < n; ++i) {
        cudaEventRecord(ev[i], s);
        const mpcb_status st = mpcb_mppi_compute_device(h, d_x, (i & 1) ? d_u1 : d_u0, NULL, MPCB_DT_F32, (i & 1) ? d_u0 : d_u1);
        if (st != MPCB_OK) return (int)st;
    }
    cudaEventRecord(ev[n], s);
    if (cudaEventSynchronize(ev[n]) != cudaSuccess) return -1;
    for (int i = 0; i < n; ++i)
        if (cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]) != cudaSuccess) return -1;
    return 0;
}
