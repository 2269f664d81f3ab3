// ukf_stream_kernel.cuh — fused predict+update of the six-state library UKF (src/ukf2.rs:44-136) in its fast
// arithmetic, restructured so that nothing of size n x (2n+1) lives in registers.
//
// Why a second kernel (ukf_kernel.cuh stays the general one: any n, o, mode, exact arithmetic): with n = 6, o = 5 the
// general fused kernel keeps the 78 propagated sigma-point doubles AND the 65 measurement doubles in registers across
// the update — 255 registers, ~2 KB of spills per thread (ncu: long_scoreboard 1.7 cycles per issue from local memory),
// and 150 KB of straight-line code because every sigma point's fx / hx / accumulation is unrolled.  Here
//   * the 2n off-centre sigma points live in shared memory, 72 doubles per thread ([row][128 threads]: a warp's access
//     is two conflict-free wavefronts; the centre point stays in registers), so fx, hx and the moment sums are REAL
//     loops over the n pairs x +- L_j — the loop body exists once in the code, and the two points of a pair are two
//     independent dependency chains for the scheduler — and
//   * both unscented transforms are one-pass and SHIFTED: with y_0 the image of the centre point, d_i = y_i - y_0,
//     s = sum_{i>=1} d_i and S = sum_{i>=1} d_i d_i^T,
//         mean = y_0 + w_i s                                  (sum of the mean weights is 1)
//         cov  = w_i S + w_i^2 ((wc_0 - wm_0) - 1) s s^T + noise
//         Pxz  = w_i sum_{i>=1} e_i dz_i^T - (wc_0 - wm_0) e_0 (w_i sz)^T,   e_i = y_i - mean
//     — algebraically the reference's two-pass sums (src/ukf2.rs:96-110, :60-66) with the weight shared by i >= 1
//     factored out; numerically better conditioned, because only differences of nearby points are squared and the
//     +-1e6 weights never multiply a full-size state.  No array of measurement sigma points exists at all.
// Everything else (eigen / Cholesky square root, fx / hx in the folded fast form, 5x5 inverse, gain, state and covariance
// update, sticky per-filter status, lower-triangle stores, tile walk) is the general kernel's code.  Differences from the
// reference order are rounding-level (per-step parity 1e-7, 1e-5 on the ill-conditioned NL6_UKF filter — the bars of
// the general fast kernels); `exact = 1`, the split predict / update calls and user models keep the general kernel.
#pragma once

#include "ukf_kernel.cuh"

namespace mpcb {

// Cyclic Jacobi of the symmetric matrix cC * P (lower triangle of P is the input) in the fast form of sym_eig_sqrt<N, true>
// (ukf_kernel.cuh: division-free rotation angle, stopping mass 3e-33 sum a_kk^2, threshold-skipped rotations) with the
// matrix held as its UPPER TRIANGLE ONLY — 21 instead of 36 doubles for n = 6 — and the result left as eigenvectors V and
// sq[j] = sqrt(|lambda_j|): the caller forms the columns V_j sq_j itself, so no third n x n array exists.  Same rotations
// in the same order as sym_eig_sqrt<N, true>, hence the same numbers.
template <int N>
__device__ __forceinline__ void sym_eig_packed(const double (&Pu)[N * (N + 1) / 2], double cC, double (&V)[N][N], double (&sq)[N]) {
    double a[N * (N + 1) / 2];
    auto A = [&](int i, int j) -> double& { return i <= j ? a[i * N - (i * (i - 1)) / 2 + (j - i)] : a[j * N - (j * (j - 1)) / 2 + (i - j)]; };
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
#pragma unroll
    for (int k = 0; k < N * (N + 1) / 2; ++k) a[k] = cC * Pu[k];
    for (int sweep = 0; sweep < 10; ++sweep) {
        double off = 0.0, dsum = 0.0;
#pragma unroll
        for (int p = 0; p < N - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < N; ++q) off += A(p, q) * A(p, q);
#pragma unroll
        for (int k = 0; k < N; ++k) dsum = fma(A(k, k), A(k, k), dsum);
        if (off <= 3.0e-33 * dsum) break;
        const double skip = (3.0e-33 / (N * (N - 1) / 2)) * dsum;
#pragma unroll
        for (int p = 0; p < N - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < N; ++q) {
                const double apq = A(p, q);
                if (apq * apq > skip) {
                    const double th = 0.5 * (A(q, q) - A(p, p));
                    const double h2 = fma(th, th, apq * apq);
                    double c, sn, t;
                    if (h2 > 1e-280 && h2 < 1e280) {
                        const double rh = rsqrt_f64_fast(h2);
                        const double c2 = fma(0.5, fabs(th) * rh, 0.5);
                        const double rc = rsqrt_f64_fast(c2);
                        c = c2 * rc;
                        const double sa = (0.5 * rc) * (fabs(apq) * rh);
                        sn = (th != 0.0 && ((th < 0.0) != (apq < 0.0))) ? -sa : sa;
                        t = sn * rc;
                    } else {
                        const double tau = (th == 0.0) ? 0.0 : th * __drcp_rn(apq);
                        const double at = fabs(tau);
                        if (at < 1e150) {
                            const double w = fma(tau, tau, 1.0);
                            t = copysign(__drcp_rn(at + w * rsqrt(w)), tau);
                        } else {
                            t = 0.5 * __drcp_rn(tau);
                        }
                        c = rsqrt(fma(t, t, 1.0));
                        sn = t * c;
                    }
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        if (k != p && k != q) {
                            const double akp = A(k, p), akq = A(k, q);
                            A(k, p) = c * akp - sn * akq;
                            A(k, q) = sn * akp + c * akq;
                        }
                    }
                    A(p, p) -= t * apq;
                    A(q, q) += t * apq;
                    A(p, q) = 0.0;
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = c * vkp - sn * vkq;
                        V[k][q] = sn * vkp + c * vkq;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) sq[j] = sqrt(fabs(A(j, j)));
}

// the 2n off-centre sigma points of every thread (the centre point stays in registers)
template <int N>
__host__ __device__ constexpr size_t ukf_stream_smem_bytes() { return (size_t)N * 2 * N * kUkfThreads * sizeof(double); }

template <int N, int O, int MODEL, int SQRT>
__global__ void __launch_bounds__(kUkfThreads, 2) ukf_stream_kernel(const __grid_constant__ UkfParams p) {
    pdl_entry();
    constexpr int M2 = 2 * N;  // columns j = x + L_j, N + j = x - L_j (library order, src/ukf2.rs:129-135, without the centre)
    extern __shared__ __align__(16) double s_sig[];  // [N * M2][kUkfThreads]: this thread's sigma points, row-major (r * M2 + c)
    const long long B = p.B;
    const int tid = threadIdx.x;
    const long long ntiles = (B + kUkfThreads - 1) / kUkfThreads;
    double* sg = s_sig + tid;
    auto SG = [&](int r, int c) -> double& { return sg[(r * M2 + c) * kUkfThreads]; };
    const double wi = p.wi;
    const double wdiff = p.wc0 - p.wm0;               // 1 - alpha^2 + beta
    const double wss = wi * wi * (wdiff - 1.0);       // weight of s s^T in the shifted covariance

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long b = (p.reverse ? ntiles - 1 - tile : tile) * kUkfThreads + tid;
        if (b >= B) continue;
        int st = p.status[b];
        // P is symmetric wherever this kernel touches it: one triangle in registers (21 doubles for n = 6), read from the
        // lower triangle in memory (the upper one may be stale, see lower_only)
        double x[N], Pu[N * (N + 1) / 2];
        auto PS = [&](int i, int j) -> double& { return i <= j ? Pu[i * N - (i * (i - 1)) / 2 + (j - i)] : Pu[j * N - (j * (j - 1)) / 2 + (i - j)]; };
#pragma unroll
        for (int r = 0; r < N; ++r) x[r] = p.x[(long long)r * B + b];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) PS(r, c) = p.P[(long long)(r * N + c) * B + b];
        bool p_full = false;  // x / P go back only once a predict has fully defined them (a failed filter keeps its state)
        double zc[O];
#pragma unroll
        for (int c = 0; c < O; ++c) zc[c] = p.z[(long long)c * B + b];

        for (int s = 0; s < p.steps && st == MPCB_OK; ++s) {
            double zn[O];
#pragma unroll
            for (int c = 0; c < O; ++c) zn[c] = (s + 1 < p.steps) ? p.z[((long long)(s + 1) * O + c) * B + b] : 0.0;
            const double u = p.has_u ? p.u[(long long)s * B + b] : p.u_scalar;

            // ---- predict: square root of C P, sigma points into shared memory ----
            if constexpr (SQRT == MPCB_SQRT_CHOLESKY) {
                double Lm[N][N];
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = 0; c < N; ++c) Lm[r][c] = p.cC * PS(r, c);
                if (!chol_lower<N, true>(Lm)) { st = MPCB_CHOLESKY_FAIL; break; }
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        SG(r, i) = x[r] + Lm[r][i];
                        SG(r, N + i) = x[r] - Lm[r][i];
                    }
            } else {
                // U sqrt(S) of src/ukf2.rs:121-124: column i = V_i sqrt(|lambda_i|)
                double V[N][N], sq[N];
                sym_eig_packed<N>(Pu, p.cC, V, sq);
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        const double l = V[r][i] * sq[i];
                        SG(r, i) = x[r] + l;
                        SG(r, N + i) = x[r] - l;
                    }
            }
            // ---- propagate through fx and take the shifted moments in the same pass: the centre point first, then the
            // points x +- L_j as a pair (two independent chains for the scheduler; d+ + d- is second-order small and is
            // summed before it meets s) ----
            double y0[N], sx[N], Sxx[N][N];
#pragma unroll
            for (int r = 0; r < N; ++r) {
                y0[r] = x[r];
                sx[r] = 0.0;
#pragma unroll
                for (int c = r; c < N; ++c) Sxx[r][c] = 0.0;
            }
            ukf_fx<MODEL, N, true>(p.mc, y0, u, p.dt);
#pragma unroll 1
            for (int j = 0; j < N; ++j) {
                double cp_[N], cm_[N];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    cp_[r] = SG(r, j);
                    cm_[r] = SG(r, N + j);
                }
                ukf_fx<MODEL, N, true>(p.mc, cp_, u, p.dt);
                ukf_fx<MODEL, N, true>(p.mc, cm_, u, p.dt);
                double dp[N], dm[N];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    SG(r, j) = cp_[r];
                    SG(r, N + j) = cm_[r];
                    dp[r] = cp_[r] - y0[r];
                    dm[r] = cm_[r] - y0[r];
                    sx[r] += dp[r] + dm[r];
                }
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = r; c < N; ++c) Sxx[r][c] = fma(dp[r], dp[c], fma(dm[r], dm[c], Sxx[r][c]));
            }
            double e0[N];  // y_0 - mean = -w_i s
#pragma unroll
            for (int r = 0; r < N; ++r) {
                e0[r] = -wi * sx[r];
                x[r] = y0[r] - e0[r];
            }
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = r; c < N; ++c) {
                    PS(r, c) = fma(wi, Sxx[r][c], fma(wss * sx[r], sx[c], p.Q[r * N + c]));
                }
            p_full = true;

            // ---- update: hx of every sigma point, shifted moments of z and of (x, z) in one pass ----
            double z0[O], sz[O], Szz[O][O], Sxz[N][O];
#pragma unroll
            for (int c = 0; c < O; ++c) {
                sz[c] = 0.0;
#pragma unroll
                for (int c2 = c; c2 < O; ++c2) Szz[c][c2] = 0.0;
            }
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < O; ++c) Sxz[r][c] = 0.0;
            {
                ukf_hx<MODEL, N, O, true>(p.mc, y0, z0);
#pragma unroll
                for (int c = 0; c < O; ++c) z0[c] = ((p.enable >> c) & 1u) ? z0[c] : 0.0;
            }
#pragma unroll 1
            for (int j = 0; j < N; ++j) {
                double cp_[N], cm_[N], zp_[O], zm_[O];
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    cp_[r] = SG(r, j);
                    cm_[r] = SG(r, N + j);
                }
                ukf_hx<MODEL, N, O, true>(p.mc, cp_, zp_);
                ukf_hx<MODEL, N, O, true>(p.mc, cm_, zm_);
                double dzp[O], dzm_[O];
#pragma unroll
                for (int c = 0; c < O; ++c) {
                    const bool on = ((p.enable >> c) & 1u) != 0u;
                    dzp[c] = (on ? zp_[c] : 0.0) - z0[c];
                    dzm_[c] = (on ? zm_[c] : 0.0) - z0[c];
                    sz[c] += dzp[c] + dzm_[c];
                }
#pragma unroll
                for (int c = 0; c < O; ++c)
#pragma unroll
                    for (int c2 = c; c2 < O; ++c2) Szz[c][c2] = fma(dzp[c], dzp[c2], fma(dzm_[c], dzm_[c2], Szz[c][c2]));
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    const double ep = cp_[r] - x[r], em = cm_[r] - x[r];
#pragma unroll
                    for (int c = 0; c < O; ++c) Sxz[r][c] = fma(ep, dzp[c], fma(em, dzm_[c], Sxz[r][c]));
                }
            }
            double zp[O], dzm[O], pz[O][O], pxz[N][O];
#pragma unroll
            for (int c = 0; c < O; ++c) {
                dzm[c] = wi * sz[c];  // mean - z_0
                zp[c] = z0[c] + dzm[c];
            }
#pragma unroll
            for (int c = 0; c < O; ++c)
#pragma unroll
                for (int c2 = c; c2 < O; ++c2) {
                    const double v = fma(wi, Szz[c][c2], fma(wss * sz[c], sz[c2], p.R[c * O + c2]));
                    pz[c][c2] = v;
                    pz[c2][c] = v;
                }
#pragma unroll
            for (int r = 0; r < N; ++r) {
                const double we = wdiff * e0[r];
#pragma unroll
                for (int c = 0; c < O; ++c) pxz[r][c] = fma(wi, Sxz[r][c], -we * dzm[c]);
            }
            // ---- gain K = Pxz Pz^-1 (src/ukf2.rs:67-68 forms the inverse): Gaussian elimination with partial pivoting on
            // [Pz | Pxz^T] — Pz is symmetric, so K^T = Pz^-1 Pxz^T — with the row swaps as predicated selects; a zero pivot
            // column is the reference's "Inverse fail".  No O x O inverse and no identity right-hand side exist. ----
            double kt[O][N];
#pragma unroll
            for (int c = 0; c < O; ++c)
#pragma unroll
                for (int r = 0; r < N; ++r) kt[c][r] = pxz[r][c];
            bool inv_ok = true;
#pragma unroll
            for (int kk = 0; kk < O; ++kk) {
                // (full-range loops with compile-time guards: the triangular forms were left partly rolled, with pz and
                // kt in local memory)
                int piv = kk;
                double best = fabs(pz[kk][kk]);
#pragma unroll
                for (int i = 0; i < O; ++i) {
                    if (i > kk) {
                        const double v = fabs(pz[i][kk]);
                        if (v > best) { best = v; piv = i; }
                    }
                }
                if (best == 0.0) inv_ok = false;
#pragma unroll
                for (int i = 0; i < O; ++i) {
                    if (i > kk) {
                        const bool sw = (piv == i);
#pragma unroll
                        for (int j = 0; j < O; ++j) {
                            if (j >= kk) {
                                const double a = pz[kk][j], bb = pz[i][j];
                                pz[kk][j] = sw ? bb : a;
                                pz[i][j] = sw ? a : bb;
                            }
                        }
#pragma unroll
                        for (int r = 0; r < N; ++r) {
                            const double a = kt[kk][r], bb = kt[i][r];
                            kt[kk][r] = sw ? bb : a;
                            kt[i][r] = sw ? a : bb;
                        }
                    }
                }
                const double ip = 1.0 / pz[kk][kk];
                pz[kk][kk] = ip;  // the diagonal keeps the reciprocal pivots for the back substitution
#pragma unroll
                for (int i = 0; i < O; ++i) {
                    if (i > kk) {
                        const double l = pz[i][kk] * ip;
#pragma unroll
                        for (int j = 0; j < O; ++j)
                            if (j > kk) pz[i][j] = fma(-l, pz[kk][j], pz[i][j]);
#pragma unroll
                        for (int r = 0; r < N; ++r) kt[i][r] = fma(-l, kt[kk][r], kt[i][r]);
                    }
                }
            }
            if (!inv_ok) { st = MPCB_INVERSE_FAIL; break; }
#pragma unroll
            for (int ii = 0; ii < O; ++ii) {
                constexpr int kLast = O - 1;
                const int i = kLast - ii;
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double acc = kt[i][r];
#pragma unroll
                    for (int j = 0; j < O; ++j)
                        if (j > i) acc = fma(-pz[i][j], kt[j][r], acc);
                    kt[i][r] = acc * pz[i][i];
                }
            }
            // x += K (z - zp);  P -= K Pz K^T = Pxz K^T (K Pz is Pxz itself), symmetric (src/ukf2.rs:69-73)
#pragma unroll
            for (int r = 0; r < N; ++r) {
                double acc = kt[0][r] * (zc[0] - zp[0]);
#pragma unroll
                for (int j = 1; j < O; ++j) acc = fma(kt[j][r], zc[j] - zp[j], acc);
                x[r] += acc;
            }
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = r; c < N; ++c) {
                    double acc = pxz[r][0] * kt[0][c];
#pragma unroll
                    for (int j = 1; j < O; ++j) acc = fma(pxz[r][j], kt[j][c], acc);
                    PS(r, c) -= acc;
                }
#pragma unroll
            for (int c = 0; c < O; ++c) zc[c] = zn[c];
        }

        if (p_full) {
#pragma unroll
            for (int r = 0; r < N; ++r) p.x[(long long)r * B + b] = x[r];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < N; ++c)
                    if (c <= r || !p.lower_only) p.P[(long long)(r * N + c) * B + b] = PS(r, c);
        }
        p.status[b] = st;
    }
}

// fused fast kernels of the six-state models in the streaming form (ukf_n6_stream.cu); nullptr = no such flavour
UkfKernelFn ukf_stream_kernel_n6(int model_id, int sqrt_mode, int order, size_t* smem_bytes);

}  // namespace mpcb
