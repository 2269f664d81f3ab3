// mppi_short_kernel.cuh — the FP64 control step for SHORT horizons (H <= HMAX = 8: BASELINE config #4, the MPPI of
// examples/mppi4-non-liner-ukf.rs, 4096 controllers x 8192 samples x 8 steps).  Same passes as mppi_rollout_kernel
// (src/mppi.rs:38-91), other bookkeeping.
//
// With eight steps per sample the per-BATCH passes of the fused kernel — block max, weights, weighted sums over the
// v tile, five block barriers per 128 samples — are 35 % of its instructions and 47 % of its stall samples (ncu source
// page on the config #4 shape).  Here a thread keeps its samples' softmax IN REGISTERS across all its
// batches: the clamped controls v[0..H) of the sample it just rolled out, a running sum of weights and H running
// weighted sums, all against a running max that is uniform in the WARP.  The max only moves when a lane beats it
// (one vote per sample; about ln(n) times in n samples), and only then the warp pays a reduction, one exp and H + 1
// multiplies.  No tile, no shared-memory weights, no block barrier inside the sample loop; the block merges its
// threads' sums ONCE, then runs the same row / arrival / merge tail as every other kernel (mppi_block_tail).
//
// Semantics are those of the FP64 fused kernel: max over finite costs (lowest sample index on ties), w = exp((c - max) /
// lambda) for every sample, NaN / +inf costs poison the sums (src/mppi.rs:71-89), -inf costs weigh 0.
#pragma once

#include "mppi_kernel.cuh"

namespace mpcb {

constexpr int kShortHorizon = 8;
#ifndef MPCB_SHORT_MINB
#define MPCB_SHORT_MINB 4  // resident blocks of 128 threads per SM the register budget is set for
#endif

template <template <typename> class ModelT, int BLOCK, int NOISE, int HMAX>
__global__ void __launch_bounds__(BLOCK, MPCB_SHORT_MINB) mppi_short_kernel(const __grid_constant__ MppiParams p) {
    pdl_entry();
    using real = double;
    static_assert(HMAX % 4 == 0 && HMAX <= 16, "whole Philox blocks, registers");
    constexpr int NW = BLOCK / 32;
    constexpr bool kReplay = (NOISE == NOISE_REPLAY);
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int H = p.H;
    const int H4a = (H + 3) & ~3;
    // the layout of mppi_rollout_kernel without its tile (mppi_smem_bytes<double>(H, BLOCK, false))
    double* scratch = reinterpret_cast<double*>(smem_raw);
    double* part_d = scratch + kScratchDoubles;
    double* U_run = part_d + mppi_part_doubles(H, BLOCK);
    const int ndbl = (H + 8 + kScratchDoubles + (int)mppi_part_doubles(H, BLOCK) + 3) & ~3;
    real* su = reinterpret_cast<real*>(scratch + ndbl);  // [H4a] u_n
    real* sui = su + H4a;                                // [H4a] u_n * sigma^-2
    __shared__ double red_sum[NW][HMAX + 2];             // the warps' sum_w, sum_w * v[0..H)

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int c = blockIdx.x / p.chunks;
    const int chunk = blockIdx.x % p.chunks;
    MPCB_TS(0);

    ModelT<real> model;
    model.load(p.mc);
    constexpr int S = ModelStateDim<ModelT<real>>::value;
    real x0[S];
    if (p.use_inline) {
#pragma unroll
        for (int i = 0; i < S; ++i) x0[i] = p.xu_inline[i];
        for (int t = tid; t < H4a; t += BLOCK) {
            const double ut = t < H ? p.xu_inline[S + t] : 0.0;
            su[t] = ut;
            sui[t] = ut * p.inv_var;
        }
    } else {
#pragma unroll
        for (int i = 0; i < S; ++i) x0[i] = p.x[(long long)c * S + i];
        for (int t = tid; t < H4a; t += BLOCK) {
            const double ut = t < H ? p.u[(long long)c * H + t] : 0.0;
            su[t] = ut;
            sui[t] = ut * p.inv_var;
        }
    }
    const real lo = p.lo, hi = p.hi;
    const float neg2s2ln2 = (float)(-2.0 * p.std_dev * p.std_dev * 0.693147180559945309417);
    const double lambda = p.lambda;

    // the softmax of this thread's samples against the warp-uniform running max m_w
    double m_w = -CUDART_INF;
    long long arg_w = kNoArg;
    int nf_t = 0;
    double s_t = 0.0;
    double a_t[HMAX];
#pragma unroll
    for (int t = 0; t < HMAX; ++t) a_t[t] = 0.0;

    const long long w_begin = p.W * chunk / p.chunks;
    const long long w_end = p.W * (chunk + 1) / p.chunks;
    __syncthreads();

    for (long long gw = w_begin + wid; gw < w_end; gw += NW) {  // this warp's sample-warps
        const long long kl = gw * 32 + lane;
        const bool valid = kl < p.K_local;
        const long long kg = p.k_offset + kl;

        // ---- PASS 1: noise, clamp (src/mppi.rs:38-45, :51) ----
        real v[HMAX];
        if constexpr (kReplay) {
            const long long row = ((long long)c * p.K_global + kg) * H;
#pragma unroll
            for (int t = 0; t < HMAX; ++t) {
                real e = 0.0;
                if (valid && t < H)
                    e = p.eps_f64 ? reinterpret_cast<const double*>(p.eps)[row + t] : (real) reinterpret_cast<const float*>(p.eps)[row + t];
                v[t] = e;
            }
        } else {
            const unsigned int c0 = (unsigned int)(kg & 0xffffffffll);
            const unsigned int khi = (unsigned int)((kg >> 32) & 0xffff) << 16;
#pragma unroll
            for (int j = 0; j < HMAX / 4; ++j) {
                if (4 * j < H) {
                    const Philox4 r = philox4x32(c0, p.call_idx, (unsigned int)c + p.c_offset, (unsigned int)j | khi, p.seed_lo, p.seed_hi);
                    float z[4];
                    philox_normal4(r, neg2s2ln2, z);
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[4 * j + i] = (real)z[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[4 * j + i] = 0.0;
                }
            }
            if constexpr (NOISE == NOISE_GENERATE_DUMP) {
                if (valid) {
                    real* dump = reinterpret_cast<real*>(p.eps_dump) + ((long long)c * p.K_local + kl) * H;
#pragma unroll
                    for (int t = 0; t < HMAX; ++t)
                        if (t < H) dump[t] = v[t];
                }
            }
        }

        // ---- PASS 2: rollout and cost (:48-63) ----
        real x[S];
#pragma unroll
        for (int i = 0; i < S; ++i) x[i] = x0[i];
        double J = 0.0, CT = 0.0;
#pragma unroll
        for (int t0 = 0; t0 < HMAX; t0 += 4) {
            if (t0 < H) {
                real u4[4], ui4[4];
                lds4(su + t0, u4);
                lds4(sui + t0, ui4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (t0 + i < H) {
                        const real vt = clampr(u4[i] + v[t0 + i], lo, hi);  // f64::clamp, NaN stays NaN
                        v[t0 + i] = vt;
                        model.step(x, vt);
                        const real ct = model.cost(x);
                        J = J + ct;          // :57 c + cost(x_n)
                        CT += ui4[i] * vt;   // :60 (u*inv)*v, summed in order
                    }
                }
            }
        }
        const double ck = -J - CT;  // :61
        if (p.costs != nullptr && valid) p.costs[(long long)c * p.K_local + kl] = ck;

        // ---- PASS 3: does a lane beat the warp's running max?  (lowest index on ties: later sample-warps only win
        // with a strictly larger cost, and inside a sample-warp warp_max_minidx takes the lowest index) ----
        const bool fin = valid && finite_f64(ck);
        nf_t += fin ? 1 : 0;
        if (__any_sync(0xffffffffu, fin && ck > m_w)) {
            double bm;
            long long ba;
            warp_max_minidx(fin ? ck : -CUDART_INF, kg, fin, &bm, &ba);
            if (m_w != -CUDART_INF) {
                const double sc = exp((m_w - bm) / lambda);
                s_t *= sc;
#pragma unroll
                for (int t = 0; t < HMAX; ++t) a_t[t] *= sc;
            }
            m_w = bm;
            arg_w = ba;
        }

        // ---- PASS 4-6: weight and running sums (:71-86).  exp(a) is exactly 0 in f64 for a < -745.14 ----
        if (valid && ck != -CUDART_INF) {
            const double arg = (ck - m_w) / lambda;  // no finite cost yet: NaN / +inf, which poison like the reference
            const double w = (arg < -746.0) ? 0.0 : exp(arg);
            s_t += w;
#pragma unroll
            for (int t = 0; t < HMAX; ++t)
                if (t < H) a_t[t] += w * v[t];
        }
    }

    // ---- the block's max, then ONE merge of the threads' sums ----
    double* red_m = scratch;                        // [16]
    long long* red_a = (long long*)(scratch + 16);  // [16]
    int* red_n = (int*)(scratch + 32);              // [16]
    {
        const int bn = __reduce_add_sync(0xffffffffu, nf_t);
        if (lane == 0) { red_m[wid] = m_w; red_a[wid] = arg_w; red_n[wid] = bn; }
    }
    __syncthreads();
    double m_run;
    long long arg_run, nfin_run;
    {
        const bool has = lane < NW;
        const double wm = has ? red_m[lane] : -CUDART_INF;
        const long long wa = has ? red_a[lane] : kNoArg;
        const int wn = has ? red_n[lane] : 0;
        warp_max_minidx(wm, wa, has && wa != kNoArg, &m_run, &arg_run);
        nfin_run = __reduce_add_sync(0xffffffffu, wn);
    }
    if (m_w != -CUDART_INF) {  // a warp without a finite cost holds zeros or poison: both stay what they are
        const double sc = exp((m_w - m_run) / lambda);
        s_t *= sc;
#pragma unroll
        for (int t = 0; t < HMAX; ++t) a_t[t] *= sc;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s_t += shfl_down_f64(s_t, off);
#pragma unroll
        for (int t = 0; t < HMAX; ++t) a_t[t] += shfl_down_f64(a_t[t], off);
    }
    if (lane == 0) {
        red_sum[wid][0] = s_t;
#pragma unroll
        for (int t = 0; t < HMAX; ++t) red_sum[wid][1 + t] = a_t[t];
    }
    __syncthreads();
    double S_run = red_sum[0][0];
#pragma unroll
    for (int w = 1; w < NW; ++w) S_run += red_sum[w][0];
    for (int t = tid; t < H; t += BLOCK) {
        double acc = red_sum[0][1 + t];
#pragma unroll
        for (int w = 1; w < NW; ++w) acc += red_sum[w][1 + t];
        U_run[t] = acc;
    }
    __syncthreads();

    mppi_block_tail<BLOCK, false>(p, c, chunk, m_run, arg_run, S_run, nfin_run, U_run, scratch, part_d);
}

// kernel tables (mppi_f64_short.cu: reference order, no FMA; mppi_f64fast_short.cu: the folded forms)
MppiKernelFn mppi_kernel_f64_short(int model_id, int noise);
MppiKernelFn mppi_kernel_f64fast_short(int model_id, int noise);
constexpr int kShortBlock = 128;

#define MPCB_SHORT_TABLE(FN, ML, MNL, MNL6)                                                              \
    template <template <typename> class M>                                                               \
    static MppiKernelFn FN##_noise(int noise) {                                                          \
        switch (noise) {                                                                                 \
            case NOISE_GENERATE: return mppi_short_kernel<M, kShortBlock, NOISE_GENERATE, kShortHorizon>; \
            case NOISE_GENERATE_DUMP: return mppi_short_kernel<M, kShortBlock, NOISE_GENERATE_DUMP, kShortHorizon>; \
            case NOISE_REPLAY: return mppi_short_kernel<M, kShortBlock, NOISE_REPLAY, kShortHorizon>;    \
            default: return nullptr;                                                                     \
        }                                                                                                \
    }                                                                                                    \
    MppiKernelFn FN(int model_id, int noise) {                                                           \
        switch (model_id) {                                                                              \
            case MPCB_MODEL_L: return FN##_noise<ML>(noise);                                             \
            case MPCB_MODEL_NL: return FN##_noise<MNL>(noise);                                           \
            case MPCB_MODEL_NL6: return FN##_noise<MNL6>(noise);                                         \
            default: return nullptr;                                                                     \
        }                                                                                                \
    }

}  // namespace mpcb
