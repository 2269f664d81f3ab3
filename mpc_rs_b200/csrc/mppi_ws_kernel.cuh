// mppi_ws_kernel.cuh — warp-specialised FP32 MPPI control step (src/mppi.rs:38-91 in one launch).
//
// What bounds the rollout on B200 (tools/pipe_bench.cu, tools/loop_bench.cu): an SM sub-partition dispatches ONE warp
// instruction per clock at best, and many forms take longer — FFMA with three changing register operands 1.6 clocks,
// FFMA2/FMUL2 (two samples) 2, LOP3 with three registers 2, IMAD 2, IMAD.WIDE ~4, MUFU 8 on its own pipe.  A
// rollout-step costs ~54 dispatch clocks of dynamics + cost (packed, two samples per thread) and ~32 of noise
// (Philox4x32-7: 14 IMAD.WIDE + 14 LOP3 per four draws, Box-Muller on MUFU) — the kernel is DISPATCH-bound, not
// latency-bound, so what matters is (a) fewer and cheaper instructions per step and (b) every scheduler of the SM
// dispatching all the time.  (b) is what this kernel shape is for: BASELINE configs[1] leaves 442 samples = 13.8
// sample-warps per SM, i.e. 4/4/3/3 (or, packed, 2/2/2/1) rollout warps on the four schedulers — the fullest
// scheduler sets the time and the others idle 14 %.  The noise of a sample (src/mppi.rs:38-45) does not depend on the
// state, so it is produced by OTHER warps, from a shared work queue:
//
//   producer warps : take items (32 samples x one chunk of steps) from an atomic counter in shared memory, draw
//                    Philox + Box-Muller (or read the caller's replay noise), v = clamp(u_n + eps), write 16-byte
//                    cells (4 steps of one sample, or 2 steps of a sample pair) into the v tile; one mbarrier per
//                    chunk says "full".  A scheduler with fewer rollout warps gives its producers more issue slots, so
//                    they take more items: the four schedulers finish together.
//   consumer warps : wait for the chunk, one (two) LDS.128 per 4 steps, dynamics + stage cost + control term —
//                    scalar (SPT = 1) or two samples per thread in packed f32x2 arithmetic (SPT = 2)
//
// After the rollouts all warps share the softmax and the weighted sums over the tile (lanes over samples, 128-bit
// loads, zero-weight groups skipped), and the first TB threads run the same row / merge / exchange tail as
// mppi_rollout_kernel (mppi_block_tail).  Arithmetic is the FP32 fast form of models.cuh.
#pragma once

#include "mppi_kernel.cuh"

namespace mpcb {

// ---- mbarrier (shared::cta) ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned int)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"((unsigned int)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    const unsigned int addr = (unsigned int)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(addr), "r"(parity)
        : "memory");
}

constexpr int kWsMaxChunks = 128;  // chunks of >= 4 steps: H <= 512

// threads that run the tail (row store, arrival, merges): the merge code is written for <= 16 warps
__host__ __device__ constexpr int mppi_ws_tail_threads(int nt) { return nt >= 512 ? 512 : (nt >= 256 ? 256 : 128); }
// doubles of the partition-sum area: what the merges need (mppi_part_doubles) and the PASS 6 partition sums
// (nq <= threads / tile rows partitions of H4a doubles, bounded by 4 * threads + H4a)
__host__ __device__ inline size_t mppi_ws_part_doubles(int H, int nt) {
    const size_t a = mppi_part_doubles(H, mppi_ws_tail_threads(nt));
    const size_t b = 4 * (size_t)nt + (size_t)(((H + 3) >> 2) * 4);
    return a > b ? a : b;
}

// shared-memory bytes: doubles [scratch, partition sums, U_run, costs], floats [u_n, u_n/sigma^2, weights, poison flags],
// then the v tile (one 16-byte cell per 4-step group and sample)
__host__ __device__ inline size_t mppi_ws_smem_bytes(int H, int ncw, int npw, int spt) {
    const int nt = (ncw + npw) * 32, sb = ncw * 32 * spt;
    const int Hq = (H + 3) >> 2, H4a = Hq * 4;
    size_t dbl = (size_t)kScratchDoubles + mppi_ws_part_doubles(H, nt) + (size_t)(H + 8) + (size_t)sb;
    dbl = (dbl + 3) & ~(size_t)3;
    size_t bytes = dbl * sizeof(double) + ((size_t)2 * H4a + 2 * (size_t)sb) * sizeof(float);
    bytes = (bytes + 15) & ~(size_t)15;
    return bytes + (size_t)Hq * spt * (size_t)(ncw * 32 + 1) * 16;  // rows of U + 1 cells
}

// NCW consumer warps, NPW producer warps, SPT samples per consumer thread; SB = 32 * NCW * SPT samples per batch.
template <template <typename> class ModelT, int NCW, int NPW, int NOISE, int SPT>
__global__ void __launch_bounds__((NCW + NPW) * 32, 1) mppi_ws_kernel(const __grid_constant__ MppiParams p) {
    pdl_entry();
    constexpr int NT = (NCW + NPW) * 32;
    constexpr int NWT = NCW + NPW;
    constexpr int NPT = NPW * 32;
    constexpr int U = NCW * 32;  // units: samples (SPT = 1) or sample pairs (k, k + U) (SPT = 2)
    constexpr int SB = U * SPT;
    static_assert(SPT == 1 || SPT == 2, "one or two samples per consumer thread");
    constexpr int LDC = U + 1;  // cells per tile row (padded: see PASS 6)
    constexpr int TB = mppi_ws_tail_threads(NT);
    constexpr int NTW = TB / 32;
    constexpr bool kReplay = (NOISE == NOISE_REPLAY);
    using areal = typename ArithT<float, SPT>::type;

    extern __shared__ __align__(32) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full_bar[kWsMaxChunks];
    __shared__ int s_poison;
    __shared__ unsigned int nz_s[32];  // per sample-warp of the batch: which samples have a non-zero weight
    __shared__ int s_next;  // producers' work queue: next item (chunk-major: all unit-warps of chunk 0, then chunk 1, ...)

    const int H = p.H;
    const int Hq = (H + 3) >> 2, H4a = Hq * 4;
    double* scratch = reinterpret_cast<double*>(smem_raw);
    double* part_d = scratch + kScratchDoubles;
    double* U_run = part_d + mppi_ws_part_doubles(H, NT);  // [H + 8]
    double* c_s = U_run + (H + 8);                       // [SB] costs c_k of the batch
    const size_t ndbl = ((size_t)kScratchDoubles + mppi_ws_part_doubles(H, NT) + (size_t)(H + 8) + (size_t)SB + 3) & ~(size_t)3;
    float* su = reinterpret_cast<float*>(scratch + ndbl);  // [H4a] u_n (0 beyond H)
    float* sui = su + H4a;                                 // [H4a] u_n * sigma^-2 (0 beyond H)
    float* w_s = sui + H4a;                                // [SB]
    int* nan_s = reinterpret_cast<int*>(w_s + SB);         // [SB] replay: the sample's noise row held a NaN
    float4* tile = reinterpret_cast<float4*>(smem_raw + ((ndbl * 8 + ((size_t)2 * H4a + 2 * (size_t)SB) * 4 + 15) & ~(size_t)15));

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool is_producer = wid < NPW;  // consumers take the higher warp ids
    const int c = blockIdx.x / p.chunks;
    const int chunk = blockIdx.x % p.chunks;

    MPCB_TS(0);
    if (p.debug_ts != nullptr && threadIdx.x == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.debug_ts[(size_t)blockIdx.x * 16 + 6] = smid + 1;
    }
    // ---- prologue: model constants, x0, u_n ----
    ModelT<areal> model;
    model.load(p.mc);
    float x0[4];
    {
        bool bad = false;
        if (tid == 0) s_poison = 0;
        __syncthreads();
        if (p.use_inline) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x0[i] = (float)p.xu_inline[i];
                bad = bad || (p.xu_inline[i] != p.xu_inline[i]);
            }
            for (int t = tid; t < H4a; t += NT) {
                const double ut = t < H ? p.xu_inline[4 + t] : 0.0;
                bad = bad || (ut != ut);
                su[t] = (float)ut;
                sui[t] = (float)(ut * p.inv_var);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double xi = p.x[(long long)c * 4 + i];
                x0[i] = (float)xi;
                bad = bad || (xi != xi);
            }
            for (int t = tid; t < H4a; t += NT) {
                const double ut = t < H ? p.u[(long long)c * H + t] : 0.0;
                bad = bad || (ut != ut);
                su[t] = (float)ut;
                sui[t] = (float)(ut * p.inv_var);
            }
        }
        if (bad) s_poison = 1;  // a NaN in x or u_n reaches every sample's cost in the reference (src/mppi.rs:53-61)
    }
    for (int t = tid; t < H; t += NT) U_run[t] = 0.0;
    const int CQ = p.ws_cq;  // 4-step groups per chunk
    const int nchunks = (Hq + CQ - 1) / CQ;
    if (tid < nchunks) mbar_init(&full_bar[tid], NCW);  // one arrival per (unit-warp, chunk) item
    const float lo = (float)p.lo, hi = (float)p.hi;
    const float neg2s2ln2 = (float)(-2.0 * p.std_dev * p.std_dev * 0.693147180559945309417);
    const double lambda = p.lambda;
    const double inv_lambda = 1.0 / lambda;

    double m_run = -CUDART_INF, S_run = 0.0;
    long long arg_run = kNoArg, nfin_run = 0;
    double* red_m = scratch;                        // [16]
    long long* red_a = (long long*)(scratch + 16);  // [16]
    int* red_n = (int*)(scratch + 32);              // [16]
    double* red_s = scratch + 48;                   // [24]: [0..15] warp sums

    const long long w_begin = p.W * chunk / p.chunks;
    const long long w_end = p.W * (chunk + 1) / p.chunks;
    if (tid == 0) s_next = 0;
    if constexpr (kReplay)
        for (int k = tid; k < SB; k += NT) nan_s[k] = 0;
    __syncthreads();
    const bool poison_all = s_poison != 0;

    unsigned int parity = 0;
    for (long long wb = w_begin; wb < w_end; wb += NCW * SPT, parity ^= 1u) {
        // sample-warps of this batch that exist; column k of the batch is local sample wb*32 + k
        long long nlw = w_end - wb;
        if (nlw > NCW * SPT) nlw = NCW * SPT;

        if (is_producer) {
            // ================= producers: noise and clamp, items (unit-warp, chunk) from the shared queue =================
            // four N(0, sigma^2) draws of batch column k for the steps of group g
            auto draw4 = [&](int k, int g, float(&e)[4]) {
                const long long kl = wb * 32 + k;
                const long long kg = p.k_offset + kl;
                if constexpr (kReplay) {
                    const long long row = ((long long)c * p.K_global + kg) * H;
                    const bool ok = kl < p.K_local;
                    bool nan = false;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        e[i] = 0.0f;
                        if (ok && 4 * g + i < H)
                            e[i] = p.eps_f64 ? (float)reinterpret_cast<const double*>(p.eps)[row + 4 * g + i]
                                             : reinterpret_cast<const float*>(p.eps)[row + 4 * g + i];
                        nan = nan || (e[i] != e[i]);
                    }
                    if (nan) nan_s[k] = 1;
                } else {
                    const unsigned int c0 = (unsigned int)(kg & 0xffffffffll);
                    const unsigned int khi = (unsigned int)((kg >> 32) & 0xffff) << 16;
                    const Philox4 r = philox4x32(c0, p.call_idx, (unsigned int)c + p.c_offset, (unsigned int)g | khi, p.seed_lo, p.seed_hi);
                    philox_normal4(r, neg2s2ln2, e);
                    if constexpr (NOISE == NOISE_GENERATE_DUMP) {
                        if (kl < p.K_local) {
                            float* dump = reinterpret_cast<float*>(p.eps_dump) + ((long long)c * p.K_local + kl) * H;
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (4 * g + i < H) dump[4 * g + i] = e[i];
                        }
                    }
                }
            };
            const int nitems = nchunks * NCW;
            for (;;) {
                int item = 0;
                if (lane == 0) item = atomicAdd(&s_next, 1);
                item = __shfl_sync(0xffffffffu, item, 0);
                if (item >= nitems) break;
                const int ch = item / NCW, uw = item - ch * NCW;
                if (uw < nlw && !(p.ws_debug & 1)) {  // unit-warps beyond the block's range only count themselves in
                    const int unit = uw * 32 + lane;
                    const int g_lo = ch * CQ;
                    int g_hi = g_lo + CQ;
                    if (g_hi > Hq) g_hi = Hq;
                    for (int g = g_lo; g < g_hi; ++g) {
                        float u4[4];
                        lds4(su + 4 * g, u4);
                        if constexpr (SPT == 1) {
                            float e[4];
                            draw4(unit, g, e);
                            tile[g * LDC + unit] = make_float4(fminf(fmaxf(u4[0] + e[0], lo), hi), fminf(fmaxf(u4[1] + e[1], lo), hi),
                                                              fminf(fmaxf(u4[2] + e[2], lo), hi), fminf(fmaxf(u4[3] + e[3], lo), hi));
                        } else {
                            // pair (k, k + U): two cells, each with two steps of both samples — what one f32x2 step reads
                            float ea[4], eb[4], va[4], vb[4];
                            draw4(unit, g, ea);
                            draw4(unit + U, g, eb);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                va[i] = fminf(fmaxf(u4[i] + ea[i], lo), hi);
                                vb[i] = fminf(fmaxf(u4[i] + eb[i], lo), hi);
                            }
                            tile[(2 * g) * LDC + unit] = make_float4(va[0], vb[0], va[1], vb[1]);
                            tile[(2 * g + 1) * LDC + unit] = make_float4(va[2], vb[2], va[3], vb[3]);
                        }
                    }
                }
                __syncwarp();  // orders the lanes' tile stores before lane 0's release
                if (lane == 0) mbar_arrive(&full_bar[ch]);
            }
            if (p.debug_ts != nullptr && lane == 0 && wid == 0) p.debug_ts[(size_t)blockIdx.x * 16 + 12] = globaltimer_ns();
        } else {
            // ================= consumers: rollout, stage cost, control term =================
            const int ctid = tid - NPT;
            const int cw = ctid >> 5;
            bool live[SPT];
#pragma unroll
            for (int s = 0; s < SPT; ++s) live[s] = (cw + s * NCW) < nlw;
            double ck[SPT];
#pragma unroll
            for (int s = 0; s < SPT; ++s) ck[s] = -CUDART_INF;
            if (live[0] && !(p.ws_debug & 2)) {  // live[1] implies live[0]
                // JC = cost + control term (src/mppi.rs:57,60): FP32 partial sums of 4 steps, added up in FP64
                double JC[SPT];
#pragma unroll
                for (int s = 0; s < SPT; ++s) JC[s] = 0.0;
                const int Hfull = H >> 2;  // groups with four real steps
                const float4* col = tile + ctid;
                if constexpr (SPT == 1) {
                    float x[4] = {x0[0], x0[1], x0[2], x0[3]};
                    float acc = 0.0f;
                    for (int ch = 0; ch < nchunks; ++ch) {
                        mbar_wait(&full_bar[ch], parity);
                        const int g_lo = ch * CQ;
                        int g_hi = g_lo + CQ;
                        if (g_hi > Hq) g_hi = Hq;
                        for (int g = g_lo; g < g_hi; ++g) {
                            const float4 v4 = col[g * LDC];
                            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
                            float ui4[4];
                            lds4(sui + 4 * g, ui4);
                            if (g < Hfull) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    model.step(x, v[i]);
                                    acc = model.cost.acc(x, fmaf(ui4[i], v[i], acc));
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    if (4 * g + i < H) {
                                        model.step(x, v[i]);
                                        acc = model.cost.acc(x, fmaf(ui4[i], v[i], acc));
                                    }
                                }
                            }
                            JC[0] += (double)acc;
                            acc = 0.0f;
                        }
                    }
                } else {
                    f2 x[4] = {splat2(x0[0]), splat2(x0[1]), splat2(x0[2]), splat2(x0[3])};
                    f2 acc = splat2(0.0f);
                    for (int ch = 0; ch < nchunks; ++ch) {
                        mbar_wait(&full_bar[ch], parity);
                        const int g_lo = ch * CQ;
                        int g_hi = g_lo + CQ;
                        if (g_hi > Hq) g_hi = Hq;
                        for (int g = g_lo; g < g_hi; ++g) {
                            const float4 a4 = col[(2 * g) * LDC], b4 = col[(2 * g + 1) * LDC];
                            const f2 v[4] = {mk2(a4.x, a4.y), mk2(a4.z, a4.w), mk2(b4.x, b4.y), mk2(b4.z, b4.w)};
                            float ui4[4];
                            lds4(sui + 4 * g, ui4);
                            if (g < Hfull) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    model.step(x, v[i]);
                                    acc = model.cost.acc(x, fma2(splat2(ui4[i]), v[i], acc));
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < 3; ++i) {
                                    if (4 * g + i < H) {
                                        model.step(x, v[i]);
                                        acc = model.cost.acc(x, fma2(splat2(ui4[i]), v[i], acc));
                                    }
                                }
                            }
                            float ja, jb;
                            un2(acc, ja, jb);
                            JC[0] += (double)ja;
                            JC[1] += (double)jb;
                            acc = splat2(0.0f);
                        }
                    }
                }
                // c_k = -cost - control term (src/mppi.rs:61)
#pragma unroll
                for (int s = 0; s < SPT; ++s) {
                    double cs = -JC[s];
                    // FP32 rollouts overflow (inf, then inf - inf = NaN) where the f64 reference still holds a huge finite
                    // cost whose weight underflows to exactly 0: with clean inputs a NaN cost is that case -> weight 0.
                    // A NaN that came in through x, u_n or the sample's replay noise poisons the sums like the reference.
                    bool in_nan = poison_all;
                    if constexpr (kReplay) in_nan = in_nan || nan_s[ctid + s * U] != 0;
                    if (in_nan) cs = (double)CUDART_NAN;
                    else if (cs != cs) cs = -CUDART_INF;
                    ck[s] = cs;
                }
            }
            if (p.debug_ts != nullptr && lane == 0 && live[0] && (cw == 0 || cw + 1 == (nlw < NCW ? nlw : NCW)))
                p.debug_ts[(size_t)blockIdx.x * 16 + (cw == 0 ? 14 : 15)] = globaltimer_ns();
            // PASS 3, first stage: the consumer warp reduces its own samples (finite costs only, lowest index on ties) right
            // here, so that after the barrier every warp only has to reduce the NCW warp results
            double tm = -CUDART_INF;
            long long ta = kNoArg;
            int nf = 0;
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const int k = ctid + s * U;
                const long long kl = wb * 32 + k;
                const bool valid = live[s] && kl < p.K_local;
                if (!valid) ck[s] = -CUDART_INF;
                c_s[k] = ck[s];
                if (p.costs != nullptr && valid) p.costs[(long long)c * p.K_local + kl] = ck[s];
                if (finite_f64(ck[s])) {
                    ++nf;
                    if (ta == kNoArg || ck[s] > tm) { tm = ck[s]; ta = p.k_offset + kl; }  // s ascending = index ascending
                }
            }
            double wm_;
            long long wa_;
            warp_max_minidx(tm, ta, ta != kNoArg, &wm_, &wa_);
            const int wn_ = __reduce_add_sync(0xffffffffu, nf);
            if (lane == 0) { red_m[cw] = wm_; red_a[cw] = wa_; red_n[cw] = wn_; }
        }
        __syncthreads();
        if (tid == 0) s_next = 0;  // the next batch's queue (its first use comes after further barriers)
        if constexpr (kReplay)
            for (int k = tid; k < SB; k += NT) nan_s[k] = 0;

        // ---- PASS 3, second stage: block max over the consumer warps' results (every warp computes it) ----
        double bm = -CUDART_INF;
        long long ba = kNoArg;
        int bn = 0;
        {
            const bool has = lane < NCW;
            const double wm = has ? red_m[lane] : -CUDART_INF;
            const long long wa = has ? red_a[lane] : kNoArg;
            const int wn = has ? red_n[lane] : 0;
            warp_max_minidx(wm, wa, has && wa != kNoArg, &bm, &ba);
            bn = __reduce_add_sync(0xffffffffu, wn);
        }
        const double m_old = m_run;
        if (bm > m_run) { m_run = bm; arg_run = ba; }
        nfin_run += bn;
        if (p.debug_ts != nullptr && tid == 0) p.debug_ts[(size_t)blockIdx.x * 16 + 13] = globaltimer_ns();

        // ---- PASS 4-5: weights against the running max (natural IEEE semantics: -inf -> 0, NaN / +inf poison) ----
        if (tid < TB) {
            float wsum = 0.0f;
            for (int k = tid; k < SB; k += TB) {
                const double cv = c_s[k];
                float w;
                if (cv == -CUDART_INF) w = 0.0f;
                else if (m_run == -CUDART_INF) w = (float)(cv - cv);  // no finite cost yet: NaN / +inf still poison
                else w = fast_exp_neg((cv - m_run) * inv_lambda);
                w_s[k] = w;
                wsum += w;
                // which samples of this sample-warp carry weight (NaN counts): PASS 6 walks the set bits only
                const unsigned int nz = __ballot_sync(0xffffffffu, w != 0.0f);
                if (lane == 0) nz_s[k >> 5] = nz;
            }
            const double wd = warp_sum_f64((double)wsum);
            if (lane == 0) red_s[wid] = wd;
        }
        __syncthreads();
        if (p.debug_ts != nullptr && tid == 0 && p.groups == 1) p.debug_ts[(size_t)blockIdx.x * 16 + 3] = globaltimer_ns();
        const double resc = (m_old == -CUDART_INF) ? 0.0 : exp((m_old - m_run) * inv_lambda);
        double bs = red_s[0];
#pragma unroll
        for (int wI = 1; wI < NTW; ++wI) bs += red_s[wI];
        S_run = S_run * resc + bs;

        // ---- PASS 6: sum_k w_k v[k][t].  Thread item (row, q): one tile row — four steps of every sample (SPT = 1) or two
        // steps of every sample pair (SPT = 2) — over the unit partition q, one 128-bit load per cell (the rows are padded
        // by one cell, so the eight lanes of a quarter warp hit all 32 banks), weights broadcast; only the cells whose
        // weights are not zero are visited (nz_s: one bit per sample, set in PASS 4-5) — samples far from the best and warps
        // beyond the range cost nothing, which leaves the sums unchanged (and keeps never-written cells out: 0 * NaN).
        // FP32 sums of <= kper terms, added up in FP64 over the partitions in order: nothing depends on timing. ----
        constexpr int TPR = 4 / SPT;       // steps per tile row
        const int rows = Hq * SPT;
        int nq = NT / rows;
        if (nq < 1) nq = 1;
        if (nq > U / 8) nq = U / 8;
        const int kper = (U + nq - 1) / nq;
        for (int item = tid; item < rows * nq; item += NT) {
            const int q = item / rows, row = item - q * rows;
            const int k_lo = q * kper;
            int k_hi = k_lo + kper;
            if (k_hi > U) k_hi = U;
            const float4* cp = tile + (size_t)row * LDC;
            float acc[TPR];
#pragma unroll
            for (int i = 0; i < TPR; ++i) acc[i] = 0.0f;
            // the units of [k_lo, k_hi) that carry weight, in ascending order (one mask word per 32 units)
            for (int wd = k_lo >> 5; k_lo < k_hi && wd <= (k_hi - 1) >> 5; ++wd) {
                unsigned int mk = nz_s[wd];
                if constexpr (SPT == 2) mk |= nz_s[wd + NCW];
                const int b_lo = k_lo - 32 * wd, b_hi = k_hi - 32 * wd;  // partition bounds relative to this word
                if (b_lo > 0) mk &= 0xffffffffu << b_lo;
                if (b_hi < 32) mk &= (1u << b_hi) - 1u;
                while (mk != 0u) {
                    const int k = 32 * wd + __ffs(mk) - 1;
                    mk &= mk - 1u;
                    const float4 v4 = cp[k];
                    if constexpr (SPT == 1) {
                        const float wk = w_s[k];
                        acc[0] = fmaf(wk, v4.x, acc[0]);
                        acc[1] = fmaf(wk, v4.y, acc[1]);
                        acc[2] = fmaf(wk, v4.z, acc[2]);
                        acc[3] = fmaf(wk, v4.w, acc[3]);
                    } else {
                        // v4 = (a[t], b[t], a[t+1], b[t+1]); a zero weight must not touch its value (0 * NaN)
                        const float wa = w_s[k], wb2 = w_s[k + U];
                        if (wa != 0.0f) {
                            acc[0] = fmaf(wa, v4.x, acc[0]);
                            acc[1] = fmaf(wa, v4.z, acc[1]);
                        }
                        if (wb2 != 0.0f) {
                            acc[0] = fmaf(wb2, v4.y, acc[0]);
                            acc[1] = fmaf(wb2, v4.w, acc[1]);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < TPR; ++i) part_d[(size_t)q * H4a + TPR * row + i] = (double)acc[i];
        }
        __syncthreads();
        if (p.debug_ts != nullptr && tid == 0 && p.groups == 1) p.debug_ts[(size_t)blockIdx.x * 16 + 4] = globaltimer_ns();
        for (int t = tid; t < H; t += NT) {
            double a = part_d[t];
            for (int q = 1; q < nq; ++q) a += part_d[(size_t)q * H4a + t];
            U_run[t] = U_run[t] * resc + a;
        }
        __syncthreads();
    }

    // ---- the block's partial row, arrival, merges: the first TB threads; whole warps beyond them are done ----
    if (tid >= TB) return;
    mppi_block_tail<TB, true>(p, c, chunk, m_run, arg_run, S_run, nfin_run, U_run, scratch, part_d);
}

// kernel entry table (mppi_ws_{L,NL,NL6}.cu): variant = index into kWsVariants
struct MppiWsVariant {
    int ncw, npw, spt;
};
constexpr MppiWsVariant kWsVariants[] = {
    {14, 14, 1},  // 448 samples per batch (configs[1]: 13-14 sample-warps per SM), one sample per consumer thread
    {7, 9, 2},    // the same samples as pairs: packed f32x2 consumers, 512 threads
    {7, 7, 2},    // packed consumers, fewer producer warps
    {7, 14, 2},   // packed consumers, more producer warps
    {16, 16, 1},  // 512 samples per batch
    {8, 8, 2},    // 512 samples per batch, packed
    {8, 8, 1},    // 256 samples per batch (long horizons: the tile is H * SB * 4 bytes)
    {4, 4, 2},    // 256 samples per batch, packed
};
constexpr int kNumWsVariants = (int)(sizeof(kWsVariants) / sizeof(kWsVariants[0]));
MppiKernelFn mppi_kernel_ws(int model_id, int variant, int noise);

}  // namespace mpcb
