// mppi_kernel.cuh — the fused MPPI control step (replaces the six rayon passes of src/mppi.rs:38-91).
//
// One thread per sample.  A block of BLOCK samples does, in one pass and without ever writing the K×H
// control samples to HBM:
//   PASS 1  noise: Philox4x32-10 + Box-Muller in registers (generate) or a coalesced tile load of the
//           caller's eps (replay);  v = clamp(u_n + eps)                                   (:38-45)
//   PASS 2  H-step rollout of the model + stage cost + control term                        (:48-63)
//   PASS 3-6 as an ONLINE softmax: block max over finite c_k (warp shuffles), w = exp((c-m)/lambda),
//           running (m, sum_w, sum_w*v[t]) rescaled when the max moves                     (:65-84)
// v stays in shared memory ([H][BLOCK+1], conflict-free both for the per-thread writes and for the
// per-t weighted sums).  Each block leaves one partial row [m, sum_w, argmax, n_finite, sum_w*v[0..H)].
// Rows are merged by a two-level ticket tree (fan-in <= 64): the last block of a group to arrive merges
// the group's rows, the last group to arrive merges the group rows and writes u_out/info — or, when the
// samples are sharded over GPUs, the merged un-normalised row for the cross-rank exchange.  Every load of
// a merge is independent, so the tail costs a few L2 round trips instead of one per row.
#pragma once

#include <math_constants.h>

#include "common.cuh"
#include "models.cuh"
#include "philox.cuh"

namespace mpcb {

constexpr int kPartialHdr = 4;   // m, sum_w, argmax (bits), n_finite (bits)
constexpr int kMergeFan = 64;    // max rows merged by one block
constexpr int kScratchDoubles = 32 + kMergeFan;

enum MppiNoise { NOISE_GENERATE = 0, NOISE_GENERATE_DUMP = 1, NOISE_REPLAY = 2 };

struct MppiParams {
    int H;
    int C;
    int chunks;      // blocks per controller (level-0 rows)
    int group_size;  // level-0 rows per group (<= kMergeFan)
    int groups;      // level-1 rows per controller (<= kMergeFan); 1 = single-level merge
    int use_inline;  // x/u come from xu_inline (C == 1, H <= kInlineHorizon)
    long long K_local;
    long long K_global;
    long long k_offset;  // global index of this rank's first sample
    long long batches_per_chunk;
    const double* x;  // [C][4]
    const double* u;  // [C][H]
    const void* eps;  // replay: [C][K_global][H]
    int eps_f64;
    int final_mode;  // 0: normalise into u_out/info; 1: merged partial row into rank_partial
    void* eps_dump;  // generate+dump: [C][K_local][H] of real
    double* costs;   // optional [C][K_local]
    unsigned int seed_lo, seed_hi, call_idx, pad0;
    double lambda, inv_var, lo, hi, std_dev;
    double* partial;         // [C][chunks + groups][kPartialHdr + H]: level-0 rows, then group rows
    unsigned int* counters;  // [C][groups + 1]
    double* u_out;           // [C][H]
    double* u_out_host;      // mapped host mirror or nullptr
    mpcb_mppi_info* info;    // [C]
    mpcb_mppi_info* info_host;
    double* rank_partial;    // [C][kPartialHdr + H]
    unsigned int* done_host;       // mapped host word; the final block stores `epoch` after u_out/info (C == 1 only)
    unsigned int epoch, pad1;
    unsigned long long* debug_ts;  // optional [blocks][8] %globaltimer stamps (diagnostics)
    ModelConsts mc;
    double xu_inline[4 + kInlineHorizon];
};

__device__ __forceinline__ double ll_as_double(long long v) { return __longlong_as_double(v); }
__device__ __forceinline__ long long double_as_ll(double v) { return __double_as_longlong(v); }

constexpr long long kNoArg = 0x7fffffffffffffffll;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define MPCB_TS(slot)                                                                       \
    do {                                                                                    \
        if (p.debug_ts != nullptr && threadIdx.x == 0) p.debug_ts[(size_t)blockIdx.x * 8 + (slot)] = globaltimer_ns(); \
    } while (0)

// gpu-scope release/acquire fence around the ticket atomics (cheaper than the sequentially consistent
// __threadfence(); the ticket pattern only needs release on the producer and acquire on the consumer side)
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Merges n_rows (<= kMergeFan) partial rows (row r at rows + r*row_stride) of one controller; the whole
// block participates.  scratch: kScratchDoubles doubles of shared memory.
//   final_mode 1: the merged, un-normalised row is written to out_row
//   final_mode 0: u_out = sum_w*v / sum_w, info, status                      (src/mppi.rs:76-91)
// All global loads of a merge are issued up front (headers and up to kMergeBatch values per thread), so a
// level costs about one L2 round trip instead of one per pass.
constexpr int kMergeBatch = 16;

template <int BLOCK>
__device__ void mppi_merge_rows(const double* rows, long long row_stride, int n_rows, int H, double lambda,
                                int final_mode, double* u_out, double* u_out_host, mpcb_mppi_info* info,
                                mpcb_mppi_info* info_host, double* out_row, double* scratch,
                                unsigned int* done_host = nullptr, unsigned int epoch = 0) {
    constexpr int NW = BLOCK / 32;
    constexpr int RPT = (kMergeFan + BLOCK - 1) / BLOCK;  // header rows per thread
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double* red_m = scratch;                        // [8]
    long long* red_a = (long long*)(scratch + 8);   // [8]
    long long* red_n = (long long*)(scratch + 16);  // [8]
    double* red_s = scratch + 24;                   // [8]
    double* sc = scratch + 32;                      // [kMergeFan] per-row scale
    __shared__ int s_status;

    // ---- issue every load of the first column pass: row headers + the first kMergeBatch rows of column tid ----
    double hm[RPT], hs[RPT];
    long long ha[RPT], hn[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = tid + i * BLOCK;
        hm[i] = -CUDART_INF; hs[i] = 0.0; ha[i] = -1; hn[i] = 0;
        if (r < n_rows) {
            const double* row = rows + (long long)r * row_stride;
            hm[i] = __ldcg(row + 0);
            hs[i] = __ldcg(row + 1);
            ha[i] = double_as_ll(__ldcg(row + 2));
            hn[i] = double_as_ll(__ldcg(row + 3));
        }
    }
    double v0[kMergeBatch];
    {
        const double* col = rows + kPartialHdr + tid;
#pragma unroll
        for (int i = 0; i < kMergeBatch; ++i)
            v0[i] = (tid < H && i < n_rows) ? __ldcg(col + (long long)i * row_stride) : 0.0;
    }

    // ---- max / argmax / finite count over the headers ----
    double m = -CUDART_INF;
    long long a = kNoArg, n = 0;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        n += hn[i];
        if (ha[i] >= 0 && (hm[i] > m || (hm[i] == m && ha[i] < a))) { m = hm[i]; a = ha[i]; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double om = shfl_down_f64(m, off);
        const long long oa = shfl_down_i64(a, off);
        n += shfl_down_i64(n, off);
        if (om > m || (om == m && oa < a)) { m = om; a = oa; }
    }
    if (lane == 0) { red_m[wid] = m; red_a[wid] = a; red_n[wid] = n; }
    __syncthreads();
    m = red_m[0]; a = red_a[0]; n = red_n[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) {
        const double om = red_m[w];
        const long long oa = red_a[w];
        n += red_n[w];
        if (om > m || (om == m && oa < a)) { m = om; a = oa; }
    }
    const bool any = (a != kNoArg);
    // ---- per-row scale exp((m_r - m)/lambda) and the merged sum of weights ----
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = tid + i * BLOCK;
        if (r < n_rows) {
            const bool empty = (hm[i] == -CUDART_INF);  // the row saw no finite cost
            const double scale = empty ? 0.0 : exp((hm[i] - m) / lambda);
            // an empty row's sum_w is 0, or NaN/inf when it saw NaN/+inf costs: keep that poison (f64 semantics)
            s += empty ? hs[i] : scale * hs[i];
            sc[r] = scale;
        }
    }
    s = warp_sum_f64(s);
    if (lane == 0) red_s[wid] = s;
    if (tid == 0) s_status = MPCB_OK;
    __syncthreads();
    s = red_s[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) s += red_s[w];
    // ---- merged weighted control sums ----
    for (int t = tid; t < H; t += BLOCK) {
        double acc = 0.0;
        const double* col = rows + kPartialHdr + t;
        for (int r0 = 0; r0 < n_rows; r0 += kMergeBatch) {
            double v[kMergeBatch];
            if (t == tid && r0 == 0) {
#pragma unroll
                for (int i = 0; i < kMergeBatch; ++i) v[i] = v0[i];
            } else {
#pragma unroll
                for (int i = 0; i < kMergeBatch; ++i)
                    v[i] = (r0 + i < n_rows) ? __ldcg(col + (long long)(r0 + i) * row_stride) : 0.0;
            }
#pragma unroll
            for (int i = 0; i < kMergeBatch; ++i)
                if (r0 + i < n_rows) acc += sc[r0 + i] * v[i];  // scale 0 * NaN = NaN keeps the reference's poisoning
        }
        if (final_mode == 1) {
            out_row[kPartialHdr + t] = acc;
        } else {
            const double uo = any ? acc / s : 0.0;
            if (t == 0) {
                int st = MPCB_OK;
                if (!any) st = MPCB_NO_FINITE_COST;            // src/mppi.rs:69
                else if (s == 0.0) st = MPCB_SUM_ZERO;          // :76-78
                else if (!finite_f64(uo)) st = MPCB_U_INVALID;  // :87-89 (element 0 only)
                s_status = st;
            }
            u_out[t] = uo;
            if (u_out_host) u_out_host[t] = uo;
        }
    }
    if (done_host) __threadfence_system();  // every writer orders its host stores before the completion word
    __syncthreads();
    if (tid == 0) {
        if (final_mode == 1) {
            out_row[0] = any ? m : -CUDART_INF;
            out_row[1] = s;
            out_row[2] = ll_as_double(any ? a : -1ll);
            out_row[3] = ll_as_double(n);
        } else {
            mpcb_mppi_info out;
            out.status = s_status;
            out.reserved = 0;
            out.argmax = any ? a : -1ll;
            out.max = any ? m : 0.0;
            out.sum = s;
            out.n_finite = n;
            *info = out;
            if (info_host) *info_host = out;
            if (done_host) {
                // results first, then the completion word the host spins on (system-scope release)
                __threadfence_system();
                *reinterpret_cast<volatile unsigned int*>(done_host) = epoch;
            }
        }
    }
}

template <typename real>
struct RealTraits;
template <>
struct RealTraits<float> {
    static constexpr bool kExact = false;
};
template <>
struct RealTraits<double> {
    static constexpr bool kExact = true;
};

// Shared-memory bytes of one block.
template <typename real>
__host__ __device__ inline size_t mppi_smem_bytes(int H, int block) {
    size_t dbl = (size_t)H + kScratchDoubles;                     // U_run + reduction/merge scratch
    size_t rl = (size_t)2 * H + block + (size_t)H * (block + 1);  // su, sui, w_s, v_s
    size_t bytes = dbl * sizeof(double) + rl * sizeof(real);
    return (bytes + 15) & ~(size_t)15;
}

template <template <typename> class ModelT, typename real, int BLOCK, int NOISE>
__global__ void __launch_bounds__(BLOCK) mppi_rollout_kernel(const __grid_constant__ MppiParams p) {
    constexpr int NW = BLOCK / 32;
    constexpr int LD = BLOCK + 1;
    constexpr bool kExact = RealTraits<real>::kExact;
    constexpr bool kReplay = (NOISE == NOISE_REPLAY);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = p.H;
    double* U_run = reinterpret_cast<double*>(smem_raw);            // [H]
    double* scratch = U_run + H;                                    // [kScratchDoubles]
    real* su = reinterpret_cast<real*>(scratch + kScratchDoubles);  // [H]  u_n
    real* sui = su + H;                                             // [H]  u_n * sigma^-2
    real* w_s = sui + H;                                            // [BLOCK]
    real* v_s = w_s + BLOCK;                                        // [H][LD]

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int c = blockIdx.x / p.chunks;      // controller
    const int chunk = blockIdx.x % p.chunks;  // sample chunk of this controller

    MPCB_TS(0);
    // ---- prologue: model constants, x0, u_n ----
    ModelT<real> model;
    model.load(p.mc);
    real x0[4];
    if (p.use_inline) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x0[i] = (real)p.xu_inline[i];
        for (int t = tid; t < H; t += BLOCK) {
            const double ut = p.xu_inline[4 + t];
            su[t] = (real)ut;
            sui[t] = (real)(ut * p.inv_var);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) x0[i] = (real)p.x[(long long)c * 4 + i];
        for (int t = tid; t < H; t += BLOCK) {
            const double ut = p.u[(long long)c * H + t];
            su[t] = (real)ut;
            sui[t] = (real)(ut * p.inv_var);
        }
    }
    for (int t = tid; t < H; t += BLOCK) U_run[t] = 0.0;
    const real lo = (real)p.lo, hi = (real)p.hi;
    const float neg2s2ln2 = (float)(-2.0 * p.std_dev * p.std_dev * 0.693147180559945309417);
    const double lambda = p.lambda;

    double m_run = -CUDART_INF, S_run = 0.0;
    long long arg_run = kNoArg, nfin_run = 0;

    double* red_m = scratch;                       // [8]
    long long* red_a = (long long*)(scratch + 8);  // [8]
    int* red_n = (int*)(scratch + 16);             // [8]
    double* red_s = scratch + 24;                  // [8]

    const long long batch0 = (long long)chunk * p.batches_per_chunk;
    const long long nbatches_total = (p.K_local + BLOCK - 1) / BLOCK;
    long long batch_end = batch0 + p.batches_per_chunk;
    if (batch_end > nbatches_total) batch_end = nbatches_total;
    __syncthreads();

    for (long long b = batch0; b < batch_end; ++b) {
        const long long kl = b * BLOCK + tid;  // local sample index
        const bool valid = kl < p.K_local;
        const long long kg = p.k_offset + kl;  // global sample index (Philox counter / replay row / argmax)

        if constexpr (kReplay) {
            // coalesced tile load: BLOCK consecutive sample rows of H values -> v_s[t][k]
            const long long k_first = b * BLOCK;
            long long nrows = p.K_local - k_first;
            if (nrows > BLOCK) nrows = BLOCK;
            const long long base = ((long long)c * p.K_global + p.k_offset + k_first) * H;
            const int total = BLOCK * H;
            const int live = (int)nrows * H;
            for (int i = tid; i < total; i += BLOCK) {
                real e = (real)0;
                if (i < live) {
                    e = p.eps_f64 ? (real) reinterpret_cast<const double*>(p.eps)[base + i]
                                  : (real) reinterpret_cast<const float*>(p.eps)[base + i];
                }
                const int kr = i / H, t = i - kr * H;
                v_s[t * LD + kr] = e;
            }
            __syncthreads();
        }

        // ---- PASS 1+2: noise, clamp, rollout, cost ----
        real x[4] = {x0[0], x0[1], x0[2], x0[3]};
        double J = 0.0, CT = 0.0;
        real cw = (real)0, cu = (real)0;
        const unsigned int c0 = (unsigned int)(kg & 0xffffffffll);
        const unsigned int khi = (unsigned int)((kg >> 32) & 0xffff) << 16;
        real* vcol = v_s + tid;

        // four N(0, sigma^2) draws for steps t0..t0+3
        auto noise4 = [&](int t0, real(&e)[4]) {
            if constexpr (kReplay) {
#pragma unroll
                for (int i = 0; i < 4; ++i) e[i] = (t0 + i < H) ? vcol[(t0 + i) * LD] : (real)0;
            } else {
                const Philox4 r = philox4x32_10(c0, p.call_idx, (unsigned int)c, (unsigned int)(t0 >> 2) | khi,
                                                p.seed_lo, p.seed_hi);
                float z[4];
                philox_normal4(r, neg2s2ln2, z);
#pragma unroll
                for (int i = 0; i < 4; ++i) e[i] = (real)z[i];
                if constexpr (NOISE == NOISE_GENERATE_DUMP) {
                    if (valid) {
                        real* dump = reinterpret_cast<real*>(p.eps_dump) + ((long long)c * p.K_local + kl) * H;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (t0 + i < H) dump[t0 + i] = e[i];
                    }
                }
            }
        };
        // one rollout step: v = clamp(u_n[t] + eps), x <- dynamics(x, v), cost, control term
        auto step = [&](int t, real eps) {
            real v = su[t] + eps;
            if constexpr (kExact) v = clampr(v, lo, hi);  // f64::clamp, NaN stays NaN
            else v = fminf(fmaxf(v, lo), hi);
            vcol[t * LD] = v;
            model.step(x, v);
            const real ct = model.cost(x);
            if constexpr (kExact) {
                J = J + ct;         // :57 c + cost(x_n)
                CT += sui[t] * v;   // :60 (u*inv)*v, summed in order
            } else {
                cw += ct;
                cu = fmaf(sui[t], v, cu);
            }
        };

        // Software pipeline: the noise of group g+1 (independent of the state) is drawn while group g rolls out,
        // so the scheduler has Philox/Box-Muller work to fill the dependency stalls of the serial dynamics chain.
        const int H4 = H & ~3;
        real e[4];
        noise4(0, e);
        for (int t0 = 0; t0 < H4; t0 += 4) {
            real en[4];
            noise4(t0 + 4, en);  // the last prefetch (t0 + 4 >= H4) feeds the tail below or is discarded
#pragma unroll
            for (int i = 0; i < 4; ++i) step(t0 + i, e[i]);
            if constexpr (!kExact) {
                // FP32 partial sums over 4 steps are flushed into the FP64 accumulators
                J += (double)cw;
                CT += (double)cu;
                cw = (real)0;
                cu = (real)0;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = en[i];
        }
        if (H4 < H) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (H4 + i < H) step(H4 + i, e[i]);
            if constexpr (!kExact) {
                J += (double)cw;
                CT += (double)cu;
            }
        }
        const double ck = -J - CT;  // :61
        if (p.costs != nullptr && valid) p.costs[(long long)c * p.K_local + kl] = ck;

        // ---- PASS 3: block max over finite c_k, lowest index on ties ----
        const bool fin = valid && finite_f64(ck);
        double bm = fin ? ck : -CUDART_INF;
        long long ba = fin ? kg : kNoArg;
        int bn = fin ? 1 : 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double om = shfl_down_f64(bm, off);
            const long long oa = shfl_down_i64(ba, off);
            bn += __shfl_down_sync(0xffffffffu, bn, off);
            if (om > bm || (om == bm && oa < ba)) { bm = om; ba = oa; }
        }
        if (lane == 0) { red_m[wid] = bm; red_a[wid] = ba; red_n[wid] = bn; }
        __syncthreads();
        bm = red_m[0]; ba = red_a[0]; bn = red_n[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
            const double om = red_m[w];
            const long long oa = red_a[w];
            bn += red_n[w];
            if (om > bm || (om == bm && oa < ba)) { bm = om; ba = oa; }
        }
        const double m_old = m_run;
        if (bm > m_run) { m_run = bm; arg_run = ba; }
        nfin_run += bn;

        // ---- PASS 4-5: weights against the running max, rescale of what was accumulated so far ----
        // exp(a) is exactly 0 in f64 for a < -745.14: skipping the call there is bit-identical and spares most
        // warps the FP64 exp (far-from-best samples dominate).
        double w;
        {
            const double arg = (ck - m_run) / lambda;
            if constexpr (kExact) {
                // reference semantics: exp((c - max)/lambda) for every sample, NaN/+inf poison the sums (:71-74)
                if (!valid || ck == -CUDART_INF) w = 0.0;
                else w = (arg < -746.0) ? 0.0 : exp(arg);
            } else {
                // FP32 rollouts can overflow where the f64 reference yields a huge finite cost whose weight
                // underflows to exactly 0: non-finite costs get weight 0 here.
                w = (fin && arg >= -746.0) ? exp(arg) : 0.0;
            }
        }
        w_s[tid] = (real)w;
        const double wsum = warp_sum_f64(w);
        if (lane == 0) red_s[wid] = wsum;
        if (tid == 0) red_s[NW] = (m_old == -CUDART_INF) ? 0.0 : exp((m_old - m_run) / lambda);  // block-uniform rescale
        __syncthreads();
        const double resc = red_s[NW];
        double bs = red_s[0];
#pragma unroll
        for (int wI = 1; wI < NW; ++wI) bs += red_s[wI];
        S_run = S_run * resc + bs;

        // ---- PASS 6: sum_k w_k * v[k][t], one thread per t, conflict-free column reads ----
        for (int t = tid; t < H; t += BLOCK) {
            const real* col = v_s + t * LD;
            real a0 = (real)0, a1 = (real)0, a2 = (real)0, a3 = (real)0;
#pragma unroll 4
            for (int k = 0; k < BLOCK; k += 4) {
                a0 += w_s[k + 0] * col[k + 0];
                a1 += w_s[k + 1] * col[k + 1];
                a2 += w_s[k + 2] * col[k + 2];
                a3 += w_s[k + 3] * col[k + 3];
            }
            U_run[t] = U_run[t] * resc + (double)((a0 + a1) + (a2 + a3));
        }
        __syncthreads();
    }

    MPCB_TS(1);
    // ---- partial row of this block, then the two-level ticket merge ----
    const int PL = kPartialHdr + H;
    double* ctrl_rows = p.partial + (long long)c * (p.chunks + p.groups) * PL;
    double* my_row = ctrl_rows + (long long)chunk * PL;
    for (int t = tid; t < H; t += BLOCK) my_row[kPartialHdr + t] = U_run[t];
    if (tid == 0) {
        my_row[0] = m_run;
        my_row[1] = S_run;
        my_row[2] = ll_as_double(arg_run == kNoArg ? -1ll : arg_run);
        my_row[3] = ll_as_double(nfin_run);
    }
    __shared__ int s_last;
    unsigned int* cnt = p.counters + (long long)c * (p.groups + 1);
    const int g = chunk / p.group_size;
    const int g_first = g * p.group_size;
    int g_rows = p.chunks - g_first;
    if (g_rows > p.group_size) g_rows = p.group_size;

    fence_acq_rel_gpu();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&cnt[g], 1u) == (unsigned int)(g_rows - 1));
    __syncthreads();
    MPCB_TS(2);
    if (!s_last) return;
    fence_acq_rel_gpu();
    double* u_out_c = p.u_out + (long long)c * H;
    double* u_host_c = p.u_out_host ? p.u_out_host + (long long)c * H : nullptr;
    mpcb_mppi_info* info_host_c = p.info_host ? p.info_host + c : nullptr;
    double* rank_row = p.rank_partial ? p.rank_partial + (long long)c * PL : nullptr;
    if (p.groups == 1) {
        mppi_merge_rows<BLOCK>(ctrl_rows, PL, p.chunks, H, lambda, p.final_mode, u_out_c, u_host_c, p.info + c,
                               info_host_c, rank_row, scratch, p.done_host, p.epoch);
        if (tid == 0) cnt[0] = 0u;  // ready for the next launch
        MPCB_TS(3);
        return;
    }
    double* group_rows = ctrl_rows + (long long)p.chunks * PL;
    mppi_merge_rows<BLOCK>(ctrl_rows + (long long)g_first * PL, PL, g_rows, H, lambda, 1, nullptr, nullptr, nullptr,
                           nullptr, group_rows + (long long)g * PL, scratch);
    if (tid == 0) cnt[g] = 0u;
    MPCB_TS(3);
    fence_acq_rel_gpu();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&cnt[p.groups], 1u) == (unsigned int)(p.groups - 1));
    __syncthreads();
    MPCB_TS(4);
    if (!s_last) return;
    fence_acq_rel_gpu();
    mppi_merge_rows<BLOCK>(group_rows, PL, p.groups, H, lambda, p.final_mode, u_out_c, u_host_c, p.info + c, info_host_c,
                           rank_row, scratch, p.done_host, p.epoch);
    if (tid == 0) cnt[p.groups] = 0u;
    MPCB_TS(5);
}

// Cross-rank merge: rows[g][c][PL] gathered from all ranks -> u_out / info.  One block per controller.
struct MppiCombineParams {
    const double* rows;  // [G][C][PL]
    int G, C, H;
    double lambda;
    double* u_out;
    double* u_out_host;
    mpcb_mppi_info* info;
    mpcb_mppi_info* info_host;
    unsigned int* done_host;
    unsigned int epoch;
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) mppi_combine_kernel(const MppiCombineParams p) {
    __shared__ double scratch[kScratchDoubles];
    const int c = blockIdx.x;
    const int PL = kPartialHdr + p.H;
    mppi_merge_rows<BLOCK>(p.rows + (long long)c * PL, (long long)p.C * PL, p.G, p.H, p.lambda, 0,
                           p.u_out + (long long)c * p.H, p.u_out_host ? p.u_out_host + (long long)c * p.H : nullptr,
                           p.info + c, p.info_host ? p.info_host + c : nullptr, nullptr, scratch, p.done_host, p.epoch);
}

// kernel entry table (defined in mppi_f32.cu / mppi_f64.cu)
using MppiKernelFn = void (*)(const MppiParams);
MppiKernelFn mppi_kernel_f32(int model_id, int block, int noise);
MppiKernelFn mppi_kernel_f64(int model_id, int block, int noise);

}  // namespace mpcb
