// mppi_kernel.cuh — the fused MPPI control step (replaces the six rayon passes of src/mppi.rs:38-91).
//
// One thread per sample.  A block of BLOCK samples does, in one pass and without ever writing the K×H
// control samples to HBM:
//   PASS 1  noise: Philox4x32-10 + Box-Muller in registers (generate) or a coalesced tile load of the
//           caller's eps (replay);  v = clamp(u_n + eps)                                   (:38-45)
//   PASS 2  H-step rollout of the model + stage cost + control term                        (:48-63)
//   PASS 3-6 as an ONLINE softmax: block max over finite c_k (redux), w = exp((c-m)/lambda),
//           running (m, sum_w, sum_w*v[t]) rescaled when the max moves                     (:65-84)
// v stays in shared memory ([H][BLOCK+4]: conflict-free scalar column writes by the sample threads and
// conflict-free 128-bit row reads by the per-t weighted sums).  Each block leaves one partial row
// [m, sum_w, argmax, n_finite, sum_w*v[0..H)].
//
// Work split: the K_local samples of a controller are cut into `chunks` contiguous ranges of whole warps
// (range i = warps [W*i/chunks, W*(i+1)/chunks)), one block per range; a block walks its range in batches of
// BLOCK samples.  When the whole controller fits on the GPU in one batch per block the host picks ONE block per
// SM sized to its range (e.g. K = 65536 on 148 SMs: 148 blocks of 512 threads, 13-14 live warps each), so every
// SM gets the same number of samples and there are only num_sms rows to merge.
//
// Rows are merged by the last block to arrive at a ticket (single level, fan-in <= kMergeFan rows handled by
// all threads of the block: columns x row-partitions), or by a two-level ticket tree when there are more rows.
// The final merge writes u_out/info — or, when the samples are sharded over GPUs, either the merged
// un-normalised row (NCCL exchange done by the host code) or it runs the cross-GPU exchange itself: the row is
// stored straight into every peer's mailbox over NVLink, a flag per (peer, controller) is released, and the same
// block waits for the peers' rows and combines them (one kernel per control step, no collective call).
#pragma once

#include <math_constants.h>

#include "common.cuh"
#include "models.cuh"
#include "philox.cuh"

namespace mpcb {

constexpr int kPartialHdr = 4;   // m, argmax (bits), sum_w, n_finite (as double)
constexpr int kMergeFan = 256;   // max rows merged by one block
constexpr int kMaxWarps = 16;    // warps per block (BLOCK <= 512)
constexpr int kFlushSteps = 4;   // FP32 cost partial sums -> FP64 accumulators every 4 steps (16 measured no faster)
// shared scratch (doubles): red_m[16] red_a[16] red_n[16] red_s[24] sc[kMergeFan]
constexpr int kScratchDoubles = 16 + 16 + 16 + 24 + kMergeFan;

enum MppiNoise { NOISE_GENERATE = 0, NOISE_GENERATE_DUMP = 1, NOISE_REPLAY = 2 };
enum MppiFinal { FINAL_NORMALISE = 0, FINAL_RANK_ROW = 1, FINAL_PEER_EXCHANGE = 2 };

struct MppiParams {
    int H;
    int Hp;          // H rounded up to a power of two (>= 8): column index space of the weighted sums
    int lgHp;
    int PL;          // doubles per partial row: kPartialHdr + H rounded up to even
    int mergers;     // >= 1: every block of the launch is resident at once, the first `mergers` blocks of a controller
                     // share the final merge column-wise and group leaders wait for their group; 0: last arriver merges
    unsigned int seq;  // 1-based launch number of this handle: arrival counters are monotonic (seq * rows)
    int C;
    int chunks;      // blocks per controller (level-0 rows)
    int group_size;  // level-0 rows per group (<= kMergeFan)
    int groups;      // level-1 rows per controller (<= kMergeFan); 1 = single-level merge
    int use_inline;  // x/u come from xu_inline (C == 1, H <= kInlineHorizon)
    long long K_local;
    long long K_global;
    long long k_offset;  // global index of this rank's first sample
    long long W;         // warps of 32 samples in K_local
    const double* x;  // [C][S]
    const double* u;  // [C][H]
    const void* eps;  // replay: [C][K_global][H]
    int eps_f64;
    int final_mode;  // MppiFinal
    void* eps_dump;  // generate+dump: [C][K_local][H] of real
    double* costs;   // optional [C][K_local]
    unsigned int seed_lo, seed_hi, call_idx;
    unsigned int c_offset;  // global index of this handle's controller 0 (part of the Philox counter)
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
    // Host hand-over without a fence (single-level warp merge): every 8-byte result word goes to mapped host
    // memory as two self-validating cells (data half | epoch << 32), one 64-bit store each; the host polls the cells
    // (mppi_api.cu finish_host).  nullptr: u_out_host / info_host + fence.sys + done_host words as before.
    unsigned long long* host_cells;  // [C][2 * (H + 5)]: u_out[0..H), then the five words of mpcb_mppi_info
    // FINAL_PEER_EXCHANGE: device tables [G] of every rank's mailbox / flag base (own entry included)
    //   mailbox: [2 parity][G source ranks][C][kPartialHdr + H] doubles;  flags: [2][G][C] u32
    double* const* peer_mbox;
    unsigned int* const* peer_flags;
    int G, rank;
    unsigned int xepoch, pad2;     // exchange epoch (same on every rank), parity = xepoch & 1
    unsigned long long* debug_ts;  // optional [blocks][8] %globaltimer stamps (diagnostics)
    long long peer_ll_offset;      // FINAL_PEER_EXCHANGE: bytes from a mailbox base to its tagged-cell region; 0 = flag protocol
    int ws_cq, ws_debug;           // warp-specialised kernels: 4-step groups per producer->consumer chunk; diagnostics bits
    ModelConsts mc;
    double xu_inline[8 + kInlineHorizon];  // x[S] (S <= kMaxStateDim = 8) then u_n[H]
};

__device__ __forceinline__ double ll_as_double(long long v) { return __longlong_as_double(v); }
__device__ __forceinline__ long long double_as_ll(double v) { return __double_as_longlong(v); }

constexpr long long kNoArg = 0x7fffffffffffffffll;

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}
#define MPCB_TS(slot)                                                                       \
    do {                                                                                    \
        if (p.debug_ts != nullptr && threadIdx.x == 0) p.debug_ts[(size_t)blockIdx.x * 16 + (slot)] = globaltimer_ns(); \
    } while (0)

__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// arrival: the barrier before it orders the block's row stores before thread 0's release (cumulativity), so no
// block-wide fence is needed; acq_rel also makes it the acquire side of the last-arriver pattern
__device__ __forceinline__ unsigned int atom_add_acq_rel_gpu_u32(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
// arrival without a return value (nobody needs to know who was last when designated mergers wait for the count)
__device__ __forceinline__ void red_add_release_gpu_u32(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Order-preserving map double -> u64 (larger value, larger key); used to run max/argmax on redux.sync.
__device__ __forceinline__ unsigned long long order_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// Warp max over the lanes with `fin`, lowest lane on ties.  Returns (in every lane) the max value, the `tag` of
// the winning lane and the number of fin lanes; value -inf / tag kNoArg when no lane is fin.
__device__ __forceinline__ void warp_argmax(double v, long long tag, bool fin, double* vmax, long long* tmax, int* nfin) {
    const unsigned int full = 0xffffffffu;
    const unsigned long long key = order_key(v);
    const unsigned int hi = fin ? (unsigned int)(key >> 32) : 0u;
    const unsigned int mh = __reduce_max_sync(full, hi);
    const bool top = fin && hi == mh;
    const unsigned int lo = top ? (unsigned int)key : 0u;
    const unsigned int ml = __reduce_max_sync(full, lo);
    const unsigned int cand = __ballot_sync(full, top && (unsigned int)key == ml);
    const unsigned int finmask = __ballot_sync(full, fin);
    *nfin = __popc(finmask);
    if (cand != 0u) {
        const int src = __ffs(cand) - 1;
        int vlo = __double2loint(v), vhi = __double2hiint(v);
        vlo = __shfl_sync(full, vlo, src);
        vhi = __shfl_sync(full, vhi, src);
        *vmax = __hiloint2double(vhi, vlo);
        int tlo = (int)(tag & 0xffffffffll), thi = (int)(tag >> 32);
        tlo = __shfl_sync(full, tlo, src);
        thi = __shfl_sync(full, thi, src);
        *tmax = ((long long)thi << 32) | (unsigned int)tlo;
    } else {
        *vmax = -CUDART_INF;
        *tmax = kNoArg;
    }
}

// Warp max of v over the lanes with `has`, and the smallest idx among the lanes holding that max.
__device__ __forceinline__ void warp_max_minidx(double v, long long idx, bool has, double* vmax, long long* imin) {
    const unsigned int full = 0xffffffffu;
    double wm;
    long long wa;
    int dummy;
    warp_argmax(v, idx, has, &wm, &wa, &dummy);
    const bool holds = has && v == wm;
    const unsigned int ahi = holds ? (unsigned int)((unsigned long long)idx >> 32) : 0xffffffffu;
    const unsigned int mh = __reduce_min_sync(full, ahi);
    const unsigned int alo = (holds && ahi == mh) ? (unsigned int)idx : 0xffffffffu;
    const unsigned int ml = __reduce_min_sync(full, alo);
    *vmax = wm;
    *imin = (wa == kNoArg) ? kNoArg : (long long)(((unsigned long long)mh << 32) | ml);
}

// 4 consecutive elements of shared memory (16-byte aligned for float, 32-byte for double)
__device__ __forceinline__ void lds4(const float* p, float (&o)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void lds4(const double* p, double (&o)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *reinterpret_cast<const double2*>(p + 2);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ bool all_zero_bits(const float (&w)[4]) {
    return ((__float_as_uint(w[0]) | __float_as_uint(w[1]) | __float_as_uint(w[2]) | __float_as_uint(w[3])) << 1) == 0u;
}
__device__ __forceinline__ bool all_zero_bits(const double (&w)[4]) {
    return (((unsigned long long)__double_as_longlong(w[0]) | (unsigned long long)__double_as_longlong(w[1]) |
             (unsigned long long)__double_as_longlong(w[2]) | (unsigned long long)__double_as_longlong(w[3])) << 1) == 0ull;
}

// Row layout (PL doubles, PL even): [0] m = max finite c_k (or -inf), [1] argmax (bits), [2] sum_w, [3] n_finite
// (as a double: exact below 2^53), [4 .. 4+H) sum_w * v[t], optional zero pad.  Seen as column PAIRS (16 bytes):
// pair -1 = (m, argmax), pair 0 = (sum_w, n_finite), pair j >= 1 = (U[2j-2], U[2j-1]).
//
// Merges n_rows (<= kMergeFan) partial rows (row r at rows + r*row_stride) of one controller; the whole block
// participates.  scratch: kScratchDoubles doubles, part: mppi_part_doubles of shared memory.
//   final_mode FINAL_RANK_ROW : the merged, un-normalised row is written to out_row
//   final_mode FINAL_NORMALISE: u_out = sum_w*v / sum_w, info, status                      (src/mppi.rs:76-91)
// The merge is COLUMN-SPLIT over nm merger blocks: merger mi owns the pairs [1 + mi*cp, 1 + (mi+1)*cp) and every
// merger also sums pair 0 (it needs sum_w to normalise) and reduces the (m, argmax) headers.  What bounds a merge
// is the bytes one SM can pull from L2 (~32 B/clk): 148 rows of H = 100 are 122 KB — 2 us through one SM, a few
// hundred ns through 8-16 of them.  Merger 0 writes info/status (it owns u[0], the element src/mppi.rs:87 checks).
// Thread item (jl, q) sums the rows of partition q for local pair jl (jl = 0: pair 0); every global load of the
// merge is issued before anything is consumed; the partitions are then added in a fixed order, so the result does
// not depend on timing.
constexpr int kMergeBatch = 8;    // 16-byte loads in flight per thread (register budget: the kernel is capped at 128)
constexpr int kMergeMaxPart = 32;  // row partitions per column pair
constexpr int kMaxMergers = 16;

// One 8-byte result word as two cells in mapped host memory: (low half | epoch << 32), (high half | epoch << 32).  A
// 64-bit store is single-copy atomic, so a cell whose upper half shows this call's epoch carries this call's data: no
// fence.sys (which waits for the posted PCIe writes to be acknowledged: ~2 us each, two per step) and no completion word.
__device__ __forceinline__ void st_host_cells(unsigned long long* cells, int idx, unsigned long long bits, unsigned int epoch) {
    const unsigned long long e = (unsigned long long)epoch << 32;
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(cells + 2 * idx), "l"((bits & 0xffffffffull) | e) : "memory");
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(cells + 2 * idx + 1), "l"((bits >> 32) | e) : "memory");
}

struct MergeOut {
    double* u_out;
    double* u_out_host;
    mpcb_mppi_info* info;
    mpcb_mppi_info* info_host;
    double* out_row;
    double* const* copy_rows;  // FINAL_RANK_ROW: device table of further destinations of the row (peer mailboxes) ...
    long long copy_offset;     // ... written at copy_rows[r] + copy_offset for r < n_copies, r != copy_skip
    int n_copies, copy_skip;
    unsigned int* done_host;  // [nm] completion words (mapped host memory) or nullptr
    unsigned int epoch;
    unsigned long long* host_cells;  // MppiParams::host_cells (mppi_warp_merge only)
    int forced_status;
    unsigned long long* ts;
};

template <int BLOCK>
__device__ __noinline__ void mppi_merge_rows(const double* rows, long long row_stride, int n_rows, int H, int mi, int nm,
                                             double inv_lambda, int final_mode, const MergeOut& o, double* scratch,
                                             double* part, double* tot) {
    constexpr int NW = BLOCK / 32;
    constexpr int RPT = (kMergeFan + BLOCK - 1) / BLOCK;  // header rows per thread
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    double* red_m = scratch;                        // [16]
    long long* red_a = (long long*)(scratch + 16);  // [16]
    double* sc = scratch + 72;                      // [kMergeFan] per-row scale
    double2* part2 = reinterpret_cast<double2*>(part);
    unsigned long long* ts = o.ts;

    const int ncol2 = (H + 3) >> 1;             // pairs 0 .. ncol2-1
    const int cp = (ncol2 - 1 + nm - 1) / nm;   // pairs owned per merger
    const int p_lo = 1 + mi * cp;
    int p_hi = p_lo + cp;
    if (p_hi > ncol2) p_hi = ncol2;
    // local pairs: pair 0 (sum_w, n_finite), for mergers other than 0 also pair 1 (it holds u[0], which decides
    // MPCB_U_INVALID: every merger must reach the same status), then the owned ones
    const int n_extra = (mi == 0) ? 1 : 2;
    const int npl = n_extra + (p_hi > p_lo ? p_hi - p_lo : 0);
    auto pair_of = [&](int jl) { return jl < n_extra ? jl : p_lo + jl - n_extra; };
    int Hpm = 2, lgHpm = 1;
    while (Hpm < npl) { Hpm <<= 1; ++lgHpm; }
    int nq = BLOCK >> lgHpm;  // row partitions (power of two)
    if (nq < 1) nq = 1;
    if (nq > kMergeMaxPart) nq = kMergeMaxPart;
    const int rq = (n_rows + nq - 1) / nq;  // rows per partition
    const int items = Hpm * nq;
    if (ts != nullptr && tid == 0) ts[7] = globaltimer_ns();

    // ---- issue every load of the first pass ----
    double2 hd[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = tid + i * BLOCK;
        hd[i] = make_double2(-CUDART_INF, ll_as_double(-1ll));
        if (r < n_rows) hd[i] = __ldcg(reinterpret_cast<const double2*>(rows + (long long)r * row_stride));
    }
    double2 v0[kMergeBatch];
    {
        const int jl = tid & (Hpm - 1), q = tid >> lgHpm;
        const int pair = pair_of(jl);
        const int r0 = q * rq;
        const double2* col = reinterpret_cast<const double2*>(rows + 2 + 2 * pair + (long long)r0 * row_stride);
        const bool live = tid < items && jl < npl;
#pragma unroll
        for (int i = 0; i < kMergeBatch; ++i)
            v0[i] = (live && i < rq && r0 + i < n_rows) ? __ldcg(col + (long long)i * (row_stride >> 1)) : make_double2(0.0, 0.0);
    }
    if (ts != nullptr && tid == 0) ts[8] = globaltimer_ns();

    // ---- max / argmax over the headers (lowest sample index on ties) ----
    double m = -CUDART_INF;
    long long a = kNoArg;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const long long ai = double_as_ll(hd[i].y);
        if (ai >= 0 && (hd[i].x > m || (hd[i].x == m && ai < a))) { m = hd[i].x; a = ai; }
    }
    // equal maxima in different rows: the row order is not the sample order after a two-level merge, so ties are
    // resolved on the sample index itself — first the max value, then the min index among its holders
    auto reduce_max_minidx = [&](double& mv, long long& av, bool has) {
        double wm;
        long long wa;
        int dummy;
        warp_argmax(mv, av, has && av != kNoArg, &wm, &wa, &dummy);
        const bool holds = has && av != kNoArg && mv == wm;
        const unsigned int ahi = holds ? (unsigned int)((unsigned long long)av >> 32) : 0xffffffffu;
        const unsigned int mh = __reduce_min_sync(0xffffffffu, ahi);
        const unsigned int alo = (holds && ahi == mh) ? (unsigned int)av : 0xffffffffu;
        const unsigned int ml = __reduce_min_sync(0xffffffffu, alo);
        mv = wm;
        av = (wa == kNoArg) ? kNoArg : (long long)(((unsigned long long)mh << 32) | ml);
    };
    reduce_max_minidx(m, a, true);
    if (lane == 0) { red_m[wid] = m; red_a[wid] = a; }
    __syncthreads();
    if (ts != nullptr && tid == 0) ts[9] = globaltimer_ns();
    m = (lane < NW) ? red_m[lane] : -CUDART_INF;
    a = (lane < NW) ? red_a[lane] : kNoArg;
    reduce_max_minidx(m, a, lane < NW);
    const bool any = (a != kNoArg);
    // ---- per-row scale exp((m_r - m)/lambda); rows that saw no finite cost get 0 (0 * NaN keeps their poison) ----
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = tid + i * BLOCK;
        if (r < n_rows) sc[r] = (hd[i].x == -CUDART_INF) ? 0.0 : exp((hd[i].x - m) * inv_lambda);
    }
    __syncthreads();
    if (ts != nullptr && tid == 0) ts[3] = globaltimer_ns();
    // ---- column pairs: partition sums ----
    for (int item = tid; item < items; item += BLOCK) {
        const int jl = item & (Hpm - 1), q = item >> lgHpm;
        if (jl >= npl) continue;
        const int pair = pair_of(jl);
        const int r0 = q * rq;
        int r1 = r0 + rq;
        if (r1 > n_rows) r1 = n_rows;
        double ax = 0.0, ay = 0.0;
        const double2* col = reinterpret_cast<const double2*>(rows + 2 + 2 * pair);
        for (int rb = r0; rb < r1; rb += kMergeBatch) {
            double2 v[kMergeBatch];
            if (item == tid && rb == r0) {
#pragma unroll
                for (int i = 0; i < kMergeBatch; ++i) v[i] = v0[i];
            } else {
#pragma unroll
                for (int i = 0; i < kMergeBatch; ++i)
                    v[i] = (rb + i < r1) ? __ldcg(col + (long long)(rb + i) * (row_stride >> 1)) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int i = 0; i < kMergeBatch; ++i) {
                if (rb + i < r1) {
                    const double f = sc[rb + i];
                    ax += f * v[i].x;                       // scale 0 * NaN = NaN keeps the reference's poisoning
                    ay += (pair == 0) ? v[i].y : f * v[i].y;  // pair 0 = (sum_w, n_finite): the count is not scaled
                }
            }
        }
        part2[q * Hpm + jl] = make_double2(ax, ay);
    }
    if (ts != nullptr && tid == 0) ts[10] = globaltimer_ns();
    __syncthreads();
    if (ts != nullptr && tid == 0) ts[4] = globaltimer_ns();
    // ---- the nq (<= 32) partition sums of a pair: one warp per pair, fixed xor tree over the lanes ----
    double2* tot2 = reinterpret_cast<double2*>(tot);  // [npl]
    for (int jl = wid; jl < npl; jl += NW) {
        double2 t0 = (lane < nq) ? part2[lane * Hpm + jl] : make_double2(0.0, 0.0);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            t0.x += __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(t0.x), off),
                                     __shfl_xor_sync(0xffffffffu, __double2loint(t0.x), off));
            t0.y += __hiloint2double(__shfl_xor_sync(0xffffffffu, __double2hiint(t0.y), off),
                                     __shfl_xor_sync(0xffffffffu, __double2loint(t0.y), off));
        }
        if (lane == 0) tot2[jl] = t0;
    }
    __shared__ int s_status;
    __syncthreads();
    const double s = tot2[0].x, nf = tot2[0].y;  // pair 0: sum_w, n_finite
    // src/mppi.rs:69 no finite cost, :76-78 sum is zero, :87-89 u[0] not finite (element 0 only; local pair 1 is pair 1)
    if (tid == 0)
        s_status = (o.forced_status != MPCB_OK) ? o.forced_status
                   : (!any ? MPCB_NO_FINITE_COST : (s == 0.0 ? MPCB_SUM_ZERO : ((final_mode != FINAL_RANK_ROW && !finite_f64(tot2[1].x / s)) ? MPCB_U_INVALID : MPCB_OK)));
    __syncthreads();
    for (int jl = tid; jl < npl; jl += BLOCK) {
        const double ax = tot2[jl].x, ay = tot2[jl].y;
        const int pair = pair_of(jl);
        if (mi != 0 && jl == 1) continue;  // pair 1 belongs to merger 0
        if (final_mode == FINAL_RANK_ROW) {
            // the row goes to out_row and, for the cross-GPU exchange, straight into the peers' mailboxes (NVLink stores)
            for (int r = -1; r < o.n_copies; ++r) {
                if (r == o.copy_skip) continue;
                double* dst = (r < 0) ? o.out_row : o.copy_rows[r] + o.copy_offset;
                if (pair == 0) {
                    if (mi == 0) {
                        *reinterpret_cast<double2*>(dst) = make_double2(any ? m : -CUDART_INF, ll_as_double(any ? a : -1ll));
                        *reinterpret_cast<double2*>(dst + 2) = make_double2(ax, ay);
                    }
                } else {
                    // for odd H the last pair's second slot is the row's zero pad
                    *reinterpret_cast<double2*>(dst + 2 + 2 * pair) = make_double2(ax, ay);
                }
            }
        } else if (pair > 0) {
            const int t0 = 2 * pair - 2;
            // a controller whose status is not OK gets a zero row, like the reference's callers fall back to zeros
            // (examples/mppi4-non-liner-ukf.rs:80-86)
            const double u0 = (s_status == MPCB_OK) ? ax / s : 0.0;
            const double u1 = (s_status == MPCB_OK) ? ay / s : 0.0;
            o.u_out[t0] = u0;
            if (o.u_out_host) o.u_out_host[t0] = u0;
            if (t0 + 1 < H) {
                o.u_out[t0 + 1] = u1;
                if (o.u_out_host) o.u_out_host[t0 + 1] = u1;
            }
        }
    }
    if (final_mode == FINAL_RANK_ROW) return;
    if (ts != nullptr && tid == 0) ts[11] = globaltimer_ns();
    if (o.done_host) __threadfence_system();  // every writer orders its host stores before the completion word
    __syncthreads();
    if (tid == 0) {
        if (mi == 0) {
            mpcb_mppi_info out;
            out.status = s_status;
            out.reserved = 0;
            out.argmax = any ? a : -1ll;
            out.max = any ? m : 0.0;
            out.sum = s;
            out.n_finite = (long long)nf;
            *o.info = out;
            if (o.info_host) *o.info_host = out;
        }
        if (o.done_host) {
            // results first, then this merger's completion word the host spins on (system-scope release)
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int*>(o.done_host + mi) = o.epoch;
        }
    }
}

// w = exp(a) for the FP32 path: a <= 0 is computed in f64 by the caller; ex2.approx (2 ulp) on a*log2(e).
// Underflows to exactly 0 below a ~ -87; the f64 reference would keep weights down to e^-745, all of which
// are < 1e-37 of the block maximum.
__device__ __forceinline__ float fast_exp_neg(double a) {
    const float t = (float)(a * 1.4426950408889634074);
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}

// ---- cross-GPU exchange of the warp merge: data that carries its own tag (NCCL's "LL" idea) ----
// A value travels as ONE 16-byte store {lo32, tag, hi32, tag}: each 8-byte half is written atomically and validates
// itself, so the receiver needs neither a flag nor the sender an acknowledged release (the round-1 exchange spent
// 2.5 us per step in st.release.sys waiting for its row stores to be acknowledged over NVLink).  tag = the exchange
// epoch (the same on every rank); slots alternate with its parity, a peer can be at most one step ahead.
// Record of (source rank, controller): cells [0..5) = m, argmax, sum_w, n_finite, sum_w*v[0] — written once per rank
// by the warp of pair 1 in merger 0 — then two cells (the pair's two sums) per column pair 1 .. ncol2-1.
struct PeerLL {
    double* const* peer_base;  // device table [G]: every rank's mailbox base (own entry included)
    long long ll_offset;       // bytes from the base to the LL region [2 parity][G source ranks][C][ncell] cells
    int G, rank, C, ncell;
    unsigned int epoch;
};
__host__ __device__ inline int mppi_ll_cells(int H) { return 3 + 2 * ((H + 3) >> 1); }  // 5 + 2 * (ncol2 - 1)
__device__ __forceinline__ void ll_store(uint4* cell, double v, unsigned int tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "r"((unsigned int)b), "r"(tag),
                 "r"((unsigned int)(b >> 32)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ bool ll_load(const uint4* cell, unsigned int tag, double* v) {
    unsigned int a, ta, b, tb;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(ta), "=r"(b), "=r"(tb) : "l"(cell) : "memory");
    *v = __longlong_as_double((long long)(((unsigned long long)b << 32) | a));
    return ta == tag && tb == tag;
}

// Single-level merge without block barriers: ONE WARP merges one column pair of the n_rows (<= kMergeFan) rows.
// Lane l takes rows l, l + 32, ...: every load of the warp — the rows' headers (m, argmax), pair 0 (sum_w, n_finite),
// pair 1 (it holds u[0], the element src/mppi.rs:87 checks) and the warp's own pair — is issued before anything is
// consumed, the max/argmax runs on redux, each lane scales and sums its rows in order and a fixed xor tree adds the
// lanes, so the result does not depend on timing and every warp of every merger block computes the same m, sum_w,
// u[0] and status.  The round-1 merge (mppi_merge_rows: all threads of the block, five block barriers, a
// partition pass through shared memory) took ~3.3 us of dependent latency for 147 rows; this takes one L2 round
// trip, the exps of the lane's rows and one shuffle tree.
// final_mode FINAL_NORMALISE: u_out[2 pair - 2 .. 2 pair - 1] = sums / sum_w — or 0 when the controller's status is not
// OK, like the reference's callers fall back to zeros (examples/mppi4-non-liner-ukf.rs:80-86); FINAL_RANK_ROW: the
// un-normalised pair goes to o.out_row (the warp of pair 1 also writes the header and pair 0).  write_info: this
// warp also writes info/status (the warp of pair 1 in merger 0).  Returns the status.
// kFastExp (FP32 kernels): the row scales use ex2.approx like the block's own weights do; the FP64 kernels call exp().
// The rows are stored PAIR-MAJOR here (slot q of row r at rows + q * pair_stride + 2 r; q = 0: header (m, argmax),
// q = 1: pair 0, q = j + 1: pair j): the rows a warp reads for one slot are consecutive 16-byte cells, i.e. coalesced
// loads (row-major rows made every load of the warp touch 32 different 128-byte lines: 2.9 us for the 20 loads).
// ll != nullptr (FINAL_PEER_EXCHANGE): the warp's un-normalised sums go to every rank's mailbox as tagged cells, the warp
// collects the G ranks' records of its pair (lane r = rank r) and combines them exactly like rows of one GPU — every
// rank does the same arithmetic on the same records, so the ranks' results agree bitwise.
// RPL = rows per lane (n_rows <= 32 RPL): the caller picks the smallest instantiation that covers its rows — 4 RPL
// 16-byte loads are in flight per lane, and at the 128-register cap of the 512-thread kernels the 32 loads of RPL = 8 were
// issued in dependent batches (configs[1] has 147 rows: RPL = 5).
template <bool kFastExp, int RPL = kMergeFan / 32>
static __device__ __forceinline__ int mppi_warp_merge(const double* rows, long long pair_stride, int n_rows, int H, int pair,
                                                      double inv_lambda, int final_mode, const MergeOut& o, bool write_info,
                                                      const PeerLL* ll = nullptr, int ctrl = 0) {
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    double2 hd[RPL], p0[RPL], p1[RPL], pv[RPL];
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
        const int r = lane + 32 * i;
        hd[i] = make_double2(-CUDART_INF, ll_as_double(-1ll));
        p0[i] = p1[i] = pv[i] = make_double2(0.0, 0.0);
        if (r < n_rows) {
            const double* cell = rows + 2 * r;
            hd[i] = __ldcg(reinterpret_cast<const double2*>(cell));
            p0[i] = __ldcg(reinterpret_cast<const double2*>(cell + pair_stride));
            p1[i] = __ldcg(reinterpret_cast<const double2*>(cell + 2 * pair_stride));
            pv[i] = __ldcg(reinterpret_cast<const double2*>(cell + (long long)(pair + 1) * pair_stride));
        }
    }
    // max over the rows, lowest sample index among equal maxima
    double m = -CUDART_INF;
    long long a = kNoArg;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
        const long long ai = double_as_ll(hd[i].y);
        if (ai >= 0 && (hd[i].x > m || (hd[i].x == m && ai < a))) { m = hd[i].x; a = ai; }
    }
    if (o.ts != nullptr && lane == 0 && m > -1e300) o.ts[8] = globaltimer_ns();  // (diagnostics) headers have arrived
    double wm;
    long long wa;
    warp_max_minidx(m, a, a != kNoArg, &wm, &wa);
    const bool any = (wa != kNoArg);
    if (o.ts != nullptr && lane == 0) o.ts[9] = globaltimer_ns();
    // the lane's rows, scaled by exp((m_r - m)/lambda); rows without a finite cost get 0 (0 * NaN keeps their poison)
    // (branch-free: the exps of a lane's rows are independent chains the scheduler interleaves; exp(-746) == 0 in f64, and
    // a row without a finite cost has m_r = -inf, i.e. exactly the scale 0 that keeps 0 * NaN = NaN)
    double fsc[RPL];
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
        double arg = (hd[i].x - wm) * inv_lambda;
        arg = (arg > -746.0) ? arg : -746.0;  // also catches -inf and NaN arguments
        if constexpr (kFastExp) fsc[i] = (double)fast_exp_neg(arg);
        else fsc[i] = exp(arg);
    }
    double s = 0.0, nf = 0.0, a0 = 0.0, ax = 0.0, ay = 0.0;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
        if (lane + 32 * i < n_rows) {
            const double f = fsc[i];
            s += f * p0[i].x;
            nf += p0[i].y;
            a0 += f * p1[i].x;
            ax += f * pv[i].x;
            ay += f * pv[i].y;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s += __hiloint2double(__shfl_xor_sync(full, __double2hiint(s), off), __shfl_xor_sync(full, __double2loint(s), off));
        nf += __hiloint2double(__shfl_xor_sync(full, __double2hiint(nf), off), __shfl_xor_sync(full, __double2loint(nf), off));
        a0 += __hiloint2double(__shfl_xor_sync(full, __double2hiint(a0), off), __shfl_xor_sync(full, __double2loint(a0), off));
        ax += __hiloint2double(__shfl_xor_sync(full, __double2hiint(ax), off), __shfl_xor_sync(full, __double2loint(ax), off));
        ay += __hiloint2double(__shfl_xor_sync(full, __double2hiint(ay), off), __shfl_xor_sync(full, __double2loint(ay), off));
    }
    if (o.ts != nullptr && lane == 0 && s == s) o.ts[10] = globaltimer_ns();  // (diagnostics) sums reduced
    bool any_r = any;
    int forced = o.forced_status;
    if (ll != nullptr) {
        // ---- send: lane r writes this rank's record of the pair (and, once per rank, the header) into rank r's mailbox ----
        const unsigned int tag = ll->epoch, par = ll->epoch & 1u;
        const long long slot_out = ((long long)(par * ll->G + ll->rank) * ll->C + ctrl) * ll->ncell;
        if (lane < ll->G) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(ll->peer_base[lane]) + ll->ll_offset) + slot_out;
            ll_store(dst + 5 + 2 * (pair - 1), ax, tag);
            ll_store(dst + 6 + 2 * (pair - 1), ay, tag);
            if (write_info) {  // the header travels once per rank
                ll_store(dst + 0, any ? wm : -CUDART_INF, tag);
                ll_store(dst + 1, ll_as_double(any ? wa : -1ll), tag);
                ll_store(dst + 2, s, tag);
                ll_store(dst + 3, nf, tag);
                ll_store(dst + 4, a0, tag);
            }
        }
        // ---- receive: lane r polls rank r's record in this rank's own mailbox until every cell carries this step's tag ----
        double rm = -CUDART_INF, rarg = ll_as_double(-1ll), rs = 0.0, rnf = 0.0, ra0 = 0.0, rax = 0.0, ray = 0.0;
        bool timed_out = false;
        if (lane < ll->G) {
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(ll->peer_base[ll->rank]) + ll->ll_offset) +
                               ((long long)(par * ll->G + lane) * ll->C + ctrl) * ll->ncell;
            unsigned long long t_start = 0;
            for (unsigned int it = 0;; ++it) {
                bool ok = ll_load(src + 0, tag, &rm);
                ok = ll_load(src + 1, tag, &rarg) && ok;
                ok = ll_load(src + 2, tag, &rs) && ok;
                ok = ll_load(src + 3, tag, &rnf) && ok;
                ok = ll_load(src + 4, tag, &ra0) && ok;
                ok = ll_load(src + 5 + 2 * (pair - 1), tag, &rax) && ok;
                ok = ll_load(src + 6 + 2 * (pair - 1), tag, &ray) && ok;
                if (ok) break;
                if ((it & 63u) == 63u) {  // bounded: a missing peer becomes MPCB_PEER_TIMEOUT, not a hang
                    const unsigned long long now = globaltimer_ns();
                    if (t_start == 0) t_start = now;
                    else if (now - t_start > 20000000000ull) { timed_out = true; break; }
                }
            }
        }
        if (__any_sync(full, timed_out)) forced = MPCB_PEER_TIMEOUT;
        // ---- combine the ranks like rows: max, lowest sample index among equal maxima, scales, fixed xor tree ----
        long long ra = double_as_ll(rarg);
        if (ra < 0) ra = kNoArg;
        warp_max_minidx(rm, ra, lane < ll->G && ra != kNoArg, &wm, &wa);
        any_r = (wa != kNoArg);
        double arg = (rm - wm) * inv_lambda;
        arg = (arg > -746.0) ? arg : -746.0;
        double f;
        if constexpr (kFastExp) f = (double)fast_exp_neg(arg);
        else f = exp(arg);
        if (lane >= ll->G) f = 0.0;
        s = f * rs; nf = rnf; a0 = f * ra0; ax = f * rax; ay = f * ray;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s += __hiloint2double(__shfl_xor_sync(full, __double2hiint(s), off), __shfl_xor_sync(full, __double2loint(s), off));
            nf += __hiloint2double(__shfl_xor_sync(full, __double2hiint(nf), off), __shfl_xor_sync(full, __double2loint(nf), off));
            a0 += __hiloint2double(__shfl_xor_sync(full, __double2hiint(a0), off), __shfl_xor_sync(full, __double2loint(a0), off));
            ax += __hiloint2double(__shfl_xor_sync(full, __double2hiint(ax), off), __shfl_xor_sync(full, __double2loint(ax), off));
            ay += __hiloint2double(__shfl_xor_sync(full, __double2hiint(ay), off), __shfl_xor_sync(full, __double2loint(ay), off));
        }
    }
    if (final_mode == FINAL_RANK_ROW) {
        if (lane == 0) {
            for (int r = -1; r < o.n_copies; ++r) {
                if (r == o.copy_skip) continue;
                double* dst = (r < 0) ? o.out_row : o.copy_rows[r] + o.copy_offset;
                if (pair == 1) {
                    *reinterpret_cast<double2*>(dst) = make_double2(any ? wm : -CUDART_INF, ll_as_double(any ? wa : -1ll));
                    *reinterpret_cast<double2*>(dst + 2) = make_double2(s, nf);
                }
                *reinterpret_cast<double2*>(dst + 2 + 2 * pair) = make_double2(ax, ay);
            }
        }
        return MPCB_OK;
    }
    // src/mppi.rs:69 no finite cost, :76-78 sum is zero, :87-89 u[0] not finite (element 0 only)
    int status = forced;
    if (status == MPCB_OK) status = !any_r ? MPCB_NO_FINITE_COST : (s == 0.0 ? MPCB_SUM_ZERO : (!finite_f64(a0 / s) ? MPCB_U_INVALID : MPCB_OK));
    if (lane == 0) {
        const int t0 = 2 * pair - 2;
        const double u0 = (status == MPCB_OK) ? ax / s : 0.0;
        const double u1 = (status == MPCB_OK) ? ay / s : 0.0;
        o.u_out[t0] = u0;
        if (o.host_cells) st_host_cells(o.host_cells, t0, (unsigned long long)__double_as_longlong(u0), o.epoch);
        else if (o.u_out_host) o.u_out_host[t0] = u0;
        if (t0 + 1 < H) {
            o.u_out[t0 + 1] = u1;
            if (o.host_cells) st_host_cells(o.host_cells, t0 + 1, (unsigned long long)__double_as_longlong(u1), o.epoch);
            else if (o.u_out_host) o.u_out_host[t0 + 1] = u1;
        }
        if (write_info) {
            mpcb_mppi_info out;
            out.status = status;
            out.reserved = 0;
            out.argmax = any_r ? wa : -1ll;
            out.max = any_r ? wm : 0.0;
            out.sum = s;
            out.n_finite = (long long)nf;
            *o.info = out;
            if (o.host_cells) {
                static_assert(sizeof(mpcb_mppi_info) == 40, "five 8-byte words");
                st_host_cells(o.host_cells, H + 0, (unsigned long long)(unsigned int)out.status, o.epoch);
                st_host_cells(o.host_cells, H + 1, (unsigned long long)out.argmax, o.epoch);
                st_host_cells(o.host_cells, H + 2, (unsigned long long)__double_as_longlong(out.max), o.epoch);
                st_host_cells(o.host_cells, H + 3, (unsigned long long)__double_as_longlong(out.sum), o.epoch);
                st_host_cells(o.host_cells, H + 4, (unsigned long long)out.n_finite, o.epoch);
            } else if (o.info_host) {
                *o.info_host = out;
            }
        }
    }
    return status;
}

// Combines the G (<= 32) rank rows of the cross-GPU exchange: same result as mppi_merge_rows, done by ONE warp with no
// block barrier — lane r holds row r's header and (sum_w, n_finite), the scales are shuffled around, and lane jl sums
// its column pair over the G rows.  Called by warp 0 of a merger block; mi/nm select the merger's column slice.
static __device__ __noinline__ void mppi_combine_ranks_warp(const double* rows, long long row_stride, int G, int H, int mi, int nm,
                                                     double inv_lambda, const MergeOut& o) {
    const unsigned int full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int ncol2 = (H + 3) >> 1;
    const int cp = (ncol2 - 1 + nm - 1) / nm;
    const int p_lo = 1 + mi * cp;
    int p_hi = p_lo + cp;
    if (p_hi > ncol2) p_hi = ncol2;
    const bool has = lane < G;
    double2 hd = make_double2(-CUDART_INF, ll_as_double(-1ll)), p0 = make_double2(0.0, 0.0), p1 = make_double2(0.0, 0.0);
    if (has) {
        hd = __ldcg(reinterpret_cast<const double2*>(rows + (long long)lane * row_stride));
        p0 = __ldcg(reinterpret_cast<const double2*>(rows + (long long)lane * row_stride + 2));
        p1 = __ldcg(reinterpret_cast<const double2*>(rows + (long long)lane * row_stride + 4));  // pair 1 holds u[0]
    }
    // max over the ranks, lowest sample index among equal maxima
    double m = hd.x, wm;
    long long a = double_as_ll(hd.y), wa;
    if (a < 0) a = kNoArg;
    int dummy;
    warp_argmax(m, a, has && a != kNoArg, &wm, &wa, &dummy);
    const bool holds = has && a != kNoArg && m == wm;
    const unsigned int ahi = holds ? (unsigned int)((unsigned long long)a >> 32) : 0xffffffffu;
    const unsigned int mh = __reduce_min_sync(full, ahi);
    const unsigned int alo = (holds && ahi == mh) ? (unsigned int)a : 0xffffffffu;
    const unsigned int ml = __reduce_min_sync(full, alo);
    const bool any = (wa != kNoArg);
    const long long arg = any ? (long long)(((unsigned long long)mh << 32) | ml) : -1ll;
    const double sc = (!has || hd.x == -CUDART_INF) ? 0.0 : exp((hd.x - wm) * inv_lambda);
    // sum_w and n_finite in rank order (a serial chain over <= 32 shuffles keeps the order fixed)
    double s = 0.0, nf = 0.0, a0 = 0.0;
    for (int r = 0; r < G; ++r) {
        const double fr = __hiloint2double(__shfl_sync(full, __double2hiint(sc), r), __shfl_sync(full, __double2loint(sc), r));
        const double sr = __hiloint2double(__shfl_sync(full, __double2hiint(p0.x), r), __shfl_sync(full, __double2loint(p0.x), r));
        const double nr = __hiloint2double(__shfl_sync(full, __double2hiint(p0.y), r), __shfl_sync(full, __double2loint(p0.y), r));
        const double ur = __hiloint2double(__shfl_sync(full, __double2hiint(p1.x), r), __shfl_sync(full, __double2loint(p1.x), r));
        s += fr * sr;  // scale 0 * NaN = NaN keeps the reference's poisoning
        nf += nr;
        a0 += fr * ur;  // the same sum, in the same order, as the owner of pair 1 forms below: every merger sees the same u[0]
    }
    // src/mppi.rs:69 no finite cost, :76-78 sum is zero, :87-89 u[0] not finite (element 0 only)
    const int status = (o.forced_status != MPCB_OK) ? o.forced_status
                       : (!any ? MPCB_NO_FINITE_COST : (s == 0.0 ? MPCB_SUM_ZERO : (!finite_f64(a0 / s) ? MPCB_U_INVALID : MPCB_OK)));
    const int npairs = p_hi > p_lo ? p_hi - p_lo : 0;
    for (int j0 = 0; j0 < npairs || j0 == 0; j0 += 32) {  // at least one pass so that the shuffles below are uniform
        const int j = j0 + lane;
        const bool live = j < npairs;
        const int pair = p_lo + j;
        double ax = 0.0, ay = 0.0;
        for (int rb = 0; rb < G; rb += 8) {  // 8 rank rows at a time: all loads in flight before the first use
            double2 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                v[i] = (live && rb + i < G)
                           ? __ldcg(reinterpret_cast<const double2*>(rows + (long long)(rb + i) * row_stride + 2 + 2 * pair))
                           : make_double2(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (rb + i < G) {  // uniform
                    const int r = rb + i;
                    const double fr = __hiloint2double(__shfl_sync(full, __double2hiint(sc), r), __shfl_sync(full, __double2loint(sc), r));
                    if (live) {
                        ax += fr * v[i].x;
                        ay += fr * v[i].y;
                    }
                }
            }
        }
        if (live) {
            const int t0 = 2 * pair - 2;
            // a controller whose status is not OK gets a zero row (examples/mppi4-non-liner-ukf.rs:80-86)
            const double u0 = (status == MPCB_OK) ? ax / s : 0.0;
            const double u1 = (status == MPCB_OK) ? ay / s : 0.0;
            o.u_out[t0] = u0;
            if (o.u_out_host) o.u_out_host[t0] = u0;
            if (t0 + 1 < H) {
                o.u_out[t0 + 1] = u1;
                if (o.u_out_host) o.u_out_host[t0 + 1] = u1;
            }
        }
        if (j0 + 32 >= npairs) break;
    }
    // u[0] lives in pair 1 = the first pair of merger 0, lane 0
    if (o.done_host) __threadfence_system();
    __syncwarp();
    if (lane == 0) {
        if (mi == 0) {
            mpcb_mppi_info out;
            out.status = status;
            out.reserved = 0;
            out.argmax = arg;
            out.max = any ? wm : 0.0;
            out.sum = s;
            out.n_finite = (long long)nf;
            *o.info = out;
            if (o.info_host) *o.info_host = out;
        }
        if (o.done_host) {
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int*>(o.done_host + mi) = o.epoch;
        }
    }
}

// state dimension S of Mppi<N,K,S>: 4 for the built-in models; run-time compiled user models declare kStateDim (1..8)
constexpr int kMaxStateDim = 8;
template <typename M, typename = void>
struct ModelStateDim {
    static constexpr int value = 4;
};
template <typename M>
struct ModelStateDim<M, decltype((void)M::kStateDim)> {
    static constexpr int value = M::kStateDim;
};

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

// arithmetic type of the rollout: the storage type, or two packed FP32 samples per thread (SPT = 2)
template <typename real, int SPT>
struct ArithT {
    using type = real;
};
template <>
struct ArithT<float, 2> {
    using type = f2;
};


__host__ __device__ inline int mppi_pow2_horizon(int H, int* lg) {
    int hp = 8, l = 3;
    while (hp < H) { hp <<= 1; ++l; }
    if (lg) *lg = l;
    return hp;
}

__host__ __device__ inline int mppi_pow2_pairs(int H, int* lg) {
    int hp = 4, l = 2;
    while (hp < ((H + 3) >> 1)) { hp <<= 1; ++l; }
    if (lg) *lg = l;
    return hp;
}
__host__ __device__ inline int mppi_partial_len(int H) { return (kPartialHdr + H + 1) & ~1; }
// doubles of the partition-sum area: the merge keeps double2 per (pair, partition), the weighted sums one real
__host__ __device__ inline size_t mppi_part_doubles(int H, int block) {
    const int Hp = mppi_pow2_horizon(H, nullptr), Hp2 = mppi_pow2_pairs(H, nullptr);
    size_t a = (size_t)(block > Hp ? block : Hp);
    size_t b = 2 * (size_t)(block > Hp2 ? block : Hp2);  // partition sums, double2 each
    return a > b ? a : b;
}

// Shared-memory bytes of one block.
// `block` = samples per batch; vt = the kernel keeps the v tile [H][block + 4] (see the VT template parameter)
template <typename real>
__host__ __device__ inline size_t mppi_smem_bytes(int H, int block, bool vt = true) {
    const size_t npart = mppi_part_doubles(H, block);
    size_t dbl = ((size_t)H + 8 + kScratchDoubles + npart + 3) & ~(size_t)3;  // U_run (+ merge totals) + scratch + partition sums
    size_t rl = (size_t)2 * ((H + 3) & ~3) + block + (vt ? (size_t)H * (block + 4) : 0);  // su, sui, w_s, v_s
    size_t bytes = dbl * sizeof(double) + rl * sizeof(real);
    return (bytes + 15) & ~(size_t)15;
}


// The end of a block's work, shared by every rollout kernel: store the block's partial row
// [m, argmax, sum_w, n_finite, sum_w*v[0..H)], count the block in, and — for the designated merger blocks (or the last
// arriver) — merge the controller's rows and write u_out / info, the rank row, or run the cross-GPU exchange.
// Called by the first BLOCK threads of the block (whole warps beyond them have exited); U_run[H] holds the block's
// weighted sums (dead afterwards: the merge totals reuse it).
template <int BLOCK, bool kFastExp>
__device__ __forceinline__ void mppi_block_tail(const MppiParams& p, int c, int chunk, double m_run, long long arg_run, double S_run,
                                                long long nfin_run, double* U_run, double* scratch, double* part_d) {
    const int tid = threadIdx.x, wid = tid >> 5;
    const int H = p.H;
    const double lambda = p.lambda;
    MPCB_TS(1);
    // ---- partial row of this block, then the merge ----
    // Arrival counters only ever grow (launch `seq` expects seq*rows arrivals), so nothing has to be reset and any
    // number of blocks can wait on one counter.  With every block of the launch resident at once (p.mergers >= 1)
    // the first block of a group merges the group and the first p.mergers blocks of the controller share the final
    // merge column-wise — the same blocks, on the same SMs, every launch; otherwise (grids of several waves) the
    // last block to arrive merges alone.
    const int PL = p.PL;
    const double inv_lambda_m = 1.0 / lambda;
    double* tot_d = U_run;  // dead once the row is stored (the arrival barrier comes first): merge totals [<= (H+5)/2 double2]
    double* ctrl_rows = p.partial + (long long)c * (p.chunks + p.groups) * PL;
    double* my_row = ctrl_rows + (long long)chunk * PL;
    // single-level merge by warps (below): the level-0 rows are stored pair-major, see mppi_warp_merge
    const bool peer_ll = (p.final_mode == FINAL_PEER_EXCHANGE && p.peer_ll_offset != 0);
    const bool pair_major = (p.groups == 1 && p.mergers >= 1 && (p.final_mode != FINAL_PEER_EXCHANGE || peer_ll));
    const long long pair_stride = 2ll * p.chunks;
    if (pair_major) {
        for (int j = tid; j < (PL >> 1); j += BLOCK) {
            double2 v;
            if (j == 0) v = make_double2(m_run, ll_as_double(arg_run == kNoArg ? -1ll : arg_run));
            else if (j == 1) v = make_double2(S_run, (double)nfin_run);
            else v = make_double2(U_run[2 * j - 4], (2 * j - 3 < H) ? U_run[2 * j - 3] : 0.0);
            *reinterpret_cast<double2*>(ctrl_rows + (long long)j * pair_stride + 2 * chunk) = v;
        }
    } else {
        for (int t = tid; t < H; t += BLOCK) my_row[kPartialHdr + t] = U_run[t];
        if (tid == 0) {
            my_row[0] = m_run;
            my_row[1] = ll_as_double(arg_run == kNoArg ? -1ll : arg_run);
            my_row[2] = S_run;
            my_row[3] = (double)nfin_run;
            if (PL > kPartialHdr + H) my_row[kPartialHdr + H] = 0.0;
        }
    }
    __shared__ int s_go;
    unsigned int* cnt = p.counters + (long long)c * (p.groups + 1);
    const int g = chunk / p.group_size;
    const int g_first = g * p.group_size;
    int g_rows = p.chunks - g_first;
    if (g_rows > p.group_size) g_rows = p.group_size;
    unsigned long long* dbg = (p.debug_ts != nullptr) ? p.debug_ts + (size_t)blockIdx.x * 16 : nullptr;
    const bool spin = p.mergers >= 1;
    const int nm = spin ? p.mergers : 1;

    // publish this block's writes and count it in; in last-arriver mode returns whether it completed the count
    auto arrive = [&](int slot, int expect) -> bool {
        __syncthreads();
        if (tid == 0) {
            const bool drop = (p.ws_debug & 4) && slot == 0 && chunk == p.chunks - 1 && p.chunks > 1;  // fault injection (tests)
            s_go = !drop && (atom_add_acq_rel_gpu_u32(&cnt[slot], 1u) + 1u == p.seq * (unsigned int)expect);
        }
        __syncthreads();
        return s_go != 0;
    };
    // wait until `expect` blocks of this launch have arrived at `slot` (bounded, see MPCB_PEER_TIMEOUT)
    // A wait that gives up (the launch is not co-resident: another context, MPS, a debugger) is recorded: the merge then
    // reports MPCB_PEER_TIMEOUT and writes zeros instead of merging stale or half-written rows.
    __shared__ int s_await_timeout;
    if (tid == 0) s_await_timeout = 0;
    auto await = [&](int slot, int expect) {
        if (tid == 0) {
            const unsigned int want = p.seq * (unsigned int)expect;
            const unsigned long long t_start = globaltimer_ns();
            while (ld_acquire_gpu_u32(&cnt[slot]) != want) {
                if (globaltimer_ns() - t_start > 5000000000ull) { s_await_timeout = 1; break; }
            }
        }
        __syncthreads();  // thread 0's acquire + this barrier order every thread's (L2) loads after the arrivals
    };

    int mi = 0;  // which column slice of the final merge this block takes
    const double* final_rows = ctrl_rows;
    int final_n = p.chunks;
    MergeOut none;
    none.u_out = nullptr; none.u_out_host = nullptr; none.info = nullptr; none.info_host = nullptr; none.out_row = nullptr;
    none.copy_rows = nullptr; none.copy_offset = 0; none.n_copies = 0; none.copy_skip = -2;
    none.done_host = nullptr; none.epoch = 0u; none.forced_status = MPCB_OK; none.ts = nullptr; none.host_cells = nullptr;
    if (pair_major) {
        // ---- single-level merge by warps, no block barrier after the arrival (mppi_warp_merge): merger block mi owns
        // the column pairs [1 + mi*cp, 1 + (mi+1)*cp), one warp per pair; every warp waits for the arrivals itself ----
        __syncthreads();  // the block's row stores before thread 0's release (cumulative)
        // fault injection (tests): ws_debug & 4 makes the last block "never arrive", as if the launch were not co-resident
        if (tid == 0 && !((p.ws_debug & 4) && chunk == p.chunks - 1 && p.chunks > 1)) red_add_release_gpu_u32(&cnt[0], 1u);
        MPCB_TS(2);
        if (chunk >= nm) return;
        mi = chunk;
        const int ncol2 = (H + 3) >> 1;
        const int cp = (ncol2 - 1 + nm - 1) / nm;
        const int p_lo = 1 + mi * cp;
        int p_hi = p_lo + cp;
        if (p_hi > ncol2) p_hi = ncol2;
        MergeOut fo = none;
        fo.u_out = p.u_out + (long long)c * H;
        fo.u_out_host = p.u_out_host ? p.u_out_host + (long long)c * H : nullptr;
        fo.info = p.info + c;
        fo.info_host = p.info_host ? p.info_host + c : nullptr;
        fo.out_row = p.rank_partial ? p.rank_partial + (long long)c * PL : nullptr;
        fo.host_cells = p.host_cells ? p.host_cells + (long long)c * 2 * (H + 5) : nullptr;
        fo.epoch = p.epoch;
        if (wid < p_hi - p_lo || (wid == 0 && p.done_host)) {  // warps with a pair (warp 0 also signs off for the block)
            const unsigned int want = p.seq * (unsigned int)p.chunks;
            bool timed_out = false;
            if ((tid & 31) == 0) {
                // poll; the clock is read only every 64 polls (one poll is one L2 round trip)
                unsigned long long t_start = 0;
                for (unsigned int it = 0; ld_acquire_gpu_u32(&cnt[0]) != want; ++it) {
                    if ((it & 63u) == 63u) {
                        const unsigned long long now = globaltimer_ns();
                        if (t_start == 0) t_start = now;
                        else if (now - t_start > 5000000000ull) { timed_out = true; break; }
                    }
                }
            }
            timed_out = __shfl_sync(0xffffffffu, timed_out ? 1 : 0, 0) != 0;
            __syncwarp();  // lane 0's acquire + the warp barrier order every lane's (L2) loads after the arrivals
            // a launch that is not co-resident (another context on the GPU) must not merge half-written rows
            fo.forced_status = timed_out ? MPCB_PEER_TIMEOUT : MPCB_OK;
            if (dbg != nullptr && tid == 0) dbg[7] = globaltimer_ns();
            fo.ts = (wid == 0) ? dbg : nullptr;
            PeerLL pll;
            pll.peer_base = p.peer_mbox;
            pll.ll_offset = p.peer_ll_offset;
            pll.G = p.G; pll.rank = p.rank; pll.C = p.C; pll.ncell = mppi_ll_cells(H);
            pll.epoch = p.xepoch;
            for (int pair = p_lo + wid; pair < p_hi; pair += BLOCK / 32) {
                if (p.chunks <= 160 && !peer_ll)  // (with the cross-GPU exchange inlined twice the RPL = 5 copy measured +6 us per step)
                    mppi_warp_merge<kFastExp, 5>(ctrl_rows, pair_stride, p.chunks, H, pair, inv_lambda_m, peer_ll ? FINAL_NORMALISE : p.final_mode,
                                                 fo, mi == 0 && pair == 1, peer_ll ? &pll : nullptr, c);
                else
                    mppi_warp_merge<kFastExp>(ctrl_rows, pair_stride, p.chunks, H, pair, inv_lambda_m, peer_ll ? FINAL_NORMALISE : p.final_mode,
                                              fo, mi == 0 && pair == 1, peer_ll ? &pll : nullptr, c);
            }
            if (p.final_mode == FINAL_RANK_ROW && mi == 0 && tid == 0 && PL > kPartialHdr + H) fo.out_row[kPartialHdr + H] = 0.0;
            if (dbg != nullptr && tid == 0) dbg[11] = globaltimer_ns();
        }
        if (p.done_host && p.host_cells == nullptr && p.final_mode != FINAL_RANK_ROW) {
            // results of every warp of this merger first, then its completion word the host spins on
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                __threadfence_system();
                *reinterpret_cast<volatile unsigned int*>(p.done_host + mi) = p.epoch;
            }
        }
        MPCB_TS(5);
        return;
    }
    if (p.groups == 1) {
        const bool last = arrive(0, p.chunks);
        MPCB_TS(2);
        if (spin) {
            if (chunk >= nm) return;
            mi = chunk;
            await(0, p.chunks);
        } else {
            if (!last) return;
        }
    } else {
        double* group_rows = ctrl_rows + (long long)p.chunks * PL;
        const bool last = arrive(g, g_rows);
        MPCB_TS(2);
        const bool leader = spin ? (chunk == g_first) : last;
        const bool merger = spin && chunk < nm;
        if (!leader && !merger) return;
        if (leader) {
            if (spin) await(g, g_rows);
            MergeOut go = none;
            go.out_row = group_rows + (long long)g * PL;
            mppi_merge_rows<BLOCK>(ctrl_rows + (long long)g_first * PL, PL, g_rows, H, 0, 1, inv_lambda_m, FINAL_RANK_ROW, go,
                                   scratch, part_d, tot_d);
            if (tid == 0 && PL > kPartialHdr + H) go.out_row[kPartialHdr + H] = 0.0;
            if (tid == 0 && s_await_timeout) go.out_row[2] = CUDART_NAN;  // a group merged from incomplete rows poisons the result
            MPCB_TS(3);
            const bool last2 = arrive(p.groups, p.groups);
            if (!spin && !last2) return;
        }
        if (spin) {
            if (!merger) return;
            mi = chunk;
            await(p.groups, p.groups);
        }
        MPCB_TS(4);
        final_rows = group_rows;
        final_n = p.groups;
    }
    // ---- final merge of this controller (this block's column slice mi of nm) ----
    MergeOut fo = none;
    fo.u_out = p.u_out + (long long)c * H;
    fo.u_out_host = p.u_out_host ? p.u_out_host + (long long)c * H : nullptr;
    fo.info = p.info + c;
    fo.info_host = p.info_host ? p.info_host + c : nullptr;
    fo.done_host = p.done_host;
    fo.epoch = p.epoch;
    if (s_await_timeout) fo.forced_status = MPCB_PEER_TIMEOUT;
    if (p.final_mode != FINAL_PEER_EXCHANGE) {
        fo.out_row = p.rank_partial ? p.rank_partial + (long long)c * PL : nullptr;
        fo.ts = (p.groups == 1 && mi == 0) ? dbg : nullptr;
        mppi_merge_rows<BLOCK>(final_rows, PL, final_n, H, mi, nm, inv_lambda_m, p.final_mode, fo, scratch, part_d, tot_d);
        if (p.final_mode == FINAL_RANK_ROW && mi == 0 && tid == 0 && PL > kPartialHdr + H) fo.out_row[kPartialHdr + H] = 0.0;
        MPCB_TS(5);
        return;
    }
    // ---- cross-GPU exchange inside the kernel: merger mi puts its column slice of this rank's merged
    // row into every rank's mailbox slot [parity][rank][c] (peer stores over NVLink) and releases its flag
    // [parity][rank][c][mi] on every peer; it then acquires, for every rank, that rank's flag mi (its own slice)
    // and flag 0 (the slice holding m, argmax, sum_w, n) and combines the G rows exactly like the rows of one GPU.
    // Slots alternate with the exchange epoch: a peer can only be one step ahead, so two slots never collide. ----
    {
        const int G = p.G;
        const unsigned int par = p.xepoch & 1u;
        const long long slot_self = ((long long)(par * G + p.rank) * p.C + c);
        double* own_box = p.peer_mbox[p.rank];
        MergeOut ro = none;
        ro.out_row = own_box + slot_self * PL;
        ro.copy_rows = p.peer_mbox;
        ro.copy_offset = slot_self * PL;
        ro.n_copies = G;
        ro.copy_skip = p.rank;
        MPCB_TS(12);
        mppi_merge_rows<BLOCK>(final_rows, PL, final_n, H, mi, nm, inv_lambda_m, FINAL_RANK_ROW, ro, scratch, part_d, tot_d);
        MPCB_TS(13);
        // the stores of the whole block come before the barrier, thread r's system-scope release after it (cumulative)
        __syncthreads();
        for (int r = tid; r < G; r += BLOCK) st_release_sys_u32(p.peer_flags[r] + slot_self * kMaxMergers + mi, p.xepoch);
        // wait for every rank's slices of this step (bounded: a missing peer becomes MPCB_PEER_TIMEOUT, not a hang)
        __shared__ int s_timeout;
        if (tid == 0) s_timeout = 0;
        __syncthreads();
        MPCB_TS(14);
        for (int r = tid; r < 2 * G; r += BLOCK) {
            const int src = r >> 1, which = (r & 1) ? mi : 0;
            const unsigned int* f = p.peer_flags[p.rank] + ((long long)(par * G + src) * p.C + c) * kMaxMergers + which;
            const unsigned long long t_start = globaltimer_ns();
            while (ld_acquire_sys_u32(f) != p.xepoch) {
                if (globaltimer_ns() - t_start > 20000000000ull) { s_timeout = 1; break; }
            }
        }
        __syncthreads();
        MPCB_TS(15);
        fo.forced_status = (s_timeout || s_await_timeout) ? MPCB_PEER_TIMEOUT : MPCB_OK;
        const double* rank_rows = own_box + (long long)(par * G) * p.C * PL + (long long)c * PL;
        if (G <= 32) {
            if (wid == 0) mppi_combine_ranks_warp(rank_rows, (long long)p.C * PL, G, H, mi, nm, inv_lambda_m, fo);
        } else {
            mppi_merge_rows<BLOCK>(rank_rows, (long long)p.C * PL, G, H, mi, nm, inv_lambda_m, FINAL_NORMALISE, fo, scratch,
                                   part_d, tot_d);
        }
        MPCB_TS(5);
    }
}

// SPT = samples per thread.  SPT = 2 (FP32 only) packs the two samples' FP32 arithmetic into f32x2 instructions
// (f32x2.cuh): thread tid of a batch owns samples tid and BLOCK + tid of the batch's SB = 2*BLOCK samples.
// VT = keep the batch's control samples v[k][t] in shared memory for the weighted sums.  With VT = false nothing of
// size K x H exists anywhere: the weighted sums REGENERATE v[k][t] = clamp(u_n[t] + eps[k][t]) from the same Philox
// counters (or re-read the replay noise), for the samples whose weight is not zero only.  That frees the ~H*4 bytes
// of shared memory per resident sample, which is what limits the occupancy of long horizons (H = 200: 2 blocks of
// 128 per SM) — the multi-batch shapes use it, so that the packed kernels keep >= 8 warps per SM at any horizon.
template <template <typename> class ModelT, typename real, int BLOCK, int NOISE, int SPT, bool VT>
__global__ void __launch_bounds__(BLOCK, (512 / BLOCK > 0 ? 512 / BLOCK : 1)) mppi_rollout_kernel(const __grid_constant__ MppiParams p) {
    pdl_entry();
    static_assert(SPT == 1 || (SPT == 2 && sizeof(real) == 4), "two samples per thread is an FP32 layout");
    constexpr int NW = BLOCK / 32;
    constexpr int SB = BLOCK * SPT;  // samples per batch
    constexpr int LD = SB + 4;
    constexpr bool kExact = RealTraits<real>::kExact;
    constexpr bool kReplay = (NOISE == NOISE_REPLAY);
    using areal = typename ArithT<real, SPT>::type;
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int H = p.H, Hp = p.Hp, lgHp = p.lgHp;
    const int H4a = (H + 3) & ~3;
    double* scratch = reinterpret_cast<double*>(smem_raw);           // [kScratchDoubles] (even)
    double* part_d = scratch + kScratchDoubles;                      // [mppi_part_doubles] (even): 16-byte aligned
    real* part_r = reinterpret_cast<real*>(part_d);
    double* U_run = part_d + mppi_part_doubles(H, SB);               // [H + 8] (the merge totals reuse it)
    const int ndbl = (H + 8 + kScratchDoubles + (int)mppi_part_doubles(H, SB) + 3) & ~3;
    real* su = reinterpret_cast<real*>(scratch + ndbl);              // [H4a]  u_n (16-byte aligned rows from here on)
    real* sui = su + H4a;                                            // [H4a]  u_n * sigma^-2
    real* w_s = sui + H4a;                                           // [SB]
    real* v_s = w_s + SB;                                            // [H][LD] (VT only)

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int c = blockIdx.x / p.chunks;      // controller
    const int chunk = blockIdx.x % p.chunks;  // sample range of this controller

    MPCB_TS(0);
    if (p.debug_ts != nullptr && threadIdx.x == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.debug_ts[(size_t)blockIdx.x * 16 + 6] = smid + 1;
    }
    // ---- prologue: model constants, x0, u_n ----
    ModelT<areal> model;
    model.load(p.mc);
    constexpr int S = ModelStateDim<ModelT<areal>>::value;
    static_assert(S >= 1 && S <= kMaxStateDim, "state dimension out of range");
    static_assert(SPT == 1 || S == 4, "the packed kernels are written for the built-in four-state models");
    real x0[S];
    // FP32 path: a NaN in x or u_n reaches every sample's cost in the reference (src/mppi.rs:53-61) and must poison the
    // step; a NaN cost that FP32 overflow produced from clean inputs must not (see the weights below)
    __shared__ int s_in_nan;
    if (tid == 0) s_in_nan = 0;
    __syncthreads();
    bool in_bad = false;
    if (p.use_inline) {
#pragma unroll
        for (int i = 0; i < S; ++i) {
            x0[i] = (real)p.xu_inline[i];
            in_bad = in_bad || (p.xu_inline[i] != p.xu_inline[i]);
        }
        for (int t = tid; t < H; t += BLOCK) {
            const double ut = p.xu_inline[S + t];
            in_bad = in_bad || (ut != ut);
            su[t] = (real)ut;
            sui[t] = (real)(ut * p.inv_var);
        }
    } else {
#pragma unroll
        for (int i = 0; i < S; ++i) {
            const double xi = p.x[(long long)c * S + i];
            x0[i] = (real)xi;
            in_bad = in_bad || (xi != xi);
        }
        for (int t = tid; t < H; t += BLOCK) {
            const double ut = p.u[(long long)c * H + t];
            in_bad = in_bad || (ut != ut);
            su[t] = (real)ut;
            sui[t] = (real)(ut * p.inv_var);
        }
    }
    if (in_bad) s_in_nan = 1;
    for (int t = tid; t < H; t += BLOCK) U_run[t] = 0.0;
    const real lo = (real)p.lo, hi = (real)p.hi;
    const float neg2s2ln2 = (float)(-2.0 * p.std_dev * p.std_dev * 0.693147180559945309417);
    const double lambda = p.lambda;
    const double inv_lambda = 1.0 / lambda;

    double m_run = -CUDART_INF, S_run = 0.0;
    long long arg_run = kNoArg, nfin_run = 0;

    double* red_m = scratch;                        // [16]
    long long* red_a = (long long*)(scratch + 16);  // [16]
    int* red_n = (int*)(scratch + 32);              // [16]
    double* red_s = scratch + 48;                   // [24]: [0..15] warp sums, [16] rescale

    // this block's range of whole warps (of 32 samples), walked in batches of NW*SPT warps
    const long long w_begin = p.W * chunk / p.chunks;
    const long long w_end = p.W * (chunk + 1) / p.chunks;
    __syncthreads();

    for (long long wb = w_begin; wb < w_end; wb += NW * SPT) {
        // sample s of this thread lives in sample-warp wb + s*NW + wid (column s*BLOCK + tid of the batch)
        long long kl[SPT], kg[SPT];  // local / global sample index (global: Philox counter, replay row, argmax)
        bool live[SPT], valid[SPT];  // whole warps beyond the range skip the rollout
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const long long gw = wb + (long long)s * NW + wid;
            live[s] = gw < w_end;
            kl[s] = gw * 32 + lane;
            valid[s] = live[s] && kl[s] < p.K_local;
            kg[s] = p.k_offset + kl[s];
        }

        if constexpr (kReplay && VT) {
            // coalesced tile load: the batch's consecutive sample rows of H values -> v_s[t][k]
            const long long k_first = wb * 32;
            long long nrows = (w_end - wb) * 32;
            if (nrows > SB) nrows = SB;
            if (nrows > p.K_local - k_first) nrows = p.K_local - k_first;
            const long long base = ((long long)c * p.K_global + p.k_offset + k_first) * H;
            const int n_live = (int)nrows * H;
            for (int i = tid; i < n_live; i += BLOCK) {
                const real e = p.eps_f64 ? (real) reinterpret_cast<const double*>(p.eps)[base + i]
                                         : (real) reinterpret_cast<const float*>(p.eps)[base + i];
                const int kr = i / H, t = i - kr * H;
                v_s[t * LD + kr] = e;
            }
            __syncthreads();
        }

        double ck[SPT];
        bool eps_nan[SPT];
#pragma unroll
        for (int s = 0; s < SPT; ++s) { ck[s] = -CUDART_INF; eps_nan[s] = false; }
        if (live[0]) {  // live[1] implies live[0]
            // ---- PASS 1+2: noise, clamp, rollout, cost ----
            real* vcol = v_s + tid;
            // four N(0, sigma^2) draws of sample s for steps t0..t0+3
            auto draw4 = [&](int s, int t0, real(&e)[4]) {
                if constexpr (kReplay && VT) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        e[i] = (valid[s] && t0 + i < H) ? vcol[(t0 + i) * LD + s * BLOCK] : (real)0;
                        eps_nan[s] = eps_nan[s] || (e[i] != e[i]);
                    }
                } else if constexpr (kReplay) {
                    // no tile: the thread reads its own row of the caller's noise (verification mode)
                    const long long row = ((long long)c * p.K_global + kg[s]) * H;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        e[i] = (real)0;
                        if (valid[s] && t0 + i < H)
                            e[i] = p.eps_f64 ? (real) reinterpret_cast<const double*>(p.eps)[row + t0 + i]
                                             : (real) reinterpret_cast<const float*>(p.eps)[row + t0 + i];
                        eps_nan[s] = eps_nan[s] || (e[i] != e[i]);
                    }
                } else {
                    const unsigned int c0 = (unsigned int)(kg[s] & 0xffffffffll);
                    const unsigned int khi = (unsigned int)((kg[s] >> 32) & 0xffff) << 16;
                    const Philox4 r = philox4x32(c0, p.call_idx, (unsigned int)c + p.c_offset, (unsigned int)(t0 >> 2) | khi,
                                                    p.seed_lo, p.seed_hi);
                    float z[4];
                    philox_normal4(r, neg2s2ln2, z);
#pragma unroll
                    for (int i = 0; i < 4; ++i) e[i] = (real)z[i];
                    if constexpr (NOISE == NOISE_GENERATE_DUMP) {
                        if (valid[s]) {
                            real* dump = reinterpret_cast<real*>(p.eps_dump) + ((long long)c * p.K_local + kl[s]) * H;
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (t0 + i < H) dump[t0 + i] = e[i];
                        }
                    }
                }
            };
            const int H4 = H & ~3;
            double J[SPT], CT[SPT];
#pragma unroll
            for (int s = 0; s < SPT; ++s) { J[s] = 0.0; CT[s] = 0.0; }

            if constexpr (SPT == 1) {
                real x[S];
#pragma unroll
                for (int i = 0; i < S; ++i) x[i] = x0[i];
                real cw = (real)0, cu = (real)0;
                // one rollout step: v = clamp(u_n[t] + eps), x <- dynamics(x, v), cost, control term
                auto step = [&](int t, real eps, real ut, real uit) {
                    real v = ut + eps;
                    if constexpr (kExact) v = clampr(v, lo, hi);  // f64::clamp, NaN stays NaN
                    else v = fminf(fmaxf(v, lo), hi);
                    if constexpr (VT) vcol[t * LD] = v;
                    model.step(x, v);
                    if constexpr (kExact) {
                        const real ct = model.cost(x);
                        J[0] = J[0] + ct;   // :57 c + cost(x_n)
                        CT[0] += uit * v;   // :60 (u*inv)*v, summed in order
                    } else {
                        cw = model.cost.acc(x, cw);
                        cu = fmaf(uit, v, cu);
                    }
                };
                // Software pipeline: the noise of group g+1 (independent of the state) is drawn while group g rolls
                // out, so the scheduler has Philox/Box-Muller work to fill the dependency stalls of the serial chain.
                real e[4];
                draw4(0, 0, e);
                for (int t0 = 0; t0 < H4; t0 += 4) {
                    real en[4], u4[4], ui4[4];
                    lds4(su + t0, u4);
                    lds4(sui + t0, ui4);
                    draw4(0, t0 + 4, en);  // the last prefetch (t0 + 4 >= H4) feeds the tail below or is discarded
#pragma unroll
                    for (int i = 0; i < 4; ++i) step(t0 + i, e[i], u4[i], ui4[i]);
                    if constexpr (!kExact) {
                        // FP32 partial sums are flushed into the FP64 accumulators every kFlushSteps steps
                        if ((t0 & (kFlushSteps - 1)) == kFlushSteps - 4) {
                            J[0] += (double)cw;
                            CT[0] += (double)cu;
                            cw = (real)0;
                            cu = (real)0;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) e[i] = en[i];
                }
                if (H4 < H) {
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (H4 + i < H) step(H4 + i, e[i], su[H4 + i], sui[H4 + i]);
                }
                if constexpr (!kExact) {
                    J[0] += (double)cw;  // whatever is left since the last flush
                    CT[0] += (double)cu;
                }
            } else {
                // two samples per thread, FP32 arithmetic packed (the second sample of a thread whose second warp is
                // beyond the range rolls out harmlessly and gets weight 0)
                f2 x[4] = {splat2(x0[0]), splat2(x0[1]), splat2(x0[2]), splat2(x0[3])};
                f2 cw = splat2(0.0f), cu = splat2(0.0f);
                auto step2 = [&](int t, f2 eps, float ut, float uit) {
                    const f2 v = clamp2(add2(splat2(ut), eps), lo, hi);
                    if constexpr (VT) {
                        float vl, vh;
                        un2(v, vl, vh);
                        vcol[t * LD] = vl;
                        vcol[t * LD + BLOCK] = vh;
                    }
                    model.step(x, v);
                    cw = model.cost.acc(x, cw);
                    cu = fma2(splat2(uit), v, cu);
                };
                auto draw4x2 = [&](int t0, f2(&e)[4]) {
                    float ea[4], eb[4];
                    draw4(0, t0, ea);
                    draw4(1, t0, eb);
#pragma unroll
                    for (int i = 0; i < 4; ++i) e[i] = mk2(ea[i], eb[i]);
                };
                auto flush = [&]() {
                    float a, b;
                    un2(cw, a, b);
                    J[0] += (double)a;
                    J[1] += (double)b;
                    un2(cu, a, b);
                    CT[0] += (double)a;
                    CT[1] += (double)b;
                    cw = splat2(0.0f);
                    cu = splat2(0.0f);
                };
                f2 e[4];
                draw4x2(0, e);
                for (int t0 = 0; t0 < H4; t0 += 4) {
                    f2 en[4];
                    float u4[4], ui4[4];
                    lds4(su + t0, u4);
                    lds4(sui + t0, ui4);
                    draw4x2(t0 + 4, en);
#pragma unroll
                    for (int i = 0; i < 4; ++i) step2(t0 + i, e[i], u4[i], ui4[i]);
                    if ((t0 & (kFlushSteps - 1)) == kFlushSteps - 4) flush();
#pragma unroll
                    for (int i = 0; i < 4; ++i) e[i] = en[i];
                }
                if (H4 < H) {
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (H4 + i < H) step2(H4 + i, e[i], su[H4 + i], sui[H4 + i]);
                }
                flush();
            }
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                ck[s] = -J[s] - CT[s];  // :61
                if constexpr (!kExact) {
                    // FP32 rollouts overflow (inf, then inf - inf = NaN) where the f64 reference still holds a huge finite
                    // cost whose weight underflows to exactly 0: with clean inputs a NaN cost is that case -> -inf, weight 0.
                    // A NaN that came in through x, u_n or the sample's replay noise poisons the sums like the reference.
                    if (s_in_nan != 0 || eps_nan[s]) ck[s] = CUDART_NAN;
                    else if (ck[s] != ck[s]) ck[s] = -CUDART_INF;
                }
                if (p.costs != nullptr && valid[s]) p.costs[(long long)c * p.K_local + kl[s]] = ck[s];
            }
        }

        // ---- PASS 3: block max over finite c_k, lowest index on ties ----
        bool fin[SPT];
        int nf_thread = 0;
        double tm = -CUDART_INF;  // this thread's best finite sample (the lower index wins a tie)
        long long ta = kNoArg;
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            fin[s] = valid[s] && finite_f64(ck[s]);
            nf_thread += fin[s] ? 1 : 0;
            if (fin[s] && (ta == kNoArg || ck[s] > tm)) { tm = ck[s]; ta = kg[s]; }
        }
        double bm;
        long long ba;
        warp_max_minidx(tm, ta, ta != kNoArg, &bm, &ba);
        int bn = __reduce_add_sync(0xffffffffu, nf_thread);
        if (lane == 0) { red_m[wid] = bm; red_a[wid] = ba; red_n[wid] = bn; }
        __syncthreads();
        {
            // every warp reduces the NW warp results again (lane l holds warp l's entry)
            const bool has = lane < NW;
            const double wm = has ? red_m[lane] : -CUDART_INF;
            const long long wa = has ? red_a[lane] : kNoArg;
            const int wn = has ? red_n[lane] : 0;
            warp_max_minidx(wm, wa, has && wa != kNoArg, &bm, &ba);
            bn = __reduce_add_sync(0xffffffffu, wn);
        }
        const double m_old = m_run;
        if (bm > m_run) { m_run = bm; arg_run = ba; }
        nfin_run += bn;

        // ---- PASS 4-5: weights against the running max, rescale of what was accumulated so far ----
        real wsum = (real)0;
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            real w;
            if constexpr (kExact) {
                // reference semantics: exp((c - max)/lambda) for every sample, NaN/+inf poison the sums (:71-74);
                // exp(a) is exactly 0 in f64 for a < -745.14, so skipping the call there is bit-identical
                const double arg = (ck[s] - m_run) / lambda;
                if (!valid[s] || ck[s] == -CUDART_INF) w = 0.0;
                else w = (arg < -746.0) ? 0.0 : exp(arg);
            } else {
                // natural IEEE semantics like the reference (src/mppi.rs:71-74): -inf -> 0, NaN / +inf poison the sums
                // (FP32-overflow NaNs were turned into -inf above)
                if (!valid[s] || ck[s] == -CUDART_INF) w = 0.0f;
                else if (m_run == -CUDART_INF) w = (float)(ck[s] - ck[s]);  // no finite cost yet: NaN / +inf still poison
                else w = fast_exp_neg((ck[s] - m_run) * inv_lambda);
            }
            w_s[s * BLOCK + tid] = w;
            wsum += w;
        }
        {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                if constexpr (kExact) wsum += shfl_down_f64(wsum, off);
                else wsum += __shfl_down_sync(0xffffffffu, wsum, off);
            }
            if (lane == 0) red_s[wid] = (double)wsum;
        }
        if (tid == 0) red_s[16] = (m_old == -CUDART_INF) ? 0.0 : exp((m_old - m_run) / lambda);  // block-uniform rescale
        __syncthreads();
        const double resc = red_s[16];
        double bs = red_s[0];
#pragma unroll
        for (int wI = 1; wI < NW; ++wI) bs += red_s[wI];
        S_run = S_run * resc + bs;

        // ---- PASS 6: sum_k w_k * v[k][t].  Item (t, q): column t, sample partition q; 128-bit shared loads of
        // four weights (broadcast) and four samples (conflict-free); groups whose weights are all zero —
        // samples far from the best, and warps beyond the range — are skipped, which leaves the sums unchanged. ----
        if constexpr (VT) {
            int nq = BLOCK >> lgHp;
            if (nq < 1) nq = 1;
            if (nq > SB / 8) nq = SB / 8;
            const int kper = SB / nq;
            const int items = Hp * nq;
            for (int item = tid; item < items; item += BLOCK) {
                const int t = item & (Hp - 1), q = item >> lgHp;
                if (t >= H) continue;
                const real* col = v_s + t * LD + q * kper;
                const real* wq = w_s + q * kper;
                real a0 = (real)0, a1 = (real)0, a2 = (real)0, a3 = (real)0;
#pragma unroll 2
                for (int k = 0; k < kper; k += 4) {
                    real w4[4], v4[4];
                    lds4(wq + k, w4);
                    if (all_zero_bits(w4)) continue;
                    lds4(col + k, v4);
                    a0 += w4[0] * v4[0];
                    a1 += w4[1] * v4[1];
                    a2 += w4[2] * v4[2];
                    a3 += w4[3] * v4[3];
                }
                part_r[q * Hp + t] = (a0 + a1) + (a2 + a3);
            }
            __syncthreads();
            for (int t = tid; t < H; t += BLOCK) {
                real acc = part_r[t];
                for (int q = 1; q < nq; ++q) acc += part_r[q * Hp + t];
                U_run[t] = U_run[t] * resc + (double)acc;
            }
            __syncthreads();
        } else {
            // ---- PASS 6 without a v tile: thread item (g, j) takes the four steps t0 = 4j .. 4j+3 of every G-th sample
            // of the batch (k = g, g+G, ...), skips the samples whose weight is zero, and rebuilds the others' controls
            // from the Philox block (global sample index, t0/4) the rollout used — or re-reads the replay noise.  The
            // G groups' sums are then added in order, so the result does not depend on timing. ----
            const int Hq = (H + 3) >> 2;
            int G = BLOCK / Hq;
            const int cap = (int)(mppi_part_doubles(H, SB) * sizeof(double) / sizeof(real)) / H4a;  // part_r rows available
            if (G > cap) G = cap;
            if (G > 8) G = 8;
            if (G < 1) G = 1;
            for (int item = tid; item < G * Hq; item += BLOCK) {
                const int g = item / Hq, j = item - g * Hq, t0 = 4 * j;
                real acc[4] = {(real)0, (real)0, (real)0, (real)0};
                real u4[4];
                lds4(su + t0, u4);
                for (int k = g; k < SB; k += G) {
                    const real wk = w_s[k];
                    real w1[4] = {wk, (real)0, (real)0, (real)0};
                    if (all_zero_bits(w1)) continue;
                    const long long kgk = p.k_offset + wb * 32 + k;  // column k of the batch is local sample wb*32 + k
                    real e[4];
                    if constexpr (kReplay) {
                        const long long row = ((long long)c * p.K_global + kgk) * H;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            e[i] = (real)0;
                            if (t0 + i < H)
                                e[i] = p.eps_f64 ? (real) reinterpret_cast<const double*>(p.eps)[row + t0 + i]
                                                 : (real) reinterpret_cast<const float*>(p.eps)[row + t0 + i];
                        }
                    } else {
                        const unsigned int c0 = (unsigned int)(kgk & 0xffffffffll);
                        const unsigned int khi = (unsigned int)((kgk >> 32) & 0xffff) << 16;
                        const Philox4 r = philox4x32(c0, p.call_idx, (unsigned int)c + p.c_offset, (unsigned int)j | khi, p.seed_lo, p.seed_hi);
                        float z[4];
                        philox_normal4(r, neg2s2ln2, z);
#pragma unroll
                        for (int i = 0; i < 4; ++i) e[i] = (real)z[i];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        real v = u4[i] + e[i];
                        if constexpr (kExact) v = clampr(v, lo, hi);
                        else v = fminf(fmaxf(v, lo), hi);
                        acc[i] += wk * v;
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) part_r[g * H4a + t0 + i] = acc[i];
            }
            __syncthreads();
            for (int t = tid; t < H; t += BLOCK) {
                real acc = part_r[t];
                for (int g = 1; g < G; ++g) acc += part_r[g * H4a + t];
                U_run[t] = U_run[t] * resc + (double)acc;
            }
            __syncthreads();
        }
    }

    mppi_block_tail<BLOCK, !kExact>(p, c, chunk, m_run, arg_run, S_run, nfin_run, U_run, scratch, part_d);
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
    __shared__ __align__(16) double part[2 * (kMaxHorizon > BLOCK ? kMaxHorizon : BLOCK)];
    __shared__ __align__(16) double tot[kMaxHorizon + 8];
    const int c = blockIdx.x;
    const int PL = mppi_partial_len(p.H);
    MergeOut o;
    o.u_out = p.u_out + (long long)c * p.H;
    o.u_out_host = p.u_out_host ? p.u_out_host + (long long)c * p.H : nullptr;
    o.info = p.info + c;
    o.info_host = p.info_host ? p.info_host + c : nullptr;
    o.out_row = nullptr;
    o.copy_rows = nullptr; o.copy_offset = 0; o.n_copies = 0; o.copy_skip = -2;
    o.done_host = p.done_host;
    o.epoch = p.epoch;
    o.forced_status = MPCB_OK;
    o.ts = nullptr;
    if (p.G <= 32) {
        if (threadIdx.x < 32) mppi_combine_ranks_warp(p.rows + (long long)c * PL, (long long)p.C * PL, p.G, p.H, 0, 1, 1.0 / p.lambda, o);
    } else {
        mppi_merge_rows<BLOCK>(p.rows + (long long)c * PL, (long long)p.C * PL, p.G, p.H, 0, 1, 1.0 / p.lambda, FINAL_NORMALISE, o,
                               scratch, part, tot);
    }
}

// kernel entry table (defined in mppi_f32*.cu / mppi_f64.cu)
using MppiKernelFn = void (*)(const MppiParams);
MppiKernelFn mppi_kernel_f32(int model_id, int block, int noise, int vt);
MppiKernelFn mppi_kernel_f32x2(int model_id, int block, int noise, int vt);  // two samples per thread (block threads = 2*block samples)
MppiKernelFn mppi_kernel_f64(int model_id, int block, int noise, int vt);
MppiKernelFn mppi_kernel_f64fast(int model_id, int block, int noise, int vt);  // MPCB_F64_FAST (models.cuh, Model*F)

}  // namespace mpcb
