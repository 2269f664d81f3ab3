// mppi_api.cu — C ABI of the MPPI controller (mpcb_mppi_*), replacing mpc::mppi::Mppi (src/mppi.rs:7-92).
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <time.h>
#include <unistd.h>

#include <new>

#include "mppi_ws_kernel.cuh"
#include "mppi_short_kernel.cuh"
#include "mppi_rtc.h"
#include "nccl_shim.h"

#include <string>

using namespace mpcb;

struct mpcb_mppi {
    mpcb_mppi_cfg cfg;
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    long long K_local = 0, k_offset = 0;
    int H = 0, C = 0, PL = 0;
    int S = 4;  // state dimension (4 for the built-in models; 1..8 for user-supplied ones)
    int block = 0, chunks = 0, group_size = 0, groups = 0;
    int Hp = 8, lgHp = 3;
    int spt = 1;            // samples per thread (2: packed f32x2 kernels)
    unsigned int c_offset = 0;  // mpcb_mppi_set_controller_offset
    int ws_variant = -1;    // >= 0: warp-specialised kernel, index into kWsVariants (mppi_ws_kernel.cuh)
    int ws_cq = 5;          // its producer->consumer chunk, in 4-step groups (20 steps: measured best of 1/2/5)
    int ws_debug = 0;       // MPCB_MPPI_WS_DEBUG: 1 = producers idle, 2 = consumers idle (timing decomposition only)
    int block_samples = 0;  // samples per block and batch = block * spt
    int mergers = 0;        // blocks sharing the final merge (0: last arriver merges alone)
    unsigned int seq = 0;   // launches so far (arrival counters are monotonic)
    long long W = 0;  // warps of 32 samples in K_local
    size_t smem = 0;
    ModelConsts mc;
    MppiKernelFn k_noise[3] = {nullptr, nullptr, nullptr};  // indexed by MppiNoise
    // user-supplied model (mpcb_mppi_create_user): CUDA source compiled at create time, kernels live in `rtc`
    bool user = false;
    std::string user_src;
    RtcModule rtc;
    // device
    double* d_x = nullptr;
    double* d_u = nullptr;
    double* d_u_out = nullptr;
    double* d_partial = nullptr;
    double* d_rank_partial = nullptr;
    double* d_gather = nullptr;
    unsigned int* d_counters = nullptr;
    mpcb_mppi_info* d_info = nullptr;
    double* d_costs = nullptr;
    void* d_eps = nullptr;
    size_t d_eps_bytes = 0;
    void* d_dump = nullptr;
    unsigned long long* d_ts = nullptr;  // diagnostics: per-block timeline stamps (MPCB_DEBUG_TS=1)
    // pinned / mapped host
    double* h_in = nullptr;              // [C][4] then [C][H]
    double* h_out = nullptr;             // mapped: [C][H]
    double* h_out_dev = nullptr;         // device alias of h_out
    mpcb_mppi_info* h_info = nullptr;    // mapped: [C]
    mpcb_mppi_info* h_info_dev = nullptr;
    unsigned int* h_done = nullptr;      // mapped completion word the host spins on (C == 1)
    unsigned int* h_done_dev = nullptr;
    // host-facing calls whose final merge is the single-level warp merge: results as self-validating cells
    // (MppiParams::host_cells) instead of fence + completion words (C == 1) or a stream sync (C > 1)
    unsigned long long* h_cells = nullptr;      // mapped: [C][2 * (H + 5)]
    unsigned long long* h_cells_dev = nullptr;
    bool use_cells = false;
    unsigned int epoch = 0;
    bool costs_valid = false;
    uint32_t call_idx = 0;
    int64_t launches = 0;
    // NCCL
    void* comm = nullptr;
    // fused peer exchange: own mailbox + flags, the peers' mappings, device tables of both
    void* d_mailbox = nullptr;  // [flags: 2*G*C*kMaxMergers u32, padded][mailbox: 2*G*C*PL doubles]
    size_t mailbox_bytes = 0, mailbox_flag_bytes = 0, mailbox_ll_offset = 0;
    int peer_ll = 0;  // 1: tagged-cell (LL) exchange inside the warp merge; 0: rows + flags (mergers must match across ranks)
    void* peer_base[kMergeFan] = {};
    bool peer_ipc[kMergeFan] = {};
    double** d_peer_mbox = nullptr;
    unsigned int** d_peer_flags = nullptr;
    bool peers_attached = false;
    unsigned int xepoch = 0;
    // host-side trace of the single-controller compute (MPCB_TRACE_HOST=1): accumulated microseconds
    bool trace = false;
    double tr_launch = 0, tr_wait = 0, tr_total = 0;
    long tr_n = 0;
};

static inline double host_now_us() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

// 128-byte blob exchanged between ranks by mpcb_mppi_peer_handle / mpcb_mppi_attach_peers
struct PeerBlob {
    uint32_t magic;
    int32_t pid;
    int32_t device;
    int32_t rank;
    uint64_t ptr;
    uint64_t bytes;
    uint64_t flag_bytes;
    int32_t world, controllers, horizon, pad;
    cudaIpcMemHandle_t ipc;  // 64 bytes
    int32_t protocol;        // 1: tagged cells (LL), 0: rows + flags — every rank must use the same one
    int32_t mergers;         // flag protocol: flags are indexed by merger, so the merger count must match too
    char reserved[MPCB_PEER_HANDLE_BYTES - 56 - 64 - 8];
};
static_assert(sizeof(PeerBlob) == MPCB_PEER_HANDLE_BYTES, "peer blob must be 128 bytes");
constexpr uint32_t kPeerMagic = 0x4d504258u;  // "MPBX"

namespace {

// spin-wait hint of the host CPU (the B200 boxes are x86-64; GB200 hosts are aarch64)
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#elif defined(__aarch64__)
    asm volatile("yield" ::: "memory");
#else
    asm volatile("" ::: "memory");
#endif
}

size_t elt_size(const mpcb_mppi* h) { return h->cfg.precision != MPCB_F32 ? sizeof(double) : sizeof(float); }

// Chooses block size, blocks per controller and merge-tree shape (see the header of mppi_kernel.cuh).
//   * If the controller's W warps fit on the GPU in ONE batch per block with about one block per SM, take
//     chunks = min(num_sms / C, W) ranges and the smallest BLOCK in {128, 256, 512} that covers a range: every SM gets
//     the same number of samples and only `chunks` rows remain to merge (K = 65536, H = 100: 148 blocks of 512).
//   * Otherwise (large K, many controllers, or a horizon whose v tile limits the block): BLOCK = 128 (or the
//     largest smaller block that fits), as many resident blocks as the GPU holds, each walking several batches.
mpcb_status pick_kernels(mpcb_mppi* h) {
    const bool f64 = h->cfg.precision != MPCB_F32;
    const bool f64_fast = h->cfg.precision == MPCB_F64_FAST;
    cudaDeviceProp prop;
    MPCB_CUDA_TRY(cudaGetDeviceProperties(&prop, h->cfg.device));
    h->num_sms = prop.multiProcessorCount;
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    h->Hp = mppi_pow2_horizon(h->H, &h->lgHp);
    h->W = (h->K_local + 31) / 32;

    struct Plan {
        int spt = 1, sb = 0, block = 0, occ = 0, vt = 1;
        long long wpc = 0;  // sample-warps per block range when there is about one block per SM
        size_t smem = 0;
        bool single_batch = false;
        long long chunks = 0;
        double warps_per_sm = 0.0;  // resident thread-warps per SM while the rollouts run
        MppiKernelFn k[3] = {nullptr, nullptr, nullptr};
    };
    // Everything is in SAMPLES per block (sb); the block has sb / spt threads.
    //   single batch  : one block per SM sized to its range, v tile in shared memory (VT kernels)
    //   several batches: 128 samples per block; horizons >= 32 without the v tile — the weighted sums regenerate the
    //                    controls of the samples with non-zero weight (NOVT kernels)
    const char* vt_env = getenv("MPCB_MPPI_VT");  // developer override: 1 = keep the tile in multi-batch plans too
    auto make_plan = [&](int spt, Plan* out) -> mpcb_status {
        Plan pl;
        pl.spt = spt;
        auto smem_of = [&](int sb, int vt) { return f64 ? mppi_smem_bytes<double>(h->H, sb, vt != 0) : mppi_smem_bytes<float>(h->H, sb, vt != 0); };
        auto kernel_of = [&](int sb, int noise, int vt) -> MppiKernelFn {
            if (sb % spt) return nullptr;
            if (h->user) {  // compiled once the shape is chosen; here only: does this flavour exist for user models
                const bool have = spt == 1 && (vt ? (sb == 128 || sb == 256) : (sb == 128 || sb == 64));
                return have ? reinterpret_cast<MppiKernelFn>(1) : nullptr;
            }
            if (f64) return f64_fast ? mppi_kernel_f64fast(h->cfg.model_id, sb, noise, vt) : mppi_kernel_f64(h->cfg.model_id, sb, noise, vt);
            return spt == 2 ? mppi_kernel_f32x2(h->cfg.model_id, sb / 2, noise, vt) : mppi_kernel_f32(h->cfg.model_id, sb, noise, vt);
        };
        auto fits = [&](int sb, int vt) { return kernel_of(sb, NOISE_GENERATE, vt) != nullptr && smem_of(sb, vt) + 1024 <= smem_max; };
        long long chunks1 = h->num_sms / h->C;
        if (chunks1 < 1) chunks1 = 1;
        if (chunks1 > h->W) chunks1 = h->W;
        const long long wpc = (h->W + chunks1 - 1) / chunks1;  // warps per range if there is about one block per SM
        pl.wpc = wpc;
        const char* force = getenv("MPCB_MPPI_BLOCK");  // developer override (samples per block, tile kept)
        if (force && fits(atoi(force), 1)) {
            pl.sb = atoi(force);
        } else {
            for (int b : {128, 256, 512}) {
                if (wpc * 32 <= b && fits(b, 1)) {
                    pl.sb = b;
                    pl.single_batch = true;
                    break;
                }
            }
            if (pl.sb == 0) {
                // several batches per block.  Measured (K = 2^20, model NL): without the tile the scalar kernels gain
                // 4-6 % at H = 100 / 200 (no store per step, weighted sums only over the few non-zero weights, 16 warps
                // per SM at any horizon); at H = 8 regenerating costs more than the tile (62 vs 47 us), and the packed
                // kernels never gain from it.  So: long horizons drop the tile, short ones keep it.
                bool want_vt = h->H < 32 || spt == 2;
                if (vt_env) want_vt = atoi(vt_env) == 1;
                if (want_vt && fits(128, 1)) {
                    pl.sb = 128;
                } else {
                    pl.vt = 0;
                    for (int b : {128, 64}) {
                        if (fits(b, 0)) {
                            pl.sb = b;
                            break;
                        }
                    }
                    if (pl.sb == 0 && fits(128, 1)) {  // no tile-less kernel of this flavour: keep the tile
                        pl.vt = 1;
                        pl.sb = 128;
                    }
                }
            }
        }
        if (pl.sb == 0) {
            set_error("no MPPI kernel configuration fits horizon %d", h->H);
            return MPCB_BAD_ARG;
        }
        pl.block = pl.sb / spt;
        pl.smem = smem_of(pl.sb, pl.vt);
        if (h->user) {
            rtc_unload(&h->rtc);
            mpcb_status rst = rtc_compile_mppi_user(h->user_src.c_str(), h->S, f64, pl.block, pl.vt != 0, true, &h->rtc);
            if (rst != MPCB_OK) return rst;
        }
        for (int noise = 0; noise < 3; ++noise) {
            pl.k[noise] = h->user ? reinterpret_cast<MppiKernelFn>(h->rtc.kernel[noise]) : kernel_of(pl.sb, noise, pl.vt);
            if (!pl.k[noise]) {
                set_error("no MPPI kernel for model %d", h->cfg.model_id);
                return MPCB_BAD_ARG;
            }
            MPCB_CUDA_TRY(cudaFuncSetAttribute((const void*)pl.k[noise], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        }
        MPCB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pl.occ, (const void*)pl.k[NOISE_GENERATE], pl.block, pl.smem));
        if (pl.occ < 1) pl.occ = 1;
        if (pl.single_batch) {
            pl.chunks = (h->W + wpc - 1) / wpc;  // ranges of wpc or wpc-1 warps, none empty
            pl.warps_per_sm = (double)((wpc + spt - 1) / spt) * h->C * pl.chunks / h->num_sms;
        } else {
            pl.chunks = (long long)pl.occ * h->num_sms / h->C;
            const long long nbatches = (h->W * 32 + pl.sb - 1) / pl.sb;
            if (pl.chunks > nbatches) pl.chunks = nbatches;
            if (pl.chunks < 1) pl.chunks = 1;
            double blocks_per_sm = (double)h->C * pl.chunks / h->num_sms;
            if (blocks_per_sm > pl.occ) blocks_per_sm = pl.occ;
            pl.warps_per_sm = blocks_per_sm * (pl.block / 32.0);
        }
        *out = pl;
        return MPCB_OK;
    };
    // FP32: two samples per thread with packed f32x2 arithmetic take ~30 % fewer issue slots per sample but halve the
    // warps.  Measured (model NL, tools/dev_spt_crossover.py): a single-batch plan is quantised by warps per scheduler —
    // scalar 19 / 25 / 31 / 37 us for 1..4 warps (H = 100), packed 27 / 35 us for 1..2 — so the packed kernels win
    // exactly when the scalar plan would put 4 warps on a scheduler (13-16 sample-warps per SM: K = 65536 on 148 SMs)
    // and lose below; multi-batch plans with a short horizon take them when >= 8 warps stay resident per SM (+35 % on
    // the config-#4 shape); long multi-batch horizons run the scalar tile-less kernels.  MPCB_MPPI_SPT=1/2 forces one.
    Plan plan;
    mpcb_status st = make_plan(1, &plan);
    if (st != MPCB_OK) return st;
    if (!f64 && !h->user) {
        const char* spt_env = getenv("MPCB_MPPI_SPT");
        const int forced = spt_env ? atoi(spt_env) : 0;
        Plan p2;
        if (forced != 1 && make_plan(2, &p2) == MPCB_OK) {
            const bool single_both = plan.single_batch && p2.single_batch;
            const bool take = single_both ? plan.wpc > 12 : (!p2.single_batch && p2.vt && p2.warps_per_sm >= 8.0 && plan.vt);
            if (forced == 2 || take) plan = p2;
        }
    }
    // FP32 built-in models: the warp-specialised kernels (mppi_ws_kernel.cuh) — noise and control term in producer
    // warps, the serial rollout chain in consumer warps — when one block per SM covers the controller in a single
    // batch (BASELINE configs[1]: 13-14 sample-warps per SM).  MPCB_MPPI_WS=<variant> forces a variant of
    // kWsVariants for any shape (several batches per block included), MPCB_MPPI_WS=-1 disables them.
    h->ws_variant = -1;
    if (const char* cq_env = getenv("MPCB_MPPI_WS_CQ")) {
        const int cq = atoi(cq_env);
        if (cq >= 1 && cq <= 128) h->ws_cq = cq;
    }
    if (const char* dbg_env = getenv("MPCB_MPPI_WS_DEBUG")) h->ws_debug = atoi(dbg_env);
    if (!f64 && !h->user) {
        const char* ws_env = getenv("MPCB_MPPI_WS");
        int want = -2;  // auto
        if (ws_env) want = atoi(ws_env);
        long long chunks1 = h->num_sms / h->C;
        if (chunks1 < 1) chunks1 = 1;
        if (chunks1 > h->W) chunks1 = h->W;
        const long long wpc = (h->W + chunks1 - 1) / chunks1;
        auto ws_fits = [&](int v) {
            if (v < 0 || v >= kNumWsVariants) return false;
            const MppiWsVariant& wv = kWsVariants[v];
            return mppi_kernel_ws(h->cfg.model_id, v, NOISE_GENERATE) != nullptr &&
                   mppi_ws_smem_bytes(h->H, wv.ncw, wv.npw, wv.spt) + 2048 <= smem_max;
        };
        int pick = -1;
        if (want >= 0) {
            if (ws_fits(want)) pick = want;
        } else if (want == -2) {
            // Measured on B200 (tools/dev_events.py, model NL, H = 100): with 13-14 sample-warps per SM the packed
            // consumers + producer queue (variant 1: 7 + 9 warps) beat the one-thread-per-sample kernels by ~1.2 us per
            // step; with fewer sample-warps per SM the consumer warps are half empty and the fused kernels win, with
            // 15-16 both tie.  Several batches per block (large K) stay with the fused kernels too.
            if ((wpc == 13 || wpc == 14) && h->C == 1 && ws_fits(1)) pick = 1;
        }
        if (pick >= 0) {
            const MppiWsVariant& wv = kWsVariants[pick];
            const int cap = wv.ncw * wv.spt;  // sample-warps per batch
            plan.spt = wv.spt;
            plan.block = (wv.ncw + wv.npw) * 32;
            plan.sb = cap * 32;
            plan.vt = 1;
            plan.smem = mppi_ws_smem_bytes(h->H, wv.ncw, wv.npw, wv.spt);
            for (int noise = 0; noise < 3; ++noise) {
                plan.k[noise] = mppi_kernel_ws(h->cfg.model_id, pick, noise);
                MPCB_CUDA_TRY(cudaFuncSetAttribute((const void*)plan.k[noise], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
            }
            plan.single_batch = wpc <= cap;
            if (plan.single_batch) {
                plan.chunks = (h->W + wpc - 1) / wpc;
            } else {
                plan.chunks = h->num_sms / h->C;  // one resident block per SM
                const long long nbatches = (h->W + cap - 1) / cap;
                if (plan.chunks > nbatches) plan.chunks = nbatches;
                if (plan.chunks < 1) plan.chunks = 1;
            }
            h->ws_variant = pick;
        }
    }
    // FP64, H <= 8, several batches per block (BASELINE config #4: 4096 controllers x 8192 samples x 8 steps): the
    // register-accumulating kernel (mppi_short_kernel.cuh) — no tile, no per-batch softmax passes.  A block that walks
    // fewer than four batches keeps the fused kernel (its one merge of 9 sums per thread would not be amortised).
    // MPCB_MPPI_SHORT=0 disables, =1 forces it for every FP64 plan with H <= 8.
    if (f64 && !h->user && h->H <= kShortHorizon) {
        const char* sh_env = getenv("MPCB_MPPI_SHORT");
        const int want = sh_env ? atoi(sh_env) : -1;  // -1: auto
        auto short_of = [&](int noise) { return f64_fast ? mppi_kernel_f64fast_short(h->cfg.model_id, noise) : mppi_kernel_f64_short(h->cfg.model_id, noise); };
        if (want != 0 && short_of(NOISE_GENERATE) != nullptr) {
            Plan ps;
            ps.spt = 1;
            ps.sb = ps.block = kShortBlock;
            ps.vt = 0;
            ps.smem = mppi_smem_bytes<double>(h->H, kShortBlock, false);
            for (int noise = 0; noise < 3; ++noise) {
                ps.k[noise] = short_of(noise);
                MPCB_CUDA_TRY(cudaFuncSetAttribute((const void*)ps.k[noise], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.smem));
            }
            MPCB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps.occ, (const void*)ps.k[NOISE_GENERATE], ps.block, ps.smem));
            if (ps.occ < 1) ps.occ = 1;
            const long long nbatches = (h->W * 32 + ps.sb - 1) / ps.sb;
            ps.chunks = (long long)ps.occ * h->num_sms / h->C;
            if (ps.chunks > nbatches) ps.chunks = nbatches;
            if (ps.chunks < 1) ps.chunks = 1;
            if (want == 1 || (!plan.single_batch && nbatches / ps.chunks >= 4)) plan = ps;
        }
    }
    h->spt = plan.spt;
    const int block = plan.block;  // threads
    h->block = block;
    h->block_samples = plan.sb;
    h->smem = plan.smem;
    for (int noise = 0; noise < 3; ++noise) h->k_noise[noise] = plan.k[noise];
    long long chunks = plan.chunks;
    if (chunks > (long long)kMergeFan * kMergeFan) chunks = (long long)kMergeFan * kMergeFan;
    h->chunks = (int)chunks;
    // designated mergers (first blocks of a controller wait for the arrival counters) only when the whole launch
    // is resident at once; the final merge is then split column-wise over up to kMaxMergers of them
    {
        int occ = 0;
        MPCB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)h->k_noise[NOISE_GENERATE], block, h->smem));
        const bool resident = ((long long)h->C * h->chunks <= (long long)occ * h->num_sms) && !getenv("MPCB_MPPI_NO_SPIN");
        const int ncol2 = (h->H + 3) / 2;
        int nm = (ncol2 - 1) / 3;  // at least three column pairs per merger
        if (nm > kMaxMergers) nm = kMaxMergers;
        if (nm > h->chunks) nm = h->chunks;
        if (nm < 1) nm = 1;
        h->mergers = resident ? nm : 0;
    }
    // merge tree: one level while one merger block can take its column slice of the rows in <= 2 load batches per
    // thread (its threads split the rows nq ways), else two levels with fan-in ~ sqrt(rows)
    {
        const int nm = h->mergers > 0 ? h->mergers : 1;
        const int ncol2 = (h->H + 3) / 2;
        const int npl = 1 + (ncol2 - 1 + nm - 1) / nm;
        int Hpm = 2;
        while (Hpm < npl) Hpm <<= 1;
        const int tail_threads = h->ws_variant >= 0 ? mppi_ws_tail_threads(block) : block;  // threads that run the merges
        int nq = tail_threads / Hpm;
        if (nq < 1) nq = 1;
        if (nq > kMergeMaxPart) nq = kMergeMaxPart;
        // (with designated mergers the single level is the barrier-free warp merge, which takes any row count <= kMergeFan)
        if (h->chunks <= kMergeFan && (h->mergers >= 1 || h->chunks <= 2 * kMergeBatch * nq)) {
            h->group_size = h->chunks;
            h->groups = 1;
        } else {
            int gs = (int)ceil(sqrt((double)h->chunks));
            if (gs > kMergeFan) gs = kMergeFan;
            h->group_size = gs;
            h->groups = (h->chunks + gs - 1) / gs;
        }
    }
    return MPCB_OK;
}

void fill_params(const mpcb_mppi* h, MppiParams* p) {
    memset(p, 0, sizeof(*p));
    p->H = h->H;
    p->Hp = h->Hp;
    p->lgHp = h->lgHp;
    p->PL = h->PL;
    p->mergers = h->mergers;

    p->W = h->W;
    p->C = h->C;
    p->chunks = h->chunks;
    p->group_size = h->group_size;
    p->groups = h->groups;
    p->K_local = h->K_local;
    p->K_global = h->cfg.samples;
    p->k_offset = h->k_offset;
    p->seed_lo = (unsigned int)(h->cfg.seed & 0xffffffffull);
    p->seed_hi = (unsigned int)(h->cfg.seed >> 32);
    p->call_idx = h->call_idx;
    p->c_offset = h->c_offset;
    p->lambda = h->cfg.lambda;
    p->inv_var = 1.0 / (h->cfg.std_dev * h->cfg.std_dev);  // std_dev.powi(-2), src/mppi.rs:48
    p->lo = h->cfg.limit_lo;
    p->hi = h->cfg.limit_hi;
    p->std_dev = h->cfg.std_dev;
    p->partial = h->d_partial;
    p->counters = h->d_counters;
    p->u_out = h->d_u_out;
    p->info = h->d_info;
    p->rank_partial = h->d_rank_partial;
    p->mc = h->mc;
    for (int i = 0; i < kModelConsts; ++i) p->mc.kf[i] = (float)h->mc.k[i];
    p->costs = h->cfg.keep_costs ? h->d_costs : nullptr;
    p->debug_ts = h->d_ts;
    p->ws_cq = h->ws_cq;
    p->ws_debug = h->ws_debug;
}

// Enqueue one fused control step.
mpcb_status launch(mpcb_mppi* h, MppiParams& p) {
    const int noise = p.eps != nullptr ? NOISE_REPLAY : (p.eps_dump != nullptr ? NOISE_GENERATE_DUMP : NOISE_GENERATE);
    MppiKernelFn fn = h->k_noise[noise];
    const dim3 grid((unsigned)(h->C * h->chunks)), block((unsigned)h->block);
    h->seq += 1;
    p.seq = h->seq;
    // explicit cudaLaunchKernel: `fn` is either a compiled-in __global__ function or the cudaKernel_t of a user model
    void* args[1] = {&p};
    MPCB_CUDA_TRY(launch_pdl(reinterpret_cast<const void*>(fn), grid, block, args, h->smem, h->stream));
    h->launches += 1;
    h->call_idx += 1;
    h->costs_valid = h->cfg.keep_costs != 0;
    return MPCB_OK;
}

mpcb_status ensure_eps(mpcb_mppi* h, size_t bytes) {
    if (h->d_eps_bytes >= bytes) return MPCB_OK;
    if (h->d_eps) MPCB_CUDA_TRY(cudaFree(h->d_eps));
    h->d_eps = nullptr;
    h->d_eps_bytes = 0;
    MPCB_CUDA_TRY(cudaMalloc(&h->d_eps, bytes));
    h->d_eps_bytes = bytes;
    return MPCB_OK;
}

// Stages host x/u_in: inline in the kernel parameters for the single-controller case, else one pinned H2D copy.
mpcb_status stage_inputs(mpcb_mppi* h, MppiParams& p, const double* x, const double* u_in) {
    const size_t nx = (size_t)h->C * h->S, nu = (size_t)h->C * h->H;
    if (h->C == 1 && h->H <= kInlineHorizon) {
        p.use_inline = 1;
        memcpy(p.xu_inline, x, (size_t)h->S * sizeof(double));
        memcpy(p.xu_inline + h->S, u_in, (size_t)h->H * sizeof(double));
        return MPCB_OK;
    }
    memcpy(h->h_in, x, nx * sizeof(double));
    memcpy(h->h_in + nx, u_in, nu * sizeof(double));
    MPCB_CUDA_TRY(cudaMemcpyAsync(h->d_x, h->h_in, nx * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    MPCB_CUDA_TRY(cudaMemcpyAsync(h->d_u, h->h_in + nx, nu * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    p.x = h->d_x;
    p.u = h->d_u;
    return MPCB_OK;
}

// Waits for the step and hands u_out/info to the caller (results were written straight into mapped host memory).
// Single controller: the final block stores an epoch word after the results; spinning on it avoids the wake-up
// latency of cudaStreamSynchronize.  After ~2 ms without completion (or for C > 1) fall back to the stream sync,
// which also surfaces any launch/runtime error.
mpcb_status finish_host(mpcb_mppi* h, double* u_out, mpcb_mppi_info* info, bool spin, int words, bool cells = false) {
    bool done = false;
    if (cells) {
        // every result word arrives as two cells (half | epoch << 32): wait for each cell to show this call's epoch;
        // controller c owns the cells [c * 2 (H + 5), (c + 1) * 2 (H + 5))
        const unsigned long long want = h->epoch;
        const int nwords = h->H + 5;
        unsigned long long out[kMaxHorizon + 5];
        long budget = 4000000;
        bool synced = false;
        for (long long c = 0; c < h->C; ++c) {
            const volatile unsigned long long* cell = h->h_cells + (size_t)c * 2 * nwords;
            int i = 0;
            while (i < nwords) {
                const unsigned long long lo = cell[2 * i], hi = cell[2 * i + 1];
                if ((lo >> 32) == want && (hi >> 32) == want) {
                    out[i] = (lo & 0xffffffffull) | (hi << 32);
                    ++i;
                    continue;
                }
                if (--budget > 0) {
                    cpu_relax();
                    continue;
                }
                if (synced) {
                    set_error("MPPI results did not reach the host cells (epoch %u, controller %lld)", h->epoch, c);
                    return MPCB_CUDA_ERROR;
                }
                // ~2 ms without the results: the stream sync surfaces a launch / runtime error, then one more look
                MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
                synced = true;
                budget = 1000;
            }
            memcpy(u_out + (size_t)c * h->H, out, (size_t)h->H * sizeof(double));
            memcpy(h->h_out + (size_t)c * h->H, out, (size_t)h->H * sizeof(double));
            memcpy(&h->h_info[c], out + h->H, sizeof(mpcb_mppi_info));  // mpcb_mppi_last_info reads the mirror
        }
        if (info) memcpy(info, h->h_info, (size_t)h->C * sizeof(mpcb_mppi_info));
        if (h->C == 1) return (mpcb_status)h->h_info[0].status;
        return MPCB_OK;
    }
    if (spin) {
        // one completion word per merger block of the final merge
        volatile unsigned int* flag = h->h_done;
        const unsigned int want = h->epoch;
        for (long i = 0; i < 4000000; ++i) {
            bool all = true;
            for (int w = 0; w < words; ++w) all = all && (flag[w] == want);
            if (all) { done = true; break; }
            cpu_relax();
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
    }
    if (!done) MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    memcpy(u_out, h->h_out, (size_t)h->C * h->H * sizeof(double));
    if (info) memcpy(info, h->h_info, (size_t)h->C * sizeof(mpcb_mppi_info));
    if (h->C == 1) return (mpcb_status)h->h_info[0].status;
    return MPCB_OK;
}

mpcb_status exchange_and_combine(mpcb_mppi* h);

// One more step of the fused peer exchange: every rank advances the same epoch counter.
void set_peer_params(mpcb_mppi* h, MppiParams& p) {
    h->xepoch += 1;
    p.final_mode = FINAL_PEER_EXCHANGE;
    p.peer_mbox = h->d_peer_mbox;
    p.peer_flags = h->d_peer_flags;
    p.G = h->cfg.world_size;
    p.rank = h->cfg.rank;
    p.xepoch = h->xepoch;
    p.peer_ll_offset = h->peer_ll ? (long long)h->mailbox_ll_offset : 0;
}

mpcb_status compute_host(mpcb_mppi* h, const double* x, const double* u_in, const void* d_eps, int eps_dtype,
                         void* d_dump, double* u_out, mpcb_mppi_info* info) {
    MPCB_REQUIRE(h && x && u_in && u_out, "null pointer");
    const double tr0 = h->trace ? host_now_us() : 0.0;
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    MppiParams p;
    fill_params(h, &p);
    mpcb_status st = stage_inputs(h, p, x, u_in);
    if (st != MPCB_OK) return st;
    p.eps = d_eps;
    p.eps_f64 = (eps_dtype == MPCB_DT_F64);
    p.eps_dump = d_dump;
    const bool sharded = h->cfg.world_size > 1;
    const bool peer = sharded && h->peers_attached;
    const bool spin = (h->C == 1);
    h->epoch += 1;
    if (sharded && !peer) {
        MPCB_REQUIRE(h->comm != nullptr,
                     "world_size > 1 needs mpcb_mppi_attach_peers or mpcb_mppi_attach_comm (or use compute_partial + combine)");
        p.final_mode = FINAL_RANK_ROW;
    } else {
        p.u_out_host = h->h_out_dev;
        p.info_host = h->h_info_dev;
        if (spin) {
            p.done_host = h->h_done_dev;
            p.epoch = h->epoch;
        }
        if (peer) set_peer_params(h, p);
    }
    // the single-level warp merge (the condition of `pair_major` in mppi_block_tail) hands the results over as cells
    const bool cells = h->use_cells && h->groups == 1 && h->mergers >= 1 && (!sharded || (peer && h->peer_ll));
    if (cells) {
        p.host_cells = h->h_cells_dev;
        p.epoch = h->epoch;
    }
    st = launch(h, p);
    if (st != MPCB_OK) return st;
    if (sharded && !peer) {
        st = exchange_and_combine(h);
        if (st != MPCB_OK) return st;
    }
    const int words = (sharded && !peer) ? 1 : (h->mergers > 0 ? h->mergers : 1);
    const double tr1 = h->trace ? host_now_us() : 0.0;
    st = finish_host(h, u_out, info, spin, words, cells);
    if (h->trace) {
        const double tr2 = host_now_us();
        h->tr_launch += tr1 - tr0;
        h->tr_wait += tr2 - tr1;
        h->tr_total += tr2 - tr0;
        h->tr_n += 1;
        if (h->tr_n % 1000 == 0) {
            fprintf(stderr, "[mpcb trace] compute: enqueue %.2f us, wait+copy-out %.2f us, total %.2f us (mean of 1000)\n",
                    h->tr_launch / 1000, h->tr_wait / 1000, h->tr_total / 1000);
            h->tr_launch = h->tr_wait = h->tr_total = 0;
        }
    }
    return st;
}

mpcb_status run_combine(mpcb_mppi* h, const double* rows, int G) {
    MppiCombineParams cp;
    cp.rows = rows;
    cp.G = G;
    cp.C = h->C;
    cp.H = h->H;
    cp.lambda = h->cfg.lambda;
    cp.u_out = h->d_u_out;
    cp.u_out_host = h->h_out_dev;
    cp.info = h->d_info;
    cp.info_host = h->h_info_dev;
    cp.done_host = (h->C == 1) ? h->h_done_dev : nullptr;
    cp.epoch = h->epoch;
    mppi_combine_kernel<128><<<h->C, 128, 0, h->stream>>>(cp);
    MPCB_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return MPCB_OK;
}

mpcb_status exchange_and_combine(mpcb_mppi* h) {
    const size_t count = (size_t)h->C * h->PL;
    mpcb_status st = nccl_all_gather(h->comm, h->d_rank_partial, h->d_gather, count, h->stream);
    if (st != MPCB_OK) return st;
    return run_combine(h, h->d_gather, h->cfg.world_size);
}

}  // namespace

extern "C" {

mpcb_status mpcb_mppi_default_cfg(int32_t model_id, mpcb_mppi_cfg* c) {
    if (!c) return MPCB_BAD_ARG;
    memset(c, 0, sizeof(*c));
    if (model_id == MPCB_MODEL_USER) {  // neutral values; the model itself comes with mpcb_mppi_create_user
        c->model_id = model_id;
        c->precision = MPCB_F32;
        c->horizon = 8;
        c->state_dim = 4;
        c->samples = 65536;
        c->controllers = 1;
        c->world_size = 1;
        c->seed = 0x6d70632d72730001ull;
        c->lambda = 1.0;
        c->std_dev = 1.0;
        c->limit_lo = -HUGE_VAL;
        c->limit_hi = HUGE_VAL;
        return MPCB_OK;
    }
    mpcb_status st = mpcb_model_defaults(model_id, &c->model);
    if (st != MPCB_OK) return st;
    c->model_id = model_id;
    // FP32 meets the 1e-5 parity bar on L and NL; NL6 at DT = 0.15 is chaotic inside its horizon and needs FP64
    // model NL6 at its shipped DT is chaotic inside the horizon: FP32 rollouts miss the 1e-5 tolerance there,
    // the FP64 fast form holds 1e-9 at half the cost of the reference-order FP64 path
    c->precision = (model_id == MPCB_MODEL_NL6) ? MPCB_F64_FAST : MPCB_F32;
    c->horizon = 8;
    c->state_dim = 4;
    c->controllers = 1;
    c->world_size = 1;
    c->seed = 0x6d70632d72730001ull;
    switch (model_id) {
        case MPCB_MODEL_L:   // examples/mppi4.rs:8-18
        case MPCB_MODEL_NL:  // examples/mppi4-non-liner.rs:8-18
            c->samples = 800000;
            c->lambda = 0.5;
            c->std_dev = 3.0;
            c->limit_lo = -20.0;
            c->limit_hi = 20.0;
            return MPCB_OK;
        case MPCB_MODEL_NL6:  // examples/mppi4-non-liner-ukf.rs:13-24
            c->samples = 500000;
            c->lambda = 1.4;
            c->std_dev = 4.0;
            c->limit_lo = -10.0;
            c->limit_hi = 10.0;
            return MPCB_OK;
        default:
            set_error("model %d is not an MPPI model", model_id);
            return MPCB_BAD_ARG;
    }
}

static mpcb_status create_impl(mpcb_mppi** out, const mpcb_mppi_cfg* cfg, const char* user_src, const double* params,
                               int32_t n_params) {
    MPCB_REQUIRE(out && cfg, "null pointer");
    *out = nullptr;
    if (user_src != nullptr) MPCB_REQUIRE(cfg->state_dim >= 1 && cfg->state_dim <= kMaxStateDim, "user models: state_dim in 1..8");
    else MPCB_REQUIRE(cfg->state_dim == 4, "the built-in models have S = 4");
    MPCB_REQUIRE(cfg->horizon >= 1 && cfg->horizon <= kMaxHorizon, "horizon out of range");
    MPCB_REQUIRE(cfg->samples >= 1, "samples must be >= 1");
    MPCB_REQUIRE(cfg->controllers >= 1, "controllers must be >= 1");
    MPCB_REQUIRE(cfg->world_size >= 1 && cfg->world_size <= kMergeFan && cfg->rank >= 0 && cfg->rank < cfg->world_size,
                 "bad rank/world_size (1..256 ranks)");
    MPCB_REQUIRE(cfg->precision == MPCB_F32 || cfg->precision == MPCB_F64 || cfg->precision == MPCB_F64_FAST, "bad precision");
    MPCB_REQUIRE(user_src == nullptr || cfg->precision != MPCB_F64_FAST, "MPCB_F64_FAST is a form of the built-in models; user models take MPCB_F32 or MPCB_F64");
    MPCB_REQUIRE(cfg->std_dev > 0.0 && cfg->lambda > 0.0, "std_dev and lambda must be positive");
    MPCB_REQUIRE(cfg->limit_lo <= cfg->limit_hi, "limit.0 > limit.1");
    int ndev = 0;
    MPCB_CUDA_TRY(cudaGetDeviceCount(&ndev));
    MPCB_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "no such CUDA device (this library has no CPU path)");
    MPCB_CUDA_TRY(cudaSetDevice(cfg->device));

    mpcb_mppi* h = new (std::nothrow) mpcb_mppi();
    MPCB_REQUIRE(h != nullptr, "out of memory");
    h->cfg = *cfg;
    h->H = cfg->horizon;
    h->S = cfg->state_dim;
    h->C = cfg->controllers;
    h->PL = mppi_partial_len(h->H);
    // contiguous shard of the global sample index
    const long long K = cfg->samples, G = cfg->world_size, r = cfg->rank;
    h->k_offset = K * r / G;
    h->K_local = K * (r + 1) / G - h->k_offset;
    mpcb_status st = MPCB_OK;
    auto fail = [&](mpcb_status s) {
        mpcb_mppi_destroy(h);
        return s;
    };
    if (h->K_local < 1) {
        set_error("rank %d of %d has no samples (K = %lld)", (int)r, (int)G, K);
        return fail(MPCB_BAD_ARG);
    }
    if (user_src != nullptr) {
        h->user = true;
        h->user_src = user_src;
        h->cfg.model_id = MPCB_MODEL_USER;
        memset(&h->mc, 0, sizeof(h->mc));
        for (int i = 0; i < n_params; ++i) {
            h->mc.k[i] = params[i];
            h->mc.kf[i] = (float)params[i];
        }
    } else {
        st = build_model_consts(cfg->model_id, cfg->model, cfg->model.dt, &h->mc);
        if (st != MPCB_OK) return fail(st);
        if (cfg->model_id != MPCB_MODEL_L && cfg->model_id != MPCB_MODEL_NL && cfg->model_id != MPCB_MODEL_NL6) {
            set_error("model %d is not a built-in MPPI model (user models: mpcb_mppi_create_user)", cfg->model_id);
            return fail(MPCB_BAD_ARG);
        }
    }
    st = pick_kernels(h);
    if (st != MPCB_OK) return fail(st);

#define TRY_OR_FAIL(expr)                                                                             \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                                \
            return fail(MPCB_CUDA_ERROR);                                                             \
        }                                                                                             \
    } while (0)
    TRY_OR_FAIL(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const size_t C = h->C, H = h->H;
    TRY_OR_FAIL(cudaMalloc(&h->d_x, C * (size_t)h->S * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_u, C * H * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_u_out, C * H * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_partial, C * (size_t)(h->chunks + h->groups) * h->PL * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_rank_partial, C * h->PL * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_gather, (size_t)cfg->world_size * C * h->PL * sizeof(double)));
    TRY_OR_FAIL(cudaMalloc(&h->d_counters, C * (size_t)(h->groups + 1) * sizeof(unsigned int)));
    TRY_OR_FAIL(cudaMemset(h->d_counters, 0, C * (size_t)(h->groups + 1) * sizeof(unsigned int)));
    TRY_OR_FAIL(cudaMalloc(&h->d_info, C * sizeof(mpcb_mppi_info)));
    TRY_OR_FAIL(cudaMemset(h->d_info, 0, C * sizeof(mpcb_mppi_info)));
    h->trace = getenv("MPCB_TRACE_HOST") != nullptr;
    if (getenv("MPCB_DEBUG_TS")) {
        TRY_OR_FAIL(cudaMalloc(&h->d_ts, C * (size_t)h->chunks * 16 * sizeof(unsigned long long)));
        TRY_OR_FAIL(cudaMemset(h->d_ts, 0, C * (size_t)h->chunks * 16 * sizeof(unsigned long long)));
    }
    if (cfg->keep_costs) TRY_OR_FAIL(cudaMalloc(&h->d_costs, C * (size_t)h->K_local * sizeof(double)));
    TRY_OR_FAIL(cudaHostAlloc(&h->h_in, C * ((size_t)h->S + H) * sizeof(double), cudaHostAllocDefault));
    TRY_OR_FAIL(cudaHostAlloc(&h->h_out, C * H * sizeof(double), cudaHostAllocMapped));
    TRY_OR_FAIL(cudaHostAlloc(&h->h_info, C * sizeof(mpcb_mppi_info), cudaHostAllocMapped));
    memset(h->h_out, 0, C * H * sizeof(double));
    memset(h->h_info, 0, C * sizeof(mpcb_mppi_info));
    TRY_OR_FAIL(cudaHostGetDevicePointer((void**)&h->h_out_dev, h->h_out, 0));
    TRY_OR_FAIL(cudaHostGetDevicePointer((void**)&h->h_info_dev, h->h_info, 0));
    TRY_OR_FAIL(cudaHostAlloc(&h->h_done, 64, cudaHostAllocMapped));
    memset(h->h_done, 0, 64);
    TRY_OR_FAIL(cudaHostGetDevicePointer((void**)&h->h_done_dev, h->h_done, 0));
    if (C * (H + 5) <= 32768) {  // (half a megabyte of cells at most: larger fleets keep the stream sync)
        const size_t cell_bytes = C * 2 * (H + 5) * sizeof(unsigned long long);
        TRY_OR_FAIL(cudaHostAlloc(&h->h_cells, cell_bytes, cudaHostAllocMapped));
        memset(h->h_cells, 0, cell_bytes);  // epoch 0 is never a call's epoch
        TRY_OR_FAIL(cudaHostGetDevicePointer((void**)&h->h_cells_dev, h->h_cells, 0));
        h->use_cells = getenv("MPCB_MPPI_HOST_CELLS") == nullptr || atoi(getenv("MPCB_MPPI_HOST_CELLS")) != 0;
    }
#undef TRY_OR_FAIL
    *out = h;
    return MPCB_OK;
}

mpcb_status mpcb_mppi_create(mpcb_mppi** out, const mpcb_mppi_cfg* cfg) { return create_impl(out, cfg, nullptr, nullptr, 0); }

mpcb_status mpcb_mppi_create_user(mpcb_mppi** out, const mpcb_mppi_cfg* cfg, const char* cuda_source, const double* params,
                                  int32_t n_params) {
    MPCB_REQUIRE(cuda_source != nullptr, "null source");
    MPCB_REQUIRE(n_params >= 0 && n_params <= MPCB_USER_PARAMS && (n_params == 0 || params != nullptr), "bad params");
    return create_impl(out, cfg, cuda_source, params, n_params);
}

mpcb_status mpcb_mppi_check_user_source(const char* cuda_source, int32_t state_dim, int32_t precision) {
    MPCB_REQUIRE(cuda_source != nullptr, "null source");
    MPCB_REQUIRE(state_dim >= 1 && state_dim <= kMaxStateDim, "state_dim in 1..8");
    MPCB_REQUIRE(precision == MPCB_F32 || precision == MPCB_F64, "bad precision");
    RtcModule m;
    return rtc_compile_mppi_user(cuda_source, state_dim, precision == MPCB_F64, 128, true, false, &m);
}

const char* mpcb_rtc_log(void) { return rtc_log(); }

void mpcb_mppi_destroy(mpcb_mppi* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    rtc_unload(&h->rtc);
    if (h->comm) nccl_destroy(h->comm);
    for (int r = 0; r < kMergeFan; ++r)
        if (h->peer_ipc[r] && h->peer_base[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
    cudaFree(h->d_peer_mbox);
    cudaFree(h->d_peer_flags);
    cudaFree(h->d_mailbox);
    cudaFree(h->d_x);
    cudaFree(h->d_u);
    cudaFree(h->d_u_out);
    cudaFree(h->d_partial);
    cudaFree(h->d_rank_partial);
    cudaFree(h->d_gather);
    cudaFree(h->d_counters);
    cudaFree(h->d_info);
    cudaFree(h->d_costs);
    cudaFree(h->d_eps);
    cudaFree(h->d_dump);
    cudaFree(h->d_ts);
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_info) cudaFreeHost(h->h_info);
    if (h->h_done) cudaFreeHost(h->h_done);
    if (h->h_cells) cudaFreeHost(h->h_cells);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
}

mpcb_status mpcb_mppi_compute(mpcb_mppi* h, const double* x, const double* u_in, double* u_out, mpcb_mppi_info* info) {
    return compute_host(h, x, u_in, nullptr, MPCB_DT_F32, nullptr, u_out, info);
}

mpcb_status mpcb_mppi_compute_replay(mpcb_mppi* h, const double* x, const double* u_in, const void* eps,
                                     int32_t eps_dtype, int32_t eps_on_device, double* u_out, mpcb_mppi_info* info) {
    MPCB_REQUIRE(h && eps, "null pointer");
    MPCB_REQUIRE(eps_dtype == MPCB_DT_F32 || eps_dtype == MPCB_DT_F64, "bad eps dtype");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    const void* d_eps = eps;
    if (!eps_on_device) {
        const size_t es = eps_dtype == MPCB_DT_F64 ? sizeof(double) : sizeof(float);
        const size_t bytes = (size_t)h->C * (size_t)h->cfg.samples * h->H * es;
        mpcb_status st = ensure_eps(h, bytes);
        if (st != MPCB_OK) return st;
        MPCB_CUDA_TRY(cudaMemcpyAsync(h->d_eps, eps, bytes, cudaMemcpyHostToDevice, h->stream));
        d_eps = h->d_eps;
    }
    return compute_host(h, x, u_in, d_eps, eps_dtype, nullptr, u_out, info);
}

mpcb_status mpcb_mppi_compute_dump(mpcb_mppi* h, const double* x, const double* u_in, void* eps_out, double* u_out,
                                   mpcb_mppi_info* info) {
    MPCB_REQUIRE(h && eps_out, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    const size_t bytes = (size_t)h->C * (size_t)h->K_local * h->H * elt_size(h);
    if (!h->d_dump) MPCB_CUDA_TRY(cudaMalloc(&h->d_dump, bytes));
    mpcb_status st = compute_host(h, x, u_in, nullptr, MPCB_DT_F32, h->d_dump, u_out, info);
    // the noise is returned even when the controller reports a numeric failure
    if (st == MPCB_BAD_ARG || st == MPCB_CUDA_ERROR || st == MPCB_NCCL_ERROR) return st;
    MPCB_CUDA_TRY(cudaMemcpy(eps_out, h->d_dump, bytes, cudaMemcpyDeviceToHost));
    return st;
}

mpcb_status mpcb_mppi_get_costs(mpcb_mppi* h, double* c_out) {
    MPCB_REQUIRE(h && c_out, "null pointer");
    MPCB_REQUIRE(h->cfg.keep_costs && h->costs_valid, "cfg.keep_costs was not set or nothing computed yet");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    MPCB_CUDA_TRY(cudaMemcpy(c_out, h->d_costs, (size_t)h->C * h->K_local * sizeof(double), cudaMemcpyDeviceToHost));
    return MPCB_OK;
}

mpcb_status mpcb_mppi_compute_device(mpcb_mppi* h, const double* d_x, const double* d_u_in, const void* d_eps,
                                     int32_t eps_dtype, double* d_u_out) {
    MPCB_REQUIRE(h && d_x && d_u_in && d_u_out, "null pointer");
    MPCB_REQUIRE(h->cfg.world_size == 1 || h->comm != nullptr || h->peers_attached,
                 "sharded handle needs mpcb_mppi_attach_peers or mpcb_mppi_attach_comm");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    MppiParams p;
    fill_params(h, &p);
    p.x = d_x;
    p.u = d_u_in;
    p.eps = d_eps;
    p.eps_f64 = (eps_dtype == MPCB_DT_F64);
    p.u_out = d_u_out;
    const bool sharded = h->cfg.world_size > 1;
    const bool peer = sharded && h->peers_attached;
    p.final_mode = sharded ? FINAL_RANK_ROW : FINAL_NORMALISE;
    if (peer) set_peer_params(h, p);
    mpcb_status st = launch(h, p);
    if (st != MPCB_OK) return st;
    if (sharded && !peer) {
        st = nccl_all_gather(h->comm, h->d_rank_partial, h->d_gather, (size_t)h->C * h->PL, h->stream);
        if (st != MPCB_OK) return st;
        MppiCombineParams cp;
        cp.rows = h->d_gather;
        cp.G = h->cfg.world_size;
        cp.C = h->C;
        cp.H = h->H;
        cp.lambda = h->cfg.lambda;
        cp.u_out = d_u_out;
        cp.u_out_host = nullptr;
        cp.info = h->d_info;
        cp.info_host = nullptr;
        cp.done_host = nullptr;
        cp.epoch = 0;
        mppi_combine_kernel<128><<<h->C, 128, 0, h->stream>>>(cp);
        MPCB_CUDA_TRY(cudaGetLastError());
        h->launches += 1;
    }
    return MPCB_OK;
}

// u0[c] = u_out[c][0]: the control each controller applies next (u_n[0], examples/mppi4-non-liner-ukf.rs:231,270)
__global__ void mppi_first_control_kernel(const double* u_out, double* u0, int C, int H) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) u0[c] = u_out[(long long)c * H];
}

mpcb_status mpcb_mppi_first_control_device(mpcb_mppi* h, const double* d_u_out, double* d_u0) {
    MPCB_REQUIRE(h && d_u_out && d_u0, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    mppi_first_control_kernel<<<(h->C + 255) / 256, 256, 0, h->stream>>>(d_u_out, d_u0, h->C, h->H);
    MPCB_CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return MPCB_OK;
}

mpcb_status mpcb_mppi_sync(mpcb_mppi* h) {
    MPCB_REQUIRE(h, "null handle");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MPCB_OK;
}

mpcb_status mpcb_mppi_last_info(mpcb_mppi* h, mpcb_mppi_info* info) {
    MPCB_REQUIRE(h && info, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    MPCB_CUDA_TRY(cudaMemcpy(info, h->d_info, (size_t)h->C * sizeof(mpcb_mppi_info), cudaMemcpyDeviceToHost));
    return MPCB_OK;
}

// Diagnostics (MPCB_DEBUG_TS=1 at create): %globaltimer stamps [blocks][16] of the last launch:
// 0 block start, 1 rollouts done, 2 ticket taken, 3 group merged, 4 second ticket, 5 final merge done.
int64_t mpcb_mppi_debug_timeline(mpcb_mppi* h, unsigned long long* out, int64_t max_blocks) {
    if (!h || !h->d_ts || !out) return 0;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->stream);
    int64_t n = (int64_t)h->C * h->chunks;
    if (n > max_blocks) n = max_blocks;
    cudaMemcpy(out, h->d_ts, (size_t)n * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return n;
}

mpcb_status mpcb_mppi_set_controller_offset(mpcb_mppi* h, int64_t first_controller) {
    MPCB_REQUIRE(h && first_controller >= 0 && first_controller <= 0x7fffffffll, "bad controller offset");
    h->c_offset = (unsigned int)first_controller;
    return MPCB_OK;
}

void* mpcb_mppi_stream(mpcb_mppi* h) { return h ? (void*)h->stream : nullptr; }
int64_t mpcb_mppi_launches(mpcb_mppi* h) { return h ? h->launches : 0; }
int64_t mpcb_mppi_local_samples(mpcb_mppi* h) { return h ? h->K_local : 0; }
int32_t mpcb_mppi_partial_len(mpcb_mppi* h) { return h ? h->PL : 0; }

mpcb_status mpcb_mppi_compute_partial(mpcb_mppi* h, const double* x, const double* u_in, const void* d_eps,
                                      int32_t eps_dtype, double* d_partial) {
    MPCB_REQUIRE(h && x && u_in && d_partial, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    MppiParams p;
    fill_params(h, &p);
    mpcb_status st = stage_inputs(h, p, x, u_in);
    if (st != MPCB_OK) return st;
    p.eps = d_eps;
    p.eps_f64 = (eps_dtype == MPCB_DT_F64);
    p.final_mode = FINAL_RANK_ROW;
    p.rank_partial = d_partial;
    st = launch(h, p);
    if (st != MPCB_OK) return st;
    MPCB_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return MPCB_OK;
}

mpcb_status mpcb_mppi_combine(mpcb_mppi* h, const double* d_partials, int32_t n_ranks, double* u_out,
                              mpcb_mppi_info* info) {
    MPCB_REQUIRE(h && d_partials && u_out, "null pointer");
    MPCB_REQUIRE(n_ranks >= 1 && n_ranks <= h->cfg.world_size, "n_ranks exceeds cfg.world_size");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    h->epoch += 1;
    mpcb_status st = run_combine(h, d_partials, n_ranks);
    if (st != MPCB_OK) return st;
    return finish_host(h, u_out, info, h->C == 1, 1);
}

mpcb_status mpcb_comm_unique_id(char id[128]) { return nccl_unique_id(id); }

mpcb_status mpcb_mppi_attach_comm(mpcb_mppi* h, const char id[128]) {
    MPCB_REQUIRE(h && id, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (h->comm) {
        nccl_destroy(h->comm);
        h->comm = nullptr;
    }
    return nccl_init_rank(&h->comm, id, h->cfg.rank, h->cfg.world_size);
}

namespace {
mpcb_status ensure_mailbox(mpcb_mppi* h) {
    if (h->d_mailbox) return MPCB_OK;
    const size_t G = h->cfg.world_size, C = h->C;
    h->mailbox_flag_bytes = ((2 * G * C * kMaxMergers * sizeof(unsigned int)) + 255) & ~(size_t)255;
    // [flags][2 parity x G x C rows of PL doubles: flag protocol][2 x G x C x ncell tagged 16-byte cells: LL protocol]
    h->mailbox_ll_offset = 2 * G * C * (size_t)h->PL * sizeof(double);  // from the end of the flags
    h->mailbox_bytes = h->mailbox_flag_bytes + h->mailbox_ll_offset + 2 * G * C * (size_t)mppi_ll_cells(h->H) * 16;
    // the tagged-cell exchange runs inside the single-level warp merge (one resident wave, <= kMergeFan rows) with lane r = rank r
    h->peer_ll = (h->groups == 1 && h->mergers >= 1 && G <= 32 && !getenv("MPCB_MPPI_PEER_FLAGS")) ? 1 : 0;
    MPCB_CUDA_TRY(cudaMalloc(&h->d_mailbox, h->mailbox_bytes));
    MPCB_CUDA_TRY(cudaMemset(h->d_mailbox, 0, h->mailbox_bytes));
    MPCB_CUDA_TRY(cudaDeviceSynchronize());  // zeroed before anyone can learn the handle
    return MPCB_OK;
}
}  // namespace

mpcb_status mpcb_mppi_peer_handle(mpcb_mppi* h, char out[MPCB_PEER_HANDLE_BYTES]) {
    MPCB_REQUIRE(h && out, "null pointer");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    mpcb_status st = ensure_mailbox(h);
    if (st != MPCB_OK) return st;
    PeerBlob b;
    memset(&b, 0, sizeof(b));
    b.magic = kPeerMagic;
    b.pid = (int32_t)getpid();
    b.device = h->cfg.device;
    b.rank = h->cfg.rank;
    b.ptr = (uint64_t)(uintptr_t)h->d_mailbox;
    b.bytes = h->mailbox_bytes;
    b.flag_bytes = h->mailbox_flag_bytes;
    b.world = h->cfg.world_size;
    b.controllers = h->C;
    b.horizon = h->H;
    b.protocol = h->peer_ll;
    b.mergers = h->mergers;
    MPCB_CUDA_TRY(cudaIpcGetMemHandle(&b.ipc, h->d_mailbox));
    memcpy(out, &b, sizeof(b));
    return MPCB_OK;
}

mpcb_status mpcb_mppi_attach_peers(mpcb_mppi* h, const char* handles) {
    MPCB_REQUIRE(h && handles, "null pointer");
    MPCB_REQUIRE(!h->peers_attached, "peers already attached");
    MPCB_CUDA_TRY(cudaSetDevice(h->cfg.device));
    mpcb_status st = ensure_mailbox(h);
    if (st != MPCB_OK) return st;
    const int G = h->cfg.world_size;
    double* mbox[kMergeFan];
    unsigned int* flags[kMergeFan];
    for (int r = 0; r < G; ++r) {
        PeerBlob b;
        memcpy(&b, handles + (size_t)r * MPCB_PEER_HANDLE_BYTES, sizeof(b));
        MPCB_REQUIRE(b.magic == kPeerMagic && b.rank == r, "handles must be the mpcb_mppi_peer_handle blobs in rank order");
        MPCB_REQUIRE(b.world == G && b.controllers == h->C && b.horizon == h->H && b.bytes == h->mailbox_bytes,
                     "peer handle was made by a controller of a different shape");
        // the exchange protocol (and, for the flag protocol, the number of merger blocks the flags are indexed by) follows from
        // each rank's own kernel plan: a mismatch would end in MPCB_PEER_TIMEOUT on every step, so it is refused here
        MPCB_REQUIRE(b.protocol == h->peer_ll && (h->peer_ll || b.mergers == h->mergers),
                     "ranks planned different exchange protocols / merger counts (different shard sizes or MPCB_MPPI_* overrides?)");
        void* base = nullptr;
        if (r == h->cfg.rank) {
            base = h->d_mailbox;
        } else if (b.pid == (int32_t)getpid()) {
            // another handle of this process: plain peer access
            if (b.device != h->cfg.device) {
                int can = 0;
                MPCB_CUDA_TRY(cudaDeviceCanAccessPeer(&can, h->cfg.device, b.device));
                MPCB_REQUIRE(can, "devices cannot access each other's memory");
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) MPCB_CUDA_TRY(e);
                cudaGetLastError();
            }
            base = (void*)(uintptr_t)b.ptr;
        } else {
            MPCB_CUDA_TRY(cudaIpcOpenMemHandle(&base, b.ipc, cudaIpcMemLazyEnablePeerAccess));
            h->peer_ipc[r] = true;
        }
        h->peer_base[r] = base;
        flags[r] = (unsigned int*)base;
        mbox[r] = (double*)((char*)base + h->mailbox_flag_bytes);
    }
    MPCB_CUDA_TRY(cudaMalloc(&h->d_peer_mbox, G * sizeof(double*)));
    MPCB_CUDA_TRY(cudaMalloc(&h->d_peer_flags, G * sizeof(unsigned int*)));
    MPCB_CUDA_TRY(cudaMemcpy(h->d_peer_mbox, mbox, G * sizeof(double*), cudaMemcpyHostToDevice));
    MPCB_CUDA_TRY(cudaMemcpy(h->d_peer_flags, flags, G * sizeof(unsigned int*), cudaMemcpyHostToDevice));
    h->peers_attached = true;
    return MPCB_OK;
}

}  // extern "C"
