// mppi_rtc.h — run-time compilation (NVRTC) of the fused MPPI kernel around a user-supplied dynamics/cost pair.
#pragma once

#include "common.cuh"

namespace mpcb {

struct RtcModule {
    void* library = nullptr;                 // cudaLibrary_t
    void* kernel[3] = {nullptr, nullptr, nullptr};  // cudaKernel_t per MppiNoise, usable wherever a `const void* func` is
};

// Compiles mppi_rollout_kernel<ModelUser, real, block, noise, 1, vt> (state dimension S = state_dim) for the three noise
// modes around `user_src` and,
// when `load` is set, loads the cubin on the current device.  The compile log (errors AND warnings) is kept per thread
// for mpcb_rtc_log().  MPCB_RTC_ERROR on a compile failure.
mpcb_status rtc_compile_mppi_user(const char* user_src, int state_dim, bool f64, int block, bool vt, bool load, RtcModule* out);
// Same for the batched UKF: ukf_kernel<n, o, MPCB_MODEL_USER_UKF, sqrt, order, mode, fast> for the three UkfMode values
// (kernel[0..2] = predict, update, fused) around the user's fx / hx.  `fast` = FMA contraction on (cfg.exact == 0).
mpcb_status rtc_compile_ukf_user(const char* user_src, int n, int o, int sqrt_mode, int order, bool fast, bool load, RtcModule* out);
void rtc_unload(RtcModule* m);
const char* rtc_log();

}  // namespace mpcb
