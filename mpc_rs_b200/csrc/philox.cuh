// philox.cuh — Philox4x32 counter-based generator (Salmon, Moraes, Dror, Shaw, SC'11) and the
// in-register Box-Muller that turns one 128-bit block into four N(0, sigma^2) draws.
// Rounds: kPhiloxRounds = 7 — the fewest rounds for which the paper reports Philox4x32 passing BigCrush
// ("Crush-resistant"; 10 is the paper's default with a safety margin).  Generate mode defines its own noise
// stream (the reference seeds from OS entropy), so the round count is a quality/speed choice, not a parity one;
// tests/test_mppi_gpu.py::test_generated_noise_statistics checks the moments, the tails and the serial correlation of the dumped stream.
// Replaces the rand_distr::Normal / Xoshiro256Plus::from_entropy draw of src/mppi.rs:38-45.
//
// Counter layout (so the sample set does not depend on how samples are sharded over GPUs or blocks):
//   ctr = (global sample index low 32, call index, controller, t/4 | sample-high-bits << 16)
//   key = (seed low 32, seed high 32)
#pragma once

#include <stdint.h>

namespace mpcb {

struct Philox4 {
    uint32_t v[4];
};

#ifdef __CUDACC__
#define MPCB_HD __host__ __device__ __forceinline__
#else
#define MPCB_HD inline
#endif

MPCB_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
    // one 32x32->64 multiply (IMAD.WIDE.U32) instead of mul.lo + mul.hi
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    *lo = (uint32_t)p;
    *hi = (uint32_t)(p >> 32);
}

#ifndef MPCB_PHILOX_ROUNDS
#define MPCB_PHILOX_ROUNDS 7
#endif
constexpr int kPhiloxRounds = MPCB_PHILOX_ROUNDS;

MPCB_HD Philox4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < kPhiloxRounds; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(M0, c0, &hi0, &lo0);
        philox_mulhilo(M1, c2, &hi1, &lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

#ifdef __CUDACC__
// Four N(0, sigma^2) draws from one Philox block.  neg2s2ln2 = -2*sigma^2*ln(2), so that
// radius = sqrt(neg2s2ln2 * log2(u1)).  MUFU lg2/sqrt/sin/cos: ~2^-21 absolute error, irrelevant for noise.
__device__ __forceinline__ void philox_normal4(const Philox4& r, float neg2s2ln2, float (&z)[4]) {
    constexpr float k2m32 = 2.3283064365386963e-10f;  // 2^-32
    constexpr float k2m33 = 1.1641532182693481e-10f;  // 2^-33
    constexpr float kTwoPi = 6.283185307179586f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        // radius word: full 32 bits -> u1 in (0, 1] (tails out to 6.7 sigma)
        const float u1 = fmaf((float)r.v[2 * i], k2m32, k2m33);
        // angle word: top 23 bits spliced into the mantissa of [1, 2) — no int->float conversion on the XU pipe
        const float f12 = __uint_as_float(0x3f800000u | (r.v[2 * i + 1] >> 9));
        const float ang = fmaf(f12, kTwoPi, -kTwoPi);  // [0, 2pi)
        float lg, rad, s, c;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(neg2s2ln2 * lg));
        asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
        asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
        z[2 * i] = rad * c;
        z[2 * i + 1] = rad * s;
    }
}
#endif

}  // namespace mpcb
