// f32x2.cuh — two FP32 values in one 64-bit register pair and the packed arithmetic sm_100 adds for them
// (add/mul/fma.rn.f32x2 -> FADD2/FMUL2/FFMA2).  A packed instruction keeps the FMA pipe busy for two cycles, so
// the FLOP rate is that of scalar FFMA (measured: 67.7 vs 66-72 TFLOP/s), but it takes ONE issue slot for two
// samples — and the MPPI rollout is bound by issue slots shared with the ALU pipe, not by the FMA pipe.
// Component .lo is the thread's first sample, .hi its second.
#pragma once

namespace mpcb {

struct f2 {
    unsigned long long v;
};

__device__ __forceinline__ f2 mk2(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void un2(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ float lo2(f2 a) {
    float l, h;
    un2(a, l, h);
    return l;
}
__device__ __forceinline__ float hi2(f2 a) {
    float l, h;
    un2(a, l, h);
    return h;
}
__device__ __forceinline__ f2 splat2(float a) { return mk2(a, a); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
// per-component clamp (FMNMX has no packed form)
__device__ __forceinline__ f2 clamp2(f2 a, float lo, float hi) {
    float l, h;
    un2(a, l, h);
    return mk2(fminf(fmaxf(l, lo), hi), fminf(fmaxf(h, lo), hi));
}

}  // namespace mpcb
