// loop_bench.cu — developer microbenchmark for the MPPI instruction diet: the producer loop (noise, clamp, control
// term) and the consumer loop (model NL step + clamped cost) of mppi_ws_kernel.cuh as stand-alone loops, in several
// formulations, timed with clock64() at W warps per SM.  Output: dispatch cycles per sample-step per scheduler
// (cycles * 4 / (W * H)), the figure the per-step budget in DESIGN.md is written in.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --cudart shared -I mpc_rs_b200/csrc -o tools/loop_bench tools/loop_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include "f32x2.cuh"
#include "philox.cuh"

using namespace mpcb;

struct Consts {
    float D, E2, T1, KTR, ML, JML, T4, dt;          // model NL (models.cuh)
    float w0, w1, w2, w3, c0, c1, k1, k2, c2;       // clamped cost
    float KU, K3, K1, DT1, DT4;                     // diet form
    float lo, hi, neg2s2ln2;
};

// default constants of examples/mppi4-non-liner.rs as compile-time values (dt = 0.008)
namespace imm {
constexpr double M1 = 0.15, RW = 0.05, M2 = 2.3 - 2.0 * 0.15 + 2.0, L = 0.2474, J1 = M1 * RW * RW, J2 = 0.2, G = 9.81, KT = 0.15, DTd = 0.008;
constexpr float D = (float)((M1 + M2 + J1 / (RW * RW)) * (M2 * L * L + J2));
constexpr float E2 = (float)(M2 * M2 * L * L);
constexpr float T1 = (float)((M1 + M2 + J1 / (RW * RW)) * M2 * G * L);
constexpr float KTR = (float)(KT / RW);
constexpr float ML = (float)(M2 * L);
constexpr float JML = (float)(J2 + M2 * L * L);
constexpr float T4 = (float)(M2 * G * L * L);
constexpr float dt = (float)DTd;
constexpr float w0 = 2.0f, w1 = 3.0f, w2 = 5.0f, w3 = 1.2f, c0 = 2.0f, c1 = 5.0f, k1 = 2.0f, k2 = 0.35f, c2 = 0.75f;
constexpr float KU = (float)((KT / RW) / (M2 * L));
constexpr float K3 = (float)((M2 * L) * (M2 * L) / ((M1 + M2 + J1 / (RW * RW)) * M2 * G * L));
constexpr float K1 = (float)((J2 + M2 * L * L) * (M2 * L) / (M2 * G * L * L));
constexpr float DT1 = (float)(DTd * (M1 + M2 + J1 / (RW * RW)) * M2 * G * L);
constexpr float DT4 = (float)(DTd * M2 * G * L * L);
}  // namespace imm

static Consts host_consts() {
    Consts c;
    c.D = imm::D; c.E2 = imm::E2; c.T1 = imm::T1; c.KTR = imm::KTR; c.ML = imm::ML; c.JML = imm::JML; c.T4 = imm::T4; c.dt = imm::dt;
    c.w0 = imm::w0; c.w1 = imm::w1; c.w2 = imm::w2; c.w3 = imm::w3; c.c0 = imm::c0; c.c1 = imm::c1; c.k1 = imm::k1; c.k2 = imm::k2; c.c2 = imm::c2;
    c.KU = imm::KU; c.K3 = imm::K3; c.K1 = imm::K1; c.DT1 = imm::DT1; c.DT4 = imm::DT4;
    c.lo = -20.0f; c.hi = 20.0f; c.neg2s2ln2 = -2.0f * 9.0f * 0.69314718f;
    return c;
}

// IMM: constants are literals (immediate operands); else they come from the kernel parameter block
template <bool IMM>
struct K {
    const Consts& p;
    __device__ __forceinline__ K(const Consts& q) : p(q) {}
#define KGET(name) __device__ __forceinline__ float name() const { if constexpr (IMM) return imm::name; else return p.name; }
    KGET(D) KGET(E2) KGET(T1) KGET(KTR) KGET(ML) KGET(JML) KGET(T4) KGET(dt)
    KGET(w0) KGET(w1) KGET(w2) KGET(w3) KGET(c0) KGET(c1) KGET(k1) KGET(k2) KGET(c2)
    KGET(KU) KGET(K3) KGET(K1) KGET(DT1) KGET(DT4)
#undef KGET
};

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ float rcp_approx(float d) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
}

// ---- sincos variants ----
// S0: the shipped one (pi/2 reduction, swap, two sign flips)
__device__ __forceinline__ void sincos_s0(float a, float* s, float* c) {
    const float magic = 12582912.0f;
    float j = fmaf(a, 0.63661977236758134308f, magic);
    const int q = __float_as_int(j);
    j -= magic;
    float r = fmaf(j, -1.5707962512969970703f, a);
    r = fmaf(j, -7.5497894158615963534e-08f, r);
    const float r2 = r * r;
    const float ps = fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f);
    const float sp = fmaf(ps, r2 * r, r);
    const float pc = fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f);
    const float cp = fmaf(pc, r2 * r2, fmaf(-0.5f, r2, 1.0f));
    const bool swap = (q & 1) != 0;
    const float ss = swap ? cp : sp;
    const float cc = swap ? sp : cp;
    *s = __int_as_float(__float_as_int(ss) ^ ((q & 2) << 30));
    *c = __int_as_float(__float_as_int(cc) ^ (((q + 1) & 2) << 30));
}
// S1: pi reduction, no swap: sin(a) = (-1)^j sin(r), cos(a) = (-1)^j cos(r), r in [-pi/2, pi/2]; degree 9 / 10
// (coefficients: least-squares fit on Chebyshev nodes, see tools/fit_sincos.py)
__device__ __forceinline__ void sincos_s1(float a, float* s, float* c) {
    const float magic = 12582912.0f;
    float j = fmaf(a, 0.31830988618379067154f, magic);
    const unsigned int sgn = __float_as_uint(j) << 31;
    j -= magic;
    float r = fmaf(j, -3.1415925025939941406f, a);
    r = fmaf(j, -1.5099578831723192707e-07f, r);
    const float r2 = r * r;
    float ps = fmaf(2.6083159809786593e-06f, r2, -1.9810690719168633e-04f);
    ps = fmaf(ps, r2, 8.3330258727073669e-03f);
    ps = fmaf(ps, r2, -1.6666665673255920e-01f);
    const float sp = fmaf(ps, r2 * r, r);
    float pc = fmaf(-2.6051615464115930e-07f, r2, 2.4760495740302590e-05f);
    pc = fmaf(pc, r2, -1.3888377165013552e-03f);
    pc = fmaf(pc, r2, 4.1666645556688309e-02f);
    pc = fmaf(pc, r2, -0.5f);
    const float cp = fmaf(pc, r2, 1.0f);
    *s = __uint_as_float(__float_as_uint(sp) ^ sgn);
    *c = __uint_as_float(__float_as_uint(cp) ^ sgn);
}
// S2: MUFU sin/cos (rejected in round 1 for accuracy; here as the lower bound)
__device__ __forceinline__ void sincos_s2(float a, float* s, float* c) {
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(*s) : "f"(a));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(*c) : "f"(a));
}

template <int SC>
__device__ __forceinline__ void sincos_v(float a, float* s, float* c) {
    if constexpr (SC == 0) sincos_s0(a, s, c);
    else if constexpr (SC == 1) sincos_s1(a, s, c);
    else sincos_s2(a, s, c);
}

// ---- consumer step variants ----
// FORM 0: shipped formulation (models.cuh ModelNL<float> + CostClamped::acc), Newton step on the reciprocal
// FORM 1: diet: no Newton, num3/num1 scaled so that two constant multiplies disappear
template <int FORM, int SC, bool IMM>
__device__ __forceinline__ float step_cost(float (&x)[4], float u, float acc, const K<IMM>& k) {
    float s, c;
    sincos_v<SC>(x[2], &s, &c);
    if constexpr (FORM == 0) {
        const float d = fmaf(-k.E2() * c, c, k.D());
        float r = rcp_approx(d);
        const float e = fmaf(-d, r, 1.0f);
        r = fmaf(r, e, r);
        const float idt = r * k.dt();
        const float q = fmaf(k.ML() * (x[3] * x[3]), s, k.KTR() * u);
        const float num3 = fmaf(-k.ML() * q, c, k.T1() * s);
        const float num1 = fmaf(k.T4() * s, c, k.JML() * q);
        const float r3 = fmaf(num3, idt, x[3]);
        const float r2 = fmaf(x[3], k.dt(), x[2]);
        const float r1 = fmaf(num1, idt, x[1]);
        const float r0 = fmaf(x[1], k.dt(), x[0]);
        x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
    } else {
        const float d = fmaf(-k.E2(), c * c, k.D());
        const float rd = rcp_approx(d);
        const float a = (x[3] * x[3]) * s;
        const float qq = fmaf(k.KU(), u, a);
        const float n3 = fmaf(-k.K3(), qq * c, s);
        const float n1 = fmaf(k.K1(), qq, s * c);
        const float r3 = fmaf(n3, rd * k.DT1(), x[3]);
        const float r1 = fmaf(n1, rd * k.DT4(), x[1]);
        const float r2 = fmaf(x[3], k.dt(), x[2]);
        const float r0 = fmaf(x[1], k.dt(), x[0]);
        x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
    }
    const float xc = clampf(x[0], -k.c0(), k.c0());
    const float a1 = clampf(fmaf(k.k1(), xc, x[1]), -k.c1(), k.c1());
    const float b = fmaf(k.k2(), clampf(x[0], -k.c2(), k.c2()), x[2]);
    if constexpr (FORM == 0) {
        acc = fmaf(k.w0() * xc, xc, acc);
        acc = fmaf(k.w1() * a1, a1, acc);
        acc = fmaf(k.w2() * b, b, acc);
        return fmaf(k.w3() * x[3], x[3], acc);
    } else {
        acc = fmaf(k.w0(), xc * xc, acc);
        acc = fmaf(k.w1(), a1 * a1, acc);
        acc = fmaf(k.w2(), b * b, acc);
        return fmaf(k.w3(), x[3] * x[3], acc);
    }
}

// packed consumer (two samples per thread), shipped formulation
__device__ __forceinline__ void sincos2(f2 a, f2* s, f2* c) {
    const float magic = 12582912.0f;
    f2 j = fma2(a, splat2(0.63661977236758134308f), splat2(magic));
    float jl, jh;
    un2(j, jl, jh);
    const int ql = __float_as_int(jl), qh = __float_as_int(jh);
    j = add2(j, splat2(-magic));
    f2 r = fma2(j, splat2(-1.5707962512969970703f), a);
    r = fma2(j, splat2(-7.5497894158615963534e-08f), r);
    const f2 r2 = mul2(r, r);
    const f2 ps = fma2(fma2(splat2(-1.9515295891e-4f), r2, splat2(8.3321608736e-3f)), r2, splat2(-1.6666654611e-1f));
    const f2 sp = fma2(ps, mul2(r2, r), r);
    const f2 pc = fma2(fma2(splat2(2.443315711809948e-5f), r2, splat2(-1.388731625493765e-3f)), r2, splat2(4.166664568298827e-2f));
    const f2 cp = fma2(pc, mul2(r2, r2), fma2(splat2(-0.5f), r2, splat2(1.0f)));
    float spl, sph, cpl, cph;
    un2(sp, spl, sph);
    un2(cp, cpl, cph);
    const float ssl = (ql & 1) ? cpl : spl, ccl = (ql & 1) ? spl : cpl;
    const float ssh = (qh & 1) ? cph : sph, cch = (qh & 1) ? sph : cph;
    *s = mk2(__int_as_float(__float_as_int(ssl) ^ ((ql & 2) << 30)), __int_as_float(__float_as_int(ssh) ^ ((qh & 2) << 30)));
    *c = mk2(__int_as_float(__float_as_int(ccl) ^ (((ql + 1) & 2) << 30)), __int_as_float(__float_as_int(cch) ^ (((qh + 1) & 2) << 30)));
}
// packed, pi reduction (S1) and the diet formulation
__device__ __forceinline__ void sincos2_s1(f2 a, f2* s, f2* c) {
    const float magic = 12582912.0f;
    f2 j = fma2(a, splat2(0.31830988618379067154f), splat2(magic));
    float jl, jh;
    un2(j, jl, jh);
    const unsigned int sl = __float_as_uint(jl) << 31, sh = __float_as_uint(jh) << 31;
    j = add2(j, splat2(-magic));
    f2 r = fma2(j, splat2(-3.1415925025939941406f), a);
    r = fma2(j, splat2(-1.5099578831723192707e-07f), r);
    const f2 r2 = mul2(r, r);
    f2 ps = fma2(splat2(2.6083159809786593e-06f), r2, splat2(-1.9810690719168633e-04f));
    ps = fma2(ps, r2, splat2(8.3330258727073669e-03f));
    ps = fma2(ps, r2, splat2(-1.6666665673255920e-01f));
    const f2 sp = fma2(ps, mul2(r2, r), r);
    f2 pc = fma2(splat2(-2.6051615464115930e-07f), r2, splat2(2.4760495740302590e-05f));
    pc = fma2(pc, r2, splat2(-1.3888377165013552e-03f));
    pc = fma2(pc, r2, splat2(4.1666645556688309e-02f));
    pc = fma2(pc, r2, splat2(-0.5f));
    const f2 cp = fma2(pc, r2, splat2(1.0f));
    float spl, sph, cpl, cph;
    un2(sp, spl, sph);
    un2(cp, cpl, cph);
    *s = mk2(__uint_as_float(__float_as_uint(spl) ^ sl), __uint_as_float(__float_as_uint(sph) ^ sh));
    *c = mk2(__uint_as_float(__float_as_uint(cpl) ^ sl), __uint_as_float(__float_as_uint(cph) ^ sh));
}
template <int FORM>
__device__ __forceinline__ f2 step_cost2(f2 (&x)[4], f2 u, f2 acc, const Consts& k) {
    f2 s, c;
    if constexpr (FORM == 0) sincos2(x[2], &s, &c);
    else sincos2_s1(x[2], &s, &c);
    if constexpr (FORM == 0) {
        const f2 nd = fma2(mul2(splat2(k.E2), c), c, splat2(-k.D));
        float l, h, rl, rh;
        un2(nd, l, h);
        rl = rcp_approx(-l);
        rh = rcp_approx(-h);
        f2 r = mk2(rl, rh);
        const f2 e = fma2(nd, r, splat2(1.0f));
        r = fma2(r, e, r);
        const f2 idt = mul2(r, splat2(k.dt));
        const f2 q = fma2(mul2(splat2(k.ML), mul2(x[3], x[3])), s, mul2(splat2(k.KTR), u));
        const f2 num3 = fma2(mul2(splat2(-k.ML), q), c, mul2(splat2(k.T1), s));
        const f2 num1 = fma2(mul2(splat2(k.T4), s), c, mul2(splat2(k.JML), q));
        const f2 r3 = fma2(num3, idt, x[3]);
        const f2 r2 = fma2(x[3], splat2(k.dt), x[2]);
        const f2 r1 = fma2(num1, idt, x[1]);
        const f2 r0 = fma2(x[1], splat2(k.dt), x[0]);
        x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
    } else {
        const f2 d = fma2(splat2(-k.E2), mul2(c, c), splat2(k.D));
        float l, h;
        un2(d, l, h);
        const f2 rd = mk2(rcp_approx(l), rcp_approx(h));
        const f2 a = mul2(mul2(x[3], x[3]), s);
        const f2 qq = fma2(splat2(k.KU), u, a);
        const f2 n3 = fma2(splat2(-k.K3), mul2(qq, c), s);
        const f2 n1 = fma2(splat2(k.K1), qq, mul2(s, c));
        const f2 r3 = fma2(n3, mul2(rd, splat2(k.DT1)), x[3]);
        const f2 r1 = fma2(n1, mul2(rd, splat2(k.DT4)), x[1]);
        const f2 r2 = fma2(x[3], splat2(k.dt), x[2]);
        const f2 r0 = fma2(x[1], splat2(k.dt), x[0]);
        x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
    }
    float x0l, x0h;
    un2(x[0], x0l, x0h);
    const f2 xc = mk2(clampf(x0l, -k.c0, k.c0), clampf(x0h, -k.c0, k.c0));
    const f2 t = fma2(splat2(k.k1), xc, x[1]);
    float tl, th;
    un2(t, tl, th);
    const f2 a1 = mk2(clampf(tl, -k.c1, k.c1), clampf(th, -k.c1, k.c1));
    const f2 b = fma2(splat2(k.k2), mk2(clampf(x0l, -k.c2, k.c2), clampf(x0h, -k.c2, k.c2)), x[2]);
    if constexpr (FORM == 0) {
        acc = fma2(mul2(splat2(k.w0), xc), xc, acc);
        acc = fma2(mul2(splat2(k.w1), a1), a1, acc);
        acc = fma2(mul2(splat2(k.w2), b), b, acc);
        return fma2(mul2(splat2(k.w3), x[3]), x[3], acc);
    } else {
        acc = fma2(splat2(k.w0), mul2(xc, xc), acc);
        acc = fma2(splat2(k.w1), mul2(a1, a1), acc);
        acc = fma2(splat2(k.w2), mul2(b, b), acc);
        return fma2(splat2(k.w3), mul2(x[3], x[3]), acc);
    }
}

// ---- kernels ----
// consumer: tile[g][tid] float4 of controls (pre-filled), H steps
template <int FORM, int SC, bool IMM>
__global__ void k_consumer(long long* cyc, float* out, const float4* vin, int H, const __grid_constant__ Consts cst) {
    extern __shared__ float4 tile[];
    const int nthr = blockDim.x, Hq = H / 4;
    for (int i = threadIdx.x; i < Hq * nthr; i += nthr) tile[i] = vin[i % 1024];
    __syncthreads();
    K<IMM> k(cst);
    float x[4] = {0.5f, 0.0f, 0.1f + 1e-4f * threadIdx.x, 0.0f};
    float acc = 0.0f;
    double J = 0.0;
    const long long t0 = clock64();
    for (int g = 0; g < Hq; ++g) {
        const float4 v4 = tile[g * nthr + threadIdx.x];
        acc = step_cost<FORM, SC, IMM>(x, v4.x, acc, k);
        acc = step_cost<FORM, SC, IMM>(x, v4.y, acc, k);
        acc = step_cost<FORM, SC, IMM>(x, v4.z, acc, k);
        acc = step_cost<FORM, SC, IMM>(x, v4.w, acc, k);
        J += (double)acc;
        acc = 0.0f;
    }
    const long long t1 = clock64();
    out[blockIdx.x * nthr + threadIdx.x] = (float)J + x[0];
    __shared__ long long mx;
    if (threadIdx.x == 0) mx = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&mx, (unsigned long long)(t1 - t0));
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = mx;
}
template <int FORM>
__global__ void k_consumer2(long long* cyc, float* out, const float4* vin, int H, const __grid_constant__ Consts cst) {
    extern __shared__ float4 tile[];
    const int nthr = blockDim.x, Hq = H / 4;
    for (int i = threadIdx.x; i < 2 * Hq * nthr; i += nthr) tile[i] = vin[i % 1024];
    __syncthreads();
    f2 x[4] = {splat2(0.5f), splat2(0.0f), mk2(0.1f + 1e-4f * threadIdx.x, 0.1f - 1e-4f * threadIdx.x), splat2(0.0f)};
    f2 acc = splat2(0.0f);
    double J0 = 0.0, J1 = 0.0;
    const long long t0 = clock64();
    for (int g = 0; g < Hq; ++g) {
        const float4 a4 = tile[(2 * g) * nthr + threadIdx.x], b4 = tile[(2 * g + 1) * nthr + threadIdx.x];
        acc = step_cost2<FORM>(x, mk2(a4.x, a4.y), acc, cst);
        acc = step_cost2<FORM>(x, mk2(a4.z, a4.w), acc, cst);
        acc = step_cost2<FORM>(x, mk2(b4.x, b4.y), acc, cst);
        acc = step_cost2<FORM>(x, mk2(b4.z, b4.w), acc, cst);
        float ja, jb;
        un2(acc, ja, jb);
        J0 += (double)ja;
        J1 += (double)jb;
        acc = splat2(0.0f);
    }
    const long long t1 = clock64();
    out[blockIdx.x * nthr + threadIdx.x] = (float)(J0 + J1) + lo2(x[0]);
    __shared__ long long mx;
    if (threadIdx.x == 0) mx = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&mx, (unsigned long long)(t1 - t0));
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = mx;
}

// producer: PV 0 = shipped (Philox + Box-Muller + clamp + control term + STS.128), 1 = Philox only (bits -> floats),
// 2 = Box-Muller only (cheap counter words), 3 = xoshiro128+ stream (seeded by one Philox block) + Box-Muller,
// 4 = shipped but 8 normals per Philox block (16-bit radius and angle words)
__device__ __forceinline__ void bm4(const unsigned int (&w)[4], float neg2s2ln2, float (&z)[4]) {
    Philox4 r;
    r.v[0] = w[0]; r.v[1] = w[1]; r.v[2] = w[2]; r.v[3] = w[3];
    philox_normal4(r, neg2s2ln2, z);
}
template <int PV>
__global__ void k_producer(long long* cyc, float* out, const float4* vin, int H, const __grid_constant__ Consts cst) {
    extern __shared__ float4 tile[];
    __shared__ float su[512], sui[512];
    const int nthr = blockDim.x, Hq = H / 4;
    for (int i = threadIdx.x; i < H; i += nthr) { su[i] = 0.01f * i; sui[i] = 0.001f * i; }
    __syncthreads();
    const unsigned int kg = blockIdx.x * nthr + threadIdx.x;
    float cu = 0.0f;
    double CT = 0.0;
    unsigned int s0 = kg * 2654435761u + 1u, s1 = kg ^ 0x9E3779B9u, s2 = kg + 0x7F4A7C15u, s3 = ~kg;
    if (PV == 3) {
        const Philox4 r = philox4x32(kg, 7u, 0u, 0u, 123u, 456u);
        s0 = r.v[0]; s1 = r.v[1]; s2 = r.v[2]; s3 = r.v[3] | 1u;
    }
    const long long t0 = clock64();
    for (int g = 0; g < Hq; ++g) {
        float e[4];
        if constexpr (PV == 0) {
            const Philox4 r = philox4x32(kg, 7u, 0u, (unsigned int)g, 123u, 456u);
            philox_normal4(r, cst.neg2s2ln2, e);
        } else if constexpr (PV == 1) {
            const Philox4 r = philox4x32(kg, 7u, 0u, (unsigned int)g, 123u, 456u);
#pragma unroll
            for (int i = 0; i < 4; ++i) e[i] = __uint_as_float(0x3f800000u | (r.v[i] >> 9)) - 1.5f;
        } else if constexpr (PV == 2) {
            const unsigned int w[4] = {s0 + g * 0x9E3779B9u, s1 ^ (g * 0x85EBCA6Bu), s2 + g * 0xC2B2AE35u, s3 ^ (g * 0x27D4EB2Fu)};
            bm4(w, cst.neg2s2ln2, e);
        } else if constexpr (PV == 3) {
            unsigned int w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {  // xoshiro128+
                w[i] = s0 + s3;
                const unsigned int t = s1 << 9;
                s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t;
                s3 = __funnelshift_l(s3, s3, 11);
            }
            bm4(w, cst.neg2s2ln2, e);
        } else {
            // 8 normals per Philox block: even groups draw, odd groups use the second half
            if ((g & 1) == 0) {
                const Philox4 r = philox4x32(kg, 7u, 0u, (unsigned int)(g >> 1), 123u, 456u);
                s0 = r.v[0]; s1 = r.v[1]; s2 = r.v[2]; s3 = r.v[3];
            }
            const unsigned int a = (g & 1) ? s2 : s0, b = (g & 1) ? s3 : s1;
            const unsigned int w[4] = {a << 16, a & 0xffff0000u, b << 16, b & 0xffff0000u};
            bm4(w, cst.neg2s2ln2, e);
        }
        const float4 u4 = *reinterpret_cast<const float4*>(su + 4 * g);
        const float4 ui4 = *reinterpret_cast<const float4*>(sui + 4 * g);
        const float v0 = clampf(u4.x + e[0], cst.lo, cst.hi), v1 = clampf(u4.y + e[1], cst.lo, cst.hi);
        const float v2 = clampf(u4.z + e[2], cst.lo, cst.hi), v3 = clampf(u4.w + e[3], cst.lo, cst.hi);
        cu = fmaf(ui4.x, v0, cu); cu = fmaf(ui4.y, v1, cu); cu = fmaf(ui4.z, v2, cu); cu = fmaf(ui4.w, v3, cu);
        tile[g * nthr + threadIdx.x] = make_float4(v0, v1, v2, v3);
        CT += (double)cu;
        cu = 0.0f;
    }
    const long long t1 = clock64();
    out[blockIdx.x * nthr + threadIdx.x] = (float)CT + tile[threadIdx.x].x;
    __shared__ long long mx;
    if (threadIdx.x == 0) mx = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)&mx, (unsigned long long)(t1 - t0));
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = mx;
}

typedef void (*kfn)(long long*, float*, const float4*, int, const Consts);
struct Test {
    const char* name;
    kfn fn;
    int samples_per_thread;
    int cells_per_group;  // float4 cells of the tile per thread and 4-step group
};

int main(int argc, char** argv) {
    Test tests[] = {
        {"consumer scalar shipped (param consts)", k_consumer<0, 0, false>, 1, 1},
        {"consumer scalar shipped (imm consts)", k_consumer<0, 0, true>, 1, 1},
        {"consumer scalar diet (param)", k_consumer<1, 0, false>, 1, 1},
        {"consumer scalar diet (imm)", k_consumer<1, 0, true>, 1, 1},
        {"consumer scalar diet + pi-sincos (param)", k_consumer<1, 1, false>, 1, 1},
        {"consumer scalar diet + pi-sincos (imm)", k_consumer<1, 1, true>, 1, 1},
        {"consumer scalar diet + MUFU sincos (imm)", k_consumer<1, 2, true>, 1, 1},
        {"consumer packed shipped", k_consumer2<0>, 2, 2},
        {"consumer packed diet + pi-sincos", k_consumer2<1>, 2, 2},
        {"producer shipped", k_producer<0>, 1, 1},
        {"producer Philox only", k_producer<1>, 1, 1},
        {"producer Box-Muller only", k_producer<2>, 1, 1},
        {"producer xoshiro128+ + Box-Muller", k_producer<3>, 1, 1},
        {"producer Philox/8 normals + Box-Muller", k_producer<4>, 1, 1},
    };
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int H = 100;
    long long* cyc;
    float* out;
    float4* vin;
    cudaMalloc(&cyc, nsm * sizeof(long long));
    cudaMalloc(&out, (size_t)nsm * 1024 * 4);
    cudaMalloc(&vin, 1024 * sizeof(float4));
    {
        float4 hv[1024];
        unsigned int s = 12345u;
        for (int i = 0; i < 1024; ++i) {
            float t[4];
            for (int j = 0; j < 4; ++j) { s = s * 1664525u + 1013904223u; t[j] = ((s >> 8) * (1.0f / 16777216.0f) - 0.5f) * 12.0f; }
            hv[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
        cudaMemcpy(vin, hv, sizeof(hv), cudaMemcpyHostToDevice);
    }
    const Consts cst = host_consts();
    long long* h = new long long[nsm];
    printf("%-44s %9s %9s %9s   dispatch cycles per sample-step per scheduler at W warps per SM (H = %d)\n", "loop", "W=8", "W=14", "W=28", H);
    for (const Test& t : tests) {
        if (argc > 1 && !strstr(t.name, argv[1])) continue;
        printf("%-44s", t.name);
        for (int W : {8, 14, 28}) {
            const int nthr = 32 * W;
            const size_t smem = (size_t)(H / 4) * nthr * 16 * t.cells_per_group;
            if (smem > 220 * 1024 || nthr > 1024) { printf(" %9s", "-"); continue; }
            cudaFuncSetAttribute((const void*)t.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int rep = 0; rep < 3; ++rep) t.fn<<<nsm, nthr, smem>>>(cyc, out, vin, H, cst);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < nsm; ++i) avg += (double)h[i];
            avg /= nsm;
            printf(" %9.1f", avg * 4.0 / ((double)W * t.samples_per_thread * H));
        }
        cudaError_t e = cudaGetLastError();
        printf("%s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
