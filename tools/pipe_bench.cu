// pipe_bench.cu — developer microbenchmark: issue rate of the instruction forms the MPPI rollout is made of, per SM
// sub-partition (SMSP), on B200.  Each test is a loop of 8 independent chains of ONE instruction form (or a fixed
// mix); the kernel runs W warps per SMSP on every SM and reports warp-instructions per cycle per SMSP from
// clock64().  The numbers are the pipe model DESIGN.md uses for the per-step instruction budget.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart shared -o tools/pipe_bench tools/pipe_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#define ITERS 2048
#define CH 8

// body(i) is PTX for chain i operating on f[i] (float), g[i] (float), u[i] (u32), d[i] (double)
#define DEFINE_TEST(NAME, NINSTR, BODY)                                                                      \
    __global__ void k_##NAME(long long* cyc, float* sink, float a, float b, unsigned int ka) {               \
        float f[CH], g[CH], h[CH];                                                                           \
        unsigned int u[CH], v[CH];                                                                           \
        double d[CH];                                                                                        \
        unsigned long long q[CH], r[CH];                                                                     \
        _Pragma("unroll") for (int i = 0; i < CH; ++i) {                                                     \
            f[i] = a + threadIdx.x * 1e-3f + i;                                                              \
            g[i] = b + i * 0.25f;                                                                            \
            h[i] = a * 0.5f + i;                                                                             \
            u[i] = ka + threadIdx.x * 977u + i;                                                              \
            v[i] = ka * 3u + i;                                                                              \
            d[i] = (double)f[i];                                                                             \
            q[i] = ((unsigned long long)__float_as_uint(f[i]) << 32) | __float_as_uint(g[i]);                \
            r[i] = ((unsigned long long)__float_as_uint(h[i]) << 32) | __float_as_uint(g[i] + 1.0f);         \
        }                                                                                                    \
        __syncthreads();                                                                                     \
        const long long t0 = clock64();                                                                      \
        for (int it = 0; it < ITERS; ++it) {                                                                 \
            _Pragma("unroll") for (int i = 0; i < CH; ++i) { BODY }                                          \
        }                                                                                                    \
        const long long t1 = clock64();                                                                      \
        float s = 0.0f;                                                                                      \
        _Pragma("unroll") for (int i = 0; i < CH; ++i)                                                       \
            s += f[i] + g[i] + h[i] + (float)u[i] + (float)v[i] + (float)d[i] + (float)(q[i] >> 40) + (float)(r[i] >> 40); \
        if (s == 12345.678f) sink[0] = s;                                                                    \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                                     \
    }                                                                                                        \
    static const int n_##NAME = NINSTR;

DEFINE_TEST(ffma_rrr, 1, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g[i]), "f"(h[i]));)
DEFINE_TEST(ffma_rrr_rot, 1, asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(f[(i + 3) % CH]), "f"(f[(i + 5) % CH]));)
DEFINE_TEST(ffma_rr_imm, 1, asm volatile("fma.rn.f32 %0, %0, %1, 0f3FC00000;" : "+f"(f[i]) : "f"(g[i]));)
DEFINE_TEST(ffma_r_imm_r, 1, asm volatile("fma.rn.f32 %0, %0, 0f3F800100, %1;" : "+f"(f[i]) : "f"(g[i]));)
DEFINE_TEST(ffma_r_imm_imm, 1, asm volatile("fma.rn.f32 %0, %0, 0f3F800100, 0f3A000000;" : "+f"(f[i]));)
DEFINE_TEST(ffma_r_param_r, 1, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(a), "f"(g[i]));)
DEFINE_TEST(fmul_rr, 1, asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));)
DEFINE_TEST(fmul_r_imm, 1, asm volatile("mul.rn.f32 %0, %0, 0f3F800100;" : "+f"(f[i]));)
DEFINE_TEST(fadd_rr, 1, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));)
DEFINE_TEST(fmnmx_rr, 1, asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));)
DEFINE_TEST(fmnmx_r_imm, 1, asm volatile("min.f32 %0, %0, 0f41A00000;" : "+f"(f[i]));)
DEFINE_TEST(fmnmx_xorsign, 1, asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i]));)
DEFINE_TEST(lop3_rrr, 1, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(v[i]), "r"(u[(i + 1) % CH]));)
DEFINE_TEST(lop3_rr, 1, asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(v[i]));)
DEFINE_TEST(iadd_rr, 1, asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(v[i]));)
DEFINE_TEST(shl_r, 1, asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));)
DEFINE_TEST(selp, 1, asm volatile("{.reg .pred p; setp.gt.f32 p, %1, 0f00000000; selp.f32 %0, %0, %2, p;}" : "+f"(f[i]) : "f"(g[i]), "f"(h[i]));)
DEFINE_TEST(imad_wide_imm, 1, asm volatile("{.reg .u64 t; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {%0, %1}, t;}" : "+r"(u[i]), "+r"(v[i]));)
DEFINE_TEST(imad_lo_imm, 1, asm volatile("mad.lo.u32 %0, %0, 0xD2511F53, %1;" : "+r"(u[i]) : "r"(v[i]));)
DEFINE_TEST(imad_hi_imm, 1, asm volatile("mul.hi.u32 %0, %0, 0xD2511F53;" : "+r"(u[i]));)
DEFINE_TEST(mufu_rcp, 1, asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));)
DEFINE_TEST(mufu_sin, 1, asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f[i]));)
DEFINE_TEST(mufu_lg2, 1, asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));)
DEFINE_TEST(mufu_sqrt, 1, asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i]));)
DEFINE_TEST(mufu_ex2, 1, asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));)
DEFINE_TEST(i2f_u32, 1, asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(u[i])); asm volatile("" : "+r"(u[i]) : "f"(f[i]));)
DEFINE_TEST(f2f_64_32, 1, asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i])); asm volatile("" : "+f"(f[i]) : "d"(d[i]));)
DEFINE_TEST(dadd, 1, asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(d[(i + 1) % CH]));)
DEFINE_TEST(dfma, 1, asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(d[(i + 1) % CH]));)
DEFINE_TEST(ffma2_rrr, 1, asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(r[i]), "l"(r[(i + 1) % CH]));)
DEFINE_TEST(fmul2_rr, 1, asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(r[i]));)
DEFINE_TEST(ffma2_r_splat_r, 1, {
    unsigned long long sp;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(sp) : "f"(a));
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(sp), "l"(r[i]));
})
// mixes (instructions per chain iteration in NINSTR)
DEFINE_TEST(mix_ffma_rrr_fmnmx, 2, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g[i]), "f"(h[i])); asm volatile("min.f32 %0, %0, %1;" : "+f"(g[i]) : "f"(h[i]));)
DEFINE_TEST(mix_ffma_imm_fmnmx, 2, asm volatile("fma.rn.f32 %0, %0, %1, 0f3FC00000;" : "+f"(f[i]) : "f"(g[i])); asm volatile("min.f32 %0, %0, 0f41A00000;" : "+f"(h[i]));)
DEFINE_TEST(mix_ffma_imm_lop3, 2, asm volatile("fma.rn.f32 %0, %0, %1, 0f3FC00000;" : "+f"(f[i]) : "f"(g[i])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(v[i]));)
DEFINE_TEST(mix_ffma_rrr_fmul_imm, 2, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g[i]), "f"(h[i])); asm volatile("mul.rn.f32 %0, %0, 0f3F800100;" : "+f"(g[i]));)
DEFINE_TEST(mix_3ffma_imm_1mufu, 4, asm volatile("fma.rn.f32 %0, %0, %1, 0f3FC00000;" : "+f"(f[i]) : "f"(g[i])); asm volatile("fma.rn.f32 %0, %0, %1, 0f3FC00000;" : "+f"(g[i]) : "f"(f[i])); asm volatile("fma.rn.f32 %0, %0, 0f3F800100, 0f3A000000;" : "+f"(f[i])); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(h[i]));)
DEFINE_TEST(mix_imadwide_lop3, 2, asm volatile("{.reg .u64 t; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {%0, %1}, t;}" : "+r"(u[i]), "+r"(v[i])); asm volatile("xor.b32 %0, %0, %1;" : "+r"(u[i]) : "r"(v[(i + 1) % CH]));)
DEFINE_TEST(mix_ffma2_fmnmx2, 3, asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(q[i]) : "l"(r[i]), "l"(r[(i + 1) % CH])); asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g[i])); asm volatile("min.f32 %0, %0, %1;" : "+f"(h[i]) : "f"(g[i]));)

typedef void (*kfn)(long long*, float*, float, float, unsigned int);
struct Test {
    const char* name;
    kfn fn;
    int ninstr;
};
#define T(NAME) {#NAME, k_##NAME, n_##NAME}

int main(int argc, char** argv) {
    Test tests[] = {T(ffma_rrr), T(ffma_rrr_rot), T(ffma_rr_imm), T(ffma_r_imm_r), T(ffma_r_imm_imm), T(ffma_r_param_r), T(fmul_rr),
                    T(fmul_r_imm), T(fadd_rr), T(fmnmx_rr), T(fmnmx_r_imm), T(fmnmx_xorsign), T(lop3_rrr), T(lop3_rr), T(iadd_rr), T(shl_r),
                    T(selp), T(imad_wide_imm), T(imad_lo_imm), T(imad_hi_imm), T(mufu_rcp), T(mufu_sin), T(mufu_lg2), T(mufu_sqrt),
                    T(mufu_ex2), T(i2f_u32), T(f2f_64_32), T(dadd), T(dfma), T(ffma2_rrr), T(fmul2_rr), T(ffma2_r_splat_r),
                    T(mix_ffma_rrr_fmnmx), T(mix_ffma_imm_fmnmx), T(mix_ffma_imm_lop3), T(mix_ffma_rrr_fmul_imm), T(mix_3ffma_imm_1mufu),
                    T(mix_imadwide_lop3), T(mix_ffma2_fmnmx2)};
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    long long* cyc;
    float* sink;
    cudaMalloc(&cyc, nsm * sizeof(long long));
    cudaMalloc(&sink, 64);
    long long* h = new long long[nsm];
    printf("%-24s %8s %8s %8s %8s   (warp-instructions per cycle per SMSP at W warps per SMSP)\n", "test", "W=1", "W=2", "W=4", "W=8");
    for (const Test& t : tests) {
        if (argc > 1 && !strstr(t.name, argv[1])) continue;
        printf("%-24s", t.name);
        for (int W : {1, 2, 4, 8}) {
            for (int rep = 0; rep < 2; ++rep) t.fn<<<nsm, 128 * W>>>(cyc, sink, 1.0001f, 0.5f, 12345u);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < nsm; ++i) avg += (double)h[i];
            avg /= nsm;
            const double instr = (double)ITERS * CH * t.ninstr * W;  // per SMSP
            printf(" %8.3f", instr / avg);
        }
        cudaError_t e = cudaGetLastError();
        printf("%s\n", e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
