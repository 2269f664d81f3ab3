// Developer microbenchmark: FP32 FMA throughput on B200 by operand form (what "FP32 peak" means for real code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fma_forms tools/fma_forms_bench.cu && /tmp/fma_forms
// Measured (profiles/README.md): FFMA with immediate/constant operands 71 TFLOP/s (the peak bench.py reports against);
// FFMA with three register operands 48; FFMA2 (fma.rn.f32x2) with three register operands 32, with one uniform
// operand 67.7; FMUL reg,reg and FFMA with one constant-bank operand run at the full rate; FMNMX at half rate.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// 3 varying register operands per FFMA: x = fma(x, y, z), y/z per-thread and per-chain
template <int ILP> __global__ void k_rrr(float* out, int iters, const float* in) {
    float x[ILP], y[ILP], z[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = in[threadIdx.x + i]; y[i] = in[threadIdx.x + 32 + i]; z[i] = in[threadIdx.x + 64 + i]; }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], y[i], z[i]);
    float s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mix: each fma uses the previous chains' values as operands (operands change every iteration, like real code)
template <int ILP> __global__ void k_mix(float* out, int iters, const float* in) {
    float x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = in[threadIdx.x + i];
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], x[(i + 3) % ILP], x[(i + 5) % ILP]);
    float s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k2_rrr(float* out, int iters, const float* in) {
    unsigned long long x[ILP], y[ILP], z[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = pk(in[threadIdx.x + i], in[threadIdx.x + i + 1]); y[i] = pk(in[threadIdx.x + 32 + i], in[threadIdx.x + 33 + i]); z[i] = pk(in[threadIdx.x + 64 + i], in[threadIdx.x + 65 + i]); }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma2(x[i], y[i], z[i]);
    float s = 0, lo, hi; for (int i = 0; i < ILP; ++i) { upk(x[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_imm(float* out, int iters, const float* in) {
    float x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = in[threadIdx.x + i];
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], 1.0001f, 1e-4f);
    float s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float *d, *in; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, ILP = 8; float ms;
    const double n = 148.0 * 8 * 256 * (double)iters * ILP;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_imm<ILP><<<148 * 8, 256>>>(d, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); double a = 2 * n / (ms * 1e-3) / 1e12;
        cudaEventRecord(e0); k_rrr<ILP><<<148 * 8, 256>>>(d, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); double b = 2 * n / (ms * 1e-3) / 1e12;
        cudaEventRecord(e0); k_mix<ILP><<<148 * 8, 256>>>(d, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); double c = 2 * n / (ms * 1e-3) / 1e12;
        cudaEventRecord(e0); k2_rrr<ILP><<<148 * 8, 256>>>(d, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); double e = 4 * n / (ms * 1e-3) / 1e12;
        printf("FFMA imm %.1f | FFMA 3 regs (fixed y,z) %.1f | FFMA 3 changing regs %.1f | FFMA2 3 regs %.1f TFLOP/s  %s\n", a, b, c, e, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
