// peak_bench.cu — measures the vector-pipe peaks MEASURED_PEAKS.json does not hold: FP32 FFMA, FP64 DFMA,
// MUFU (ex2) and F32<->F64 conversion throughput, with dependent-chain ILP high enough to saturate issue.
// Used as the roofline denominators for the MPPI (FP32) and UKF (FP64) kernels.  Output: one JSON line.
#include <cuda_runtime.h>
#include <stdio.h>

template <typename T, int ILP>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void mufu_kernel(float* out, int iters) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = 0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void cvt_kernel(float* out, int iters) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = 1.0f + 0.001f * (threadIdx.x + i);
    double d = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            double t;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(acc[i]));
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(acc[i]) : "d"(t));
        }
    }
    float s = (float)d;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const int threads = 256, blocks = sms * 8;
    const int iters = 4096;
    constexpr int ILP = 8;
    void* out;
    cudaMalloc(&out, (size_t)threads * blocks * sizeof(double));
    const double n = (double)threads * blocks * iters * ILP;
    double ms32 = time_ms([&] { fma_kernel<float, ILP><<<blocks, threads>>>((float*)out, iters, 1.0001f, 0.5f); }, 10);
    double ms64 = time_ms([&] { fma_kernel<double, ILP><<<blocks, threads>>>((double*)out, iters, 1.0001, 0.5); }, 10);
    double msmu = time_ms([&] { mufu_kernel<ILP><<<blocks, threads>>>((float*)out, iters); }, 10);
    double mscv = time_ms([&] { cvt_kernel<ILP><<<blocks, threads>>>((float*)out, iters); }, 10);
    cudaError_t e = cudaDeviceSynchronize();
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"fp32_tflops\": %.2f, \"fp64_tflops\": %.2f, "
           "\"mufu_gops\": %.1f, \"cvt_pair_gops\": %.1f, \"err\": \"%s\"}\n",
           prop.name, sms, prop.clockRate, 2.0 * n / (ms32 * 1e-3) / 1e12, 2.0 * n / (ms64 * 1e-3) / 1e12,
           n / (msmu * 1e-3) / 1e9, n / (mscv * 1e-3) / 1e9, cudaGetErrorString(e));
    return 0;
}
