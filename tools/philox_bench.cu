// philox_bench.cu — developer microbenchmark: dispatch cost of one Philox4x32 round on B200 in several instruction
// forms (the round is 2 32x32->64 multiplies + 2 three-input xors).  4 independent states per thread, W warps/SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --cudart shared -o tools/philox_bench tools/philox_bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define ITERS 512
#define ROUNDS 7
#define NS 4
template <int FORM>
__device__ __forceinline__ void round_(uint32_t (&c)[4], uint32_t k0, uint32_t k1, uint32_t m0, uint32_t m1) {
    uint32_t hi0, lo0, hi1, lo1;
    if (FORM == 0) {  // mul.wide with immediate
        uint64_t p0 = (uint64_t)c[0] * 0xD2511F53u, p1 = (uint64_t)c[2] * 0xCD9E8D57u;
        lo0 = (uint32_t)p0; hi0 = (uint32_t)(p0 >> 32); lo1 = (uint32_t)p1; hi1 = (uint32_t)(p1 >> 32);
    } else if (FORM == 1) {  // mul.wide with the multiplier in a register
        uint64_t p0 = (uint64_t)c[0] * m0, p1 = (uint64_t)c[2] * m1;
        lo0 = (uint32_t)p0; hi0 = (uint32_t)(p0 >> 32); lo1 = (uint32_t)p1; hi1 = (uint32_t)(p1 >> 32);
    } else if (FORM == 2) {  // mul.lo + mul.hi
        lo0 = c[0] * 0xD2511F53u; hi0 = __umulhi(c[0], 0xD2511F53u); lo1 = c[2] * 0xCD9E8D57u; hi1 = __umulhi(c[2], 0xCD9E8D57u);
    } else {  // only the low halves (NOT Philox: lower bound of a multiply-xor round)
        lo0 = c[0] * 0xD2511F53u; hi0 = lo0 >> 7; lo1 = c[2] * 0xCD9E8D57u; hi1 = lo1 >> 9;
    }
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
template <int FORM>
__global__ void k(long long* cyc, uint32_t* out, uint32_t seed0, uint32_t seed1, uint32_t m0, uint32_t m1) {
    uint32_t c[NS][4];
    for (int s = 0; s < NS; ++s) { c[s][0] = threadIdx.x + s * 1000; c[s][1] = blockIdx.x; c[s][2] = s; c[s][3] = 7; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            uint32_t k0 = seed0, k1 = seed1;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) { round_<FORM>(c[s], k0, k1, m0, m1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
            c[s][3] += it;
        }
    }
    const long long t1 = clock64();
    uint32_t x = 0;
    for (int s = 0; s < NS; ++s) x ^= c[s][0] ^ c[s][1] ^ c[s][2] ^ c[s][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    long long* cyc; uint32_t* out; cudaMalloc(&cyc, nsm * 8); cudaMalloc(&out, nsm * 1024 * 4);
    long long* h = new long long[nsm];
    const char* names[] = {"mul.wide imm", "mul.wide reg", "mul.lo + mul.hi", "mul.lo only (not Philox)"};
    void (*fns[])(long long*, uint32_t*, uint32_t, uint32_t, uint32_t, uint32_t) = {k<0>, k<1>, k<2>, k<3>};
    printf("%-28s %8s %8s %8s  cycles per Philox round per warp (per SMSP)\n", "form", "W=1", "W=2", "W=4");
    for (int f = 0; f < 4; ++f) {
        printf("%-28s", names[f]);
        for (int W : {1, 2, 4}) {
            for (int rep = 0; rep < 2; ++rep) fns[f]<<<nsm, 128 * W>>>(cyc, out, 123u, 456u, 0xD2511F53u, 0xCD9E8D57u);
            cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < nsm; ++i) avg += h[i]; avg /= nsm;
            printf(" %8.2f", avg / ((double)ITERS * NS * ROUNDS * W));
        }
        printf("\n");
    }
    return 0;
}
