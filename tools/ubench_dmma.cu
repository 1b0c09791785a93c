// FP64 tensor-core (DMMA, mma.sync.aligned.m8n8k4.f64) micro-benchmarks on B200 (sm_100a):
//   dmma<ILP, 0, 0>   : ILP independent accumulator chains per warp -> latency and peak rate of the DMMA
//   dmma<ILP, F, 0>   : F independent DFMA per DMMA                 -> do DMMA and DFMA share the FP64 pipe?
//   dmma<ILP, 0, K>   : K integer ops per DMMA                      -> does a DMMA hold the issue port like a DFMA does?
// One DMMA = 8 x 8 x 4 = 256 FMA = 8 warp-wide DFMA worth of arithmetic.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_dmma tools/ubench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP, int F, int K>
__global__ void __launch_bounds__(128) kern(double *out, int iters, double m, double b, unsigned salt) {
    double c0[ILP], c1[ILP], f[F > 0 ? F * ILP : 1];
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = 1e-3 * threadIdx.x + i; c1[i] = 2e-3 * threadIdx.x + i; }
#pragma unroll
    for (int i = 0; i < (F > 0 ? F * ILP : 1); ++i) f[i] = 1.0 + threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 7u + i + salt;
    const double av = m * 1e-3 * ((threadIdx.x & 3) + 1), bv = b + 1e-3 * (threadIdx.x >> 2);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                dmma(c0[i], c1[i], av, bv);
#pragma unroll
                for (int j = 0; j < F; ++j) f[i * F + j] = fma(f[i * F + j], m, b);
#pragma unroll
                for (int k = 0; k < K; ++k) x[(i * K + k) & 7] = (x[(i * K + k) & 7] ^ (x[(i * K + k + 1) & 7] >> 3)) + 0x9e3779b9u;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < (F > 0 ? F * ILP : 1); ++i) s += f[i];
    unsigned xs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) xs ^= x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s + (double)xs;
}

template <int ILP, int F, int K>
void run(int sms, int warps_per_smsp, double *out) {
    const int iters = 1024;
    const int blocks = sms * warps_per_smsp;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<ILP, F, K><<<blocks, 128>>>(out, 16, 0.999999, 1e-9, 1u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<ILP, F, K><<<blocks, 128>>>(out, iters, 0.999999, 1e-9, 1u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double dmma_per_warp = (double)iters * 8 * ILP;
    const double cycles = best * 1e-3 * khz * 1e3;
    const double cyc_per_dmma_smsp = cycles / (dmma_per_warp * warps_per_smsp);
    const double tflops_mma = 2.0 * 256 * dmma_per_warp * 4 * warps_per_smsp * sms / (best * 1e-3) / 1e12;
    const double tflops_fma = 2.0 * 32 * F * dmma_per_warp * 4 * warps_per_smsp * sms / (best * 1e-3) / 1e12;
    printf("{\"bench\": \"dmma\", \"ilp\": %d, \"dfma_per_dmma\": %d, \"int_per_dmma\": %d, \"warps_per_smsp\": %d, \"ms\": %.4f, "
           "\"cycles_per_dmma_per_smsp\": %.3f, \"cycles_per_dmma_per_warp\": %.3f, \"tflops_dmma\": %.2f, \"tflops_dfma\": %.2f}\n",
           ILP, F, K, warps_per_smsp, best, cyc_per_dmma_smsp, cycles / dmma_per_warp, tflops_mma, tflops_fma);
    fflush(stdout);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 16 * 128 * sizeof(double));
    const int sms = p.multiProcessorCount;
    for (int w : {1, 2, 4, 8}) run<1, 0, 0>(sms, w, out);
    for (int w : {1, 4, 8}) run<2, 0, 0>(sms, w, out);
    for (int w : {1, 4, 8}) run<4, 0, 0>(sms, w, out);
    // DMMA + DFMA: shared pipe => time = sum of the two; separate => max
    for (int w : {4, 8}) run<2, 2, 0>(sms, w, out);
    for (int w : {4, 8}) run<2, 4, 0>(sms, w, out);
    for (int w : {4, 8}) run<2, 8, 0>(sms, w, out);
    for (int w : {4, 8}) run<2, 16, 0>(sms, w, out);
    // DMMA + integer work: issue-port occupancy of a DMMA
    for (int w : {4, 8}) run<2, 0, 4>(sms, w, out);
    for (int w : {4, 8}) run<2, 0, 8>(sms, w, out);
    for (int w : {4, 8}) run<2, 0, 16>(sms, w, out);
    // all three: 1 DMMA + 4 DFMA + 8 int
    for (int w : {4, 8}) run<2, 4, 8>(sms, w, out);
    return 0;
}
