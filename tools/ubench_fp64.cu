// FP64 pipe micro-benchmarks for the kernel design decisions in DESIGN.md (B200, sm_100a):
//   dep<ILP>      : ILP independent DFMA chains per thread  -> dependent-issue latency and peak rate
//   mix<ILP, K>   : ILP DFMA chains + K independent integer (LOP3/IADD) ops per DFMA -> does integer work
//                   ride for free in the issue slots the 2-cycle FP64 dispatch leaves open?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int K>
__global__ void __launch_bounds__(128) kern(double *out, int iters, double m, double b, unsigned salt) {
    double a[ILP];
    unsigned x[8];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = 1.0 + threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 7u + i + salt;
    double mm[ILP], bb[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { mm[i] = m + 1e-9 * i; bb[i] = b + 1e-12 * i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                a[i] = fma(a[i], mm[i], bb[i]);
#pragma unroll
                for (int k = 0; k < K; ++k) x[(i * K + k) & 7] = (x[(i * K + k) & 7] ^ (x[(i * K + k + 1) & 7] >> 3)) + 0x9e3779b9u;
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    unsigned xs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) xs ^= x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s + (double)xs;
}

template <int ILP, int K>
void run(const char *name, int sms, int warps_per_smsp, double *out) {
    const int iters = 2048;
    const int blocks = sms * warps_per_smsp;      // 128 threads = 4 warps = 1 warp per SMSP per block
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<ILP, K><<<blocks, 128>>>(out, 16, 0.999999, 1e-9, 1u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<ILP, K><<<blocks, 128>>>(out, iters, 0.999999, 1e-9, 1u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double dfma_per_warp = (double)iters * 8 * ILP;
    const double cycles = best * 1e-3 * khz * 1e3;
    // per SMSP: warps_per_smsp warps each issuing dfma_per_warp DFMA
    const double cyc_per_dfma_smsp = cycles / (dfma_per_warp * warps_per_smsp);
    const double tflops = 2.0 * 32 * dfma_per_warp * 4 * warps_per_smsp * sms / (best * 1e-3) / 1e12;
    printf("{\"bench\": \"%s\", \"ilp\": %d, \"int_per_dfma\": %d, \"warps_per_smsp\": %d, \"ms\": %.4f, \"cycles_per_dfma_per_smsp\": %.3f, "
           "\"cycles_per_dfma_per_warp\": %.3f, \"tflops\": %.2f}\n", name, ILP, K, warps_per_smsp, best, cyc_per_dfma_smsp,
           cycles / dfma_per_warp, tflops);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 16 * 128 * sizeof(double));
    const int sms = p.multiProcessorCount;
    for (int w : {1, 2, 4, 5, 8}) run<1, 0>("dep", sms, w, out);
    for (int w : {1, 2, 4, 8}) run<2, 0>("dep", sms, w, out);
    for (int w : {1, 4, 8}) run<4, 0>("dep", sms, w, out);
    for (int w : {1, 4, 8}) run<8, 0>("dep", sms, w, out);
    for (int w : {4, 8}) run<4, 1>("mix", sms, w, out);
    for (int w : {4, 8}) run<4, 2>("mix", sms, w, out);
    for (int w : {4, 8}) run<4, 3>("mix", sms, w, out);
    for (int w : {4}) run<1, 1>("mix", sms, w, out);
    for (int w : {4}) run<1, 2>("mix", sms, w, out);
    return 0;
}
