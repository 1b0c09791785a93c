// Synthetic multi-view keypoint streams generated ON THE DEVICE: a pure function of (seed, unit, camera).
//
// BASELINE configs[4] (10 M frames x up to 32 cameras) is 100 GB of observations — not something
// to make on the host and push over PCIe, so every rank generates its own shard.  The NumPy twin
// (pose2sim_b200/synth_philox.py) evaluates the same Philox4x32-10 counters and the same IEEE operations in the same
// order (explicit round-to-nearest adds / multiplies / divides, no fused multiply-add, no transcendental function), so
// any subsample can be regenerated bit for bit for the CPU oracle (tests/test_gpu_synth.py).
// This is workload generation for the benchmarks and tests; it produces INPUTS only.
#include "p2s_internal.h"
#include "p2s_math.cuh"

namespace p2s {

struct SynthArgs {
    long long unit0, n_units;
    int n_cams, n_keypoints;
    unsigned int seed;
    double sigma, p_out, p_low;
    const double *kp_off;          // device [n_keypoints][3]
    const double *circle;          // device [600][2]   2 cos / 2 sin of the walk circle
    const double *dirs;            // device [256][2]   outlier directions
    float *x, *y, *lik;            // device [n_units][n_cams]
    double *truth;                 // device [n_units][3] or null
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&r)[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

__device__ __forceinline__ double u24(uint32_t r) { return __dmul_rn((double)(r >> 8), 1.0 / 16777216.0); }

// (sum of four 16-bit uniforms, centred) * sqrt(3) / 65536: unit variance, exact integer sum
__device__ __forceinline__ double ih4(uint32_t a, uint32_t b) {
    const int s = (int)(a & 0xFFFFu) + (int)(a >> 16) + (int)(b & 0xFFFFu) + (int)(b >> 16) - 131070;
    return __dmul_rn((double)s, 1.7320508075688772 / 65536.0);
}

__global__ void __launch_bounds__(256) synth_kernel(const CamParams<P2S_MAX_CAMS> cams, const SynthArgs a) {
    const uint32_t k1 = 0x5032534Du;
    const long long n_el = a.n_units * a.n_cams;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (long long)gridDim.x * blockDim.x) {
        const long long ul = e / a.n_cams;
        const int c = (int)(e - ul * a.n_cams);
        const long long u = a.unit0 + ul;
        const long long f = u / a.n_keypoints;
        const int k = (int)(u - f * a.n_keypoints);
        const uint32_t ulo = (uint32_t)(u & 0xFFFFFFFFLL), uhi = (uint32_t)(u >> 32);
        uint32_t r[4], r2[4];
        // 3D truth of the unit: walk circle + keypoint offset + 2 cm jitter
        philox4x32_10(ulo, uhi, 0xFFFFu, 0u, a.seed, k1, r);
        philox4x32_10(ulo, uhi, 0xFFFFu, 1u, a.seed, k1, r2);
        const int fi = (int)(f % 600);
        const double X = __dadd_rn(__dadd_rn(a.circle[2 * fi], a.kp_off[3 * k]), __dmul_rn(ih4(r[0], r[1]), 0.02));
        const double Y = __dadd_rn(__dadd_rn(a.circle[2 * fi + 1], a.kp_off[3 * k + 1]), __dmul_rn(ih4(r[2], r[3]), 0.02));
        const double Z = __dadd_rn(a.kp_off[3 * k + 2], __dmul_rn(ih4(r2[0], r2[1]), 0.02));
        if (a.truth != nullptr && c == 0) { a.truth[3 * ul] = X; a.truth[3 * ul + 1] = Y; a.truth[3 * ul + 2] = Z; }
        // projection (((P0 X + P1 Y) + P2 Z) + P3), observation noise, outliers, likelihoods
        const double *P = cams.P[c];
        const double hu = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[0], X), __dmul_rn(P[1], Y)), __dmul_rn(P[2], Z)), P[3]);
        const double hv = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[4], X), __dmul_rn(P[5], Y)), __dmul_rn(P[6], Z)), P[7]);
        const double hd = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(P[8], X), __dmul_rn(P[9], Y)), __dmul_rn(P[10], Z)), P[11]);
        uint32_t ra[4], rb[4];
        philox4x32_10(ulo, uhi, (uint32_t)c, 0u, a.seed, k1, ra);
        philox4x32_10(ulo, uhi, (uint32_t)c, 1u, a.seed, k1, rb);
        double x = __dadd_rn(__ddiv_rn(hu, hd), __dmul_rn(ih4(ra[0], ra[1]), a.sigma));
        double y = __dadd_rn(__ddiv_rn(hv, hd), __dmul_rn(ih4(ra[2], ra[3]), a.sigma));
        const bool is_out = u24(rb[0]) < a.p_out;
        const double mag = __dadd_rn(50.0, __dmul_rn(250.0, u24(rb[1])));
        const int di = (int)(rb[2] & 0xFFu);
        if (is_out) {
            x = __dadd_rn(x, __dmul_rn(mag, a.dirs[2 * di]));
            y = __dadd_rn(y, __dmul_rn(mag, a.dirs[2 * di + 1]));
        }
        const double ul_ = u24(rb[3]);
        double lik = is_out ? __dadd_rn(0.3, __dmul_rn(0.4, ul_)) : __dadd_rn(0.5, __dmul_rn(0.5, ul_));
        if (u24(rb[2]) < a.p_low) lik = __dmul_rn(0.3, ul_);
        a.x[e] = __double2float_rn(x);
        a.y[e] = __double2float_rn(y);
        a.lik[e] = __double2float_rn(lik);
    }
}

cudaError_t launch_synth(const double *P, int n_cams, int n_keypoints, unsigned int seed, long long unit0, long long n_units,
                         double sigma, double p_out, double p_low, const double *kp_off, const double *circle,
                         const double *dirs, float *x, float *y, float *lik, double *truth, int sm_count, cudaStream_t stream) {
    CamParams<P2S_MAX_CAMS> cams;
    for (int c = 0; c < P2S_MAX_CAMS; ++c)
        for (int j = 0; j < 12; ++j) cams.P[c][j] = (c < n_cams) ? P[c * 12 + j] : 0.0;
    SynthArgs a;
    a.unit0 = unit0; a.n_units = n_units; a.n_cams = n_cams; a.n_keypoints = n_keypoints; a.seed = seed;
    a.sigma = sigma; a.p_out = p_out; a.p_low = p_low; a.kp_off = kp_off; a.circle = circle; a.dirs = dirs;
    a.x = x; a.y = y; a.lik = lik; a.truth = truth;
    const long long n_el = n_units * n_cams;
    long long grid = (n_el + 255) / 256;
    if (grid > (long long)sm_count * 8) grid = (long long)sm_count * 8;
    if (grid < 1) grid = 1;
    synth_kernel<<<(unsigned)grid, 256, 0, stream>>>(cams, a);
    return cudaGetLastError();
}

}  // namespace p2s
