// Multi-person cross-view association on the device: one CTA per frame.
//
// Replaces, per frame, the `multi_person = true` branch of associate_all
// (Pose2Sim/personAssociation.py:783-801):
//   compute_rays / compute_affinity   :277-408   Plücker rays of every joint of every detection, affinity
//                                                = 1 - d / d_max with d the likelihood-weighted mean
//                                                |reciprocal product| of two detections of different views
//   circular_constraint               :411-428   one person per view
//   matchSVT                          :450-509   <= 20 ADMM steps: singular-value shrinkage of X + Y/mu,
//                                                projection on [0,1] / zero same-view blocks / unit
//                                                diagonal / symmetry, dual update, mu doubling/halving
//   person_index_per_cam (first half) :526-533   per row and view the arg-max detection, -1 if none > 0
// The remaining, discrete half of person_index_per_cam (unique rows, ordering by multiplicity, duplicate
// and min-cameras filters) stays on the host: it is a few integer rows per frame.
//
// The SVD of the symmetric N x N matrix (N = detections in the frame, <= 64) is a one-sided Jacobi
// (Hestenes) in shared memory, FP64: round-robin pairing gives N/2 disjoint column pairs per round, one
// half-warp per pair (16 consecutive rows per access), three dot products by xor-shuffles, the rotation
// applied to A and V; every ADMM step after the first warm-starts from the previous step's V.
// U S V^T with S shrunk by tau is then sum_k max(s_k - tau, 0)/s_k a_k v_k^T — only the few columns above
// tau (about one per person) contribute.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "p2s_math.cuh"
#include "p2s_internal.h"

namespace p2s {

struct RayCam { double iK[9], Rt[9], T[3], ctr[3]; };
struct RayCams { RayCam cam[P2S_MAX_CAMS]; };

struct MpArgs {
    const float *obs;             // [n_frames][n_cams][max_persons][3 * n_joints]
    const int32_t *count;         // [n_frames][n_cams] detections per camera
    long long n_frames;
    int n_cams, max_persons, n_joints, n_max;      // n_max: largest sum of detections over the frames
    double d_max, min_affinity;
    int max_iter;
    double w_rank, tol, w_sparse;
    int8_t *out_rows;             // [n_frames][n_max][n_cams]: arg-max detection per view or -1
    double *out_affinity;         // [n_frames][n_max][n_max] final (thresholded) affinity, or null
    int32_t *out_iters;           // [n_frames] ADMM iterations used, or null
    unsigned int *tile_counter;
};

constexpr int kMpMaxWarps = 16;      // the widest team: 512 threads = 32 half-warps = the column pairs of 64 detections

template <int kMpWarps>
__device__ __forceinline__ double block_sum(double v, double *red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(P2S_FULL, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kMpWarps; ++w) s += red[w];                 // same order in every thread
    return s;
}

// Plücker coordinates of the camera->keypoint ray (personAssociation.py:301-315): unit direction, moment
// about the origin; returns false (zero weight) when anything is NaN.
__device__ __forceinline__ bool joint_ray(const RayCam &c, float fx, float fy, float fl, double *l, double *m) {
    const double x = (double)fx, y = (double)fy;
    double q[3], w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) q[i] = c.iK[3 * i] * x + c.iK[3 * i + 1] * y + c.iK[3 * i + 2] - c.T[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = c.Rt[3 * i] * q[0] + c.Rt[3 * i + 1] * q[1] + c.Rt[3 * i + 2] * q[2] - c.ctr[i];
    const double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) l[i] = w[i] / n;
    m[0] = c.ctr[1] * l[2] - c.ctr[2] * l[1];
    m[1] = c.ctr[2] * l[0] - c.ctr[0] * l[2];
    m[2] = c.ctr[0] * l[1] - c.ctr[1] * l[0];
    const double s = l[0] + l[1] + l[2] + m[0] + m[1] + m[2] + (double)fl;
    return s == s;                                                     // any NaN -> the joint carries no weight
}

// Shared-memory layout: [X packed][W packed][Y][sig][red][views, cum, flags][schedule] then the union region
// {A, V, Qp | observations}.  Byte offset of the union region (16-byte aligned):
__host__ __device__ __forceinline__ size_t mp_union_offset(int n_max) {
    const size_t LD = (size_t)(n_max | 1), tri = ((size_t)n_max * (n_max + 1)) >> 1;
    size_t b = (2 * tri + (size_t)n_max * LD + (size_t)n_max + kMpMaxWarps) * sizeof(double) +
               ((size_t)n_max + P2S_MAX_CAMS + 1 + 2) * sizeof(int) + (size_t)(n_max | 1) * 32 * sizeof(unsigned short);
    return (b + 15) & ~(size_t)15;
}

// X (the iterate) and W (w_sparse - affinity) are symmetric BIT FOR BIT — both triangles are written with the same
// value — so they are stored as packed upper triangles; the frame's observations are only read while the affinity is
// built and share their region with A, V and Qp, which are first written after it.  That brings a frame of 48
// detections from 131 KB to 98 KB of shared memory: TWO frames are resident per SM (RESIDENT = 2, 64 registers),
// and one frame's round barriers and rotation chains are covered by the other's work.
__device__ __forceinline__ int tri_index(int i, int j, int N2) {       // N2 = 2 N - 1; (i, j) and (j, i) -> the same slot
    const int a = min(i, j), b = max(i, j);
    return ((a * (N2 - a)) >> 1) + b;                                   // a N - a (a - 1) / 2 + (b - a)
}

// kMpThreads: one half-warp per column pair of a Jacobi round, so 128 / 256 / 512 threads serve frames of up to 16 / 32 / 64
// detections; the narrower teams leave room for 8 / 4 resident frames per SM instead of 2.
template <int kMpThreads, int RESIDENT>
__global__ void __launch_bounds__(kMpThreads, RESIDENT) mp_associate_kernel(const RayCams cams, const MpArgs a) {
    constexpr int kMpWarps = kMpThreads / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int C = a.n_cams, NP = a.max_persons, J = a.n_joints, NM = a.n_max;
    const int LD = NM | 1;                                              // odd leading dimension: fewer bank conflicts
    const int TRI = (NM * (NM + 1)) >> 1, N2 = 2 * NM - 1;
    double *X = reinterpret_cast<double *>(smem_raw);                  // packed upper triangle, tri_index
    double *W = X + TRI;                                               // packed upper triangle
    double *Y = W + TRI;
    double *sig = Y + (size_t)NM * LD;                                 // NM shrink factors
    double *red = sig + NM;                                            // one per warp
    int *s_view = reinterpret_cast<int *>(red + kMpMaxWarps);          // NM: view of each detection
    int *s_cum = s_view + NM;                                          // C + 1
    int *s_flag = s_cum + P2S_MAX_CAMS + 1;                            // [0] frame, [1] rotations in the sweep
    unsigned short *s_sched = reinterpret_cast<unsigned short *>(s_flag + 2);   // [rounds][32] pair schedule: p | q << 8, 0xffff = idle
    double *A = reinterpret_cast<double *>(smem_raw + mp_union_offset(NM));     // column-major: A[col * LD + row]
    double *V = A + (size_t)NM * LD;
    double *Qp = V + (size_t)NM * LD;
    float *sobs = reinterpret_cast<float *>(A);                        // [N][3 J], dead before A / V / Qp are born

    for (;;) {
        __syncthreads();
        if (tid == 0) s_flag[0] = (int)atomicAdd(a.tile_counter, 1u);
        __syncthreads();
        const long long f = s_flag[0];
        if (f >= a.n_frames) break;

        // ---- detections of the frame -----------------------------------------------------------------------
        if (tid == 0) {
            int acc = 0;
            for (int c = 0; c < C; ++c) {
                s_cum[c] = acc;
                int n = a.count[f * C + c];
                n = max(0, min(n, NP));
                n = min(n, NM - acc);                                   // never past the shared-memory regions sized for NM
                                                                        // detections (the host entry rejects such frames)
                for (int p = 0; p < n; ++p) s_view[acc + p] = c;
                acc += n;
            }
            s_cum[C] = acc;
        }
        __syncthreads();
        const int N = s_cum[C];
        for (int i = tid; i < N * 3 * J; i += kMpThreads) {
            const int n = i / (3 * J), r = i - n * 3 * J;
            const int c = s_view[n], p = n - s_cum[c];
            sobs[i] = a.obs[((f * C + c) * NP + p) * (size_t)(3 * J) + r];
        }
        __syncthreads();

        // ---- affinity between detections of different views (:347-408), times the view constraint -----------
        for (int e = tid; e < N * N; e += kMpThreads) {
            const int i = e / N, j = e - i * N;
            if (i > j) continue;
            double aff = 0.0;                                           // same view (and the diagonal): distance 2 d_max -> 0
            const int ci = s_view[i], cj = s_view[j];
            if (ci != cj) {
                double num = 0.0, den = 0.0;
                const float *oi = sobs + (size_t)i * 3 * J, *oj = sobs + (size_t)j * 3 * J;
                for (int k = 0; k < J; ++k) {
                    double li[3], mi[3], lj[3], mj[3];
                    const bool vi = joint_ray(cams.cam[ci], oi[3 * k], oi[3 * k + 1], oi[3 * k + 2], li, mi);
                    const bool vj = joint_ray(cams.cam[cj], oj[3 * k], oj[3 * k + 1], oj[3 * k + 2], lj, mj);
                    if (vi && vj) {
                        const double prod = (li[0] * mj[0] + li[1] * mj[1] + li[2] * mj[2]) + (lj[0] * mi[0] + lj[1] * mi[1] + lj[2] * mi[2]);
                        const double w = sqrt((double)oi[3 * k + 2] * (double)oj[3 * k + 2]);
                        num += fabs(prod) * w;
                        den += w;
                    }
                }
                double d = num / (1e-5 + den);
                if (d > a.d_max) d = a.d_max;
                aff = 1.0 - d / a.d_max;
            }
            X[tri_index(i, j, N2)] = aff;
        }
        __syncthreads();
        // matchSVT start (:468-476): zero diagonal, Y = 0, W = w_sparse - X
        for (int e = tid; e < N * N; e += kMpThreads) {
            const int i = e / N, j = e - i * N;
            Y[i * LD + j] = 0.0;
            if (i > j) continue;
            const int t = tri_index(i, j, N2);
            if (i == j) X[t] = 0.0;
            W[t] = a.w_sparse - ((i == j) ? 0.0 : X[t]);
        }
        __syncthreads();

        // Round-robin (circle method) schedule of the one-sided Jacobi, once per frame: round r pairs column
        // n_even-1 with r and (r + k) with (r - k) modulo n_even-1 — the rounds then cost one shared-memory read
        // instead of two integer divisions per thread.
        {
            const int n_even = (N + 1) & ~1, m = n_even - 1;
            for (int e = tid; e < m * 32; e += kMpThreads) {
                const int r = e >> 5, k = e & 31;
                unsigned short v = 0xffffu;
                if (k < n_even / 2) {
                    int p, q;
                    if (k == 0) { p = n_even - 1; q = r; }
                    else { p = (r + k) % m; q = (r - k + m) % m; }
                    if (p > q) { const int t = p; p = q; q = t; }
                    if (q < N) v = (unsigned short)(p | (q << 8));      // q >= N: the padding column of an odd N
                }
                s_sched[e] = v;
            }
        }
        __syncthreads();
        double mu = 64.0;
        int iters = 0;
        for (int it = 0; it < a.max_iter && N > 0; ++it) {
            iters = it + 1;
            // ---- B = X + Y / mu (symmetric), held in Qp until the shrinkage overwrites it ------------------------
            // First step: A = B, V = I.  Later steps WARM-START the Jacobi SVD from the previous step's V:
            // A = B V_prev has nearly orthogonal columns already (B moves little between ADMM steps), so one or
            // two sweeps finish where a cold start needs eight to ten.  B = (B V) V^T, so the SVD is the same.
            for (int e = tid; e < N * N; e += kMpThreads) {
                const int i = e / N, j = e - i * N;
                const double b = X[tri_index(i, j, N2)] + Y[i * LD + j] * 1.0 / mu;
                Qp[i * LD + j] = b;
                if (it == 0) {
                    A[j * LD + i] = b;
                    V[j * LD + i] = (i == j) ? 1.0 : 0.0;
                }
            }
            __syncthreads();
            if (it > 0) {
                for (int e = tid; e < N * N; e += kMpThreads) {
                    const int k = e / N, i = e - k * N;                 // A[:, k] = B V[:, k]; lanes walk i: Qp row reads are
                    const double *vk = V + (size_t)k * LD;             // conflict-free (LD odd), V[k][j] is a broadcast
                    double acc = 0.0;
                    for (int j = 0; j < N; ++j) acc = fma(Qp[i * LD + j], vk[j], acc);
                    A[(size_t)k * LD + i] = acc;
                }
                __syncthreads();
            }
            const double tau = a.w_rank / mu;
            // ---- one-sided Jacobi: orthogonalise the columns of A, accumulate V ----------------------------------
            // Round-robin pairing: n_even / 2 disjoint column pairs per round (<= 32), ONE HALF-WARP per pair.  Its 16
            // lanes read 16 consecutive rows of a column at a time (one conflict-free 128-byte wavefront), keep the
            // pair's A and V entries in registers between the dot products and the rotation, and meet by four
            // xor-shuffles.  The rotation's scalar chain uses the MUFU-seeded reciprocal / square root of p2s_math.cuh:
            // c^2 + s^2 = 1 to a few ulp is what orthogonality needs, the angle itself only steers convergence.
            // A sweep whose largest rotation was below 1e-7 ends the SVD: Jacobi converges quadratically, so the
            // columns are orthogonal to ~1e-14 after it and a confirming sweep would only cost time.
            const int n_even = (N + 1) & ~1;
            const int pair = tid >> 4, l16 = tid & 15;
            for (int sweep = 0; sweep < 40; ++sweep) {
                if (tid == 0) s_flag[1] = 0;
                // squared column norms, refreshed once per sweep and carried through the rotations in between
                // (|a_p'|^2 = |a_p|^2 - t g, |a_q'|^2 = |a_q|^2 + t g): the rounds only need the cross product g
                for (int k0 = 0; k0 < N; k0 += kMpThreads / 16) {       // trip count uniform over the CTA: full-mask shuffles inside
                    const int k = k0 + pair;
                    double n2 = 0.0;
                    if (k < N) {
                        const double *ak = A + (size_t)k * LD;
                        for (int i = l16; i < N; i += 16) n2 = fma(ak[i], ak[i], n2);
                    }
#pragma unroll
                    for (int off = 8; off > 0; off >>= 1) n2 += __shfl_xor_sync(P2S_FULL, n2, off);
                    if (k < N && l16 == 0) sig[k] = n2;
                }
                __syncthreads();
                for (int r = 0; r < n_even - 1; ++r) {
                    const unsigned int pq = s_sched[r * 32 + pair];
                    const int p = (pq == 0xffffu) ? -1 : (int)(pq & 0xffu), q = (int)(pq >> 8);
                    double al = 0.0, be = 0.0, ga = 0.0;
                    double ru[4], rv[4];
                    if (p >= 0) {
                        const double *ap = A + (size_t)p * LD, *aq = A + (size_t)q * LD;
                        al = sig[p]; be = sig[q];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int i = l16 + 16 * t;
                            const bool in = i < N;
                            ru[t] = in ? ap[i] : 0.0;
                            rv[t] = in ? aq[i] : 0.0;
                            ga = fma(ru[t], rv[t], ga);
                        }
                    }
#pragma unroll
                    for (int off = 8; off > 0; off >>= 1) ga += __shfl_xor_sync(P2S_FULL, ga, off);
                    // A pair whose columns together carry less than tau^2 spans only singular values below the
                    // shrinkage threshold: whatever basis it ends in contributes exactly zero, so it is left alone.
                    const double ab = al * be, g2 = ga * ga;
                    if (p >= 0 && g2 > 1e-30 * ab && ga != 0.0 && al + be > 0.25 * tau * tau) {
                        // Jacobi angle with |theta| <= pi/4 from two reciprocal square roots:
                        //   d = be - al, h = hypot(d, 2 g), cos 2theta = |d| / h, sin 2theta = sign(d) 2 g / h,
                        //   c = sqrt((1 + cos 2theta) / 2), s = sin 2theta / (2 c)          (c^2 + s^2 = 1 identically)
                        const double d = be - al;
                        const double h2 = fma(d, d, 4.0 * g2);
                        double rh;
                        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rh) : "d"(h2));
                        { const double e = fma(-h2 * rh, rh, 1.0); rh = fma(rh * e, fma(0.375, e, 0.5), rh); }       // 1 / h
                        const double c2 = fma(0.5 * fabs(d), rh, 0.5);                                           // c^2 in [0.5, 1]
                        double rc;
                        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rc) : "d"(c2));
                        { const double e = fma(-c2 * rc, rc, 1.0); rc = fma(rc * e, fma(0.375, e, 0.5), rc); }       // 1 / c
                        const double cs = c2 * rc;
                        const double sn = (d >= 0.0 ? ga : -ga) * rh * rc;
                        double *ap = A + (size_t)p * LD, *aq = A + (size_t)q * LD;
                        double *vp = V + (size_t)p * LD, *vq = V + (size_t)q * LD;
#pragma unroll
                        for (int tt = 0; tt < 4; ++tt) {
                            const int i = l16 + 16 * tt;
                            if (i < N) {
                                ap[i] = fma(cs, ru[tt], -sn * rv[tt]); aq[i] = fma(sn, ru[tt], cs * rv[tt]);
                                const double x = vp[i], y = vq[i];
                                vp[i] = fma(cs, x, -sn * y); vq[i] = fma(sn, x, cs * y);
                            }
                        }
                        if (l16 == 0) {
                            const double tg = sn * rc * ga;                 // t g, t = s / c
                            sig[p] = fmax(al - tg, 0.0); sig[q] = be + tg;
                            if (g2 > 1e-14 * ab) s_flag[1] = 1;             // a rotation above 1e-7: sweep again
                        }
                    }
                    __syncthreads();
                }
                if (s_flag[1] == 0) break;
                __syncthreads();
            }
            // ---- shrink: factor_k = max(s_k - tau, 0) / s_k, s_k = |a_k| --------------------------------------------
            for (int k = tid; k < N; k += kMpThreads) {
                double s2 = 0.0;
                for (int i = 0; i < N; ++i) s2 = fma(A[(size_t)k * LD + i], A[(size_t)k * LD + i], s2);
                const double s = sqrt(s2);
                sig[k] = (s > tau) ? (s - tau) / s : 0.0;
            }
            __syncthreads();
            // ---- Qp = U max(S - tau, 0) V^T, then the projection / dual update / residuals -------------------------
            for (int e = tid; e < N * N; e += kMpThreads) {
                const int i = e / N, j = e - i * N;
                double acc = 0.0;
                for (int k = 0; k < N; ++k) {
                    const double fk = sig[k];
                    if (fk != 0.0) acc = fma(fk * A[(size_t)k * LD + i], V[(size_t)k * LD + j], acc);
                }
                Qp[i * LD + j] = acc;
            }
            __syncthreads();
            double pr = 0.0, dr = 0.0;
            for (int e = tid; e < N * N; e += kMpThreads) {
                const int i = e / N, j = e - i * N;
                if (i > j) continue;
                const bool same = s_view[i] == s_view[j];
                const int t = tri_index(i, j, N2);
                const double wij = W[t];
                double xij = Qp[i * LD + j] - (wij + Y[i * LD + j]) / mu;
                double xji = Qp[j * LD + i] - (wij + Y[j * LD + i]) / mu;
                if (same) { xij = 0.0; xji = 0.0; }
                if (i == j) { xij = 1.0; xji = 1.0; }
                if (xij < 0.0) xij = 0.0; if (xij > 1.0) xij = 1.0;
                if (xji < 0.0) xji = 0.0; if (xji > 1.0) xji = 1.0;
                const double xs = (xij + xji) / 2.0;
                const double oij = X[t], oji = oij;
                const double eij = xs - Qp[i * LD + j], eji = xs - Qp[j * LD + i];
                Y[i * LD + j] += mu * eij;
                pr += eij * eij;
                dr += (xs - oij) * (xs - oij);
                if (i != j) {
                    Y[j * LD + i] += mu * eji;
                    pr += eji * eji;
                    dr += (xs - oji) * (xs - oji);
                }
                X[t] = xs;
            }
            const double pres = sqrt(block_sum<kMpWarps>(pr, red)) / N;
            const double dres = mu * sqrt(block_sum<kMpWarps>(dr, red)) / N;
            __syncthreads();
            if (pres < a.tol && dres < a.tol) break;
            if (pres > 10.0 * dres) mu = 2.0 * mu;
            else if (dres > 10.0 * pres) mu = mu / 2.0;
        }

        // ---- min_affinity threshold (:800) and the per-row / per-view arg-max (:526-533) ---------------------------
        for (int e = tid; e < N * N; e += kMpThreads) {
            const int i = e / N, j = e - i * N;
            if (i > j) continue;
            const int t = tri_index(i, j, N2);
            double x = X[t];
            if (x < a.min_affinity) { x = 0.0; X[t] = 0.0; }
            if (a.out_affinity) { a.out_affinity[(f * NM + i) * NM + j] = x; a.out_affinity[(f * NM + j) * NM + i] = x; }
        }
        __syncthreads();
        for (int e = tid; e < N * C; e += kMpThreads) {
            const int r = e / C, v = e - r * C;
            int best = -1;
            double bv = 0.0;
            for (int j = s_cum[v]; j < s_cum[v + 1]; ++j) {
                const double x = X[tri_index(r, j, N2)];
                if (x > bv) { bv = x; best = j - s_cum[v]; }                // first maximum, and only if > 0
            }
            a.out_rows[(f * NM + r) * C + v] = (int8_t)best;
        }
        if (tid == 0 && a.out_iters) a.out_iters[f] = iters;
        }
}

size_t mp_smem_bytes(int n_max, int n_joints) {
    const size_t LD = (size_t)(n_max | 1);
    const size_t mats = 3 * (size_t)n_max * LD * sizeof(double);                 // A, V, Qp
    const size_t obs = (size_t)n_max * 3 * n_joints * sizeof(float);             // alive only before them
    return mp_union_offset(n_max) + (mats > obs ? mats : obs);
}

template <int kMpThreads, int RESIDENT>
static cudaError_t launch_mp_variant(const RayCams &cams, const MpArgs &a, size_t smem, int sm_count, cudaStream_t stream, int *grid_out) {
    cudaError_t e = cudaFuncSetAttribute(mp_associate_kernel<kMpThreads, RESIDENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mp_associate_kernel<kMpThreads, RESIDENT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mp_associate_kernel<kMpThreads, RESIDENT>, kMpThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > RESIDENT) per_sm = RESIDENT;
    long long grid = (long long)sm_count * per_sm;
    if (grid > a.n_frames) grid = a.n_frames;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    mp_associate_kernel<kMpThreads, RESIDENT><<<(unsigned)grid, kMpThreads, smem, stream>>>(cams, a);
    return cudaGetLastError();
}

cudaError_t launch_mp_associate(const MpLaunch &L, int *grid_out) {
    RayCams cams;
    std::memset(&cams, 0, sizeof cams);
    for (int c = 0; c < L.n_cams; ++c) {
        const p2s_camera_model &m = L.cams[c];
        RayCam &o = cams.cam[c];
        // inverse of K (common.py:282 np.linalg.inv) by the adjugate; K is upper triangular in practice
        const double *K = m.K;
        const double det = K[0] * (K[4] * K[8] - K[5] * K[7]) - K[1] * (K[3] * K[8] - K[5] * K[6]) + K[2] * (K[3] * K[7] - K[4] * K[6]);
        o.iK[0] = (K[4] * K[8] - K[5] * K[7]) / det; o.iK[1] = (K[2] * K[7] - K[1] * K[8]) / det; o.iK[2] = (K[1] * K[5] - K[2] * K[4]) / det;
        o.iK[3] = (K[5] * K[6] - K[3] * K[8]) / det; o.iK[4] = (K[0] * K[8] - K[2] * K[6]) / det; o.iK[5] = (K[2] * K[3] - K[0] * K[5]) / det;
        o.iK[6] = (K[3] * K[7] - K[4] * K[6]) / det; o.iK[7] = (K[1] * K[6] - K[0] * K[7]) / det; o.iK[8] = (K[0] * K[4] - K[1] * K[3]) / det;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) o.Rt[3 * i + j] = m.R[3 * j + i];
        for (int i = 0; i < 3; ++i) o.T[i] = m.T[i];
        for (int i = 0; i < 3; ++i) o.ctr[i] = -(o.Rt[3 * i] * m.T[0] + o.Rt[3 * i + 1] * m.T[1] + o.Rt[3 * i + 2] * m.T[2]);
    }
    MpArgs a;
    a.obs = L.obs; a.count = L.count; a.n_frames = L.n_frames; a.n_cams = L.n_cams; a.max_persons = L.max_persons;
    a.n_joints = L.n_joints; a.n_max = L.n_max; a.d_max = L.d_max; a.min_affinity = L.min_affinity;
    a.max_iter = 20; a.w_rank = 50.0; a.tol = 1e-4; a.w_sparse = 0.1;            // matchSVT's call-site constants (:799)
    a.out_rows = L.out_rows; a.out_affinity = L.out_affinity; a.out_iters = L.out_iters; a.tile_counter = L.tile_counter;
    const size_t smem = mp_smem_bytes(L.n_max, L.n_joints);
    // Team width from the largest frame (one half-warp per column pair), resident frames per SM from the register file at
    // 64 registers per thread (8 / 4 / 2); a 512-thread team whose frame does not fit twice into the SM's shared memory
    // (228 KB less 1 KB per CTA) runs alone with the full register file.
    // (P2S_MP_ONE_RESIDENT in the environment: A/B switch of tests/perf/mp_bench.py — the round-2 configuration)
    static const bool force_one = std::getenv("P2S_MP_ONE_RESIDENT") != nullptr;
    if (force_one) return launch_mp_variant<512, 1>(cams, a, smem, L.sm_count, L.stream, grid_out);
    if (L.n_max <= 16) return launch_mp_variant<128, 8>(cams, a, smem, L.sm_count, L.stream, grid_out);
    if (L.n_max <= 32) return launch_mp_variant<256, 4>(cams, a, smem, L.sm_count, L.stream, grid_out);
    const bool two = 2 * (smem + 1024) <= L.smem_per_sm;
    return two ? launch_mp_variant<512, 2>(cams, a, smem, L.sm_count, L.stream, grid_out)
               : launch_mp_variant<512, 1>(cams, a, smem, L.sm_count, L.stream, grid_out);
}

}  // namespace p2s
