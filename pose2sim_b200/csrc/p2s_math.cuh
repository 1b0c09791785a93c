// Device-side leaf math of the triangulation hot path (sm_100a, FP64 CUDA-core pipe).
//
// What is computed (reference: Pose2Sim/common.py:327-403, restated in SURVEY.md §8(a) `solve`):
//   M   = sum over valid cameras c, ascending, of r1 r1^T + r2 r2^T,
//         r1 = (P_c[0] - x_c P_c[2]) w_c,  r2 = (P_c[1] - y_c P_c[2]) w_c          (common.py:344-345)
//   Q   = v[0:3]/v[3], v = eigenvector of M for its smallest eigenvalue
//         (= right singular vector of A for the smallest singular value, common.py:348-350)
//   err = mean over valid cameras of hypot(x_c - P_c[0].Q~/P_c[2].Q~, y_c - P_c[1].Q~/P_c[2].Q~)
//                                                    (common.py:357-375, triangulation.py:485-489)
//
// How the eigenvector is obtained.  With M = [[A, b], [b^T, c]] (A 3x3) and v = (q, 1), the smallest
// eigenpair satisfies (A - lam I) q = -b and the secular equation
//        f(lam) = c - lam + b.q(lam) = 0,   f'(lam) = -(1 + |q|^2),
// whose smallest root lies left of A's smallest eigenvalue.  f is concave and decreasing there, so
// Newton from lam = 0 overshoots once and then converges monotonically and quadratically from the
// right; a pivot check on the 3x3 LDL^T (positive definite <=> lam left of the pole) with bisection
// back towards the last lam known to be left of the root makes it unconditional.  With the chord step
// and the first-order final update below it needs 2.03 factorisations of a 3x3 per candidate on cfg2
// (measured; about 45 FP64 instructions each) instead of the ~2000 flops + 100 divisions/square roots
// of six cyclic Jacobi sweeps, returns q already de-homogenised, and agrees with cv2.SVDecomp to
// ~1e-12 m (measured).  The Jacobi variant is kept below for A/B.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define P2S_FULL 0xffffffffu

namespace p2s {

template <int CMAX>
struct CamParams {           // passed BY VALUE as a kernel parameter => lives in the constant bank,
    double P[CMAX][12];      // read as immediate-offset constant operands inside the unrolled loops
};

__device__ __forceinline__ double nan64() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ double inf64() { return __longlong_as_double(0x7ff0000000000000LL); }

// 1/d: MUFU.RCP64H seed (relative error < 2^-22) + one Newton step (2 DFMA) -> < 2^-44 (6e-14).  That is ample:
// the reciprocals feed the pivots of a Newton iteration that corrects itself and a mean over <= 32 distances whose
// tolerance is 1e-6 px.  P2S_RCP_CUBIC (A/B switch): the cubic correction (3 DFMA, ~1 ulp).
__device__ __forceinline__ double rcp_fast(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
#ifdef P2S_RCP_CUBIC
    double t = fma(e, e, e);
    return fma(y, t, y);
#else
    return fma(y, e, y);
#endif
}

// sqrt(a) for a >= 0 to ~1 ulp: MUFU.RSQ64H seed + Halley step + one residual correction.
__device__ __forceinline__ double sqrt_fast(double a) {
    double s = fmax(a, 1e-300);                     // a == 0 -> result 0 without a 0*inf NaN
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
    double t = s * r;
    double e = fma(-t, r, 1.0);                     // 1 - s r^2
    double p = fma(0.375, e, 0.5);
    r = fma(r * e, p, r);                           // r (1 + e/2 + 3 e^2/8)
    double q = a * r;
    double res = fma(-q, q, a);
    return fma(res * 0.5, r, q);
}

struct Sym4 {                                       // upper triangle of the symmetric 4x4 normal matrix
    double m00, m01, m02, m03, m11, m12, m13, m22, m23, m33;
};

__device__ __forceinline__ void sym4_zero(Sym4 &M) {
    M.m00 = M.m01 = M.m02 = M.m03 = M.m11 = M.m12 = M.m13 = M.m22 = M.m23 = M.m33 = 0.0;
}

// M += r r^T
__device__ __forceinline__ void sym4_rank1(Sym4 &M, double r0, double r1, double r2, double r3) {
    M.m00 = fma(r0, r0, M.m00); M.m01 = fma(r0, r1, M.m01); M.m02 = fma(r0, r2, M.m02); M.m03 = fma(r0, r3, M.m03);
    M.m11 = fma(r1, r1, M.m11); M.m12 = fma(r1, r2, M.m12); M.m13 = fma(r1, r3, M.m13);
    M.m22 = fma(r2, r2, M.m22); M.m23 = fma(r2, r3, M.m23);
    M.m33 = fma(r3, r3, M.m33);
}

// Adds camera c's two weighted DLT rows to M (common.py:344-345).
__device__ __forceinline__ void accumulate_camera(Sym4 &M, const double *Pc, double x, double y, double w) {
    double a0 = fma(-x, Pc[8], Pc[0]) * w, a1 = fma(-x, Pc[9], Pc[1]) * w;
    double a2 = fma(-x, Pc[10], Pc[2]) * w, a3 = fma(-x, Pc[11], Pc[3]) * w;
    sym4_rank1(M, a0, a1, a2, a3);
    double b0 = fma(-y, Pc[8], Pc[4]) * w, b1 = fma(-y, Pc[9], Pc[5]) * w;
    double b2 = fma(-y, Pc[10], Pc[6]) * w, b3 = fma(-y, Pc[11], Pc[7]) * w;
    sym4_rank1(M, b0, b1, b2, b3);
}

// Smallest eigenvector of M, de-homogenised: safeguarded Newton on the secular equation.
//
// Newton on f is Rayleigh-quotient iteration for v = (q, 1): lam + f/g is the Rayleigh quotient of the
// current v, so convergence is cubic.  Two refinements keep the number of 3x3 factorisations near two:
//   * chord step (first factorisation only): with w = (A - lam I)^-1 q = dq/dlam from the factors at
//     hand, q~ = q + dl w solves (A - lam - dl) q~ = -b up to the residual r = -dl^2 w, and the exact
//     Rayleigh quotient of (q~, 1) is lam + dl + (f~ - dl^2 q~.w) / (1 + |q~|^2); its error is the
//     square of q~'s, i.e. O((dl/dmin)^4), so the SECOND factorisation already sits at the root to
//     ~1e-8 relative for any candidate whose cameras roughly agree;
//   * final step: once |dl| ||(A - lam I)^-1|| <= 1e-6, q(lam + dl) = q + dl w to O(1e-12) relative, so no
//     factorisation is spent on confirming convergence.
// Returns the number of factorisations used.
// Iteration cap: almost every candidate needs two factorisations.  The exception is a smallest eigenvector whose last
// component is almost zero (a solution "at infinity": two nearly parallel rays): the secular root then sits within
// ~1e-3 relative of the pole, Newton keeps overshooting it and the safeguard degenerates into bisection — about forty
// steps.  With the former cap of 24 such a candidate came back unconverged, its error was wrong, and a FAILED unit
// could report another arg-min camera set than the reference (found by tests/perf/fuzz_parity.py).
#ifndef P2S_SOLVER_MAX_ITERS
#define P2S_SOLVER_MAX_ITERS 100
#endif
#ifndef P2S_DONE_TOL
#define P2S_DONE_TOL 1e-6        /* the first-order final update leaves O(tol^2) = 1e-12 relative in q; A/B: 1e-8 costs 5 % */
#endif

// What the first factorisation (lam = 0) yields beyond the iterate, for the branch-and-bound of the exclusion search:
//   s   = c - b^T A^-1 b = min over q of (q,1)^T M (q,1)  (the inhomogeneous least-squares minimum: q0 = -A^-1 b attains it),
//         so (q,1)^T M (q,1) >= s for EVERY q, in particular for the eigenvector-derived point the solver ends at;
//   g   = 1 + |q0|^2, s / g = the Rayleigh quotient of (q0, 1) >= the smallest eigenvalue lam*;
//   tau = trace(A^-1) >= 1 / (smallest eigenvalue of A), from the LDL^T factors at hand.
struct SecularBound { double s, g, tau; };

// Iteration 0 of the safeguarded Newton iteration (lam = 0, lo = 0), including the chord step.
// Returns 0: to be continued by secular_rest(M, lam, ...) — x holds q(lam = 0) + dl w;  1: finished, x final;
//         2: A is not positive definite (or NaN input): x = NaN, like the reference's NaN SVD.
__device__ __forceinline__ int secular_first(const Sym4 &M, double &lam, double &x0, double &x1, double &x2, SecularBound &B) {
    const double a00 = M.m00, a11 = M.m11, a22 = M.m22;
    const double r0 = rcp_fast(a00);
    const double l10 = M.m01 * r0, l20 = M.m02 * r0;
    const double d1 = fma(-l10, M.m01, a11), t21 = fma(-l20, M.m01, M.m12);
    const double r1 = rcp_fast(d1);
    const double l21 = t21 * r1;
    const double d2 = fma(-l21, t21, fma(-l20, M.m02, a22));
    lam = 0.0;
    if (!(a00 > 0.0 && d1 > 0.0 && d2 > 0.0)) {
        x0 = x1 = x2 = nan64();
        B.s = B.g = B.tau = nan64();
        return 2;
    }
    const double r2 = rcp_fast(d2);
    double z0 = -M.m03;
    double z1 = fma(-l10, z0, -M.m13);
    double z2 = fma(-l21, z1, fma(-l20, z0, -M.m23));
    x2 = z2 * r2;
    x1 = fma(-l21, x2, z1 * r1);
    x0 = fma(-l20, x2, fma(-l10, x1, z0 * r0));
    const double f = fma(M.m03, x0, fma(M.m13, x1, fma(M.m23, x2, M.m33)));
    const double g = fma(x0, x0, fma(x1, x1, fma(x2, x2, 1.0)));
    const double dl = f * rcp_fast(g);
    {   // trace(A^-1) = sum_k r_k |row k of L^-1|^2,  L^-1 = [[1,0,0],[-l10,1,0],[l10 l21 - l20, -l21, 1]]
        const double t = fma(l10, l21, -l20);
        B.tau = fma(r2, fma(t, t, fma(l21, l21, 1.0)), fma(r1, fma(l10, l10, 1.0), r0));
        B.s = f; B.g = g;
    }
    const double y1 = fma(-l10, x0, x1);
    const double y2 = fma(-l21, y1, fma(-l20, x0, x2));
    const double w2 = y2 * r2;
    const double w1 = fma(-l21, w2, y1 * r1);
    const double w0 = fma(-l20, w2, fma(-l10, w1, x0 * r0));
    const bool done = fabs(dl) * (r0 + r1 + r2) <= P2S_DONE_TOL;
    x0 = fma(dl, w0, x0); x1 = fma(dl, w1, x1); x2 = fma(dl, w2, x2);      // q(lam + dl), first order
    if (done) return 1;
    // chord step: Rayleigh quotient of (q~, 1)
    const double f1 = fma(M.m03, x0, fma(M.m13, x1, fma(M.m23, x2, M.m33 - dl)));
    const double g1 = fma(x0, x0, fma(x1, x1, fma(x2, x2, 1.0)));
    const double qw = fma(x0, w0, fma(x1, w1, x2 * w2));
    lam = dl + fma(-dl * dl, qw, f1) * rcp_fast(g1);
    return 0;
}

// Iterations 1.. of the same iteration (lo = 0 after iteration 0).  Returns the number of factorisations used in total.
__device__ __forceinline__ int secular_rest(const Sym4 &M, double lam, double &x0, double &x1, double &x2) {
    double lo = 0.0;
    int it = 1;
#pragma unroll 1
    for (; it < P2S_SOLVER_MAX_ITERS; ++it) {
        const double a00 = M.m00 - lam, a11 = M.m11 - lam, a22 = M.m22 - lam;
        const double r0 = rcp_fast(a00);
        const double l10 = M.m01 * r0, l20 = M.m02 * r0;
        const double d1 = fma(-l10, M.m01, a11), t21 = fma(-l20, M.m01, M.m12);
        const double r1 = rcp_fast(d1);
        const double l21 = t21 * r1;
        const double d2 = fma(-l21, t21, fma(-l20, M.m02, a22));
        if (!(a00 > 0.0 && d1 > 0.0 && d2 > 0.0)) {          // right of the pole
            if (!(lam > lo)) break;                          // not even positive definite at lo: give up
            lam = 0.5 * (lam + lo);
            continue;
        }
        const double r2 = rcp_fast(d2);
        double z0 = -M.m03;
        double z1 = fma(-l10, z0, -M.m13);
        double z2 = fma(-l21, z1, fma(-l20, z0, -M.m23));
        x2 = z2 * r2;
        x1 = fma(-l21, x2, z1 * r1);
        x0 = fma(-l20, x2, fma(-l10, x1, z0 * r0));
        const double f = fma(M.m03, x0, fma(M.m13, x1, fma(M.m23, x2, M.m33 - lam)));
        const double g = fma(x0, x0, fma(x1, x1, fma(x2, x2, 1.0)));
        const double dl = f * rcp_fast(g);
        if (f > 0.0) lo = lam;
        const bool done = fabs(dl) * (r0 + r1 + r2) <= P2S_DONE_TOL;
        if (done) {
            // w = (A - lam I)^-1 q from the factors at hand; q(lam + dl) = q + dl w, first order
            const double y1 = fma(-l10, x0, x1);
            const double y2 = fma(-l21, y1, fma(-l20, x0, x2));
            const double w2 = y2 * r2;
            const double w1 = fma(-l21, w2, y1 * r1);
            const double w0 = fma(-l20, w2, fma(-l10, w1, x0 * r0));
            x0 = fma(dl, w0, x0); x1 = fma(dl, w1, x1); x2 = fma(dl, w2, x2);
            ++it;
            break;
        }
        lam += dl;
    }
    return it;
}

__device__ __forceinline__ int smallest_eigvec_secular(const Sym4 &M, double &qx, double &qy, double &qz) {
    double lam;
    SecularBound B;
    const int r = secular_first(M, lam, qx, qy, qz, B);
    if (r == 2) return 0;
    if (r == 1) return 1;
    return secular_rest(M, lam, qx, qy, qz);
}

// Cyclic Jacobi on the 4x4 (north-star's nominal solver), eigenvectors accumulated; A/B only.
static __device__ __noinline__ int smallest_eigvec_jacobi(const Sym4 &M, double &qx, double &qy, double &qz) {
    double a[4][4] = {{M.m00, M.m01, M.m02, M.m03}, {M.m01, M.m11, M.m12, M.m13},
                      {M.m02, M.m12, M.m22, M.m23}, {M.m03, M.m13, M.m23, M.m33}};
    double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    int sweeps = 0;
#pragma unroll 1
    for (; sweeps < 12; ++sweeps) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[0][3]) + fabs(a[1][2]) + fabs(a[1][3]) + fabs(a[2][3]);
        double dia = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]) + fabs(a[3][3]);
        if (!(off > 1e-22 * dia)) break;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                double apq = a[p][q];
                if (apq != 0.0) {
                    double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
                    double t = 1.0 / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
                    t = theta < 0.0 ? -t : t;
                    double c = 1.0 / sqrt(fma(t, t, 1.0)), s = t * c;
                    a[p][p] = fma(-t, apq, a[p][p]);
                    a[q][q] = fma(t, apq, a[q][q]);
                    a[p][q] = a[q][p] = 0.0;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (r != p && r != q) {
                            double arp = a[r][p], arq = a[r][q];
                            a[r][p] = a[p][r] = fma(c, arp, -s * arq);
                            a[r][q] = a[q][r] = fma(s, arp, c * arq);
                        }
                        double vrp = v[r][p], vrq = v[r][q];
                        v[r][p] = fma(c, vrp, -s * vrq);
                        v[r][q] = fma(s, vrp, c * vrq);
                    }
                }
            }
        }
    }
    int k = 0;
    double dmin = a[0][0];
#pragma unroll
    for (int i = 1; i < 4; ++i) if (a[i][i] < dmin) { dmin = a[i][i]; k = i; }
    double w = 1.0 / (k == 0 ? v[3][0] : k == 1 ? v[3][1] : k == 2 ? v[3][2] : v[3][3]);
    qx = (k == 0 ? v[0][0] : k == 1 ? v[0][1] : k == 2 ? v[0][2] : v[0][3]) * w;
    qy = (k == 0 ? v[1][0] : k == 1 ? v[1][1] : k == 2 ? v[1][2] : v[1][3]) * w;
    qz = (k == 0 ? v[2][0] : k == 1 ? v[2][1] : k == 2 ? v[2][2] : v[2][3]) * w;
    return sweeps;
}

// ---- wide likelihood spread: factorisation of A instead of A^T A ------------------------------------------------
// The reference takes the SVD of the weighted 2m x 4 matrix A itself (common.py:347-350, cv2.SVDecomp).  The normal
// matrix squares the condition number: with valid likelihoods spanning w_max / w_min = s the eigenvector of A^T A
// is off by ~2e-16 s^2 metres on ring rigs (measured on reference-generated units: 2e-8 m at s = 1e4, 1.7e-4 m at
// s = 1e6, at most 2e-11 m for s <= 512 on the reference-generated units of tests/golden/tri_wide_likelihood.npz), so units with s > P2S_WIDE_SPREAD — only possible when the
// likelihood threshold is near 0 — take this path instead, off the common one and still on the device:
//   1. A = Q R by Givens rotations, the 2m rows streamed through a 4x4 upper-triangular R (10 registers; each row is
//      rebuilt from P, x, y, w on the fly), which is backward stable column by column;
//   2. one-sided Jacobi (Hestenes) on R accumulating V: the right singular vector of the smallest singular value of R
//      is that of A.  ~5 sweeps of 6 column pairs.
// Agreement with the reference's outputs on tests/golden/tri_wide_likelihood.npz (spreads up to 1e6): 5e-14 m.
#ifndef P2S_WIDE_SPREAD
#define P2S_WIDE_SPREAD 256.0f
#endif
struct Tri4 {                                       // upper-triangular 4x4
    double r00, r01, r02, r03, r11, r12, r13, r22, r23, r33;
};
__device__ __forceinline__ void tri4_zero(Tri4 &R) {
    R.r00 = R.r01 = R.r02 = R.r03 = R.r11 = R.r12 = R.r13 = R.r22 = R.r23 = R.r33 = 0.0;
}

// R <- the triangular factor of [R; a^T]  (four Givens rotations; a NaN row makes R NaN, like the reference's SVD)
static __device__ __noinline__ void givens_add_row(Tri4 &R, double a0, double a1, double a2, double a3) {
    if (a0 != 0.0) {
        const double rho = sqrt(fma(R.r00, R.r00, a0 * a0)), c = R.r00 / rho, s = a0 / rho;
        R.r00 = rho;
        double t = R.r01; R.r01 = fma(c, t, s * a1); a1 = fma(-s, t, c * a1);
        t = R.r02; R.r02 = fma(c, t, s * a2); a2 = fma(-s, t, c * a2);
        t = R.r03; R.r03 = fma(c, t, s * a3); a3 = fma(-s, t, c * a3);
    }
    if (a1 != 0.0) {
        const double rho = sqrt(fma(R.r11, R.r11, a1 * a1)), c = R.r11 / rho, s = a1 / rho;
        R.r11 = rho;
        double t = R.r12; R.r12 = fma(c, t, s * a2); a2 = fma(-s, t, c * a2);
        t = R.r13; R.r13 = fma(c, t, s * a3); a3 = fma(-s, t, c * a3);
    }
    if (a2 != 0.0) {
        const double rho = sqrt(fma(R.r22, R.r22, a2 * a2)), c = R.r22 / rho, s = a2 / rho;
        R.r22 = rho;
        const double t = R.r23; R.r23 = fma(c, t, s * a3); a3 = fma(-s, t, c * a3);
    }
    if (a3 != 0.0) R.r33 = sqrt(fma(R.r33, R.r33, a3 * a3));
}

// camera c's two weighted DLT rows (common.py:344-345) into R
__device__ __forceinline__ void givens_add_camera(Tri4 &R, const double *Pc, double x, double y, double w) {
    givens_add_row(R, fma(-x, Pc[8], Pc[0]) * w, fma(-x, Pc[9], Pc[1]) * w, fma(-x, Pc[10], Pc[2]) * w, fma(-x, Pc[11], Pc[3]) * w);
    givens_add_row(R, fma(-y, Pc[8], Pc[4]) * w, fma(-y, Pc[9], Pc[5]) * w, fma(-y, Pc[10], Pc[6]) * w, fma(-y, Pc[11], Pc[7]) * w);
}

// Right singular vector of R's smallest singular value, de-homogenised.  Returns the number of sweeps.
static __device__ __noinline__ int smallest_singvec_jacobi(const Tri4 &R, double &qx, double &qy, double &qz) {
    double g[4][4] = {{R.r00, R.r01, R.r02, R.r03}, {0.0, R.r11, R.r12, R.r13}, {0.0, 0.0, R.r22, R.r23}, {0.0, 0.0, 0.0, R.r33}};
    double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    int sweeps = 0;
#pragma unroll 1
    for (; sweeps < 16; ++sweeps) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const double al = fma(g[0][p], g[0][p], fma(g[1][p], g[1][p], fma(g[2][p], g[2][p], g[3][p] * g[3][p])));
                const double be = fma(g[0][q], g[0][q], fma(g[1][q], g[1][q], fma(g[2][q], g[2][q], g[3][q] * g[3][q])));
                const double ga = fma(g[0][p], g[0][q], fma(g[1][p], g[1][q], fma(g[2][p], g[2][q], g[3][p] * g[3][q])));
                if (ga * ga > 1e-30 * al * be) {               // |gamma| > 1e-15 sqrt(alpha beta); false for NaN
                    rotated = true;
                    const double z = (be - al) / (2.0 * ga);
                    double t = 1.0 / (fabs(z) + sqrt(fma(z, z, 1.0)));
                    t = z < 0.0 ? -t : t;
                    const double c = 1.0 / sqrt(fma(t, t, 1.0)), s = c * t;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double gp = g[i][p], gq = g[i][q];
                        g[i][p] = fma(c, gp, -s * gq);
                        g[i][q] = fma(s, gp, c * gq);
                        const double vp = v[i][p], vq = v[i][q];
                        v[i][p] = fma(c, vp, -s * vq);
                        v[i][q] = fma(s, vp, c * vq);
                    }
                }
            }
        }
        if (!rotated) break;
    }
    double n[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) n[j] = fma(g[0][j], g[0][j], fma(g[1][j], g[1][j], fma(g[2][j], g[2][j], g[3][j] * g[3][j])));
    int k = 0;
    double nmin = n[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) if (n[j] < nmin) { nmin = n[j]; k = j; }
    const double s4 = n[0] + n[1] + n[2] + n[3];
    if (!(s4 == s4)) { qx = qy = qz = nan64(); return sweeps; }       // NaN in A: cv2.SVDecomp returns NaN
    const double w = 1.0 / (k == 0 ? v[3][0] : k == 1 ? v[3][1] : k == 2 ? v[3][2] : v[3][3]);
    qx = (k == 0 ? v[0][0] : k == 1 ? v[0][1] : k == 2 ? v[0][2] : v[0][3]) * w;
    qy = (k == 0 ? v[1][0] : k == 1 ? v[1][1] : k == 2 ? v[1][2] : v[1][3]) * w;
    qz = (k == 0 ? v[2][0] : k == 1 ? v[2][1] : k == 2 ? v[2][2] : v[2][3]) * w;
    return sweeps;
}

// Pixel distance between the observation (x, y) and the reprojection of Q~ = (qx, qy, qz, 1):
//   hypot(x - u/d, y - v/d) = sqrt(N) / |d| = N * rsqrt(N d^2),   N = (u - x d)^2 + (v - y d)^2,
// i.e. ONE reciprocal square root (MUFU.RSQ64H seed + a cubic correction, ~1.5 ulp) instead of a
// reciprocal plus a square root.  The 1e-300 keeps N = 0 at distance 0 (0 * rsqrt(tiny)).
__device__ __forceinline__ double reproj_distance(const double *Pc, double qx, double qy, double qz, double x, double y) {
    const double u = fma(Pc[0], qx, fma(Pc[1], qy, fma(Pc[2], qz, Pc[3])));
    const double v = fma(Pc[4], qx, fma(Pc[5], qy, fma(Pc[6], qz, Pc[7])));
    const double d = fma(Pc[8], qx, fma(Pc[9], qy, fma(Pc[10], qz, Pc[11])));
    const double dx = fma(-x, d, u), dy = fma(-y, d, v);
    const double N = fma(dx, dx, dy * dy);
    const double S = fma(N, d * d, 1e-300);
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(S));
    const double t = S * r;
    const double e = fma(-t, r, 1.0);                // 1 - S r^2
#ifdef P2S_RCP_CUBIC
    const double p = fma(0.375, e, 0.5);
    r = fma(r * e, p, r);                            // r (1 + e/2 + 3 e^2/8)
#else
    r = fma(0.5 * r, e, r);                          // Newton: r (1 + e/2), relative error 3/8 e^2 < 3e-14
#endif
    return N * r;
}

// Lens model of one camera for `undistort_points = true` (OpenCV's pinhole + radial/tangential model).
struct LensParams {
    double R[9], T[3];           // world -> camera
    double fx, fy, cx, cy;       // ORIGINAL intrinsics (cv2.projectPoints re-projects with these)
    double k[8];                 // k1 k2 p1 p2 k3 k4 k5 k6
};
template <int CMAX>
struct LensSet { LensParams cam[CMAX]; };

// Pixel distance between the (undistorted) observation and the DISTORTED re-projection of Q — what
// the reference compares when undistort_points is on (triangulation.py:472-476, cv2.projectPoints with
// the original K and distortion coefficients).
__device__ __forceinline__ double reproj_distance_distorted(const LensParams &L, double qx, double qy, double qz,
                                                            double ox, double oy) {
    const double X = fma(L.R[0], qx, fma(L.R[1], qy, fma(L.R[2], qz, L.T[0])));
    const double Y = fma(L.R[3], qx, fma(L.R[4], qy, fma(L.R[5], qz, L.T[1])));
    const double Z = fma(L.R[6], qx, fma(L.R[7], qy, fma(L.R[8], qz, L.T[2])));
    const double iz = rcp_fast(Z);
    const double x = X * iz, y = Y * iz;
    const double r2 = fma(x, x, y * y), r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2.0 * x * y, a2 = fma(2.0 * x, x, r2), a3 = fma(2.0 * y, y, r2);
    const double cdist = fma(L.k[4], r6, fma(L.k[1], r4, fma(L.k[0], r2, 1.0)));
    const double icd2 = rcp_fast(fma(L.k[7], r6, fma(L.k[6], r4, fma(L.k[5], r2, 1.0))));
    const double s = cdist * icd2;
    const double xd = fma(x, s, fma(L.k[2], a1, L.k[3] * a2));
    const double yd = fma(y, s, fma(L.k[2], a3, L.k[3] * a1));
    const double dx = ox - fma(xd, L.fx, L.cx), dy = oy - fma(yd, L.fy, L.cy);
    const double N = fma(dx, dx, dy * dy);
    const double S = N + 1e-300;
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(S));
    const double t = S * r;
    const double e = fma(-t, r, 1.0);
    r = fma(r * e, fma(0.375, e, 0.5), r);
    return N * r;
}

// sum / m for a small positive integer m without the IEEE division slow path: reciprocal + one
// residual correction (correctly rounded except in rare half-way cases).
__device__ __forceinline__ double div_small(double sum, double m) {
    const double r = rcp_fast(m);
    const double e = sum * r;
    return fma(fma(-m, e, sum), r, e);
}

// Camera c's contribution to the normal matrix as a block: B = a a^T + b b^T with the two weighted
// DLT rows a = (P[0] - x P[2]) w, b = (P[1] - y P[2]) w (common.py:344-345).  Pc may point to shared
// memory (dynamic camera index).  blk = {m00, m01, m02, m03, m11, m12, m13, m22, m23, m33}.
__device__ __forceinline__ void camera_block(const double *Pc, double x, double y, double w, double *blk) {
    const double a0 = fma(-x, Pc[8], Pc[0]) * w, a1 = fma(-x, Pc[9], Pc[1]) * w;
    const double a2 = fma(-x, Pc[10], Pc[2]) * w, a3 = fma(-x, Pc[11], Pc[3]) * w;
    const double b0 = fma(-y, Pc[8], Pc[4]) * w, b1 = fma(-y, Pc[9], Pc[5]) * w;
    const double b2 = fma(-y, Pc[10], Pc[6]) * w, b3 = fma(-y, Pc[11], Pc[7]) * w;
    blk[0] = fma(b0, b0, a0 * a0); blk[1] = fma(b0, b1, a0 * a1); blk[2] = fma(b0, b2, a0 * a2); blk[3] = fma(b0, b3, a0 * a3);
    blk[4] = fma(b1, b1, a1 * a1); blk[5] = fma(b1, b2, a1 * a2); blk[6] = fma(b1, b3, a1 * a3);
    blk[7] = fma(b2, b2, a2 * a2); blk[8] = fma(b2, b3, a2 * a3);
    blk[9] = fma(b3, b3, a3 * a3);
}

// One candidate camera subset `valid` (bit c = camera c used).  `fetch(c)` returns the unit's
// observation {x, y, likelihood, -} for camera c (from the warp's shared-memory slab).
// Follows SURVEY.md §8(a) `solve`:
//   no camera -> (NaN, NaN)   [mean of an empty list];  one camera -> (NaN, +inf)  [common.py:351, :394-396]
template <int CMAX, int SOLVER, class Fetch>
__device__ __forceinline__ int solve_subset(const CamParams<CMAX> &cams, Fetch fetch, int n_cams, uint32_t valid,
                                            double &qx, double &qy, double &qz, double &err) {
    const int m = __popc(valid);
    if (m < 2) {
        qx = qy = qz = nan64();
        err = (m == 0) ? nan64() : inf64();
        return 0;
    }
    Sym4 M;
    sym4_zero(M);
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < n_cams && ((valid >> c) & 1u)) {
            const float4 o = fetch(c);
            accumulate_camera(M, cams.P[c], (double)o.x, (double)o.y, (double)o.z);
        }
    }
    int iters;
    if (SOLVER == 0) iters = smallest_eigvec_secular(M, qx, qy, qz);
    else iters = smallest_eigvec_jacobi(M, qx, qy, qz);
    double sum = 0.0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        if (c < n_cams && ((valid >> c) & 1u)) {
            const float4 o = fetch(c);
            sum += reproj_distance(cams.P[c], qx, qy, qz, (double)o.x, (double)o.y);
        }
    }
    err = div_small(sum, (double)m);
    return iters;
}

// Order-preserving key for the arg-min over candidate errors (errors are >= 0): finite values by
// value, +inf after them, NaN after that (np.nanargmin skips NaN; when EVERY candidate is NaN the
// reference would raise — here the first candidate wins with a NaN error), "no candidate" last.
#define P2S_KEY_NAN 0xfffffffffffffffeULL
#define P2S_KEY_EMPTY 0xffffffffffffffffULL
__device__ __forceinline__ unsigned long long err_key(double e) {
    return (e != e) ? P2S_KEY_NAN : (unsigned long long)__double_as_longlong(e);
}
// Triangulation: a NaN candidate error is the reference's +inf — its euclidean_distance (common.py:394-399) returns inf
// when the re-projection is NaN and never NaN — so it ties with +inf and the first candidate index wins.
__device__ __forceinline__ unsigned long long err_key_inf(double e) {
    return (e != e) ? 0x7ff0000000000000ULL : (unsigned long long)__double_as_longlong(e);
}
__device__ __forceinline__ double key_err(unsigned long long k) {
    return (k >= P2S_KEY_NAN) ? nan64() : __longlong_as_double((long long)k);
}

// ---- subset enumeration beyond the tabulated levels ------------------------------------------------
__device__ __forceinline__ uint32_t binom_u32(int n, int k) {
    if (k < 0 || k > n) return 0;
    if (k > n - k) k = n - k;
    unsigned long long r = 1;
    for (int i = 1; i <= k; ++i) r = r * (unsigned)(n - k + i) / (unsigned)i;
    return r > 0xffffffffULL ? 0xffffffffu : (uint32_t)r;
}

// rank -> mask of the rank-th k-subset of {0..n-1} in lexicographic order (itertools.combinations).
static __device__ __noinline__ uint32_t unrank_subset(int n, int k, uint32_t rank) {
    uint32_t mask = 0;
    int x = 0;
    for (int i = 0; i < k; ++i) {
        for (;; ++x) {
            uint32_t cnt = binom_u32(n - 1 - x, k - 1 - i);
            if (rank < cnt) break;
            rank -= cnt;
        }
        mask |= 1u << x;
        ++x;
    }
    return mask;
}

}  // namespace p2s
