/*
 * TEST INFRASTRUCTURE — plain-C restatement of the reference's triangulation hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library; the
 * product (pose2sim_b200/) never does.  It exists so that millions of units can be checked against
 * the CUDA path in seconds (the NumPy restatement oracle/p2s_oracle.py manages ~1.5 k units/s) and
 * so that bench.py can quote a *compiled, multi-threaded* CPU number beside the Python one.
 *
 * Parity pinning: validated against the UNMODIFIED reference through tests/golden/*.npz
 * (tests/test_oracle_golden.py::test_c_oracle_*), like the NumPy restatement.
 *
 * It deliberately does NOT share the CUDA path's numerics: the DLT is solved like the reference
 * does (common.py:348, cv2.SVDecomp), by a one-sided Jacobi SVD of the 2m x 4 matrix A — OpenCV's
 * own algorithm for small matrices (JacobiSVDImpl_ in modules/core/src/lapack.cpp, opencv-python
 * unpinned in the reference's pyproject.toml:54, 4.13.0 in this image) — not by an eigen-solve of
 * A^T A.  Agreement of the two is therefore a real check.
 *
 * Reference lines followed:
 *   weighted DLT rows ............. Pose2Sim/common.py:341-345
 *   >= 4 rows, V[:3,3]/V[3,3] ...... Pose2Sim/common.py:347-352
 *   reprojection ................... Pose2Sim/common.py:357-375
 *   distance, all-NaN -> inf ....... Pose2Sim/common.py:378-403
 *   exclusion search ............... Pose2Sim/triangulation.py:408-505, :588-602
 *   handle_LR_swap swapped pass .... Pose2Sim/triangulation.py:509-579 (as executed, see p2s_oracle.py::swapped_pass)
 *   association search ............. Pose2Sim/personAssociation.py:67-99, :154-257
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXC 32

/* Right singular vector of the smallest singular value of A (n x 4, row-major), by one-sided
 * (Hestenes) Jacobi: rotate column pairs of A until mutually orthogonal, accumulating V. */
static void smallest_right_singular_vector(double *A, int n, double v_out[4]) {
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 60; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < 3; ++p) {
            for (int q = p + 1; q < 4; ++q) {
                double a = 0, b = 0, g = 0;
                for (int i = 0; i < n; ++i) {
                    double x = A[i * 4 + p], y = A[i * 4 + q];
                    a += x * x; b += y * y; g += x * y;
                }
                if (fabs(g) <= 1e-17 * sqrt(a * b) || g == 0.0) continue;
                rotated = 1;
                double zeta = (b - a) / (2.0 * g);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < n; ++i) {
                    double x = A[i * 4 + p], y = A[i * 4 + q];
                    A[i * 4 + p] = c * x - s * y;
                    A[i * 4 + q] = s * x + c * y;
                }
                for (int i = 0; i < 4; ++i) {
                    double x = V[i][p], y = V[i][q];
                    V[i][p] = c * x - s * y;
                    V[i][q] = s * x + c * y;
                }
            }
        }
        if (!rotated) break;
    }
    int k = 0;
    double best = INFINITY;
    for (int j = 0; j < 4; ++j) {
        double nrm = 0;
        for (int i = 0; i < n; ++i) nrm += A[i * 4 + j] * A[i * 4 + j];
        if (nrm < best) { best = nrm; k = j; }
    }
    for (int i = 0; i < 4; ++i) v_out[i] = V[i][k];
}

/* DLT + reprojection + mean pixel error over the cameras in `valid` (ascending).
 * handle_LR_swap (triangulation.py:509-579 as executed, see p2s_oracle.py::swapped_pass): with xs != NULL the first
 * `n_first` valid cameras take the partner keypoint's coordinates (xs, ys) and the error is the mean over those
 * first `n_first` cameras only (:557-559); the DLT still runs over all valid cameras with the original likelihoods. */
static void solve_subset_swapped(const double *P, const double *x, const double *y, const double *w, const double *xs,
                                 const double *ys, int n_first, int C, uint32_t valid, double Q[3], double *err) {
    int m = __builtin_popcount(valid);
    Q[0] = Q[1] = Q[2] = NAN;
    if (m == 0) { *err = NAN; return; }             /* np.mean([]) */
    if (m == 1) { *err = INFINITY; return; }        /* < 4 rows -> NaN point -> distance inf */
    double A[2 * MAXC * 4];
    int r = 0;
    for (int c = 0; c < C; ++c) {
        if (!((valid >> c) & 1u)) continue;
        const double *Pc = P + c * 12;
        const double xc = (xs && r < n_first) ? xs[c] : x[c], yc = (xs && r < n_first) ? ys[c] : y[c];
        for (int j = 0; j < 4; ++j) {
            A[(2 * r) * 4 + j] = (Pc[j] - xc * Pc[8 + j]) * w[c];
            A[(2 * r + 1) * 4 + j] = (Pc[4 + j] - yc * Pc[8 + j]) * w[c];
        }
        ++r;
    }
    double v[4];
    smallest_right_singular_vector(A, 2 * m, v);
    Q[0] = v[0] / v[3]; Q[1] = v[1] / v[3]; Q[2] = v[2] / v[3];
    double sum = 0;
    int pos = 0;
    for (int c = 0; c < C; ++c) {
        if (!((valid >> c) & 1u)) continue;
        if (xs && pos >= n_first) break;
        const double *Pc = P + c * 12;
        const double xc = xs ? xs[c] : x[c], yc = xs ? ys[c] : y[c];
        double u = Pc[0] * Q[0] + Pc[1] * Q[1] + Pc[2] * Q[2] + Pc[3];
        double vv = Pc[4] * Q[0] + Pc[5] * Q[1] + Pc[6] * Q[2] + Pc[7];
        double d = Pc[8] * Q[0] + Pc[9] * Q[1] + Pc[10] * Q[2] + Pc[11];
        double dx = xc - u / d, dy = yc - vv / d;
        double dist;
        if (isnan(dx) && isnan(dy)) dist = INFINITY;
        else dist = sqrt((isnan(dx) ? 0 : dx * dx) + (isnan(dy) ? 0 : dy * dy));
        sum += dist;
        ++pos;
    }
    *err = sum / pos;
}

static void solve_subset(const double *P, const double *x, const double *y, const double *w, int C,
                         uint32_t valid, double Q[3], double *err) {
    solve_subset_swapped(P, x, y, w, NULL, NULL, 0, C, valid, Q, err);
}

/* next k-subset of {0..n-1} in lexicographic order; returns 0 when exhausted */
static int next_comb(int *idx, int n, int k) {
    int i = k - 1;
    while (i >= 0 && idx[i] == n - k + i) --i;
    if (i < 0) return 0;
    ++idx[i];
    for (int j = i + 1; j < k; ++j) idx[j] = idx[j - 1] + 1;
    return 1;
}

static void triangulate_unit(const double *P, const float *xf, const float *yf, const float *wf, const float *xsf,
                             const float *ysf, int C, double thr, int min_cams, double Q[3], double *err_out,
                             uint8_t *nexcl_out, uint32_t *mask_out, int *last_level, long long *n_cands) {
    double x[MAXC], y[MAXC], w[MAXC], xs[MAXC], ys[MAXC];
    uint32_t nan0 = 0, inv0 = 0;
    const uint32_t cmask = C >= 32 ? 0xffffffffu : ((1u << C) - 1u);
    for (int c = 0; c < C; ++c) {
        x[c] = xf[c]; y[c] = yf[c]; w[c] = wf[c];
        xs[c] = xsf ? xsf[c] : NAN; ys[c] = ysf ? ysf[c] : NAN;
        if (isnan(w[c])) { nan0 |= 1u << c; inv0 |= 1u << c; }
        else if (w[c] == 0.0) inv0 |= 1u << c;
    }
    double err_min = INFINITY;
    double Qb[3] = {NAN, NAN, NAN};
    uint32_t ids = cmask;
    int nexcl = C, evaluated = -1;
    for (int k = 0; err_min > thr && C - k >= min_cams; ++k) {
        /* break rule (:437-441): worst candidate's NaN-or-zero count */
        int idx[MAXC];
        for (int i = 0; i < k; ++i) idx[i] = i;
        int worst = 0;
        do {
            uint32_t cm = 0;
            for (int i = 0; i < k; ++i) cm |= 1u << idx[i];
            int cnt = __builtin_popcount(inv0 | cm);
            if (cnt > worst) worst = cnt;
        } while (next_comb(idx, C, k));
        if (worst > C - min_cams) break;
        for (int i = 0; i < k; ++i) idx[i] = i;
        double best = NAN;
        int have = 0;
        uint32_t bnan = nan0;
        int bexcl = 0;
        double bQ[3] = {NAN, NAN, NAN};
        int first = 1;
        do {
            uint32_t cm = 0;
            for (int i = 0; i < k; ++i) cm |= 1u << idx[i];
            double q[3], e;
            solve_subset(P, x, y, w, C, cmask & ~(inv0 | cm), q, &e);
            if (n_cands) ++*n_cands;
            if (first) { bnan = nan0 | cm; bexcl = __builtin_popcount(inv0 | cm); first = 0; }
            if (!isnan(e) && (!have || e < best)) {      /* nanargmin: first index of the minimum */
                have = 1; best = e; bnan = nan0 | cm; bexcl = __builtin_popcount(inv0 | cm);
                bQ[0] = q[0]; bQ[1] = q[1]; bQ[2] = q[2];
            }
        } while (next_comb(idx, C, k));
        err_min = have ? best : NAN;
        Qb[0] = bQ[0]; Qb[1] = bQ[1]; Qb[2] = bQ[2];
        ids = bnan; nexcl = bexcl; evaluated = k;
        /* handle_LR_swap: one swapped evaluation per candidate while the level is still above the threshold and
         * 1 < (n_cams - nb_cams_off_tot) / 2 (:509-513); np.min / argmin over the candidates (:565-566); error, Q and the
         * id list become the swapped winner's, nb_cams_excluded keeps the un-swapped winner's count (:574-577) */
        const int n_first = C - worst;
        if (xsf && err_min > thr && n_first > 2) {
            for (int i = 0; i < k; ++i) idx[i] = i;
            double sbest = INFINITY, sQ[3] = {NAN, NAN, NAN};
            uint32_t sids = 0;
            int shave = 0, snan = 0;
            do {
                uint32_t cm = 0;
                for (int i = 0; i < k; ++i) cm |= 1u << idx[i];
                double q[3], e;
                solve_subset_swapped(P, x, y, w, xs, ys, n_first, C, cmask & ~(inv0 | cm), q, &e);
                if (isnan(e)) snan = 1;                  /* np.min would be NaN and `NaN < error_min` false; mean of distances is never NaN */
                if (!shave || e < sbest) { shave = 1; sbest = e; sids = nan0 | cm; sQ[0] = q[0]; sQ[1] = q[1]; sQ[2] = q[2]; }
            } while (next_comb(idx, C, k));
            if (!snan && sbest < err_min) {
                err_min = sbest; ids = sids;
                Qb[0] = sQ[0]; Qb[1] = sQ[1]; Qb[2] = sQ[2];
            }
        }
    }
    if (err_min > thr) { err_min = NAN; Qb[0] = Qb[1] = Qb[2] = NAN; }
    Q[0] = Qb[0]; Q[1] = Qb[1]; Q[2] = Qb[2];
    *err_out = err_min; *nexcl_out = (uint8_t)nexcl; *mask_out = ids;
    if (last_level) *last_level = evaluated;
}

/* x, y, w: [n_units][C] float32 (NaN = invalid); P: C x 12 float64.  Threads via OpenMP if built with it. */
void p2s_oracle_triangulate(const float *x, const float *y, const float *w, const double *P, long long n_units,
                            int C, double thr, int min_cams, double *Q, double *err, uint8_t *nexcl,
                            uint32_t *mask, int32_t *level, long long *n_candidates) {
    long long total = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : total)
    for (long long u = 0; u < n_units; ++u) {
        int lv;
        long long nc = 0;
        triangulate_unit(P, x + u * C, y + u * C, w + u * C, NULL, NULL, C, thr, min_cams, Q + u * 3, err + u, nexcl + u,
                         mask + u, &lv, &nc);
        if (level) level[u] = lv;
        total += nc;
    }
    if (n_candidates) *n_candidates = total;
}

/* handle_LR_swap = true: units ordered (.., keypoint); the swapped coordinates of unit u are those of unit
 * u - u % n_keypoints + partner[u % n_keypoints] (triangulation.py:838, keypoints_idx_swapped :742-745). */
void p2s_oracle_triangulate_lrswap(const float *x, const float *y, const float *w, const int32_t *partner,
                                   int n_keypoints, const double *P, long long n_units, int C, double thr,
                                   int min_cams, double *Q, double *err, uint8_t *nexcl, uint32_t *mask) {
#pragma omp parallel for schedule(dynamic, 64)
    for (long long u = 0; u < n_units; ++u) {
        const long long up = u - u % n_keypoints + partner[u % n_keypoints];
        triangulate_unit(P, x + u * C, y + u * C, w + u * C, x + up * C, y + up * C, C, thr, min_cams, Q + u * 3,
                         err + u, nexcl + u, mask + u, NULL, NULL);
    }
}

/* obs: [n_frames][C][NP][4] float32; count: [n_frames][C].  out_comb: int8 [n_frames][C], -1 = off. */
void p2s_oracle_associate(const float *obs, const int32_t *count, const double *P, long long n_frames, int C,
                          int NP, double thr, double lik_thr, int min_cams, double *out_err, int8_t *out_comb,
                          double *out_Q) {
#pragma omp parallel for schedule(dynamic, 8)
    for (long long f = 0; f < n_frames; ++f) {
        const float *fo = obs + f * C * NP * 4;
        int n[MAXC];
        uint32_t ok[MAXC], present = 0;
        unsigned long long total_rows = 1;
        for (int c = 0; c < C; ++c) {
            n[c] = count[f * C + c];
            if (n[c] < 0) n[c] = 0;
            if (n[c] > NP) n[c] = NP;
            ok[c] = 0;
            for (int p = 0; p < n[c]; ++p) {
                double l = fo[(c * NP + p) * 4 + 2];
                if (!(l < lik_thr) && !(l == 0.0)) ok[c] |= 1u << p;
            }
            if (n[c]) present |= 1u << c;
            total_rows *= (unsigned long long)(n[c] ? n[c] : 1);
        }
        int n_missing = C - __builtin_popcount(present);
        double err_last = INFINITY, best = INFINITY;
        int have_best = 0;
        double bQ[3] = {NAN, NAN, NAN};
        int8_t bcomb[MAXC];
        for (int c = 0; c < C; ++c) bcomb[c] = -1;
        for (int k = 0; err_last > thr && C - (n_missing + k) >= min_cams; ++k) {
            for (unsigned long long r = 0; r < total_rows; ++r) {
                int dig[MAXC];
                unsigned long long q = r;
                uint32_t active = 0;
                double x[MAXC], y[MAXC], w[MAXC];
                for (int c = C - 1; c >= 0; --c) {
                    int p = 0;
                    if (n[c] > 1) { p = (int)(q % n[c]); q /= n[c]; }
                    dig[c] = p;
                    if (n[c] && ((ok[c] >> p) & 1u)) active |= 1u << c;
                    x[c] = fo[(c * NP + p) * 4]; y[c] = fo[(c * NP + p) * 4 + 1]; w[c] = fo[(c * NP + p) * 4 + 2];
                }
                int na = __builtin_popcount(active);
                if (na < min_cams || k > na) continue;
                /* k-subsets of the active cameras, lexicographic */
                int act[MAXC], idx[MAXC], m = 0;
                for (int c = 0; c < C; ++c) if ((active >> c) & 1u) act[m++] = c;
                for (int i = 0; i < k; ++i) idx[i] = i;
                int have = 0;
                double rbest = NAN, rQ[3] = {NAN, NAN, NAN};
                uint32_t rvalid = 0;
                do {
                    uint32_t cm = 0;
                    for (int i = 0; i < k; ++i) cm |= 1u << act[idx[i]];
                    double qq[3], e;
                    solve_subset(P, x, y, w, C, active & ~cm, qq, &e);
                    if (!isnan(e) && (!have || e < rbest)) {
                        have = 1; rbest = e; rvalid = active & ~cm;
                        rQ[0] = qq[0]; rQ[1] = qq[1]; rQ[2] = qq[2];
                    }
                } while (next_comb(idx, m, k));
                if (!have) continue;
                err_last = rbest;
                if (rbest < best) {
                    best = rbest; have_best = 1;
                    bQ[0] = rQ[0]; bQ[1] = rQ[1]; bQ[2] = rQ[2];
                    for (int c = 0; c < C; ++c) bcomb[c] = ((rvalid >> c) & 1u) ? (int8_t)dig[c] : -1;
                }
                if (rbest < thr) break;
            }
        }
        out_err[f] = have_best ? best : INFINITY;
        for (int c = 0; c < C; ++c) out_comb[f * C + c] = have_best ? bcomb[c] : -1;
        out_Q[f * 3] = bQ[0]; out_Q[f * 3 + 1] = bQ[1]; out_Q[f * 3 + 2] = bQ[2];
    }
}

int p2s_oracle_max_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
