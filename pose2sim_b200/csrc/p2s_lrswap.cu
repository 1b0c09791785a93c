// `handle_LR_swap = true` (Pose2Sim/triangulation.py:509-579): the exclusion search with the left/right
// swapped evaluation after every level whose error is still above the threshold.
//
// What the reference EXECUTES (restated in oracle/p2s_oracle.py::swapped_pass, checked against the live
// reference): `[[x] * n for x in x_files_filt]` (:518-519) makes every "sub-configuration" of a candidate the
// SAME array object, so the assignments at :525-526 accumulate in place — after the loops the first
// T' = n_cams - nb_cams_off_tot compacted positions of every candidate hold the partner keypoint's
// coordinates, whatever `n_cams_swapped` is.  Per candidate that is ONE extra evaluation: DLT over all its
// valid cameras (first T' of them, in ascending order, with the swapped coordinates, original likelihoods
// :529), error = mean distance over those first T' only (:557-559); np.min / argmin over the candidates
// (:565-566); when it beats the level's error, error / Q / id list become the swapped winner's while
// nb_cams_excluded stays the un-swapped winner's (:574-577).  The swap loop runs iff 1 < T' / 2 (:513).
//
// This mode is off in every shipped configuration (the fork's notes recommend it off), so it gets a compact
// kernel of its own rather than more template axes on the main one: a warp owns a tile of 32 units (lane = unit
// for the search state), and every level's (unit, candidate) pairs are dealt to the 32 lanes as work items
// (`level_pass`), with the same leaf math as the main kernel (p2s_math.cuh).  Input is the STAGED buffer (gated,
// undistorted if asked), so the partner's coordinates are already NaN where ITS likelihood failed the gate, like
// the reference's slices (:838).
#include <cmath>
#include <cstring>

#include "p2s_internal.h"
#include "p2s_math.cuh"

namespace p2s {

struct SwapArgs {
    const float4 *obs;            // staged [n_cams][n_units]
    const int32_t *partner;       // [n_keypoints] keypoint index of the left/right partner (itself when none)
    int n_keypoints;
    long long n_units;
    int n_cams, min_cams;
    double thr;
    const uint32_t *cand_masks;
    uint32_t level_off[P2S_MAX_CAMS + 2];
    int max_table_level;
    uint32_t ncand[P2S_MAX_CAMS + 1];
    double *out_Q, *out_err;
    uint8_t *out_nexcl;
    uint32_t *out_mask;
};

// One candidate.  The first `n_swapped` valid cameras (ascending) take (xs, ys); the error is the mean over the
// first `n_err` valid cameras (n_err >= m: all of them).
template <int CMAX, bool DISTORT>
__device__ __noinline__ void evaluate_candidate(const double *sP, const LensParams *sL,
                                                const float *x, const float *y, const float *w, const float *xs,
                                                const float *ys, int n_cams, uint32_t valid, int n_swapped, int n_err,
                                                double &qx, double &qy, double &qz, double &err) {
    const int m = __popc(valid);
    if (m < 2) {
        qx = qy = qz = nan64();
        err = (m == 0) ? nan64() : inf64();
        return;
    }
    Sym4 M;
    sym4_zero(M);
    int pos = 0;
    float wlo = __int_as_float(0x7f800000), whi = 0.f;
#pragma unroll 1
    for (int c = 0; c < n_cams; ++c) {
        if ((valid >> c) & 1u) {
            const bool sw = pos < n_swapped;
            accumulate_camera(M, sP + c * 12, (double)(sw ? xs[c] : x[c]), (double)(sw ? ys[c] : y[c]), (double)w[c]);
            wlo = fminf(wlo, fabsf(w[c])); whi = fmaxf(whi, fabsf(w[c]));
            ++pos;
        }
    }
    if (whi > P2S_WIDE_SPREAD * wlo) {
        // wide likelihood spread: factorisation of A instead of the normal matrix (p2s_math.cuh)
        Tri4 T;
        tri4_zero(T);
        pos = 0;
#pragma unroll 1
        for (int c = 0; c < n_cams; ++c) {
            if ((valid >> c) & 1u) {
                const bool sw = pos < n_swapped;
                givens_add_camera(T, sP + c * 12, (double)(sw ? xs[c] : x[c]), (double)(sw ? ys[c] : y[c]), (double)w[c]);
                ++pos;
            }
        }
        smallest_singvec_jacobi(T, qx, qy, qz);
    } else {
        smallest_eigvec_secular(M, qx, qy, qz);
    }
    double sum = 0.0;
    pos = 0;
#pragma unroll 1
    for (int c = 0; c < n_cams; ++c) {
        if ((valid >> c) & 1u) {
            if (pos < n_err) {
                const bool sw = pos < n_swapped;
                const double ox = (double)(sw ? xs[c] : x[c]), oy = (double)(sw ? ys[c] : y[c]);
                if (DISTORT) sum += reproj_distance_distorted(sL[DISTORT ? c : 0], qx, qy, qz, ox, oy);
                else sum += reproj_distance(sP + c * 12, qx, qy, qz, ox, oy);
            }
            ++pos;
        }
    }
    err = div_small(sum, (double)(m < n_err ? m : n_err));
}

#ifndef P2S_SWAP_WARPS
#define P2S_SWAP_WARPS 1            /* warps per CTA; 4 measured 2.19 ms, see DESIGN §4.6 */
#endif

// Per-warp exchange area: lane = unit slot of the warp's 32-unit tile.
struct SwapWarp {
    unsigned long long rkey[32];   // this round's smallest error key per unit
    unsigned long long skey[32];   // the pass's smallest error key per unit (first candidate index among equals)
    long long partner_unit[32];    // unit index of the left/right partner keypoint
    double q[32][3];               // the pass winner's point
    uint32_t rcand[32];            // this round's smallest candidate index among the holders of rkey
    uint32_t cm[32];               // the pass winner's excluded-camera mask
    uint32_t inv0[32];             // cameras without a usable likelihood (NaN or 0)
    uint32_t list[32];             // slots of the units taking part in the pass, compacted
};

// One pass (un-swapped or swapped) of exclusion level k over the units in `active`.  Every (unit, candidate) pair is
// one work item; the items are dealt to the 32 lanes round by round, so a unit that needs a deep level does not hold
// up its 31 neighbours (thread-per-unit measured 5.1 ms on the 520 k-unit benchmark, this form see DESIGN §4.6).
// Per round and unit: atomicMin of the error key, then of the candidate index among its holders; that lane updates
// the pass's running best on a STRICT `<` — candidates of a unit arrive in ascending order over the rounds, so this
// is np.nanargmin's "first index wins" (triangulation.py:565-566).
template <int CMAX, bool DISTORT>
__device__ __forceinline__ void level_pass(const SwapArgs &a, SwapWarp &S, const double *sP, const LensParams *sL,
                                           long long tile0, uint32_t active, int k, bool swapped, int lane) {
    const int C = a.n_cams;
    const uint32_t cmask = (C >= 32) ? 0xffffffffu : ((1u << C) - 1u);
    const uint32_t ncand = a.ncand[k];
    const uint32_t *table = a.cand_masks + a.level_off[k <= a.max_table_level ? k : 0];
    if ((active >> lane) & 1u) {
        S.list[__popc(active & ((1u << lane) - 1u))] = (uint32_t)lane;
        S.skey[lane] = P2S_KEY_EMPTY;
    }
    __syncwarp();
    const unsigned long long n_items = (unsigned long long)__popc(active) * ncand;
    float x[CMAX], y[CMAX], w[CMAX], xs[CMAX], ys[CMAX];
    int loaded = -1;
    for (unsigned long long base = 0; base < n_items; base += 32) {
        S.rkey[lane] = P2S_KEY_EMPTY;
        S.rcand[lane] = 0xffffffffu;
        __syncwarp();
        const unsigned long long item = base + (unsigned)lane;
        const bool live = item < n_items;
        int slot = 0;
        uint32_t cand = 0, cm = 0;
        unsigned long long key = P2S_KEY_EMPTY;
        double cx = 0, cy = 0, cz = 0;
        if (live) {
            slot = (int)S.list[item / ncand];
            cand = (uint32_t)(item % ncand);
            cm = (k == 0) ? 0u : (k <= a.max_table_level) ? table[cand] : unrank_subset(C, k, cand);
            if (slot != loaded) {                                  // consecutive items of a lane mostly stay on one unit
                const long long u = tile0 + slot, up = S.partner_unit[slot];
                for (int c = 0; c < C; ++c) {
                    const float4 o = a.obs[(long long)c * a.n_units + u];
                    const float4 p = a.obs[(long long)c * a.n_units + up];
                    x[c] = o.x; y[c] = o.y; w[c] = o.z; xs[c] = p.x; ys[c] = p.y;
                }
                loaded = slot;
            }
            const uint32_t inv0 = S.inv0[slot];
            const int n_first = C - min(C, __popc(inv0) + k);       // n_cams - nb_cams_off_tot (:437, :513)
            double e;
            evaluate_candidate<CMAX, DISTORT>(sP, sL, x, y, w, xs, ys, C, cmask & ~(inv0 | cm), swapped ? n_first : 0,
                                              swapped ? n_first : C, cx, cy, cz, e);
            key = err_key_inf(e);
            atomicMin(&S.rkey[slot], key);
        }
        __syncwarp();
        if (live && key == S.rkey[slot]) atomicMin(&S.rcand[slot], cand);
        __syncwarp();
        if (live && key == S.rkey[slot] && cand == S.rcand[slot] && key < S.skey[slot]) {
            S.skey[slot] = key;
            S.q[slot][0] = cx; S.q[slot][1] = cy; S.q[slot][2] = cz;
            S.cm[slot] = cm;
        }
        __syncwarp();
    }
}

template <int CMAX, bool DISTORT>
__global__ void __launch_bounds__(32 * P2S_SWAP_WARPS) lrswap_kernel(const CamParams<CMAX> cams, const LensSet<DISTORT ? CMAX : 1> lens,
                                                     const SwapArgs a) {
    // projection rows and lens models in shared memory: the candidate evaluation indexes them by camera
    __shared__ double sP[CMAX * 12];
    __shared__ LensParams sL[DISTORT ? CMAX : 1];
    __shared__ SwapWarp sW[P2S_SWAP_WARPS];
    for (int i = threadIdx.x; i < CMAX * 12; i += blockDim.x) sP[i] = (&cams.P[0][0])[i];
    if (DISTORT) {
        const double *src = reinterpret_cast<const double *>(&lens);
        double *dst = reinterpret_cast<double *>(sL);
        for (int i = threadIdx.x; i < (int)(sizeof(LensParams) / sizeof(double)) * CMAX; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int C = a.n_cams;
    const int lane = threadIdx.x & 31;
    SwapWarp &S = sW[threadIdx.x >> 5];
    const uint32_t cmask = (C >= 32) ? 0xffffffffu : ((1u << C) - 1u);
    const long long n_tiles = (a.n_units + 31) / 32;
    const long long warp0 = (long long)blockIdx.x * P2S_SWAP_WARPS + (threadIdx.x >> 5), n_warps = (long long)gridDim.x * P2S_SWAP_WARPS;
    for (long long tile = warp0; tile < n_tiles; tile += n_warps) {               // warp-uniform loop
        const long long tile0 = tile * 32, u = tile0 + lane;
        const bool in = u < a.n_units;
        uint32_t nan0 = 0, inv0 = 0;
        if (in) {
            for (int c = 0; c < C; ++c) {
                const float wl = a.obs[(long long)c * a.n_units + u].z;
                if (wl != wl) nan0 |= 1u << c;
                if (wl != wl || wl == 0.f) inv0 |= 1u << c;
            }
            const int kp = (int)(u % a.n_keypoints);
            S.partner_unit[lane] = u - kp + a.partner[kp];
        }
        S.inv0[lane] = inv0;
        __syncwarp();
        const int ninv0 = __popc(inv0);
        double err_min = inf64(), qx = nan64(), qy = nan64(), qz = nan64();
        uint32_t ids = cmask, nexcl = (uint32_t)C;
        for (int k = 0;; ++k) {
            const int T = min(C, ninv0 + k);                        // nb_cams_off_tot: the worst candidate's count (:437)
            const bool go = in && err_min > a.thr && C - k >= a.min_cams && T <= C - a.min_cams;    // :408, :440-441
            const uint32_t active = __ballot_sync(0xffffffffu, go);
            if (!active) break;
            level_pass<CMAX, DISTORT>(a, S, sP, sL, tile0, active, k, false, lane);
            if (go) {
                const unsigned long long key = S.skey[lane];
                if (key != P2S_KEY_EMPTY) {
                    qx = S.q[lane][0]; qy = S.q[lane][1]; qz = S.q[lane][2];
                    ids = nan0 | S.cm[lane]; nexcl = (uint32_t)__popc(inv0 | S.cm[lane]);
                }
                err_min = key_err(key);
            }
            __syncwarp();
            const bool go_sw = go && err_min > a.thr && C - T > 2;  // :509, :513 with n_cams_swapped = 1
            const uint32_t active_sw = __ballot_sync(0xffffffffu, go_sw);
            if (active_sw) {
                level_pass<CMAX, DISTORT>(a, S, sP, sL, tile0, active_sw, k, true, lane);
                if (go_sw) {
                    const double e_sw = key_err(S.skey[lane]);
                    if (e_sw < err_min) {                           // :574-577: nb_cams_excluded keeps the un-swapped count
                        err_min = e_sw;
                        qx = S.q[lane][0]; qy = S.q[lane][1]; qz = S.q[lane][2];
                        ids = nan0 | S.cm[lane];
                    }
                }
                __syncwarp();
            }
        }
        if (in) {
            const bool failed = err_min > a.thr;
            double *q = a.out_Q + u * 3;
            q[0] = failed ? nan64() : qx; q[1] = failed ? nan64() : qy; q[2] = failed ? nan64() : qz;
            a.out_err[u] = failed ? nan64() : err_min;
            a.out_nexcl[u] = (uint8_t)nexcl;
            a.out_mask[u] = ids;
        }
        __syncwarp();
    }
}

template <int CMAX>
static cudaError_t launch_swap(const SwapLaunch &L) {
    CamParams<CMAX> cams;
    for (int c = 0; c < CMAX; ++c)
        for (int j = 0; j < 12; ++j) cams.P[c][j] = (c < L.n_cams) ? L.P[c * 12 + j] : 0.0;
    SwapArgs a;
    a.obs = (const float4 *)L.obs; a.partner = L.partner; a.n_keypoints = L.n_keypoints;
    a.n_units = L.n_units; a.n_cams = L.n_cams; a.min_cams = L.min_cams; a.thr = L.thr;
    a.cand_masks = L.cand_masks;
    for (int i = 0; i < P2S_MAX_CAMS + 2; ++i) a.level_off[i] = L.level_off[i];
    a.max_table_level = L.max_table_level;
    for (int k = 0; k <= P2S_MAX_CAMS; ++k) {
        unsigned long long r = (k <= L.n_cams) ? 1ULL : 0ULL;
        for (int i = 1; i <= k && k <= L.n_cams; ++i) {
            r = r * (unsigned)(L.n_cams - k + i) / (unsigned)i;
            if (r > 0xffffffffULL) { r = 0xffffffffULL; break; }
        }
        a.ncand[k] = (uint32_t)r;
    }
    a.out_Q = L.out_Q; a.out_err = L.out_err; a.out_nexcl = L.out_nexcl; a.out_mask = L.out_mask;
    // one CTA per tile, no persistence: tiles differ a lot in cost (a unit may stop at level 0 or walk every level
    // twice), so the block scheduler does the balancing (a static stride left 32 % of the SM time idle at the tail,
    // and four-warp CTAs held their slot for the slowest warp: 11.7 of 20 possible warps per SM active)
    const long long per_cta = 32 * P2S_SWAP_WARPS;
    long long grid = (L.n_units + per_cta - 1) / per_cta;
    if (grid > 0x7fffffffLL) grid = 0x7fffffffLL;                 // the tile loop strides over the rest
    if (grid < 1) grid = 1;
    if (L.lens) {
        LensSet<CMAX> lens;
        std::memset(&lens, 0, sizeof lens);
        for (int c = 0; c < L.n_cams; ++c) {
            const p2s_camera_model &m = L.lens[c];
            LensParams &o = lens.cam[c];
            for (int j = 0; j < 9; ++j) o.R[j] = m.R[j];
            for (int j = 0; j < 3; ++j) o.T[j] = m.T[j];
            o.fx = m.K[0]; o.fy = m.K[4]; o.cx = m.K[2]; o.cy = m.K[5];
            for (int j = 0; j < 8; ++j) o.k[j] = m.dist[j];
        }
        lrswap_kernel<CMAX, true><<<(unsigned)grid, 32 * P2S_SWAP_WARPS, 0, L.stream>>>(cams, lens, a);
    } else {
        LensSet<1> none;
        std::memset(&none, 0, sizeof none);
        lrswap_kernel<CMAX, false><<<(unsigned)grid, 32 * P2S_SWAP_WARPS, 0, L.stream>>>(cams, none, a);
    }
    return cudaGetLastError();
}

cudaError_t launch_lrswap(const SwapLaunch &L) {
    if (L.n_cams <= 8) return launch_swap<8>(L);
    if (L.n_cams <= 16) return launch_swap<16>(L);
    return launch_swap<32>(L);
}

}  // namespace p2s
