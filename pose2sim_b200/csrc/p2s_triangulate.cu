// Triangulation with reprojection-error camera-exclusion search, one launch for all units.
//
// Replaces the per-unit Python call `triangulation_from_best_cameras`
// (Pose2Sim/triangulation.py:363-604; handle_LR_swap / undistort_points off) — restated in
// SURVEY.md §8(a) `triangulate_unit`:
//
//   k = 0; err_min = inf
//   while err_min > thr and C - k >= min_cams:                                (:408)
//       candidates = lexicographic k-subsets of ALL cameras                    (:411)
//       if max_i |inv0 U cand_i| > C - min_cams: break                         (:437-441)
//       solve every candidate on its valid cameras; arg-min, first index wins  (:469-505)
//       k += 1
//   ids / nexcl from the last evaluated level's best candidate, or all cameras (:588-596)
//   err_min > thr  ->  Q = NaN, err = NaN                                      (:600-602)
//
// Mapping to the machine (persistent grid, warp-autonomous tiles of 32 units):
//   * a warp stages the tile's float4 observations into its private shared-memory slab with
//     coalesced 512-byte LDG.128 rows (one per camera);
//   * level 0 has exactly one candidate per unit, so it runs THREAD-PER-UNIT (32 units per warp,
//     all lanes busy) — the north-star's warp-per-unit mapping would idle 31 lanes here;
//   * levels k >= 1 run LANES-ENUMERATE-SUBSETS: the warp's still-failing units are processed
//     G = 32/W at a time, W = min(32, pow2 >= C(C,k)) lanes each; a lane walks candidates
//     sub, sub+W, ...; candidate masks come from a lexicographic table (coalesced LDG); the
//     (error, index) arg-min and the runner-up for the eps-band statistics are reduced with
//     warp shuffles; the winning lane publishes Q / error / masks to the unit's slot in shared
//     memory;
//   * projection matrices are a by-value kernel parameter => constant-bank operands.
#include "p2s_math.cuh"
#include "p2s_internal.h"

namespace p2s {

struct TriArgs {
    const float4 *obs;            // [n_cams][n_units]
    long long n_units;
    int n_cams;
    int min_cams;
    double thr;
    double band_eps;
    const uint32_t *cand_masks;   // lexicographic subset table, level k at level_off[k]
    uint32_t level_off[P2S_MAX_CAMS + 2];
    int max_table_level;          // levels above this are unranked arithmetically
    double *out_Q;
    double *out_err;
    uint8_t *out_nexcl;
    uint32_t *out_mask;
    unsigned long long *stats;
    unsigned int *tile_counter;
};

template <int CMAX>
struct WarpSlab {                 // per-warp shared memory
    float4 obs[CMAX][32];         // staged observations of the tile
    double r_err[32];             // level results published by the winning lane of each unit
    double r_qx[32], r_qy[32], r_qz[32];
    uint32_t r_nan[32];           // NaN-camera set of the winner (id_excluded_cams)
    uint32_t r_flags[32];         // bit0..7: excl count, bit 8: argmin band hit
    uint32_t nan0[32];            // cameras whose likelihood is NaN
    uint32_t inv0[32];            // NaN or zero likelihood
};

template <int CMAX, int SOLVER>
__global__ void __launch_bounds__(128, 4) triangulate_kernel(const CamParams<CMAX> cams, const TriArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpSlab<CMAX> &S = reinterpret_cast<WarpSlab<CMAX> *>(smem_raw)[warp];

    const int C = a.n_cams;
    const uint32_t cmask = (C >= 32) ? 0xffffffffu : ((1u << C) - 1u);
    const long long n_tiles = (a.n_units + 31) >> 5;

    // per-warp statistics, flushed once at the end
    unsigned long long st_cands = 0, st_camsolves = 0, st_iters = 0;
    unsigned int st_failed = 0, st_noeval = 0, st_band_thr = 0, st_band_arg = 0;
    unsigned int st_level[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // levels >= 7 go straight to global

    for (;;) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(a.tile_counter, 1u);
        tile = __shfl_sync(P2S_FULL, tile, 0);
        if ((long long)tile >= n_tiles) break;

        const long long u = (long long)tile * 32 + lane;
        const bool active = u < a.n_units;

        // ---- stage the tile: one coalesced 512 B row per camera ------------------------------
        uint32_t nan0 = 0, inv0 = 0;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (c < C) {
                float4 o = make_float4(0.f, 0.f, __int_as_float(0x7fc00000), 0.f);
                if (active) o = __ldg(a.obs + (long long)c * a.n_units + u);
                S.obs[c][lane] = o;
                const bool isn = o.z != o.z;
                nan0 |= (uint32_t)isn << c;
                inv0 |= (uint32_t)(isn || o.z == 0.f) << c;
            }
        }
        S.nan0[lane] = nan0;
        S.inv0[lane] = inv0;
        __syncwarp();

        // ---- per-unit state (owner lane) -----------------------------------------------------
        double err_min = inf64();
        double qx = nan64(), qy = qx, qz = qx;
        uint32_t ids = cmask, nexcl = (uint32_t)C;
        int last_level = -1;
        bool band_thr = false, band_arg = false;
        const int ninv0 = __popc(inv0);

        // ---- level 0: thread per unit -----------------------------------------------------------
        {
            const bool go = active && (C >= a.min_cams) && !(ninv0 > C - a.min_cams);
            if (go) {
                double e;
                auto fetch = [&](int c) -> float4 { return S.obs[c][lane]; };
                int it = solve_subset<CMAX, SOLVER>(cams, fetch, C, cmask & ~inv0, qx, qy, qz, e);
                err_min = e;
                ids = nan0;
                nexcl = (uint32_t)ninv0;
                last_level = 0;
                st_cands += 1; st_camsolves += (unsigned)(C - ninv0); st_iters += (unsigned)it;
                band_thr |= fabs(e - a.thr) < a.band_eps;
            }
        }

        // ---- levels k >= 1: lanes enumerate subsets ---------------------------------------------
        for (int k = 1; k <= C; ++k) {
            // reference loop condition (:408) and break rule (:437-441) in closed form:
            // max_i |inv0 U cand_i| = min(C, |inv0| + k)
            const bool pend = active && last_level == k - 1 && (err_min > a.thr) && (C - k >= a.min_cams) &&
                              !(min(C, ninv0 + k) > C - a.min_cams);
            const uint32_t pmask = __ballot_sync(P2S_FULL, pend);
            if (pmask == 0) break;
            const int npend = __popc(pmask);
            const uint32_t ncand = (k <= a.max_table_level) ? (a.level_off[k + 1] - a.level_off[k]) : binom_u32(C, k);
            int W = 32;
            if (ncand <= 16) { W = 1; while ((uint32_t)W < ncand) W <<= 1; }
            const int G = 32 / W;
            const int grp = lane / W, sub = lane - grp * W;
            const uint32_t *table = a.cand_masks + a.level_off[k <= a.max_table_level ? k : 0];

            for (int base = 0; base < npend; base += G) {
                const int idx = base + grp;
                const bool on = idx < npend;
                const int ul = on ? (int)__fns(pmask, 0, idx + 1) : 0;     // owner lane of my unit
                const uint32_t u_nan0 = S.nan0[ul], u_inv0 = S.inv0[ul];

                unsigned long long bkey = P2S_KEY_EMPTY, skey = P2S_KEY_EMPTY;
                uint32_t bcand = 0xffffffffu, bnan = 0, bexcl = 0;
                double bqx = nan64(), bqy = bqx, bqz = bqx;
                if (on) {
                    for (uint32_t cand = (uint32_t)sub; cand < ncand; cand += (uint32_t)W) {
                        const uint32_t cm = (k <= a.max_table_level) ? __ldg(table + cand) : unrank_subset(C, k, cand);
                        const uint32_t nanset = u_nan0 | cm;
                        const uint32_t invset = u_inv0 | cm;
                        double cqx, cqy, cqz, e;
                        auto fetch = [&](int c) -> float4 { return S.obs[c][ul]; };
                        int it = solve_subset<CMAX, SOLVER>(cams, fetch, C, cmask & ~invset, cqx, cqy, cqz, e);
                        st_cands += 1; st_camsolves += (unsigned)(C - __popc(invset)); st_iters += (unsigned)it;
                        const unsigned long long key = err_key(e);
                        if (key < bkey) {                       // ascending cand per lane: strict < keeps the first
                            skey = bkey;
                            bkey = key; bcand = cand; bnan = nanset; bexcl = (uint32_t)__popc(invset);
                            bqx = cqx; bqy = cqy; bqz = cqz;
                        } else if (key > bkey && key < skey) {
                            skey = key;
                        }
                    }
                }
                // ---- (error, index) arg-min + runner-up across the W lanes of the group -------------
                for (int off = W >> 1; off > 0; off >>= 1) {
                    const unsigned long long okey = __shfl_xor_sync(P2S_FULL, bkey, off);
                    const unsigned long long oskey = __shfl_xor_sync(P2S_FULL, skey, off);
                    const uint32_t ocand = __shfl_xor_sync(P2S_FULL, bcand, off);
                    const bool take = (okey < bkey) || (okey == bkey && ocand < bcand);
                    // runner-up: smallest key strictly above the new best
                    const unsigned long long nb = take ? okey : bkey;
                    unsigned long long ns = P2S_KEY_EMPTY;
                    if (bkey > nb && bkey < ns) ns = bkey;
                    if (okey > nb && okey < ns) ns = okey;
                    if (skey > nb && skey < ns) ns = skey;
                    if (oskey > nb && oskey < ns) ns = oskey;
                    skey = ns;
                    bkey = nb;
                    bcand = take ? ocand : bcand;
                }
                // after the butterfly every lane of the group knows (bkey, bcand); the lane that
                // evaluated bcand (sub == bcand % W) still holds its Q / masks
                if (on && bcand != 0xffffffffu && (uint32_t)sub == (bcand & (uint32_t)(W - 1))) {
                    const double e = key_err(bkey);
                    S.r_err[ul] = e;
                    S.r_qx[ul] = bqx; S.r_qy[ul] = bqy; S.r_qz[ul] = bqz;
                    S.r_nan[ul] = bnan;
                    bool barg = false;
                    // runner-up among DISTINCT errors (duplicates of the winner are bitwise equal)
                    barg = (key_err(skey) - e) < a.band_eps;              // NaN / inf compare false
                    S.r_flags[ul] = bexcl | (barg ? 0x100u : 0u);
                }
            }
            __syncwarp();
            if (pend) {
                err_min = S.r_err[lane];
                qx = S.r_qx[lane]; qy = S.r_qy[lane]; qz = S.r_qz[lane];
                ids = S.r_nan[lane];
                const uint32_t fl = S.r_flags[lane];
                nexcl = fl & 0xffu;
                band_arg |= (fl & 0x100u) != 0;
                band_thr |= fabs(err_min - a.thr) < a.band_eps;
                last_level = k;
            }
            __syncwarp();
        }

        // ---- finalise (:588-602) and write -------------------------------------------------------
        if (active) {
            double e_out = err_min;
            const bool failed = err_min > a.thr;
            if (failed) { e_out = nan64(); qx = qy = qz = nan64(); }
            double *q = a.out_Q + u * 3;
            q[0] = qx; q[1] = qy; q[2] = qz;
            a.out_err[u] = e_out;
            a.out_nexcl[u] = (uint8_t)nexcl;
            a.out_mask[u] = ids;
            st_failed += failed ? 1u : 0u;
            st_noeval += last_level < 0 ? 1u : 0u;
            st_band_thr += band_thr ? 1u : 0u;
            st_band_arg += band_arg ? 1u : 0u;
            if (last_level >= 0) {
                if (last_level < 7) {
#pragma unroll
                    for (int l = 0; l < 7; ++l) st_level[l] += (last_level == l) ? 1u : 0u;
                } else if (a.stats) {
                    atomicAdd(a.stats + P2S_STAT_LEVEL0 + last_level, 1ULL);
                }
            }
        }
        __syncwarp();
    }

    // ---- flush statistics: warp-reduce, one atomic per counter per warp ---------------------------
    if (a.stats) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            st_cands += __shfl_xor_sync(P2S_FULL, st_cands, off);
            st_camsolves += __shfl_xor_sync(P2S_FULL, st_camsolves, off);
            st_iters += __shfl_xor_sync(P2S_FULL, st_iters, off);
            st_failed += __shfl_xor_sync(P2S_FULL, st_failed, off);
            st_noeval += __shfl_xor_sync(P2S_FULL, st_noeval, off);
            st_band_thr += __shfl_xor_sync(P2S_FULL, st_band_thr, off);
            st_band_arg += __shfl_xor_sync(P2S_FULL, st_band_arg, off);
#pragma unroll
            for (int l = 0; l < 7; ++l) st_level[l] += __shfl_xor_sync(P2S_FULL, st_level[l], off);
        }
        if (lane == 0) {
            if (st_cands) atomicAdd(a.stats + P2S_STAT_CANDIDATES, st_cands);
            if (st_camsolves) atomicAdd(a.stats + P2S_STAT_CAM_SOLVES, st_camsolves);
            if (st_iters) atomicAdd(a.stats + P2S_STAT_NEWTON_STEPS, st_iters);
            if (st_failed) atomicAdd(a.stats + P2S_STAT_FAILED, (unsigned long long)st_failed);
            if (st_noeval) atomicAdd(a.stats + P2S_STAT_NOT_EVALUATED, (unsigned long long)st_noeval);
            if (st_band_thr) atomicAdd(a.stats + P2S_STAT_BAND_THRESHOLD, (unsigned long long)st_band_thr);
            if (st_band_arg) atomicAdd(a.stats + P2S_STAT_BAND_ARGMIN, (unsigned long long)st_band_arg);
#pragma unroll
            for (int l = 0; l < 7; ++l)
                if (st_level[l]) atomicAdd(a.stats + P2S_STAT_LEVEL0 + l, (unsigned long long)st_level[l]);
        }
    }
}

// ---- staging: [U][C] planes -> float4 [C][U] with the likelihood gate (triangulation.py:817-821) ----
__global__ void __launch_bounds__(256) stage_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                    const float *__restrict__ lik, long long n_units, int n_cams,
                                                    double lik_thr, int gate, float4 *__restrict__ out) {
    extern __shared__ float sh[];                 // 3 planes of 256 * n_cams floats (+1 pad per row)
    const int C = n_cams;
    const int ld = C + 1;
    float *sx = sh, *sy = sh + 256 * ld, *sl = sh + 512 * ld;
    const long long n_blocks = (n_units + 255) / 256;
    for (long long b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const long long u0 = b * 256;
        const int nu = (int)min((long long)256, n_units - u0);
        const int n = nu * C;
        for (int i = threadIdx.x; i < n; i += 256) {       // coalesced reads of the row-major planes
            const int uu = i / C, cc = i - uu * C;
            sx[uu * ld + cc] = x[u0 * C + i];
            sy[uu * ld + cc] = y[u0 * C + i];
            sl[uu * ld + cc] = lik[u0 * C + i];
        }
        __syncthreads();
        if ((int)threadIdx.x < nu) {
            for (int c = 0; c < C; ++c) {                   // coalesced 16 B stores per camera row
                float vx = sx[threadIdx.x * ld + c], vy = sy[threadIdx.x * ld + c], vl = sl[threadIdx.x * ld + c];
                if (gate && (double)vl < lik_thr) { vx = vy = vl = __int_as_float(0x7fc00000); }
                out[(long long)c * n_units + u0 + threadIdx.x] = make_float4(vx, vy, vl, 0.f);
            }
        }
        __syncthreads();
    }
}

// ---- FP64 FMA peak microbenchmark ------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---- host-side launchers ----------------------------------------------------------------------------
template <int CMAX>
static cudaError_t launch_tri(const TriLaunch &L) {
    CamParams<CMAX> cams;
    for (int c = 0; c < CMAX; ++c)
        for (int j = 0; j < 12; ++j) cams.P[c][j] = (c < L.n_cams) ? L.P[c * 12 + j] : 0.0;
    TriArgs a;
    a.obs = (const float4 *)L.obs; a.n_units = L.n_units; a.n_cams = L.n_cams; a.min_cams = L.min_cams;
    a.thr = L.thr; a.band_eps = L.band_eps; a.cand_masks = L.cand_masks;
    for (int i = 0; i < P2S_MAX_CAMS + 2; ++i) a.level_off[i] = L.level_off[i];
    a.max_table_level = L.max_table_level;
    a.out_Q = L.out_Q; a.out_err = L.out_err; a.out_nexcl = L.out_nexcl; a.out_mask = L.out_mask;
    a.stats = L.stats; a.tile_counter = L.tile_counter;
    const size_t smem = sizeof(WarpSlab<CMAX>) * 4;
    cudaError_t e;
    if (L.solver == 0) {
        auto kern = triangulate_kernel<CMAX, 0>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        long long want = ((L.n_units + 31) / 32 + 3) / 4;
        long long grid = (long long)L.sm_count * per_sm;
        if (grid > want) grid = want;
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, 128, smem, L.stream>>>(cams, a);
    } else {
        auto kern = triangulate_kernel<CMAX, 1>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        long long want = ((L.n_units + 31) / 32 + 3) / 4;
        long long grid = (long long)L.sm_count * per_sm;
        if (grid > want) grid = want;
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, 128, smem, L.stream>>>(cams, a);
    }
    return cudaGetLastError();
}

cudaError_t launch_triangulate(const TriLaunch &L) {
    if (L.n_cams <= 4) return launch_tri<4>(L);
    if (L.n_cams <= 8) return launch_tri<8>(L);
    if (L.n_cams <= 16) return launch_tri<16>(L);
    return launch_tri<32>(L);
}

cudaError_t launch_stage(const float *x, const float *y, const float *lik, long long n_units, int n_cams,
                         double lik_thr, void *out, int sm_count, cudaStream_t stream) {
    const int gate = (lik_thr == lik_thr) && !(lik_thr == -INFINITY);
    const size_t smem = (size_t)3 * 256 * (n_cams + 1) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long blocks = (n_units + 255) / 256;
    long long grid = (long long)sm_count * 8;
    if (grid > blocks) grid = blocks;
    if (grid < 1) grid = 1;
    stage_kernel<<<(unsigned)grid, 256, smem, stream>>>(x, y, lik, n_units, n_cams, lik_thr, gate, (float4 *)out);
    return cudaGetLastError();
}

cudaError_t launch_fp64_peak(double *out, int blocks, int iters, cudaStream_t stream) {
    fp64_peak_kernel<<<blocks, 256, 0, stream>>>(out, iters, 1.0);
    return cudaGetLastError();
}

}  // namespace p2s
