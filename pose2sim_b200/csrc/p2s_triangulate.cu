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
// Mapping to the machine (persistent grid, warp-autonomous tiles of 32 units, two-deep tile dispenser):
//   * staging: on the fused raw-plane path lane 0 issues cp.async.bulk (TMA) copies of the NEXT tile's planes into
//     the warp's shared-memory buffer while the current tile is searched; the warp then gates the likelihoods and
//     transposes into its float2 xy[C][32] + float w[C][32] slab (staged-buffer API / C > 8: coalesced LDG.128
//     after an L2 prefetch one tile ahead);
//   * level 0 has exactly one candidate per unit, so it runs THREAD-PER-UNIT (32 units per warp,
//     all lanes busy) — the north-star's warp-per-unit mapping would idle 31 lanes here; its normal matrix is kept
//     (m0) as the sum of the unit's valid camera blocks for the deeper levels;
//   * levels k >= 1 run LANES-ENUMERATE-SUBSETS: the warp's still-failing units are processed
//     G = 32/W at a time, W = min(32, pow2 >= C(C,k)) lanes each; lane c of a group builds camera c's block, then
//     a lane walks candidates sub, sub+W, ... (masks from a lexicographic table, coalesced LDG) with
//     M = m0 - excluded blocks; the arg-min is one redux.sync + ballot in the common case (full 64-bit
//     (error, index) reduction with the runner-up for the eps-band statistics otherwise); the winning lane
//     publishes Q / error / masks to the unit's slot in shared memory;
//   * projection matrices are a by-value kernel parameter => constant-bank operands;
//   * a full tile's outputs leave as 16-byte vectors from a shared-memory staging area — locally or, on the push
//     path, straight into the consumer GPU's memory, followed by an arrival flag from the launch's last CTA.
#include <cstddef>
#include <cstdlib>
#include <cstring>

#include "p2s_math.cuh"
#include "p2s_internal.h"

namespace p2s {

struct TriArgs {
    const float4 *obs;            // staged [n_cams][n_units], or null when the raw planes are given
    const float *px, *py, *pl;    // raw planes [n_units][n_cams] (x, y, likelihood) or null
    double lik_thr;               // gate for the raw-plane path
    float lik_thr_f;              // smallest float >= lik_thr: `fl < lik_thr_f` in float <=> `(double)fl < lik_thr`
    int gate;
    long long n_units;
    int n_cams;
    int min_cams;
    double thr;
    double band_eps;
    const uint32_t *cand_masks;   // lexicographic subset table, level k at level_off[k]
    uint32_t level_off[P2S_MAX_CAMS + 2];
    int max_table_level;          // levels above this are unranked arithmetically
    uint32_t ncand[P2S_MAX_CAMS + 1];       // candidates of level k: C(n_cams, k) (saturating)
    unsigned char lw[P2S_MAX_CAMS + 1];     // log2 of the lanes per unit at level k: W = min(32, pow2 >= ncand)
    double rinv[P2S_MAX_CAMS + 1];          // 1 / m for the mean over m valid cameras
    double *out_Q;
    double *out_err;
    uint8_t *out_nexcl;
    uint32_t *out_mask;
    unsigned long long *stats;
    unsigned int *tile_counter;   // [0] tile dispenser, [1] fix-up CTAs finished, [2] tiles holding wide-spread units
                                  // (all zeroed before the launch)
    int vec_out;                  // output planes 16-byte aligned: full tiles are written as 16-byte vectors
    int bulk_out;                 // ... by cp.async.bulk (TMA) stores from the staging area instead of st.global.v4
    // multi-GPU push (outputs may live in a PEER's memory, p2s_triangulate_planes_push_device):
    const unsigned int *wait_flag;   // local: no output is written before *wait_flag >= wait_value (back-pressure)
    unsigned int wait_value;
    unsigned int *done_flag;         // local or peer: set to done_value once every output of the launch is visible
    unsigned int done_value;
    unsigned int *err_word;          // local: bit 0 set when the wait timed out
    // deep levels (deep_search_kernel below): a unit that is pending at a level with >= deep_min candidates is parked in
    // deep_list ((unit << 8) | level, counted in tile_counter[3]) instead of being walked by one warp; null = never
    unsigned long long *deep_list;
    unsigned int deep_cap;
    uint32_t deep_min;
};

__device__ __forceinline__ void prefetch_l2(const void *p) {
#ifndef P2S_NO_PREFETCH                                        /* A/B switch, tools/kernel_ab.py */
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// ---- TMA bulk copy global -> shared with a transaction barrier (sm_90+ PTX) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

#ifndef P2S_TRI_MIN_BLOCKS
#define P2S_TRI_MIN_BLOCKS 4        /* 4 vs 5 resident CTAs per SM measured equal (tools/kernel_ab.py); 4 has no spills.
                                       Wider slabs only fit 3 (24 cameras) / 2 (32 cameras) CTAs per SM anyway: the
                                       register cap follows, which removes the spills of the 32-camera unrolls */
#endif

#ifdef P2S_NO_DOWNDATE_EXACT                                  /* A/B switch, tools/kernel_ab.py: exact-count kernels keep both forms */
#define P2S_ALWAYS_DOWNDATE(exact) false
#else
#define P2S_ALWAYS_DOWNDATE(exact) (exact)                     /* cfg3 (16 cameras): 5.02 -> 4.95 ms */
#endif
template <int CMAX, bool STATS = false>
struct alignas(16) WarpSlab {     // per-warp shared memory
    // next tile's raw planes x | y | likelihood ([32 units][C] floats each), written by cp.async.bulk (TMA) while the
    // current tile is searched; only for CMAX <= 8 (3 KB per warp) — wider slabs would cost a resident CTA
    float raw[(CMAX <= 8) ? 3 * 32 * CMAX : 4];
    float2 xy[CMAX][32];          // staged observations of the tile: pixel coordinates ...
    float w[CMAX][32];            // ... and likelihood (NaN = invalid)
    unsigned long long mbar;      // transaction barrier of the raw buffer
    unsigned long long pad_;      // keeps blk (read back as 16-byte vectors by the output staging) 16-byte aligned
    double blk[32 * 10 + 32];     // camera blocks of the current group pass: block (group, camera) at
                                  // group * (10 C + 2) + 10 camera (the +2 staggers the groups over the banks)
    double2 gxy[32];              // levels >= 1: pixel coordinates of the current group pass as doubles, entry
                                  // (group, camera) at group * C + camera (G * C <= 32 because W >= C)
    double m0[10][32];            // level 0's normal matrix of every unit of the tile (entry-major): the sum over the
                                  // unit's valid cameras, from which levels >= 1 subtract the excluded blocks
    unsigned long long r_key[32]; // levels >= 1: error key of the unit's running best candidate of the level ...
    double r_qx[32], r_qy[32], r_qz[32];   // ... and its point, published by the winning lane of a pass
    unsigned long long st64[8];   // per-warp statistics: candidates, camera-solves, solver steps, solved,
                                  // direct cameras, blocks, entry additions
    uint32_t st32[12];            // level histogram [0..7], failed, not evaluated, threshold band, arg-min band
    uint32_t r_nan[32];           // NaN-camera set of the winner (id_excluded_cams)
    uint32_t r_flags[32];         // bit0..7: excl count, bit 8: argmin band hit
    uint32_t nan0[32];            // cameras whose likelihood is NaN
    uint32_t inv0[32];            // NaN or zero likelihood
    uint32_t plist[32];           // owner lanes of the units pending at the current level, compacted
};

static_assert(sizeof(WarpSlab<8>) % 16 == 0 && offsetof(WarpSlab<8>, blk) % 16 == 0 && offsetof(WarpSlab<8>, xy) % 16 == 0, "slab alignment");
static_assert(sizeof(WarpSlab<8, true>) % 16 == 0 && sizeof(WarpSlab<16, true>) % 16 == 0, "slab alignment");
static_assert(sizeof(WarpSlab<32>) % 16 == 0 && offsetof(WarpSlab<32>, blk) % 16 == 0, "slab alignment");

// Weighted-DLT normal matrix of ONE unit accumulated straight from the observations (level 0:
// thread per unit).  Invalid cameras hold x = y = w = 0 in the slab (written once when the tile is staged), i.e. exact
// zeros are added — no branch and no select per camera, so the unrolled cameras interleave in the FP64 pipe.
// "Poisoned" cameras: valid (likelihood neither NaN nor 0) but x or y is NaN.  The reference keeps such a camera (only the
// likelihood decides validity, triangulation.py:435-436), its DLT rows are NaN, cv2.SVDecomp returns NaN, and every
// distance of that candidate is +inf (common.py:394-396) — until the exclusion search drops the camera.  Nothing is
// tested here: a kept poisoned camera makes M, Q and the error NaN, which err_key_inf() orders as that +inf.
template <int CMAX>
__device__ __forceinline__ void accumulate_direct(Sym4 &M, const CamParams<CMAX> &cams, const float2 (*xy)[32],
                                                  const float (*wt)[32], int ul) {
    sym4_zero(M);
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        const float2 o = xy[c][ul];
        accumulate_camera(M, cams.P[c], (double)o.x, (double)o.y, (double)wt[c][ul]);
    }
}

// Level-0 fix-up for a unit with poisoned cameras (off the common path, so a rolled loop: small code): the level-0 matrix
// accumulate_direct() stored is NaN; rebuild it from the clean valid cameras, in the same order, so that the deeper levels
// subtract finite blocks from a finite sum (the block pass gives poisoned cameras a zero block).
__device__ __forceinline__ void rebuild_without_poisoned(double (*m0)[32], const double *sP, const float2 (*xy)[32],
                                                      const float (*wt)[32], int ul, uint32_t valid, int n_cams) {
    Sym4 M;
    sym4_zero(M);
    bool any = false;
#pragma unroll 1
    for (int c = 0; c < n_cams; ++c) {
        const float2 o = xy[c][ul];
        const bool v = (valid >> c) & 1u;
        const bool clean = v && (o.x == o.x) && (o.y == o.y);
        any |= v && !clean;
        if (clean) accumulate_camera(M, sP + c * 12, (double)o.x, (double)o.y, (double)wt[c][ul]);
    }
    if (any) {
        m0[0][ul] = M.m00; m0[1][ul] = M.m01; m0[2][ul] = M.m02; m0[3][ul] = M.m03; m0[4][ul] = M.m11;
        m0[5][ul] = M.m12; m0[6][ul] = M.m13; m0[7][ul] = M.m22; m0[8][ul] = M.m23; m0[9][ul] = M.m33;
    }
}

// ---- units whose valid likelihoods span more than P2S_WIDE_SPREAD ------------------------------------------------
// Their whole exclusion search runs here, one lane per unit, candidate after candidate, every candidate solved from a
// factorisation of A itself (Givens QR streamed over the cameras in ascending order + one-sided Jacobi on R,
// p2s_math.cuh) instead of the normal matrix.  Same loop rules as the main level loop (triangulation.py:408-505).
// Only reachable with a likelihood threshold near 0, so it lives in a kernel of its own (wide_fixup_kernel below, which
// returns at once when the main kernel counted no such unit): inside the main kernel an ABI call, even one placed after
// the level loop, cost the common path 11 registers and 4.9 % of cfg2's time (profiles/r2a_kernel_ab_wide.jsonl).
struct WideArgs {
    const double *sP;                  // projection matrices (shared memory)
    const LensParams *lens;            // DISTORT: lens models (shared memory), else null
    const float2 (*xy)[32];
    const float (*wt)[32];
    const uint32_t *table;             // lexicographic subset table, levels concatenated from level 0
    int max_table_level, n_cams, min_cams, ul;
    double thr;
    uint32_t nan0, inv0;
};
struct WideRes {
    double qx, qy, qz, err;
    uint32_t ids, nexcl;
    int last_level;
    uint32_t cands, cams, sweeps, solved;
};

template <bool DISTORT>
static __device__ __forceinline__ void search_wide_unit(const WideArgs &A, WideRes &R) {
    const int C = A.n_cams, ul = A.ul;
    const uint32_t cmask = (C >= 32) ? 0xffffffffu : ((1u << C) - 1u);
    const int ninv0 = __popc(A.inv0);
    double err_min = inf64();
    uint32_t off = 0;
    R.qx = R.qy = R.qz = nan64();
    R.ids = cmask; R.nexcl = (uint32_t)C; R.last_level = -1;
    R.cands = R.cams = R.sweeps = R.solved = 0;
#pragma unroll 1
    for (int k = 0; k <= C; ++k) {
        if (!(err_min > A.thr) || C - k < A.min_cams || min(C, ninv0 + k) > C - A.min_cams) break;
        const uint32_t ncand = binom_u32(C, k);
        unsigned long long bkey = P2S_KEY_EMPTY;
        uint32_t bnan = 0, bexcl = 0;
        double bqx = nan64(), bqy = bqx, bqz = bqx;
#pragma unroll 1
        for (uint32_t cand = 0; cand < ncand; ++cand) {
            const uint32_t cm = (k == 0) ? 0u : (k <= A.max_table_level) ? __ldg(A.table + off + cand) : unrank_subset(C, k, cand);
            const uint32_t invset = A.inv0 | cm;
            const uint32_t valid = cmask & ~invset;
            const int m = __popc(valid);
            double qx, qy, qz, e;
            if (m < 2) {
                qx = qy = qz = nan64();
                e = (m == 0) ? nan64() : inf64();
            } else {
                Tri4 T;
                tri4_zero(T);
#pragma unroll 1
                for (int c = 0; c < C; ++c) {
                    if (!((valid >> c) & 1u)) continue;
                    const float2 o = A.xy[c][ul];
                    givens_add_camera(T, A.sP + c * 12, (double)o.x, (double)o.y, (double)A.wt[c][ul]);
                }
                R.sweeps += (uint32_t)smallest_singvec_jacobi(T, qx, qy, qz);
                R.solved += 1u;
                double sum = 0.0;
#pragma unroll 1
                for (int c = 0; c < C; ++c) {
                    if (!((valid >> c) & 1u)) continue;
                    const float2 o = A.xy[c][ul];
                    if (DISTORT) sum += reproj_distance_distorted(A.lens[c], qx, qy, qz, (double)o.x, (double)o.y);
                    else sum += reproj_distance(A.sP + c * 12, qx, qy, qz, (double)o.x, (double)o.y);
                }
                e = sum * (1.0 / (double)m);
            }
            R.cands += 1u; R.cams += (uint32_t)m;
            const unsigned long long key = err_key_inf(e);
            if (key < bkey) {                                   // strict <: the first index wins (np.nanargmin)
                bkey = key; bnan = A.nan0 | cm; bexcl = (uint32_t)__popc(invset);
                bqx = qx; bqy = qy; bqz = qz;
            }
        }
        err_min = key_err(bkey);
        R.qx = bqx; R.qy = bqy; R.qz = bqz; R.ids = bnan; R.nexcl = bexcl; R.last_level = k;
        off += ncand;
    }
    R.err = err_min;
}

// min over the aligned group of W = 2^k lanes this lane belongs to (all 32 lanes take part)
// P2S_SHFL_ARGMIN (A/B switch): xor-shuffle butterfly; default: ONE redux.sync over the group's member mask
// (every group of the warp executes the same instruction with its own mask, like a cooperative-groups tile).
__device__ __forceinline__ uint32_t group_min(uint32_t v, int W, uint32_t gmask) {
#ifdef P2S_SHFL_ARGMIN
    for (int off = W >> 1; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(P2S_FULL, v, off));
    return v;
#else
    return __reduce_min_sync(gmask, v);
#endif
}

// Mean reprojection distance over the cameras in `valid` (all cameras are evaluated, the excluded ones are dropped
// by a predicated add: no branch, full instruction-level parallelism across cameras).  GXY: the unit's pixel
// coordinates come as doubles from the group's array (levels >= 1: converted once per unit by the camera-parallel
// block pass instead of twice per camera and candidate); otherwise from the float slab (level 0).
template <int CMAX, bool DISTORT, bool GXY>
__device__ __forceinline__ double mean_reproj_error(const CamParams<CMAX> &cams, const LensSet<DISTORT ? CMAX : 1> &lens,
                                                    const float2 (*obs)[32], const double2 *gxy, int ul, uint32_t valid,
                                                    double rinv_m, double qx, double qy, double qz, const double *sP) {
    double sum = 0.0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        double2 o;
        if (GXY) o = gxy[c];
        else { const float2 f = obs[c][ul]; o = make_double2((double)f.x, (double)f.y); }
        double dist;
        if (DISTORT) dist = reproj_distance_distorted(lens.cam[DISTORT ? c : 0], qx, qy, qz, o.x, o.y);
#ifdef P2S_SMEM_P                                              /* A/B switch: projection rows from shared memory */
        else dist = reproj_distance(sP + c * 12, qx, qy, qz, o.x, o.y);
#else
        else dist = reproj_distance(cams.P[c], qx, qy, qz, o.x, o.y);
#endif
        if ((valid >> c) & 1u) sum += dist;                      // compiles to two FSEL + DADD (an inline-PTX predicated
                                                                 // add is if-converted to the same selects, measured)
    }
    return sum * rinv_m;                                        // mean: 1/m from the host table (exactly rounded 1/m)
}

// DISTORT: `undistort_points = true` — observations were undistorted by the stage kernel, P is built on
// the optimal new camera matrix, and the error is measured against the distorted re-projection.
// EXACT: n_cams == CMAX, so the camera count is a compile-time constant (index arithmetic of the tile transposition
// and the camera loops fold).
// STATS: the statistics block (work counters, level histogram, eps-band counts incl. the arg-min runner-up) is
// wanted; the lean variant compiles all of that bookkeeping out of the candidate loop.
// the unit's C observations of one plane as registers (two / one 16-byte shared-memory loads)
template <int CMAX>
__device__ __forceinline__ void load_row(const float *plane, int lane, float (&v)[CMAX]) {
#pragma unroll
    for (int j = 0; j < CMAX; j += 4) {
        const float4 t = *reinterpret_cast<const float4 *>(plane + lane * CMAX + j);
        v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
    }
}

// RAW: raw planes with a compile-time camera count of 4 or 8 (what the TMA staging serves; implies EXACT).  The tile's
// planes land in S.raw unit-major, so a lane reads ITS OWN unit's row with 16-byte shared-memory loads and level 0 runs
// from registers: no transposition pass, and only the units that go on to level 1 write their slab column.
template <int CMAX, int SOLVER, bool DISTORT, bool EXACT, bool STATS, bool RAW = false>
__global__ void __launch_bounds__(128, CMAX <= 16 ? P2S_TRI_MIN_BLOCKS : CMAX <= 24 ? 3 : 2) triangulate_kernel(const CamParams<CMAX> cams,
                                                                               const LensSet<DISTORT ? CMAX : 1> lens,
                                                                               const TriArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double *sP = reinterpret_cast<double *>(smem_raw);                 // projection matrices for dynamic camera index
    WarpSlab<CMAX, STATS> &S = reinterpret_cast<WarpSlab<CMAX, STATS> *>(smem_raw + CMAX * 12 * sizeof(double))[warp];

    const int C = EXACT ? CMAX : a.n_cams;
    const uint32_t cmask = (C >= 32) ? 0xffffffffu : ((1u << C) - 1u);
    const uint32_t lt_mask = (1u << lane) - 1u;
    const long long n_tiles = (a.n_units + 31) >> 5;

    for (int i = threadIdx.x; i < CMAX * 12; i += blockDim.x) sP[i] = (&cams.P[0][0])[i];
    if (lane < 8) S.st64[lane] = 0ULL;
    if (lane < 12) S.st32[lane] = 0u;
    if (a.wait_flag != nullptr && threadIdx.x == 0) {
        // back-pressure of the push path: the consumer has released this output buffer once the flag reaches
        // wait_value.  Bounded (2 s): a lost peer must not hang the GPU; the time-out is reported, not hidden.
        const volatile unsigned int *wf = a.wait_flag;
        const unsigned long long t0 = global_ns();
        while ((int)(*wf - a.wait_value) < 0) {
            __nanosleep(200);
            if (global_ns() - t0 > 2000000000ULL) { atomicOr(a.err_word, 1u); break; }
        }
        __threadfence_system();
    }
    __syncthreads();

    // Tile dispenser, two deep: the tile after next is claimed (one atomic, result not touched for a whole tile)
    // and the next tile's input lines are pulled into L2 while the current tile is searched, so neither the
    // atomic's round trip nor HBM latency sits on the warp's critical path.
    unsigned int t1 = 0, t2 = 0;
    if (lane == 0) { t1 = atomicAdd(a.tile_counter, 1u); t2 = atomicAdd(a.tile_counter, 1u); }
    // TMA staging (raw planes, compile-time camera count <= 8): lane 0 issues three cp.async.bulk copies of the NEXT
    // tile's 32 C floats per plane into S.raw, completion is counted by the warp's transaction barrier.
#ifdef P2S_NO_TMA                                              /* A/B switch, tools/kernel_ab.py */
    const bool tma = false;
#else
    const bool tma = RAW || (EXACT && CMAX <= 8 && (CMAX * 4) % 16 == 0 && a.px != nullptr);   // bulk copies move multiples of 16 bytes
#endif
    static_assert(!RAW || (EXACT && !DISTORT && (CMAX == 4 || CMAX == 8)), "RAW: exact camera count of 4 or 8, no lens model");
    const uint32_t bar = smem_u32(&S.mbar);
    uint32_t phase = 0;
    auto issue_tile = [&](unsigned int t) {                    // lane 0 only
        const long long e0t = (long long)t * 32 * C;
        const long long left = a.n_units * C - e0t;
        const uint32_t bytes = (uint32_t)(left < 32LL * C ? left : 32LL * C) * 4u;      // multiple of 16: C in {4, 8}
        mbar_expect_tx(bar, 3u * bytes);
        bulk_g2s(smem_u32(S.raw), a.px + e0t, bytes, bar);
        bulk_g2s(smem_u32(S.raw + 32 * CMAX), a.py + e0t, bytes, bar);
        bulk_g2s(smem_u32(S.raw + 64 * CMAX), a.pl + e0t, bytes, bar);
    };
    if (tma) {
        if (lane == 0) {
            mbar_init(bar, 1u);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            fence_proxy_async();
            if ((long long)t1 < n_tiles) issue_tile(t1);
        }
        __syncwarp();
    }
    for (;;) {
        const unsigned int tile = __shfl_sync(P2S_FULL, t1, 0);
        if ((long long)tile >= n_tiles) break;
        if (!RAW && a.bulk_out) {                               // the previous tile's bulk stores have read the staging area
            if (lane == 0) bulk_wait_read();
            __syncwarp();
        }
        t1 = t2;
        if (lane == 0) t2 = atomicAdd(a.tile_counter, 1u);
        const unsigned int nt = __shfl_sync(P2S_FULL, t1, 0);   // the tile after this one
        if (!tma && (long long)nt < n_tiles) {
            if (a.px == nullptr) {
                for (int i = lane; i < 4 * C; i += 32) {          // 128-byte lines of the C staged rows of the tile
                    const long long e = (long long)(i >> 2) * a.n_units + (long long)nt * 32 + (i & 3) * 8;
                    if ((long long)nt * 32 + (i & 3) * 8 < a.n_units) prefetch_l2(a.obs + e);
                }
            } else {
                const long long e0n = (long long)nt * 32 * C, e_endn = a.n_units * C;
                for (int off = lane * 32; off < 32 * C; off += 1024) {   // 32 floats = one 128-byte line
                    if (e0n + off < e_endn) { prefetch_l2(a.px + e0n + off); prefetch_l2(a.py + e0n + off); prefetch_l2(a.pl + e0n + off); }
                }
            }
        }

        const long long u = (long long)tile * 32 + lane;
        const bool active = u < a.n_units;

        // ---- stage the tile ---------------------------------------------------------------------
        uint32_t nan0 = 0, inv0 = 0;
        float rx[RAW ? CMAX : 1], ry[RAW ? CMAX : 1], rl[RAW ? CMAX : 1];   // RAW: my unit's observations
        if constexpr (RAW) {
            mbar_wait(bar, phase);                                     // this tile's planes have landed in S.raw
            load_row<CMAX>(S.raw, lane, rx);
            load_row<CMAX>(S.raw + 32 * CMAX, lane, ry);
            load_row<CMAX>(S.raw + 64 * CMAX, lane, rl);
            __syncwarp();
            // every lane has its row in registers: hand the buffer back to the copy engine for the next tile
            phase ^= 1u;
            if (lane == 0 && (long long)nt < n_tiles) { fence_proxy_async(); issue_tile(nt); }
        } else if (a.px == nullptr) {
            // staged float4 SoA [C][U]: one coalesced 512 B row per camera
#pragma unroll
            for (int c = 0; c < CMAX; ++c) {
                float4 o = make_float4(0.f, 0.f, __int_as_float(0x7fc00000), 0.f);
                if (c < C && active) o = __ldg(a.obs + (long long)c * a.n_units + u);
                S.xy[c][lane] = make_float2(o.x, o.y);
                S.w[c][lane] = o.z;
            }
        } else {
            // raw planes x, y, lik [U][C] (what the host hands over): the tile is 32 C contiguous floats per
            // plane, read as coalesced float4 and transposed into the float4 SoA slab, with the likelihood
            // gate of triangulation.py:817-821 fused in (compared in double like the reference)
            const long long e0 = (long long)tile * 32 * C;             // first element of the tile
            const long long e_end = a.n_units * C;
            const float nanf_ = __int_as_float(0x7fc00000);
            __syncwarp();
            for (int c = 0; c < CMAX; ++c) { S.xy[c][lane] = make_float2(0.f, 0.f); S.w[c][lane] = nanf_; }
            __syncwarp();
            if (tma) mbar_wait(bar, phase);                            // this tile's planes have landed in S.raw
            for (int j = lane * 4; j < 32 * C; j += 128) {
                float vx[4], vy[4], vl[4];
                if (tma) {
                    const float4 tx = *reinterpret_cast<const float4 *>(S.raw + j);
                    const float4 ty = *reinterpret_cast<const float4 *>(S.raw + 32 * CMAX + j);
                    const float4 tl = *reinterpret_cast<const float4 *>(S.raw + 64 * CMAX + j);
                    vx[0] = tx.x; vx[1] = tx.y; vx[2] = tx.z; vx[3] = tx.w;
                    vy[0] = ty.x; vy[1] = ty.y; vy[2] = ty.z; vy[3] = ty.w;
                    vl[0] = tl.x; vl[1] = tl.y; vl[2] = tl.z; vl[3] = tl.w;
                    if (e0 + j + 3 >= e_end) {                         // beyond the last unit: stale bytes of an older tile
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (e0 + j + t >= e_end) { vx[t] = 0.f; vy[t] = 0.f; vl[t] = nanf_; }
                    }
                } else if (e0 + j + 3 < e_end) {
                    const float4 tx = __ldg(reinterpret_cast<const float4 *>(a.px + e0 + j));
                    const float4 ty = __ldg(reinterpret_cast<const float4 *>(a.py + e0 + j));
                    const float4 tl = __ldg(reinterpret_cast<const float4 *>(a.pl + e0 + j));
                    vx[0] = tx.x; vx[1] = tx.y; vx[2] = tx.z; vx[3] = tx.w;
                    vy[0] = ty.x; vy[1] = ty.y; vy[2] = ty.z; vy[3] = ty.w;
                    vl[0] = tl.x; vl[1] = tl.y; vl[2] = tl.z; vl[3] = tl.w;
                } else {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const bool in = e0 + j + t < e_end;
                        vx[t] = in ? a.px[e0 + j + t] : 0.f;
                        vy[t] = in ? a.py[e0 + j + t] : 0.f;
                        vl[t] = in ? a.pl[e0 + j + t] : nanf_;
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int uu = (j + t) / C, cc = (j + t) - uu * C;
                    float fx = vx[t], fy = vy[t], fl = vl[t];
                    // the reference compares the float64 likelihood with the float64 threshold (triangulation.py:817-821);
                    // for a float32 likelihood that is exactly `fl < (smallest float >= threshold)`
                    if (a.gate && fl < a.lik_thr_f) { fx = fy = fl = nanf_; }
                    S.xy[cc][uu] = make_float2(fx, fy);
                    S.w[cc][uu] = fl;
                }
            }
            __syncwarp();
            if (tma) {
                // every lane has consumed S.raw: hand the buffer back to the copy engine for the next tile
                phase ^= 1u;
                if (lane == 0 && (long long)nt < n_tiles) { fence_proxy_async(); issue_tile(nt); }
            }
        }
        float wlo = __int_as_float(0x7f800000), whi = 0.f;         // smallest / largest valid |likelihood| of the unit
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if constexpr (RAW) {
                const float nanf_ = __int_as_float(0x7fc00000);
                if (!active) { rx[c] = 0.f; ry[c] = 0.f; rl[c] = nanf_; }          // beyond the last unit: stale bytes
                if (a.gate && rl[c] < a.lik_thr_f) { rx[c] = ry[c] = rl[c] = nanf_; }
            }
            const float lz = RAW ? rl[RAW ? c : 0] : S.w[c][lane];
            const bool isn = lz != lz;
            const bool inv = isn || lz == 0.f;
            nan0 |= (uint32_t)isn << c;
            inv0 |= (uint32_t)inv << c;
            const float la = fabsf(lz);
            whi = fmaxf(whi, la);                                  // fmaxf / fminf skip NaN
            wlo = fminf(wlo, la > 0.f ? la : wlo);
            // invalid cameras hold exact zeros from here on: they add nothing to a normal matrix, so no pass below
            // needs a select per camera (validity lives in the nan0 / inv0 masks)
            if constexpr (RAW) {
                if (inv) { rx[c] = 0.f; ry[c] = 0.f; rl[c] = 0.f; }
            } else {
                if (inv) { S.xy[c][lane] = make_float2(0.f, 0.f); S.w[c][lane] = 0.f; }
            }
        }
        nan0 &= cmask; inv0 &= cmask;
        S.nan0[lane] = nan0;
        S.inv0[lane] = inv0;
        // units whose valid likelihoods span more than P2S_WIDE_SPREAD solve every candidate from a factorisation of A
        // (p2s_math.cuh, "wide likelihood spread"); never the case with a likelihood threshold >= 1 / 256
#ifdef P2S_NO_WIDE                                             /* A/B switch, tools/kernel_ab.py */
        const bool wide = false;
#else
        const bool wide = whi > P2S_WIDE_SPREAD * wlo;
#endif
        // such units are left to wide_fixup_kernel, which runs right after this kernel when the counter is non-zero
        if (__ballot_sync(P2S_FULL, wide && active) != 0u && lane == 0) atomicAdd(a.tile_counter + 2, 1u);

        // ---- per-unit state (owner lane) -----------------------------------------------------
        double err_min = inf64();
        double qx = nan64(), qy = qx, qz = qx;
        uint32_t ids = cmask, nexcl = (uint32_t)C;
        int last_level = -1;
        bool band_thr = false, band_arg = false;
        const int ninv0 = __popc(inv0);
        uint32_t t_cands = 0, t_cams = 0, t_iters = 0;      // per-tile work counters of this lane
        uint32_t t_solved = 0, t_direct = 0, t_blocks = 0, t_adds = 0;

        Sym4 M0;                                                // RAW: level-0 matrix until the unit's column is written
        sym4_zero(M0);
        // ---- level 0: one candidate per unit, THREAD PER UNIT, straight line ---------------------------------------
        // reference loop condition (:408) and break rule (:437-441) in closed form: max_i |inv0 U cand_i| = min(C, |inv0| + k)
        if (active && !wide && (err_min > a.thr) && (C >= a.min_cams) && !(min(C, ninv0) > C - a.min_cams)) {
            const uint32_t valid = cmask & ~inv0;
            const int m = C - ninv0;
            if (m >= 2) {
                Sym4 M;
                int it;
                double e;
                if constexpr (RAW) {
                    sym4_zero(M);
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) accumulate_camera(M, cams.P[c], (double)rx[c], (double)ry[c], (double)rl[c]);
                    M0 = M;
                    it = smallest_eigvec_secular(M, qx, qy, qz);
                    double sum = 0.0;
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) {
                        const double dist = reproj_distance(cams.P[c], qx, qy, qz, (double)rx[c], (double)ry[c]);
                        if ((valid >> c) & 1u) sum += dist;
                    }
                    e = sum * a.rinv[m];
                } else {
                    accumulate_direct<CMAX>(M, cams, S.xy, S.w, lane);
                    // kept (entry-major, conflict-free column): the sum of the unit's valid camera blocks, from which the deeper
                    // levels subtract the blocks they exclude
                    S.m0[0][lane] = M.m00; S.m0[1][lane] = M.m01; S.m0[2][lane] = M.m02; S.m0[3][lane] = M.m03; S.m0[4][lane] = M.m11;
                    S.m0[5][lane] = M.m12; S.m0[6][lane] = M.m13; S.m0[7][lane] = M.m22; S.m0[8][lane] = M.m23; S.m0[9][lane] = M.m33;
                    if (SOLVER == 0) it = smallest_eigvec_secular(M, qx, qy, qz);
                    else it = smallest_eigvec_jacobi(M, qx, qy, qz);
                    e = mean_reproj_error<CMAX, DISTORT, false>(cams, lens, S.xy, nullptr, lane, valid, a.rinv[m], qx, qy, qz, sP);
                }
                err_min = (e != e) ? inf64() : e;            // a NaN error is the reference's +inf (err_key_inf)
                t_iters += (uint32_t)it; t_solved += 1u; t_direct += (uint32_t)m;
            }                                                // m < 2: Q = NaN, error +inf (common.py:351, :394-396)
            t_cands += 1u; t_cams += (uint32_t)m;
            ids = nan0; nexcl = (uint32_t)ninv0; last_level = 0;
            band_thr |= fabs(err_min - a.thr) < a.band_eps;
            if constexpr (RAW) {
                // only a unit that goes on to level 1 needs its slab column (observations, level-0 matrix)
                if ((err_min > a.thr) && (C - 1 >= a.min_cams) && !(min(C, ninv0 + 1) > C - a.min_cams)) {
#pragma unroll
                    for (int c = 0; c < CMAX; ++c) { S.xy[c][lane] = make_float2(rx[c], ry[c]); S.w[c][lane] = rl[c]; }
                    S.m0[0][lane] = M0.m00; S.m0[1][lane] = M0.m01; S.m0[2][lane] = M0.m02; S.m0[3][lane] = M0.m03; S.m0[4][lane] = M0.m11;
                    S.m0[5][lane] = M0.m12; S.m0[6][lane] = M0.m13; S.m0[7][lane] = M0.m22; S.m0[8][lane] = M0.m23; S.m0[9][lane] = M0.m33;
                    if (!(err_min < inf64())) rebuild_without_poisoned(S.m0, sP, S.xy, S.w, lane, valid, C);
                }
            } else {
                // level 0 came out +inf: a unit with poisoned cameras gets its level-0 matrix rebuilt without them
                if (!(err_min < inf64())) rebuild_without_poisoned(S.m0, sP, S.xy, S.w, lane, valid, C);
            }
        }
        __syncwarp();

        // ---- levels k >= 1: lanes enumerate subsets, W = min(32, pow2 >= C(C,k)) lanes per unit, G = 32 / W units per pass
        for (int k = 1; k < C; ++k) {
            bool pend = active && !wide && last_level == k - 1 && (err_min > a.thr) && (C - k >= a.min_cams) &&
                        !(min(C, ninv0 + k) > C - a.min_cams);
            if constexpr (!STATS && SOLVER == 0 && !DISTORT && !RAW) {
                // a level of thousands of candidates is not walked by ONE warp (at 16 cameras a unit that reaches level 7
                // costs 26 k candidates = ~800 rounds, longer than the rest of the launch takes an SM): the unit is parked
                // and deep_search_kernel, behind this kernel, gives it a cluster of 512-thread CTAs.  Same arithmetic, same result.
                if (a.deep_list != nullptr && a.ncand[k] >= a.deep_min) {
                    const uint32_t pm = __ballot_sync(P2S_FULL, pend);
                    if (pm != 0u) {
                        unsigned int base = 0;
                        if (lane == 0) base = atomicAdd(a.tile_counter + 3, (unsigned int)__popc(pm));
                        base = __shfl_sync(P2S_FULL, base, 0);
                        const unsigned int slot = base + (unsigned int)__popc(pm & lt_mask);
                        if (pend && slot < a.deep_cap) {         // list full: the unit is searched here, as before
                            a.deep_list[slot] = ((unsigned long long)u << 8) | (unsigned long long)k;
                            pend = false;
                        }
                    }
                }
            }
            const uint32_t pmask = __ballot_sync(P2S_FULL, pend);
            if (pmask == 0) break;
            const int npend = __popc(pmask);
            if (pend) {
                S.plist[__popc(pmask & lt_mask)] = (uint32_t)lane;
                S.r_key[lane] = P2S_KEY_EMPTY;
            }
            const uint32_t ncand = a.ncand[k];
            const int lw = a.lw[k];                          // W = 2^lw lanes per unit (host table), W >= C
            const int W = 1 << lw;
            const int G = 32 >> lw;
            const int grp = lane >> lw, sub = lane & (W - 1);
            const uint32_t gmask = (W >= 32) ? P2S_FULL : (((1u << W) - 1u) << (grp << lw));   // lanes of my group
            const bool tabled = RAW || k <= a.max_table_level;   // <= 8 cameras: every level is in the table (256 entries)
            const uint32_t *table = a.cand_masks + a.level_off[tabled ? k : 0];
            double *gblk = S.blk + grp * (C * 10 + 2);
            // M = M_all - excluded blocks, else sum of the kept blocks.  Exact-count kernels: always the downdate — the levels
            // that drop more cameras than they keep are reached by ~1e-5 of the units, and one path less is code the
            // candidate loop does not have to carry (the zero-initialisation of M was hoisted in front of the branch)
            const bool subtract = RAW || P2S_ALWAYS_DOWNDATE(EXACT) || 2 * k <= C;
            __syncwarp();

            for (int base = 0; base < npend; base += G) {
                const int idx = base + grp;
                const bool on = idx < npend;
                const int ul = on ? (int)S.plist[idx] : 0;                 // owner lane of my unit
                const uint32_t u_nan0 = S.nan0[ul], u_inv0 = S.inv0[ul];

                // camera-parallel: lane `sub` builds the block of camera `sub` of its group's unit (the sum of the
                // valid blocks, M_all, is level 0's normal matrix, kept in S.m0)
                __syncwarp();
                if (on && sub < C) {
                    const float2 o = S.xy[sub][ul];
                    const float ow = S.w[sub][ul];
                    // A valid camera whose x or y is NaN ("poisoned", see accumulate_direct) gets a ZERO block, so M and
                    // its downdates stay finite, but keeps its NaN in gxy: the distance of every candidate that keeps
                    // the camera is NaN, which err_key_inf() orders as the reference's +inf.  (Invalid cameras hold zeros.)
                    const bool clean = (o.x == o.x) && (o.y == o.y);
                    const double ox = (double)o.x, oy = (double)o.y;
                    S.gxy[grp * C + sub] = make_double2(ox, oy);
                    double b[10];
                    camera_block(sP + sub * 12, clean ? ox : 0.0, clean ? oy : 0.0, (double)(clean ? ow : 0.f), b);
                    if (STATS) t_blocks += ((u_inv0 >> sub) & 1u) ? 0u : 1u;
                    double2 *dst = reinterpret_cast<double2 *>(gblk + sub * 10);
#pragma unroll
                    for (int e = 0; e < 5; ++e) dst[e] = make_double2(b[2 * e], b[2 * e + 1]);
                }
                __syncwarp();

                // a lane walks candidates sub, sub + W, ... and keeps its own best (first = smallest index among equal keys);
                // levels with at most W candidates — the common ones — run the body once
                unsigned long long bkey = P2S_KEY_EMPTY, bskey = P2S_KEY_EMPTY;
                uint32_t bcand = 0xffffffffu, bcm = 0;
                double bqx, bqy, bqz;
                for (uint32_t cand = (uint32_t)sub; on && cand < ncand; cand += (uint32_t)W) {
                    const uint32_t cm = tabled ? __ldg(table + cand) : unrank_subset(C, k, cand);
                    const uint32_t valid = cmask & ~(u_inv0 | cm);
                    const int m = __popc(valid);
                    double cqx = nan64(), cqy = cqx, cqz = cqx;
                    double e = inf64();                      // m < 2: common.py:351, :394-396 / mean of an empty list,
                    if (m >= 2) {                            // both ordered as +inf
                        Sym4 M;
                        uint32_t bits;
                        double sgn;
                        if (subtract) {
                            M.m00 = S.m0[0][ul]; M.m01 = S.m0[1][ul]; M.m02 = S.m0[2][ul]; M.m03 = S.m0[3][ul]; M.m11 = S.m0[4][ul];
                            M.m12 = S.m0[5][ul]; M.m13 = S.m0[6][ul]; M.m22 = S.m0[7][ul]; M.m23 = S.m0[8][ul]; M.m33 = S.m0[9][ul];
                            bits = cm & ~u_inv0 & cmask;
                            sgn = -1.0;
                        } else {
                            sym4_zero(M);
                            bits = valid;
                            sgn = 1.0;
                        }
                        if (STATS) t_adds += 10u * (uint32_t)__popc(bits);
                        while (bits) {                        // ascending camera order
                            const int c = __ffs(bits) - 1;
                            bits &= bits - 1;
                            const double2 *src = reinterpret_cast<const double2 *>(gblk + c * 10);
                            const double2 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4];
                            M.m00 = fma(sgn, v0.x, M.m00); M.m01 = fma(sgn, v0.y, M.m01); M.m02 = fma(sgn, v1.x, M.m02);
                            M.m03 = fma(sgn, v1.y, M.m03); M.m11 = fma(sgn, v2.x, M.m11); M.m12 = fma(sgn, v2.y, M.m12);
                            M.m13 = fma(sgn, v3.x, M.m13); M.m22 = fma(sgn, v3.y, M.m22); M.m23 = fma(sgn, v4.x, M.m23);
                            M.m33 = fma(sgn, v4.y, M.m33);
                        }
                        int it;
                        if (SOLVER == 0) it = smallest_eigvec_secular(M, cqx, cqy, cqz);
                        else it = smallest_eigvec_jacobi(M, cqx, cqy, cqz);
                        e = mean_reproj_error<CMAX, DISTORT, true>(cams, lens, S.xy, S.gxy + grp * C, ul, valid, a.rinv[m], cqx, cqy, cqz, sP);
                        if (STATS) { t_iters += (uint32_t)it; t_solved += 1u; }
                    }
                    if (STATS) { t_cands += 1u; t_cams += (uint32_t)m; }
                    const unsigned long long key = err_key_inf(e);
                    if (key < bkey) {                           // ascending cand per lane: strict < keeps the first
                        if (STATS) bskey = bkey;
                        bkey = key; bcand = cand; bcm = cm;
                        bqx = cqx; bqy = cqy; bqz = cqz;
                    } else if (STATS && key > bkey && key < bskey) {
                        bskey = key;
                    }
                }
                // ---- arg-min across the W lanes of the group: keys are 64-bit — min of the high words (ONE redux.sync),
                // then, only if several lanes hold it (duplicates, all-inf levels, errors closer than 2^-20 relative), min of
                // the low words among them and the smallest candidate index among those (np.nanargmin's first index)
                const uint32_t hi = (uint32_t)(bkey >> 32), lo = (uint32_t)bkey;
                const uint32_t mh = group_min(hi, W, gmask);
                const bool have = bcand != 0xffffffffu;
                uint32_t holders = __ballot_sync(P2S_FULL, have && hi == mh) & gmask;
                bool winner = holders != 0u && lane == __ffs(holders) - 1;
                bool barg = false;
                if (__any_sync(P2S_FULL, STATS || __popc(holders) > 1)) {
                    const uint32_t ml = group_min(hi == mh ? lo : 0xffffffffu, W, gmask);
                    const bool is_min = have && hi == mh && lo == ml;
                    const uint32_t mc = group_min(is_min ? bcand : 0xffffffffu, W, gmask);
                    winner = is_min && bcand == mc;
                    if (STATS) {
                        // runner-up: smallest key strictly above the minimum (duplicates of the winner are bitwise equal)
                        const unsigned long long rk = is_min ? bskey : bkey;
                        const uint32_t rh = (uint32_t)(rk >> 32), rl = (uint32_t)rk;
                        const uint32_t sh = group_min(rh, W, gmask);
                        const uint32_t sl = group_min(rh == sh ? rl : 0xffffffffu, W, gmask);
                        barg = (key_err(((unsigned long long)sh << 32) | sl) - key_err(bkey)) < a.band_eps;   // NaN / inf compare false
                    }
                }
                if (winner) {
                    S.r_key[ul] = bkey;
                    S.r_qx[ul] = bqx; S.r_qy[ul] = bqy; S.r_qz[ul] = bqz;
                    S.r_nan[ul] = u_nan0 | bcm;
                    S.r_flags[ul] = (uint32_t)__popc(u_inv0 | bcm) | (barg ? 0x100u : 0u);
                }
            }
            __syncwarp();
            if (pend) {
                const unsigned long long bk = S.r_key[lane];
                err_min = key_err(bk);
                qx = S.r_qx[lane]; qy = S.r_qy[lane]; qz = S.r_qz[lane];
                ids = S.r_nan[lane];
                nexcl = S.r_flags[lane] & 0xffu;
                if (STATS) band_arg |= (S.r_flags[lane] & 0x100u) != 0u;
                band_thr |= fabs(err_min - a.thr) < a.band_eps;
                last_level = k;
            }
            __syncwarp();
        }

        // ---- finalise (:588-602) and write -------------------------------------------------------
        const bool failed = active && (err_min > a.thr);
        double e_out = err_min;
        if (failed) { e_out = nan64(); qx = qy = qz = nan64(); }
        if (a.vec_out && (long long)tile * 32 + 32 <= a.n_units) {
            // full tile: the four output planes are contiguous per tile (768 + 256 + 128 + 32 bytes); stage them in
            // the warp's scratch and write whole 16-byte vectors — full sectors, which is what keeps NVLink packets
            // large when the planes live in a peer's memory, and 3 store instructions instead of 6 locally
            double *stg = S.blk;
            stg[3 * lane] = qx; stg[3 * lane + 1] = qy; stg[3 * lane + 2] = qz;
            stg[96 + lane] = e_out;
            reinterpret_cast<uint32_t *>(stg + 128)[lane] = ids;
            reinterpret_cast<uint8_t *>(stg + 144)[lane] = (uint8_t)nexcl;
            if (!RAW && a.bulk_out) {                            // (RAW instantiations are not launched with bulk stores)
                // TMA stores: four bulk copies per tile (768 + 256 + 128 + 32 bytes) issued by one lane; the copy engine
                // moves them while the warp goes on — no store instructions, whole tile records on the link
                fence_proxy_async();                             // my staging writes, visible to the async proxy
                __syncwarp();
                if (lane == 0) {
                    const uint32_t s0 = smem_u32(stg);
                    bulk_s2g(a.out_Q + (long long)tile * 96, s0, 768u);
                    bulk_s2g(a.out_err + (long long)tile * 32, s0 + 768u, 256u);
                    bulk_s2g(a.out_mask + (long long)tile * 32, s0 + 1024u, 128u);
                    bulk_s2g(a.out_nexcl + (long long)tile * 32, s0 + 1152u, 32u);
                    bulk_commit();
                }
                __syncwarp();
            } else {
            __syncwarp();
            const float4 *src = reinterpret_cast<const float4 *>(stg);
            float4 *dq = reinterpret_cast<float4 *>(a.out_Q + (long long)tile * 96);
            float4 *de = reinterpret_cast<float4 *>(a.out_err + (long long)tile * 32);
            float4 *dm = reinterpret_cast<float4 *>(a.out_mask + (long long)tile * 32);
            float4 *dn = reinterpret_cast<float4 *>(a.out_nexcl + (long long)tile * 32);
            dq[lane] = src[lane];
            if (lane < 16) dq[32 + lane] = src[32 + lane];
            else de[lane - 16] = src[48 + lane - 16];
            if (lane < 8) dm[lane] = src[64 + lane];
            else if (lane < 10) dn[lane - 8] = src[72 + lane - 8];
            __syncwarp();
            }
        } else if (active) {
            double *q = a.out_Q + u * 3;
            q[0] = qx; q[1] = qy; q[2] = qz;
            a.out_err[u] = e_out;
            a.out_nexcl[u] = (uint8_t)nexcl;
            a.out_mask[u] = ids;
        }
        // ---- per-tile statistics: warp-wide counts, lane 0 keeps the warp's totals in shared ----------
        if (STATS && a.stats != nullptr) {
            const uint32_t s_c = __reduce_add_sync(P2S_FULL, t_cands);
            const uint32_t s_m = __reduce_add_sync(P2S_FULL, t_cams);
            const uint32_t s_i = __reduce_add_sync(P2S_FULL, t_iters);
            const uint32_t s_s = __reduce_add_sync(P2S_FULL, t_solved);
            const uint32_t s_d = __reduce_add_sync(P2S_FULL, t_direct);
            const uint32_t s_b = __reduce_add_sync(P2S_FULL, t_blocks);
            const uint32_t s_a = __reduce_add_sync(P2S_FULL, t_adds);
            const uint32_t n_fail = __popc(__ballot_sync(P2S_FULL, failed && !wide));
            const uint32_t n_noev = __popc(__ballot_sync(P2S_FULL, active && !wide && last_level < 0));
            const uint32_t n_bthr = __popc(__ballot_sync(P2S_FULL, active && band_thr));
            const uint32_t n_barg = __popc(__ballot_sync(P2S_FULL, active && band_arg));
            uint32_t hist = 0;
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                const uint32_t n = __popc(__ballot_sync(P2S_FULL, active && !wide && last_level == l));
                if (lane == l) hist = n;
            }
            if (lane < 8) S.st32[lane] += hist;
            if (lane == 0) {
                S.st64[0] += s_c; S.st64[1] += s_m; S.st64[2] += s_i; S.st64[3] += s_s;
                S.st64[4] += s_d; S.st64[5] += s_b; S.st64[6] += s_a;
                S.st32[8] += n_fail; S.st32[9] += n_noev; S.st32[10] += n_bthr; S.st32[11] += n_barg;
            }
            if (active && !wide && last_level >= 8) atomicAdd(a.stats + P2S_STAT_LEVEL0 + last_level, 1ULL);
        }
        __syncwarp();
    }

    // ---- flush statistics: one atomic per counter per warp ---------------------------------------------
    if (STATS && a.stats != nullptr) {
        __syncwarp();
        if (lane < 8 && S.st32[lane]) atomicAdd(a.stats + P2S_STAT_LEVEL0 + lane, (unsigned long long)S.st32[lane]);
        if (lane == 8) {
            if (S.st64[0]) atomicAdd(a.stats + P2S_STAT_CANDIDATES, S.st64[0]);
            if (S.st64[1]) atomicAdd(a.stats + P2S_STAT_CAM_SOLVES, S.st64[1]);
            if (S.st64[2]) atomicAdd(a.stats + P2S_STAT_NEWTON_STEPS, S.st64[2]);
            if (S.st64[3]) atomicAdd(a.stats + P2S_STAT_SOLVED, S.st64[3]);
            if (S.st64[4]) atomicAdd(a.stats + P2S_STAT_DIRECT_CAMS, S.st64[4]);
            if (S.st64[5]) atomicAdd(a.stats + P2S_STAT_BLOCKS, S.st64[5]);
            if (S.st64[6]) atomicAdd(a.stats + P2S_STAT_ENTRY_ADDS, S.st64[6]);
            if (S.st32[8]) atomicAdd(a.stats + P2S_STAT_FAILED, (unsigned long long)S.st32[8]);
            if (S.st32[9]) atomicAdd(a.stats + P2S_STAT_NOT_EVALUATED, (unsigned long long)S.st32[9]);
            if (S.st32[10]) atomicAdd(a.stats + P2S_STAT_BAND_THRESHOLD, (unsigned long long)S.st32[10]);
            if (S.st32[11]) atomicAdd(a.stats + P2S_STAT_BAND_ARGMIN, (unsigned long long)S.st32[11]);
        }
    }

    if (!RAW && a.bulk_out) {                                   // every bulk store of this warp has been performed
        if (lane == 0) bulk_wait_all();
        __syncwarp();
    }
    // the kernel behind this one (wide_fixup_kernel) may be scheduled onto the SMs as the CTAs of this one retire
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- pooled variant: level-1 work of SEVERAL tiles shares its passes ---------------------------------------------------
// In triangulate_kernel a tile's pending units go through ceil(n / G) passes of G units per level; at 8 cameras ~9.3 of a
// tile's 32 units reach level 1 (G = 4), so the last pass of almost every tile is partly empty: 227 k of the launch's 447 k
// warp passes on cfg2 where 187 k would do.  Here the slab is a POOL of 32 unit SLOTS instead of the image of one tile:
//   * level 0 runs thread per unit straight from the raw planes the TMA copies landed (no transposition of the 71 % of the
//     units that end at level 0) and the tile's outputs leave as 16-byte vectors at once;
//   * a unit that needs level 1 is copied into a free slot (observations, level-0 normal matrix, validity masks, unit index);
//   * level-1 passes take G OCCUPIED SLOTS whatever tile they came from and only run full — up to G - 1 slots wait for the
//     next tile — except when the stream ends or the pool overflows; the owner lane of a slot (lane == slot) then walks the
//     deeper levels exactly like the tile kernel does and overwrites the unit's four output entries (the tile store of the
//     same warp precedes it in program order, with a __syncwarp between).
// Same arithmetic per candidate and the same arg-min rules, so the results are the tile kernel's bit for bit
// (tests/test_gpu_triangulate.py::test_pooled_kernel_equals_tile_kernel).  Lean secular kernel on raw planes with a
// compile-time camera count of 4 or 8 only (what the TMA staging serves); outputs in device memory (the scattered 8-byte
// overwrites would make poor PCIe / NVLink packets, so the zero-copy host path and the push path keep the tile kernel).
#ifndef P2S_POOL_HIGH
#define P2S_POOL_HIGH 18
#endif
template <int CMAX>
struct alignas(16) PoolSlab {
    float raw[3 * 32 * CMAX];     // the tile's raw planes x | y | likelihood, [32 units][C] floats each (TMA destination)
    float2 xy[CMAX][32];          // by SLOT: observations of the pooled units ...
    float w[CMAX][32];            // ... and likelihoods (invalid cameras hold zeros)
    unsigned long long mbar;
    unsigned long long pad_;
    double blk[32 * 10 + 32];     // camera blocks of the current group pass / output staging of a tile
    double2 gxy[32];
    double m0[10][32];            // by slot: level-0 normal matrix
    unsigned long long r_key[32];
    double r_qx[32], r_qy[32], r_qz[32];
    uint32_t r_nan[32], r_flags[32], nan0[32], inv0[32], plist[32];
    uint32_t uid[32];             // by slot: unit index
};
static_assert(sizeof(PoolSlab<8>) % 16 == 0 && offsetof(PoolSlab<8>, blk) % 16 == 0 && offsetof(PoolSlab<8>, xy) % 16 == 0, "slab alignment");
static_assert(sizeof(PoolSlab<4>) % 16 == 0 && offsetof(PoolSlab<4>, blk) % 16 == 0, "slab alignment");

template <int CMAX>
__global__ void __launch_bounds__(128, P2S_TRI_MIN_BLOCKS) triangulate_pool_kernel(const CamParams<CMAX> cams, const TriArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int C = CMAX;
    constexpr bool DISTORT = false, STATS = false;
    const LensSet<1> lens = {};
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double *sP = reinterpret_cast<double *>(smem_raw);
    PoolSlab<CMAX> &S = reinterpret_cast<PoolSlab<CMAX> *>(smem_raw + CMAX * 12 * sizeof(double))[warp];
    constexpr uint32_t cmask = (1u << C) - 1u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const long long n_tiles = (a.n_units + 31) >> 5;
    const float nanf_ = __int_as_float(0x7fc00000);

    for (int i = threadIdx.x; i < CMAX * 12; i += blockDim.x) sP[i] = (&cams.P[0][0])[i];
    __syncthreads();

    unsigned int t1 = 0, t2 = 0;
    if (lane == 0) { t1 = atomicAdd(a.tile_counter, 1u); t2 = atomicAdd(a.tile_counter, 1u); }
    const uint32_t bar = smem_u32(&S.mbar);
    uint32_t phase = 0;
    auto issue_tile = [&](unsigned int t) {                    // lane 0 only
        const long long e0t = (long long)t * 32 * C;
        const long long left = a.n_units * C - e0t;
        const uint32_t bytes = (uint32_t)(left < 32LL * C ? left : 32LL * C) * 4u;
        mbar_expect_tx(bar, 3u * bytes);
        bulk_g2s(smem_u32(S.raw), a.px + e0t, bytes, bar);
        bulk_g2s(smem_u32(S.raw + 32 * CMAX), a.py + e0t, bytes, bar);
        bulk_g2s(smem_u32(S.raw + 64 * CMAX), a.pl + e0t, bytes, bar);
    };
    if (lane == 0) {
        mbar_init(bar, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        if ((long long)t1 < n_tiles) issue_tile(t1);
    }
    __syncwarp();

    const int G1 = 32 >> a.lw[1];                               // slots per level-1 pass
    // the pool is a ring: slots are handed out in increasing order (mod 32) and the passes take them in the same order
    uint32_t q_head = 0;                                        // next slot to hand out (the same value in every lane)
    int q_count = 0;                                            // occupied slots: q_head - q_count ... q_head - 1 (mod 32)
    uint32_t todo = 0;                                          // lanes of the current tile still waiting for a slot
    bool have_tile = true, raw_held = false;
    unsigned int nt = 0;
    Sym4 M0;                                                    // my unit's level-0 matrix until it has a slot
    sym4_zero(M0);
    uint32_t nan0 = 0, inv0 = 0;
    long long u = 0;

    for (;;) {
        if (todo == 0u && have_tile) {
            const unsigned int tile = __shfl_sync(P2S_FULL, t1, 0);
            if ((long long)tile >= n_tiles) {
                have_tile = false;
            } else {
                t1 = t2;
                if (lane == 0) t2 = atomicAdd(a.tile_counter, 1u);
                nt = __shfl_sync(P2S_FULL, t1, 0);
                u = (long long)tile * 32 + lane;
                const bool active = u < a.n_units;
                mbar_wait(bar, phase);                          // this tile's planes have landed in S.raw
                raw_held = true;

                // ---- level 0: thread per unit, straight from the raw planes --------------------------------------
                float fx[CMAX], fy[CMAX], fl[CMAX];
                load_row<CMAX>(S.raw, lane, fx);
                load_row<CMAX>(S.raw + 32 * CMAX, lane, fy);
                load_row<CMAX>(S.raw + 64 * CMAX, lane, fl);
                nan0 = 0; inv0 = 0;
                float wlo = __int_as_float(0x7f800000), whi = 0.f;
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if (!active) { fx[c] = 0.f; fy[c] = 0.f; fl[c] = nanf_; }     // beyond the last unit: stale bytes
                    if (a.gate && fl[c] < a.lik_thr_f) { fx[c] = fy[c] = fl[c] = nanf_; }
                    const float lz = fl[c];
                    const bool isn = lz != lz;
                    const bool inv = isn || lz == 0.f;
                    nan0 |= (uint32_t)isn << c;
                    inv0 |= (uint32_t)inv << c;
                    const float la = fabsf(lz);
                    whi = fmaxf(whi, la);
                    wlo = fminf(wlo, la > 0.f ? la : wlo);
                    if (inv) { fx[c] = 0.f; fy[c] = 0.f; fl[c] = 0.f; }
                }
#ifdef P2S_NO_WIDE
                const bool wide = false;
#else
                const bool wide = whi > P2S_WIDE_SPREAD * wlo;
#endif
                if (__ballot_sync(P2S_FULL, wide && active) != 0u && lane == 0) atomicAdd(a.tile_counter + 2, 1u);

                double err_min = inf64();
                double qx = nan64(), qy = qx, qz = qx;
                uint32_t ids = cmask, nexcl = (uint32_t)C;
                int last_level = -1;
                const int ninv0 = __popc(inv0);
                if (active && !wide && (C >= a.min_cams) && !(min(C, ninv0) > C - a.min_cams)) {
                    const uint32_t valid = cmask & ~inv0;
                    const int m = C - ninv0;
                    if (m >= 2) {
                        sym4_zero(M0);
#pragma unroll
                        for (int c = 0; c < CMAX; ++c) accumulate_camera(M0, cams.P[c], (double)fx[c], (double)fy[c], (double)fl[c]);
                        smallest_eigvec_secular(M0, qx, qy, qz);
                        double sum = 0.0;
#pragma unroll
                        for (int c = 0; c < CMAX; ++c) {
                            const double dist = reproj_distance(cams.P[c], qx, qy, qz, (double)fx[c], (double)fy[c]);
                            if ((valid >> c) & 1u) sum += dist;
                        }
                        const double e = sum * a.rinv[m];
                        err_min = (e != e) ? inf64() : e;
                    }
                    ids = nan0; nexcl = (uint32_t)ninv0; last_level = 0;
                }
                const bool pend1 = active && !wide && last_level == 0 && (err_min > a.thr) && (C - 1 >= a.min_cams) &&
                                   !(min(C, ninv0 + 1) > C - a.min_cams);

                // ---- the tile's outputs (units that go on are overwritten by their slot's owner later) --------------
                const bool failed = active && (err_min > a.thr);
                double e_out = err_min;
                if (failed) { e_out = nan64(); qx = qy = qz = nan64(); }
                if (a.vec_out && (long long)tile * 32 + 32 <= a.n_units) {
                    double *stg = S.blk;
                    stg[3 * lane] = qx; stg[3 * lane + 1] = qy; stg[3 * lane + 2] = qz;
                    stg[96 + lane] = e_out;
                    reinterpret_cast<uint32_t *>(stg + 128)[lane] = ids;
                    reinterpret_cast<uint8_t *>(stg + 144)[lane] = (uint8_t)nexcl;
                    __syncwarp();
                    const float4 *src = reinterpret_cast<const float4 *>(stg);
                    float4 *dq = reinterpret_cast<float4 *>(a.out_Q + (long long)tile * 96);
                    float4 *de = reinterpret_cast<float4 *>(a.out_err + (long long)tile * 32);
                    float4 *dm = reinterpret_cast<float4 *>(a.out_mask + (long long)tile * 32);
                    float4 *dn = reinterpret_cast<float4 *>(a.out_nexcl + (long long)tile * 32);
                    dq[lane] = src[lane];
                    if (lane < 16) dq[32 + lane] = src[32 + lane];
                    else de[lane - 16] = src[48 + lane - 16];
                    if (lane < 8) dm[lane] = src[64 + lane];
                    else if (lane < 10) dn[lane - 8] = src[72 + lane - 8];
                    __syncwarp();
                } else if (active) {
                    double *q = a.out_Q + u * 3;
                    q[0] = qx; q[1] = qy; q[2] = qz;
                    a.out_err[u] = e_out;
                    a.out_nexcl[u] = (uint8_t)nexcl;
                    a.out_mask[u] = ids;
                }
                todo = __ballot_sync(P2S_FULL, pend1);
            }
        }

        // ---- pending units of the tile move into free slots ---------------------------------------------------------
        if (todo != 0u) {
            const int nfree = 32 - q_count;
            const bool mine = (todo >> lane) & 1u;
            const int rank = __popc(todo & lt_mask);
            const bool take = mine && rank < nfree;
            if (take) {
                const int slot = (int)((q_head + (uint32_t)rank) & 31u);
                // observations again from the raw planes (still the tile's: the buffer goes back to the copy engine below)
                float fx[CMAX], fy[CMAX], fl[CMAX];
                load_row<CMAX>(S.raw, lane, fx);
                load_row<CMAX>(S.raw + 32 * CMAX, lane, fy);
                load_row<CMAX>(S.raw + 64 * CMAX, lane, fl);
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    const bool inv = (inv0 >> c) & 1u;
                    S.xy[c][slot] = make_float2(inv ? 0.f : fx[c], inv ? 0.f : fy[c]);
                    S.w[c][slot] = inv ? 0.f : fl[c];
                }
                if (!(M0.m33 == M0.m33)) {
                    // level 0 came out +inf through a NaN coordinate under a valid likelihood (a NaN x or y makes m33 = sum of
                    // (w x)^2-free terms NaN as well): the deeper levels need the sum of the CLEAN cameras
                    sym4_zero(M0);
#pragma unroll 1
                    for (int c = 0; c < C; ++c) {
                        const float2 o = S.xy[c][slot];
                        if ((o.x == o.x) && (o.y == o.y))
                            accumulate_camera(M0, sP + c * 12, (double)o.x, (double)o.y, (double)S.w[c][slot]);
                    }
                }
                S.m0[0][slot] = M0.m00; S.m0[1][slot] = M0.m01; S.m0[2][slot] = M0.m02; S.m0[3][slot] = M0.m03; S.m0[4][slot] = M0.m11;
                S.m0[5][slot] = M0.m12; S.m0[6][slot] = M0.m13; S.m0[7][slot] = M0.m22; S.m0[8][slot] = M0.m23; S.m0[9][slot] = M0.m33;
                S.nan0[slot] = nan0;
                S.inv0[slot] = inv0;
                S.uid[slot] = (uint32_t)u;
            }
            const uint32_t taken = __ballot_sync(P2S_FULL, take);
            q_head += (uint32_t)__popc(taken);
            q_count += __popc(taken);
            todo &= ~taken;
            __syncwarp();
        }
        if (raw_held && todo == 0u) {
            // every lane is done with S.raw: hand the buffer back to the copy engine for the next tile
            __syncwarp();
            raw_held = false;
            phase ^= 1u;
            if (lane == 0 && (long long)nt < n_tiles) { fence_proxy_async(); issue_tile(nt); }
        }

        // ---- passes over occupied slots: full groups of G1, everything when the stream ends or the pool overflowed ----
        // The passes start when the pool is well filled (P2S_POOL_HIGH slots: two or three tiles' worth at 8 cameras) and then
        // run back to back until less than one group is left: the level code stays in the instruction cache for 4-6 pass
        // sets in a row instead of alternating with the level-0 code after every tile.
        const bool drain = (todo != 0u) || !have_tile;
        const bool running = drain || q_count >= P2S_POOL_HIGH;
        while (running && q_count != 0 && (q_count >= G1 || drain)) {
            // every full group at once (a partial one only when draining): ONE level loop over up to 32 pooled units, so the
            // per-level set-up is shared by 4-6 level-1 passes
            const int npass = drain ? q_count : (q_count / G1) * G1;
            const uint32_t low = (npass >= 32) ? 0xffffffffu : ((1u << npass) - 1u);
            const uint32_t pm = __funnelshift_l(low, low, (q_head - (uint32_t)q_count) & 31u);   // the npass oldest slots
            q_count -= npass;
            const bool owner = (pm >> lane) & 1u;
            const uint32_t o_nan0 = S.nan0[lane], o_inv0 = S.inv0[lane];
            const int o_ninv0 = __popc(o_inv0);
            double err_min = inf64();
            double qx = nan64(), qy = qx, qz = qx;
            uint32_t ids = o_nan0, nexcl = (uint32_t)o_ninv0;
            int last_level = 0;

            for (int k = 1; k < C; ++k) {
                const bool pend = owner && last_level == k - 1 && (err_min > a.thr) && (C - k >= a.min_cams) &&
                                  !(min(C, o_ninv0 + k) > C - a.min_cams);
                const uint32_t pmask = __ballot_sync(P2S_FULL, pend);
                if (pmask == 0) break;
                const int npend = __popc(pmask);
                if (pend) {
                    S.plist[__popc(pmask & lt_mask)] = (uint32_t)lane;
                    S.r_key[lane] = P2S_KEY_EMPTY;
                }
                const uint32_t ncand = a.ncand[k];
                const int lw = a.lw[k];
                const int W = 1 << lw;
                const int G = 32 >> lw;
                const int grp = lane >> lw, sub = lane & (W - 1);
                const uint32_t gmask = (W >= 32) ? P2S_FULL : (((1u << W) - 1u) << (grp << lw));
                const bool tabled = true;                               // <= 8 cameras: every level is in the table
                const uint32_t *table = a.cand_masks + a.level_off[tabled ? k : 0];
                double *gblk = S.blk + grp * (C * 10 + 2);
                const bool subtract = true;                             // like the RAW tile kernel
                __syncwarp();

                for (int base = 0; base < npend; base += G) {
                    const int idx = base + grp;
                    const bool on = idx < npend;
                    const int ul = on ? (int)S.plist[idx] : 0;
                    const uint32_t u_nan0 = S.nan0[ul], u_inv0 = S.inv0[ul];
                    __syncwarp();
                    if (on && sub < C) {
                        const float2 o = S.xy[sub][ul];
                        const float ow = S.w[sub][ul];
                        const bool clean = (o.x == o.x) && (o.y == o.y);
                        const double ox = (double)o.x, oy = (double)o.y;
                        S.gxy[grp * C + sub] = make_double2(ox, oy);
                        double b[10];
                        camera_block(sP + sub * 12, clean ? ox : 0.0, clean ? oy : 0.0, (double)(clean ? ow : 0.f), b);
                        double2 *dst = reinterpret_cast<double2 *>(gblk + sub * 10);
#pragma unroll
                        for (int e = 0; e < 5; ++e) dst[e] = make_double2(b[2 * e], b[2 * e + 1]);
                    }
                    __syncwarp();

                    unsigned long long bkey = P2S_KEY_EMPTY;
                    uint32_t bcand = 0xffffffffu, bcm = 0;
                    double bqx, bqy, bqz;
                    for (uint32_t cand = (uint32_t)sub; on && cand < ncand; cand += (uint32_t)W) {
                        const uint32_t cm = tabled ? __ldg(table + cand) : unrank_subset(C, k, cand);
                        const uint32_t valid = cmask & ~(u_inv0 | cm);
                        const int m = __popc(valid);
                        double cqx = nan64(), cqy = cqx, cqz = cqx;
                        double e = inf64();
                        if (m >= 2) {
                            Sym4 M;
                            uint32_t bits;
                            double sgn;
                            if (subtract) {
                                M.m00 = S.m0[0][ul]; M.m01 = S.m0[1][ul]; M.m02 = S.m0[2][ul]; M.m03 = S.m0[3][ul]; M.m11 = S.m0[4][ul];
                                M.m12 = S.m0[5][ul]; M.m13 = S.m0[6][ul]; M.m22 = S.m0[7][ul]; M.m23 = S.m0[8][ul]; M.m33 = S.m0[9][ul];
                                bits = cm & ~u_inv0 & cmask;
                                sgn = -1.0;
                            } else {
                                sym4_zero(M);
                                bits = valid;
                                sgn = 1.0;
                            }
                            while (bits) {
                                const int c = __ffs(bits) - 1;
                                bits &= bits - 1;
                                const double2 *src = reinterpret_cast<const double2 *>(gblk + c * 10);
                                const double2 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4];
                                M.m00 = fma(sgn, v0.x, M.m00); M.m01 = fma(sgn, v0.y, M.m01); M.m02 = fma(sgn, v1.x, M.m02);
                                M.m03 = fma(sgn, v1.y, M.m03); M.m11 = fma(sgn, v2.x, M.m11); M.m12 = fma(sgn, v2.y, M.m12);
                                M.m13 = fma(sgn, v3.x, M.m13); M.m22 = fma(sgn, v3.y, M.m22); M.m23 = fma(sgn, v4.x, M.m23);
                                M.m33 = fma(sgn, v4.y, M.m33);
                            }
                            smallest_eigvec_secular(M, cqx, cqy, cqz);
                            e = mean_reproj_error<CMAX, DISTORT, true>(cams, lens, S.xy, S.gxy + grp * C, ul, valid, a.rinv[m], cqx, cqy, cqz, sP);
                        }
                        const unsigned long long key = err_key_inf(e);
                        if (key < bkey) {
                            bkey = key; bcand = cand; bcm = cm;
                            bqx = cqx; bqy = cqy; bqz = cqz;
                        }
                    }
                    const uint32_t hi = (uint32_t)(bkey >> 32), lo = (uint32_t)bkey;
                    const uint32_t mh = group_min(hi, W, gmask);
                    const bool have = bcand != 0xffffffffu;
                    uint32_t holders = __ballot_sync(P2S_FULL, have && hi == mh) & gmask;
                    bool winner = holders != 0u && lane == __ffs(holders) - 1;
                    if (__any_sync(P2S_FULL, __popc(holders) > 1)) {
                        const uint32_t ml = group_min(hi == mh ? lo : 0xffffffffu, W, gmask);
                        const bool is_min = have && hi == mh && lo == ml;
                        const uint32_t mc = group_min(is_min ? bcand : 0xffffffffu, W, gmask);
                        winner = is_min && bcand == mc;
                    }
                    if (winner) {
                        S.r_key[ul] = bkey;
                        S.r_qx[ul] = bqx; S.r_qy[ul] = bqy; S.r_qz[ul] = bqz;
                        S.r_nan[ul] = u_nan0 | bcm;
                        S.r_flags[ul] = (uint32_t)__popc(u_inv0 | bcm);
                    }
                }
                __syncwarp();
                if (pend) {
                    err_min = key_err(S.r_key[lane]);
                    qx = S.r_qx[lane]; qy = S.r_qy[lane]; qz = S.r_qz[lane];
                    ids = S.r_nan[lane];
                    nexcl = S.r_flags[lane] & 0xffu;
                    last_level = k;
                }
                __syncwarp();
            }
            // ---- finalise (:588-602): the slot's owner overwrites the unit's outputs ---------------------------------
            if (owner) {
                const long long uu = (long long)S.uid[lane];
                double e_out = err_min;
                if (err_min > a.thr) { e_out = nan64(); qx = qy = qz = nan64(); }
                double *q = a.out_Q + uu * 3;
                q[0] = qx; q[1] = qy; q[2] = qz;
                a.out_err[uu] = e_out;
                a.out_nexcl[uu] = (uint8_t)nexcl;
                a.out_mask[uu] = ids;
            }
            __syncwarp();
        }

        if (todo != 0u) {
            // pool overflow: the waiting lanes' matrices did not survive the passes in registers — rebuild them from the raw
            // planes (rolled loop, off the common path); poisoned cameras are handled when the lane takes its slot
            sym4_zero(M0);                                       // (every lane: nothing of M0 is live across the passes)
#pragma unroll 1
            for (int c = 0; c < C; ++c) {
                const float x = S.raw[lane * C + c], y = S.raw[32 * CMAX + lane * C + c];
                if (((todo & ~lt_mask) & (1u << lane)) != 0u && !((inv0 >> c) & 1u))
                    accumulate_camera(M0, sP + c * 12, (double)x, (double)y, (double)S.raw[64 * CMAX + lane * C + c]);
            }
            continue;
        }
        if (!have_tile) break;
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- deep levels: a cluster of 512-thread CTAs per parked unit -------------------------------------------------------
// triangulate_kernel walks a level's candidates with the 32 lanes of ONE warp.  That is the right shape while levels hold
// tens or hundreds of candidates, but C(16, 5..8) = 4 368 .. 12 870 and C(32, 4) = 35 960: a unit that goes that deep keeps
// its warp for hundreds of rounds, and the SM that drew it is still busy when every other SM has run out of tiles (cfg3
// shard, ncu: SMs active 73 % of the kernel's duration, the longest 9.5 M cycles against a mean of 6.9 M).  Such units are
// rare (6e-4 of cfg3's), so the main kernel parks them — (unit, level) in a list, at the first level with >= deep_min
// candidates — and this kernel, launched behind it, re-stages a parked unit and walks the rest of its search with 512
// threads per CTA (P2S_DEEP_CLUSTER CTAs per unit, below): candidates strided over the threads, the (error key, candidate index) arg-min reduced per warp and then across
// the 16 warps, the level rules of triangulation.py:408-505 applied by every thread on the published result.
// The arithmetic is the main kernel's, statement for statement (level-0 matrix accumulated over the cameras in ascending
// order, camera blocks from camera_block(), M = M_all - excluded blocks or the sum of the kept ones by the same rule,
// the same solver and distance functions, the mean from the same 1/m table), so parking changes no output bit
// (tests/test_gpu_triangulate.py::test_deep_levels_*).  STATS / Jacobi / lens-model launches never park.
struct DeepArgs {
    const float4 *obs;
    const float *px, *py, *pl;
    float lik_thr_f;
    int gate;
    long long n_units;
    int n_cams, min_cams;
    double thr;
    const uint32_t *cand_masks;
    uint32_t level_off[P2S_MAX_CAMS + 2];
    int max_table_level;
    int always_downdate;              // the main kernel's `subtract` rule: exact-count kernels always downdate
    double rinv[P2S_MAX_CAMS + 1];
    double *out_Q, *out_err;
    uint8_t *out_nexcl;
    uint32_t *out_mask;
    const unsigned long long *list;   // (unit << 8) | level
    unsigned int cap;
    const unsigned int *count;        // tile_counter + 3: units the main kernel tried to park (may exceed cap)
};

constexpr int kDeepThreads = 512;
constexpr int kDeepClusterMax = 8;
// CTAs of the deep-level kernel per SM's worth of grid.  A parked unit costs between ~9 rounds (one level of 4 368 subsets)
// and several hundred, and one CTA is resident per SM: with the list strided over 2 x SMs CTAs every CTA walked ~7 units
// of cfg3's ~2 000 and the kernel waited for the unluckiest sum (ncu: SMs active 46 % of its 297 us).  With 16 x SMs CTAs a
// CTA walks one or two units and the hardware's block scheduler does the balancing; a CTA beyond the parked count exits
// on its first comparison.
#ifndef P2S_DEEP_GRID_MULT
#define P2S_DEEP_GRID_MULT 16
#endif
// CTAs per parked unit (thread-block cluster; 1 = one CTA per unit).  The kernel lasted as long as its longest unit
// (levels 5-7 of a 16-camera rig: 23 816 subsets = 47 rounds of 512 threads, ~290 us of the kernel's 297; the grid above
// changed nothing), so a unit gets the SMs of a cluster: its CTAs each stage the unit, split every level's subsets (warp
// interleaved), exchange their level winners through distributed shared memory (each CTA stores its winner into every
// CTA's slab, one cluster barrier per level) and all apply the level rules to the same merged winner.  (key, subset index)
// is a total order, so the winner does not depend on the split.  How many: a round is latency-bound (four warps per
// scheduler on a dependent DFMA chain), it takes about as long with two warps as with sixteen, so a level that fills
// fewer rounds than CTAs x 512 threads costs SM-time in proportion to the cluster size.  Same-box A/B on the cfg3 shard
// (profiles/r3v_kernel_ab_deep.jsonl), 1 / 2 / 4 / 8 CTAs: 3.954 / 3.885 / 3.909 / 4.029 ms, identical checksums.
#ifndef P2S_DEEP_CLUSTER
#define P2S_DEEP_CLUSTER 2
#endif

struct DeepSlab {
    double sP[P2S_MAX_CAMS * 12];
    double blk[P2S_MAX_CAMS * 10];
    double2 gxy[P2S_MAX_CAMS];
    double m0[10];
    float2 xy[P2S_MAX_CAMS];
    float w[P2S_MAX_CAMS];
    unsigned long long wkey[kDeepThreads / 32];
    double wq[kDeepThreads / 32][3];
    uint32_t wcand[kDeepThreads / 32], wcm[kDeepThreads / 32];
    unsigned long long rkey;
    double rq[3];
    uint32_t rcm, nan0, inv0;
    // cluster form (KC > 1): every CTA of the cluster stores its level winner into slot [phase][its rank] of EVERY CTA's
    // slab (distributed shared memory), phase alternating per level so that a slot is rewritten two cluster barriers later
    unsigned long long ckey[2][kDeepClusterMax];
    double cq[2][kDeepClusterMax][3];
    uint32_t ccand[2][kDeepClusterMax], ccm[2][kDeepClusterMax];
};

// distributed-shared-memory stores / cluster barrier of the cluster form
__device__ __forceinline__ uint32_t dsmem_addr(const void *own_smem, uint32_t cta_rank) {
    uint32_t local = (uint32_t)__cvta_generic_to_shared(own_smem), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cta_rank));
    return remote;
}
__device__ __forceinline__ void dsmem_st_u64(uint32_t addr, unsigned long long v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st_f64(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void dsmem_st_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// (key, candidate) arg-min over the lanes in `mask`: smallest 64-bit key, then the smallest candidate index among its
// holders (np.nanargmin's first index).  Returns true on the one winning lane; false everywhere when no lane holds a candidate.
__device__ __forceinline__ bool deep_argmin(uint32_t mask, unsigned long long key, uint32_t cand) {
    const uint32_t hi = (uint32_t)(key >> 32), lo = (uint32_t)key;
    const uint32_t mh = __reduce_min_sync(mask, hi);
    const uint32_t ml = __reduce_min_sync(mask, hi == mh ? lo : 0xffffffffu);
    const bool is_min = cand != 0xffffffffu && hi == mh && lo == ml;
    const uint32_t mc = __reduce_min_sync(mask, is_min ? cand : 0xffffffffu);
    return is_min && cand == mc;
}

template <int KC>
__global__ void __launch_bounds__(kDeepThreads, 1) deep_search_kernel(const CamParams<P2S_MAX_CAMS> cams, const DeepArgs a) {
    static_assert(KC >= 1 && KC <= kDeepClusterMax, "cluster size");
    __shared__ DeepSlab S;
    // programmatic dependent launch: nothing the search kernel wrote is read before it has completed
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned int n = *reinterpret_cast<const volatile unsigned int *>(a.count);
    if (n > a.cap) n = a.cap;
    // KC CTAs (one thread-block cluster) share a parked unit; the whole cluster leaves together
    const unsigned int team = blockIdx.x / (unsigned)KC, n_teams = gridDim.x / (unsigned)KC;
    const uint32_t crank = (KC > 1) ? cluster_cta_rank() : 0u;
    if (team >= n) return;                                     // (n == 0 included) more teams than parked units: nothing to stage
    uint32_t phase = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = a.n_cams;
    const uint32_t cmask = (C >= 32) ? 0xffffffffu : ((1u << C) - 1u);
    for (int i = tid; i < C * 12; i += kDeepThreads) S.sP[i] = (&cams.P[0][0])[i];
    for (unsigned int j = team; j < n; j += n_teams) {
        __syncthreads();                                       // sP is there; the previous unit's slab has been read
        const unsigned long long rec = a.list[j];
        const long long u = (long long)(rec >> 8);
        int k = (int)(rec & 0xffu);
        // ---- stage the unit like the main kernel stages a tile: gate, validity masks, zeros for invalid cameras,
        // camera blocks (a poisoned camera — valid, NaN coordinate — gets a zero block and keeps its NaN in gxy)
        if (warp == 0) {
            const float nanf_ = __int_as_float(0x7fc00000);
            float fx = 0.f, fy = 0.f, fl = nanf_;
            if (lane < C) {
                if (a.px == nullptr) {
                    const float4 o = __ldg(a.obs + (long long)lane * a.n_units + u);
                    fx = o.x; fy = o.y; fl = o.z;
                } else {
                    fx = a.px[u * C + lane]; fy = a.py[u * C + lane]; fl = a.pl[u * C + lane];
                    if (a.gate && fl < a.lik_thr_f) { fx = fy = fl = nanf_; }
                }
            }
            const bool isn = fl != fl;
            const bool inv = isn || fl == 0.f;
            const uint32_t nan0 = __ballot_sync(P2S_FULL, isn) & cmask;
            const uint32_t inv0 = __ballot_sync(P2S_FULL, inv) & cmask;
            if (inv) { fx = 0.f; fy = 0.f; fl = 0.f; }
            if (lane < C) {
                const bool clean = (fx == fx) && (fy == fy);
                const double ox = (double)fx, oy = (double)fy;
                S.xy[lane] = make_float2(fx, fy);
                S.w[lane] = fl;
                S.gxy[lane] = make_double2(ox, oy);
                double b[10];
                camera_block(S.sP + lane * 12, clean ? ox : 0.0, clean ? oy : 0.0, (double)(clean ? fl : 0.f), b);
#pragma unroll
                for (int e = 0; e < 10; ++e) S.blk[lane * 10 + e] = b[e];
            }
            __syncwarp();
            if (lane == 0) {
                // level 0's normal matrix: accumulate_direct(), then rebuild_without_poisoned() when it applies
                S.nan0 = nan0; S.inv0 = inv0;
                const uint32_t valid0 = cmask & ~inv0;
                Sym4 M;
                sym4_zero(M);
                bool poisoned = false;
#pragma unroll 1
                for (int c = 0; c < C; ++c) {
                    const float2 o = S.xy[c];
                    accumulate_camera(M, S.sP + c * 12, (double)o.x, (double)o.y, (double)S.w[c]);
                    poisoned |= ((valid0 >> c) & 1u) && !((o.x == o.x) && (o.y == o.y));
                }
                if (poisoned) {
                    sym4_zero(M);
#pragma unroll 1
                    for (int c = 0; c < C; ++c) {
                        const float2 o = S.xy[c];
                        if (((valid0 >> c) & 1u) && (o.x == o.x) && (o.y == o.y))
                            accumulate_camera(M, S.sP + c * 12, (double)o.x, (double)o.y, (double)S.w[c]);
                    }
                }
                S.m0[0] = M.m00; S.m0[1] = M.m01; S.m0[2] = M.m02; S.m0[3] = M.m03; S.m0[4] = M.m11;
                S.m0[5] = M.m12; S.m0[6] = M.m13; S.m0[7] = M.m22; S.m0[8] = M.m23; S.m0[9] = M.m33;
            }
        }
        __syncthreads();
        const uint32_t u_nan0 = S.nan0, u_inv0 = S.inv0;
        const int ninv0 = __popc(u_inv0);
        double err_min = inf64();
        // ---- levels k, k + 1, ...: the parked level is known to be evaluated (the main kernel's `pend`)
#pragma unroll 1
        for (;; ++k) {
            const uint32_t ncand = binom_u32(C, k);
            const bool tabled = k <= a.max_table_level;
            const uint32_t *table = a.cand_masks + a.level_off[tabled ? k : 0];
            const bool subtract = a.always_downdate || 2 * k <= C;
            unsigned long long bkey = P2S_KEY_EMPTY;
            uint32_t bcand = 0xffffffffu, bcm = 0;
            double bqx = nan64(), bqy = bqx, bqz = bqx;
            // KC == 1: candidates strided over the threads.  Cluster: strided over the KC * 512 threads with the CTAs
            // interleaved at warp granularity, so a level's last, partial round is spread over all the SMs of the cluster
            const uint32_t cand0 = (KC > 1) ? (uint32_t)((warp * KC + (int)crank) * 32 + lane) : (uint32_t)tid;
#pragma unroll 1
            for (uint32_t cand = cand0; cand < ncand; cand += (uint32_t)(kDeepThreads * KC)) {
                const uint32_t cm = tabled ? __ldg(table + cand) : unrank_subset(C, k, cand);
                const uint32_t valid = cmask & ~(u_inv0 | cm);
                const int m = __popc(valid);
                double cqx = nan64(), cqy = cqx, cqz = cqx;
                double e = inf64();
                if (m >= 2) {
                    Sym4 M;
                    uint32_t bits;
                    double sgn;
                    if (subtract) {
                        M.m00 = S.m0[0]; M.m01 = S.m0[1]; M.m02 = S.m0[2]; M.m03 = S.m0[3]; M.m11 = S.m0[4];
                        M.m12 = S.m0[5]; M.m13 = S.m0[6]; M.m22 = S.m0[7]; M.m23 = S.m0[8]; M.m33 = S.m0[9];
                        bits = cm & ~u_inv0 & cmask;
                        sgn = -1.0;
                    } else {
                        sym4_zero(M);
                        bits = valid;
                        sgn = 1.0;
                    }
                    while (bits) {                            // ascending camera order
                        const int c = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const double *v = S.blk + c * 10;
                        M.m00 = fma(sgn, v[0], M.m00); M.m01 = fma(sgn, v[1], M.m01); M.m02 = fma(sgn, v[2], M.m02);
                        M.m03 = fma(sgn, v[3], M.m03); M.m11 = fma(sgn, v[4], M.m11); M.m12 = fma(sgn, v[5], M.m12);
                        M.m13 = fma(sgn, v[6], M.m13); M.m22 = fma(sgn, v[7], M.m22); M.m23 = fma(sgn, v[8], M.m23);
                        M.m33 = fma(sgn, v[9], M.m33);
                    }
                    smallest_eigvec_secular(M, cqx, cqy, cqz);
                    double sum = 0.0;
#pragma unroll 4
                    for (int c = 0; c < C; ++c) {
                        const double2 o = S.gxy[c];
                        const double dist = reproj_distance(S.sP + c * 12, cqx, cqy, cqz, o.x, o.y);
                        if ((valid >> c) & 1u) sum += dist;
                    }
                    e = sum * a.rinv[m];
                }
                const unsigned long long key = err_key_inf(e);
                if (key < bkey) {                               // ascending cand per thread: strict < keeps the first
                    bkey = key; bcand = cand; bcm = cm;
                    bqx = cqx; bqy = cqy; bqz = cqz;
                }
            }
            const uint32_t holders = __ballot_sync(P2S_FULL, bcand != 0xffffffffu);
            if (deep_argmin(P2S_FULL, bkey, bcand)) {
                S.wkey[warp] = bkey; S.wcand[warp] = bcand; S.wcm[warp] = bcm;
                S.wq[warp][0] = bqx; S.wq[warp][1] = bqy; S.wq[warp][2] = bqz;
            }
            if (holders == 0u && lane == 0) { S.wkey[warp] = P2S_KEY_EMPTY; S.wcand[warp] = 0xffffffffu; }   // fewer candidates than threads
            __syncthreads();
            double win_qx, win_qy, win_qz;
            uint32_t win_cm;
            if constexpr (KC == 1) {
                if (warp == 0) {
                    const bool in = lane < kDeepThreads / 32;
                    const unsigned long long wk = in ? S.wkey[lane] : P2S_KEY_EMPTY;
                    const uint32_t wc = in ? S.wcand[lane] : 0xffffffffu;
                    if (deep_argmin(P2S_FULL, wk, wc)) {
                        S.rkey = wk; S.rcm = S.wcm[lane];
                        S.rq[0] = S.wq[lane][0]; S.rq[1] = S.wq[lane][1]; S.rq[2] = S.wq[lane][2];
                    }
                }
                __syncthreads();
                err_min = key_err(S.rkey);
                win_qx = S.rq[0]; win_qy = S.rq[1]; win_qz = S.rq[2]; win_cm = S.rcm;
            } else {
                if (warp == 0) {
                    const bool in = lane < kDeepThreads / 32;
                    const unsigned long long wk = in ? S.wkey[lane] : P2S_KEY_EMPTY;
                    const uint32_t wc = in ? S.wcand[lane] : 0xffffffffu;
                    const bool won = deep_argmin(P2S_FULL, wk, wc);
                    const bool none = __ballot_sync(P2S_FULL, won) == 0u;      // this CTA drew no candidate of the level
                    if (won || (none && lane == 0)) {
                        const unsigned long long key = won ? wk : P2S_KEY_EMPTY;
                        const uint32_t cand = won ? wc : 0xffffffffu, cmw = won ? S.wcm[lane] : 0u;
                        const double q0 = won ? S.wq[lane][0] : nan64(), q1 = won ? S.wq[lane][1] : nan64(),
                                     q2 = won ? S.wq[lane][2] : nan64();
#pragma unroll
                        for (uint32_t p = 0; p < (uint32_t)KC; ++p) {
                            dsmem_st_u64(dsmem_addr(&S.ckey[phase][crank], p), key);
                            dsmem_st_u32(dsmem_addr(&S.ccand[phase][crank], p), cand);
                            dsmem_st_u32(dsmem_addr(&S.ccm[phase][crank], p), cmw);
                            dsmem_st_f64(dsmem_addr(&S.cq[phase][crank][0], p), q0);
                            dsmem_st_f64(dsmem_addr(&S.cq[phase][crank][1], p), q1);
                            dsmem_st_f64(dsmem_addr(&S.cq[phase][crank][2], p), q2);
                        }
                    }
                }
                cluster_barrier();                             // release / acquire: every CTA's slots of this phase are complete
                // every thread of every CTA picks the same winner: smallest key, then the smallest candidate index
                unsigned long long gk = P2S_KEY_EMPTY;
                uint32_t gc = 0xffffffffu;
                int gp = 0;
#pragma unroll
                for (int p = 0; p < KC; ++p) {
                    const unsigned long long pk = S.ckey[phase][p];
                    const uint32_t pc = S.ccand[phase][p];
                    if (pc != 0xffffffffu && (gc == 0xffffffffu || pk < gk || (pk == gk && pc < gc))) { gk = pk; gc = pc; gp = p; }
                }
                err_min = key_err(gk);
                win_qx = S.cq[phase][gp][0]; win_qy = S.cq[phase][gp][1]; win_qz = S.cq[phase][gp][2]; win_cm = S.ccm[phase][gp];
                phase ^= 1u;
            }
            // the reference's loop condition (:408) and break rule (:437-441) for the next level
            const bool go_on = (err_min > a.thr) && (C - (k + 1) >= a.min_cams) && !(min(C, ninv0 + k + 1) > C - a.min_cams);
            if (!go_on) {
                if (tid == 0 && crank == 0u) {
                    const bool failed = err_min > a.thr;
                    double *q = a.out_Q + u * 3;
                    q[0] = failed ? nan64() : win_qx; q[1] = failed ? nan64() : win_qy; q[2] = failed ? nan64() : win_qz;
                    a.out_err[u] = failed ? nan64() : err_min;
                    a.out_nexcl[u] = (uint8_t)__popc(u_inv0 | win_cm);
                    a.out_mask[u] = u_nan0 | win_cm;
                }
                break;
            }
            if constexpr (KC == 1) __syncthreads();            // every thread has read the level's result
        }
    }
}

// ---- second kernel of every launch: wide-spread units + the push path's arrival flag ---------------------------
// Runs right behind triangulate_kernel on the same stream.  When that kernel counted no unit with a wide likelihood
// spread (tile_counter[2] == 0: always, unless the likelihood threshold is below 1 / P2S_WIDE_SPREAD) it only publishes the
// push path's arrival flag and returns (~2 us).  Otherwise every warp re-reads the likelihoods of its tiles, finds the
// wide units again (same float comparison as the main kernel) and runs their whole search on the factorisation of A,
// one lane per unit, overwriting the placeholder the main kernel wrote.
struct FixArgs {
    const float4 *obs;
    const float *px, *py, *pl;
    float lik_thr_f;
    int gate;
    long long n_units;
    int n_cams, min_cams;
    double thr, band_eps;
    const uint32_t *cand_masks;
    int max_table_level;
    double *out_Q, *out_err;
    uint8_t *out_nexcl;
    uint32_t *out_mask;
    unsigned long long *stats;
    unsigned int *tile_counter;
    unsigned int *done_flag;
    unsigned int done_value;
};

template <bool DISTORT>
__global__ void __launch_bounds__(128) wide_fixup_kernel(const CamParams<P2S_MAX_CAMS> cams, const LensSet<DISTORT ? P2S_MAX_CAMS : 1> lens,
                                                         const FixArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = a.n_cams;
    // launched with programmatic stream serialisation: the grid may be scheduled while the search kernel drains; nothing of
    // that kernel is read before it has completed and its writes are visible
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const bool any = *reinterpret_cast<const volatile unsigned int *>(a.tile_counter + 2) != 0u;
    if (any) {
        double *sP = reinterpret_cast<double *>(smem_raw);
        LensParams *sLens = reinterpret_cast<LensParams *>(smem_raw + P2S_MAX_CAMS * 12 * sizeof(double));
        unsigned char *slab = smem_raw + P2S_MAX_CAMS * 12 * sizeof(double) + (DISTORT ? sizeof(LensSet<P2S_MAX_CAMS>) : 0) +
                              (size_t)warp * C * 32 * (sizeof(float2) + sizeof(float));
        float2 (*xy)[32] = reinterpret_cast<float2 (*)[32]>(slab);
        float (*wt)[32] = reinterpret_cast<float (*)[32]>(slab + (size_t)C * 32 * sizeof(float2));
        for (int i = threadIdx.x; i < C * 12; i += blockDim.x) sP[i] = (&cams.P[0][0])[i];
        if (DISTORT) {
            const double *src = reinterpret_cast<const double *>(&lens);
            double *dst = reinterpret_cast<double *>(sLens);
            for (int i = threadIdx.x; i < (int)(C * sizeof(LensParams) / sizeof(double)); i += blockDim.x) dst[i] = src[i];
        }
        __syncthreads();
        const uint32_t cmask = (C >= 32) ? 0xffffffffu : ((1u << C) - 1u);
        const float nanf_ = __int_as_float(0x7fc00000);
        const long long n_tiles = (a.n_units + 31) >> 5;
        const long long w0 = (long long)blockIdx.x * 4 + warp, nw = (long long)gridDim.x * 4;
        for (long long tile = w0; tile < n_tiles; tile += nw) {
            const long long u = tile * 32 + lane;
            const bool active = u < a.n_units;
            uint32_t nan0 = 0, inv0 = 0;
            float wlo = __int_as_float(0x7f800000), whi = 0.f;
            for (int c = 0; c < C; ++c) {
                float fx = 0.f, fy = 0.f, fl = nanf_;
                if (active) {
                    if (a.px == nullptr) {
                        const float4 o = __ldg(a.obs + (long long)c * a.n_units + u);
                        fx = o.x; fy = o.y; fl = o.z;
                    } else {
                        fx = a.px[u * C + c]; fy = a.py[u * C + c]; fl = a.pl[u * C + c];
                        if (a.gate && fl < a.lik_thr_f) { fx = fy = fl = nanf_; }
                    }
                }
                xy[c][lane] = make_float2(fx, fy);
                wt[c][lane] = fl;
                const bool isn = fl != fl;
                nan0 |= (uint32_t)isn << c;
                inv0 |= (uint32_t)(isn || fl == 0.f) << c;
                const float la = fabsf(fl);
                whi = fmaxf(whi, la);
                wlo = fminf(wlo, la > 0.f ? la : wlo);
            }
            nan0 &= cmask; inv0 &= cmask;
            const bool wide = active && whi > P2S_WIDE_SPREAD * wlo;
            WideRes wr;
            wr.cands = wr.cams = wr.sweeps = wr.solved = 0; wr.last_level = -1; wr.err = inf64();
            if (wide) {
                WideArgs wa;
                wa.sP = sP; wa.lens = DISTORT ? sLens : nullptr; wa.xy = xy; wa.wt = wt;
                wa.table = a.cand_masks; wa.max_table_level = a.max_table_level; wa.n_cams = C; wa.min_cams = a.min_cams;
                wa.ul = lane; wa.thr = a.thr; wa.nan0 = nan0; wa.inv0 = inv0;
                search_wide_unit<DISTORT>(wa, wr);
                const bool failed = wr.err > a.thr;
                double *q = a.out_Q + u * 3;
                q[0] = failed ? nan64() : wr.qx; q[1] = failed ? nan64() : wr.qy; q[2] = failed ? nan64() : wr.qz;
                a.out_err[u] = failed ? nan64() : wr.err;
                a.out_nexcl[u] = (uint8_t)wr.nexcl;
                a.out_mask[u] = wr.ids;
                if (a.stats != nullptr) {
                    atomicAdd(a.stats + P2S_STAT_CANDIDATES, (unsigned long long)wr.cands);
                    atomicAdd(a.stats + P2S_STAT_CAM_SOLVES, (unsigned long long)wr.cams);
                    atomicAdd(a.stats + P2S_STAT_SOLVED, (unsigned long long)wr.solved);
                    atomicAdd(a.stats + P2S_STAT_WIDE_UNITS, 1ULL);
                    if (wr.last_level >= 0) atomicAdd(a.stats + P2S_STAT_LEVEL0 + wr.last_level, 1ULL);
                    else atomicAdd(a.stats + P2S_STAT_NOT_EVALUATED, 1ULL);
                    if (failed) atomicAdd(a.stats + P2S_STAT_FAILED, 1ULL);
                    if (fabs(wr.err - a.thr) < a.band_eps) atomicAdd(a.stats + P2S_STAT_BAND_THRESHOLD, 1ULL);
                }
            }
            __syncwarp();
        }
    }
    // ---- push path: publish "every output of this launch is visible" to the consumer (possibly a peer GPU) ------
    if (a.done_flag != nullptr) {
        __syncthreads();                                       // the CTA's stores are all issued
        if (threadIdx.x == 0) {
            __threadfence_system();                            // ... and ordered before the count at system scope
            const unsigned int prev = atomicAdd(a.tile_counter + 1, 1u);
            if (prev == gridDim.x - 1) {                       // last CTA of the launch (the main kernel has retired:
                __threadfence_system();                        //  same stream, its stores are visible device-wide)
                *reinterpret_cast<volatile unsigned int *>(a.done_flag) = a.done_value;
            }
        }
    }
}

// Consumer side of the push path (rank 0): wait until every producer's arrival flag reached `value`, then
// release the buffer by writing `value` to each producer's acknowledgement flag (peer memory).
struct CollectArgs {
    const unsigned int *arrive;                 // local [n]
    unsigned int *ack[P2S_MAX_PEERS];           // local or peer
    int n;
    unsigned int value, ack_value;
    unsigned int *err_word;
};

__global__ void __launch_bounds__(32) collect_kernel(const CollectArgs a) {
    const int lane = threadIdx.x;
    if (lane < a.n) {
        const volatile unsigned int *f = a.arrive + lane;
        const unsigned long long t0 = global_ns();
        while ((int)(*f - a.value) < 0) {
            __nanosleep(200);
            if (global_ns() - t0 > 2000000000ULL) { atomicOr(a.err_word, 2u); break; }
        }
    }
    __syncwarp();
    __threadfence_system();
    if (lane < a.n && a.ack[lane] != nullptr) *reinterpret_cast<volatile unsigned int *>(a.ack[lane]) = a.ack_value;
}

// ---- staging: [U][C] planes -> float4 [C][U] with the likelihood gate (triangulation.py:817-821) ----
// UNDISTORT: the points are first undistorted like cv2.undistortPoints(points.astype('float32'), K, dist,
// None, optim_K) (triangulation.py:808-813): OpenCV's 5 fixed-point iterations in double, re-projection
// with the new camera matrix, result rounded to float32.  The arithmetic is written with explicit
// round-to-nearest multiplies / adds in OpenCV's expression order (no FMA contraction), so the float32
// results equal cv2's bit for bit (tests/test_gpu_undistort.py).
struct UndistortCam {
    double cx, cy, ifx, ify;
    double k[8];
    double nk[9];                // new camera matrix, row-major
};
struct UndistortSet { UndistortCam cam[P2S_MAX_CAMS]; };

__device__ __forceinline__ void undistort_point(const UndistortCam &c, float &px, float &py) {
    const double u = (double)px, v = (double)py;
    const double x0 = __dmul_rn(__dsub_rn(u, c.cx), c.ifx), y0 = __dmul_rn(__dsub_rn(v, c.cy), c.ify);
    double x = x0, y = y0;
#pragma unroll 1
    for (int it = 0; it < 5; ++it) {
        const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
        const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k[7], r2), c.k[6]), r2), c.k[5]), r2));
        const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k[4], r2), c.k[1]), r2), c.k[0]), r2));
        const double icdist = __ddiv_rn(num, den);
        if (icdist < 0.0) { x = x0; y = y0; break; }
        // deltaX = 2 p1 x y + p2 (r2 + 2 x x);  deltaY = p1 (r2 + 2 y y) + 2 p2 x y   (k[2] = p1, k[3] = p2)
        const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, c.k[2]), x), y),
                                    __dmul_rn(c.k[3], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
        const double dY = __dadd_rn(__dmul_rn(c.k[2], __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                    __dmul_rn(__dmul_rn(__dmul_rn(2.0, c.k[3]), x), y));
        x = __dmul_rn(__dsub_rn(x0, dX), icdist);
        y = __dmul_rn(__dsub_rn(y0, dY), icdist);
    }
    const double xx = __dadd_rn(__dadd_rn(__dmul_rn(c.nk[0], x), __dmul_rn(c.nk[1], y)), c.nk[2]);
    const double yy = __dadd_rn(__dadd_rn(__dmul_rn(c.nk[3], x), __dmul_rn(c.nk[4], y)), c.nk[5]);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(c.nk[6], x), __dmul_rn(c.nk[7], y)), c.nk[8]));
    px = (float)__dmul_rn(xx, ww);
    py = (float)__dmul_rn(yy, ww);
}

template <bool UNDISTORT>
__global__ void __launch_bounds__(256) stage_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                                    const float *__restrict__ lik, long long n_units, int n_cams,
                                                    double lik_thr, int gate, const UndistortSet lens,
                                                    float4 *__restrict__ out) {
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
            float vx = x[u0 * C + i], vy = y[u0 * C + i];
            if (UNDISTORT) undistort_point(lens.cam[cc], vx, vy);
            sx[uu * ld + cc] = vx;
            sy[uu * ld + cc] = vy;
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
template <class Kern, class... Args>
static cudaError_t launch_persistent(Kern kern, size_t smem, const TriLaunch &L, int *grid_out, Args... args) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long want = ((L.n_units + 31) / 32 + 3) / 4;
    long long grid = (long long)L.sm_count * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    kern<<<(unsigned)grid, 128, smem, L.stream>>>(args...);
    return cudaGetLastError();
}

// Does this launch park its deep levels (see deep_search_kernel)?  Only the lean secular kernels without a lens model do,
// and only when some level the search can reach holds at least deep_min candidates.
static bool deep_applies(const TriLaunch &L) {
    if (L.deep_list == nullptr || L.deep_cap == 0 || L.deep_min == 0 || L.stats != nullptr || L.solver != 0 || L.lens != nullptr) return false;
    for (int k = 1; k <= L.n_cams - L.min_cams; ++k) {
        unsigned long long r = 1;
        for (int i = 1; i <= k; ++i) r = r * (unsigned)(L.n_cams - k + i) / (unsigned)i;
        if (r >= L.deep_min) return true;
    }
    return false;
}

// FULLSET = false (the camera counts 6, 12, 24 between the powers of two): only the exact-count secular kernels are
// instantiated — they are what a 6 / 12 / 24-camera rig runs; everything else goes to the next power of two.
template <int CMAX, bool FULLSET>
static cudaError_t launch_tri(const TriLaunch &L, int *grid_out) {
    CamParams<CMAX> cams;
    for (int c = 0; c < CMAX; ++c)
        for (int j = 0; j < 12; ++j) cams.P[c][j] = (c < L.n_cams) ? L.P[c * 12 + j] : 0.0;
    TriArgs a;
    a.obs = (const float4 *)L.obs; a.n_units = L.n_units; a.n_cams = L.n_cams; a.min_cams = L.min_cams;
    a.px = L.px; a.py = L.py; a.pl = L.pl; a.lik_thr = L.lik_thr;
    a.gate = (L.lik_thr == L.lik_thr) && !(L.lik_thr == -INFINITY);
    {
        float tf = (float)L.lik_thr;                            // round to nearest
        if (a.gate && (double)tf < L.lik_thr) tf = nextafterf(tf, INFINITY);
        a.lik_thr_f = tf;
    }
    a.thr = L.thr; a.band_eps = L.band_eps; a.cand_masks = L.cand_masks;
    for (int i = 0; i < P2S_MAX_CAMS + 2; ++i) a.level_off[i] = L.level_off[i];
    a.max_table_level = L.max_table_level;
    for (int k = 0; k <= P2S_MAX_CAMS; ++k) {
        unsigned long long r = (k <= L.n_cams) ? 1ULL : 0ULL;
        for (int i = 1; i <= k && k <= L.n_cams; ++i) {
            r = r * (unsigned)(L.n_cams - k + i) / (unsigned)i;
            if (r > 0xffffffffULL) { r = 0xffffffffULL; break; }
        }
        a.ncand[k] = (uint32_t)r;
        int lw = 5;
        if (r <= 16) { lw = 0; while ((1ULL << lw) < r) ++lw; }
        a.lw[k] = (unsigned char)lw;
    }
    a.rinv[0] = 0.0;
    for (int m = 1; m <= P2S_MAX_CAMS; ++m) a.rinv[m] = 1.0 / (double)m;
    a.out_Q = L.out_Q; a.out_err = L.out_err; a.out_nexcl = L.out_nexcl; a.out_mask = L.out_mask;
    a.stats = L.stats; a.tile_counter = L.tile_counter;
    a.vec_out = ((((uintptr_t)L.out_Q | (uintptr_t)L.out_err | (uintptr_t)L.out_nexcl | (uintptr_t)L.out_mask) & 15u) == 0) ? 1 : 0;
    a.bulk_out = (a.vec_out && L.bulk_out) ? 1 : 0;
    a.wait_flag = L.wait_flag; a.wait_value = L.wait_value; a.done_flag = L.done_flag; a.done_value = L.done_value;
    a.err_word = L.err_word;
    a.deep_list = deep_applies(L) ? L.deep_list : nullptr; a.deep_cap = L.deep_cap; a.deep_min = L.deep_min;
    const size_t smem = (size_t)CMAX * 12 * sizeof(double) + sizeof(WarpSlab<CMAX, true>) * 4;
    const size_t smem_lean = (size_t)CMAX * 12 * sizeof(double) + sizeof(WarpSlab<CMAX, false>) * 4;
    if constexpr (CMAX == 4 || CMAX == 8) {
        // the pooled variant (opt-in: p2s_set_output_mode(h, 2), or P2S_POOL in the environment for tools/kernel_ab.py /
        // level_time.py): lean secular search on raw planes with exactly CMAX cameras, outputs in this device's memory
        static const bool env_pool = std::getenv("P2S_POOL") != nullptr;
        if (L.allow_pool && (L.pool || env_pool) && L.px != nullptr && L.lens == nullptr && L.solver == 0 && L.n_cams == CMAX && L.stats == nullptr &&
            !a.bulk_out && L.wait_flag == nullptr && L.done_flag == nullptr && L.n_units < 0xffffffffLL) {
            const size_t smem_pool = (size_t)CMAX * 12 * sizeof(double) + sizeof(PoolSlab<CMAX>) * 4;
            return launch_persistent(triangulate_pool_kernel<CMAX>, smem_pool, L, grid_out, cams, a);
        }
    }
    if constexpr (!FULLSET) {
        LensSet<1> none;
        std::memset(&none, 0, sizeof none);
        if (L.stats == nullptr) return launch_persistent(triangulate_kernel<CMAX, 0, false, true, false>, smem_lean, L, grid_out, cams, none, a);
        return launch_persistent(triangulate_kernel<CMAX, 0, false, true, true>, smem, L, grid_out, cams, none, a);
    } else {
    if (L.lens) {                                             // undistort_points: distorted re-projection
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
        return launch_persistent(triangulate_kernel<CMAX, 0, true, false, true>, smem, L, grid_out, cams, lens, a);
    }
    LensSet<1> none;
    std::memset(&none, 0, sizeof none);
    const bool st = L.stats != nullptr, exact = L.n_cams == CMAX;
#ifndef P2S_NO_TMA
    if constexpr (CMAX == 4 || CMAX == 8) {
        // raw planes, exactly CMAX cameras: level 0 straight from the TMA-landed rows (P2S_NO_RAW_L0: A/B switch, kernel_ab.py)
        static const bool no_raw = std::getenv("P2S_NO_RAW_L0") != nullptr;
        if (L.solver == 0 && exact && L.px != nullptr && !no_raw && !a.bulk_out) {
            if (!st) return launch_persistent(triangulate_kernel<CMAX, 0, false, true, false, true>, smem_lean, L, grid_out, cams, none, a);
            return launch_persistent(triangulate_kernel<CMAX, 0, false, true, true, true>, smem, L, grid_out, cams, none, a);
        }
    }
#endif
    if (L.solver == 0 && exact && !st) return launch_persistent(triangulate_kernel<CMAX, 0, false, true, false>, smem_lean, L, grid_out, cams, none, a);
    if (L.solver == 0 && exact) return launch_persistent(triangulate_kernel<CMAX, 0, false, true, true>, smem, L, grid_out, cams, none, a);
    if (L.solver == 0 && !st) return launch_persistent(triangulate_kernel<CMAX, 0, false, false, false>, smem_lean, L, grid_out, cams, none, a);
    if (L.solver == 0) return launch_persistent(triangulate_kernel<CMAX, 0, false, false, true>, smem, L, grid_out, cams, none, a);
    return launch_persistent(triangulate_kernel<CMAX, 1, false, false, true>, smem, L, grid_out, cams, none, a);
    }
}

static cudaError_t launch_main(const TriLaunch &L, int *grid_out) {
    const bool plain = L.solver == 0 && L.lens == nullptr;
    if (plain && L.n_cams == 6) return launch_tri<6, false>(L, grid_out);
    if (plain && L.n_cams == 12) return launch_tri<12, false>(L, grid_out);
    if (plain && L.n_cams == 24) return launch_tri<24, false>(L, grid_out);
    if (L.n_cams <= 4) return launch_tri<4, true>(L, grid_out);
    if (L.n_cams <= 8) return launch_tri<8, true>(L, grid_out);
    if (L.n_cams <= 16) return launch_tri<16, true>(L, grid_out);
    return launch_tri<32, true>(L, grid_out);
}

// Programmatic dependent launch (sm_90+): the kernel's CTAs may be scheduled while the previous kernel of the stream is
// still draining; it orders itself behind that kernel with griddepcontrol.wait.  Hides the ~3 us launch gap of the
// second kernel of every step (P2S_NO_PDL: plain launch, A/B).
template <class Kern, class... Args>
static cudaError_t launch_dependent_block(Kern kern, unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
#ifndef P2S_NO_PDL
    cfg.attrs = attr; cfg.numAttrs = 1;
#endif
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
template <class Kern, class... Args>
static cudaError_t launch_dependent(Kern kern, unsigned grid, size_t smem, cudaStream_t stream, Args... args) {
    return launch_dependent_block(kern, grid, 128u, smem, stream, args...);
}

// The deep-level kernel between the search kernel and the fix-up kernel (only when deep_applies()).  Every CTA reads the
// parked count first and returns at once when its index is beyond it.
static cudaError_t launch_deep(const TriLaunch &L) {
    CamParams<P2S_MAX_CAMS> cams;
    for (int c = 0; c < P2S_MAX_CAMS; ++c)
        for (int j = 0; j < 12; ++j) cams.P[c][j] = (c < L.n_cams) ? L.P[c * 12 + j] : 0.0;
    DeepArgs a;
    std::memset(&a, 0, sizeof a);
    a.obs = (const float4 *)L.obs; a.px = L.px; a.py = L.py; a.pl = L.pl;
    a.gate = (L.lik_thr == L.lik_thr) && !(L.lik_thr == -INFINITY);
    {
        float tf = (float)L.lik_thr;
        if (a.gate && (double)tf < L.lik_thr) tf = nextafterf(tf, INFINITY);
        a.lik_thr_f = tf;
    }
    a.n_units = L.n_units; a.n_cams = L.n_cams; a.min_cams = L.min_cams; a.thr = L.thr;
    a.cand_masks = L.cand_masks; a.max_table_level = L.max_table_level;
    for (int i = 0; i < P2S_MAX_CAMS + 2; ++i) a.level_off[i] = L.level_off[i];
    const int n = L.n_cams;
#ifdef P2S_NO_DOWNDATE_EXACT
    a.always_downdate = 0;
#else
    a.always_downdate = (n == 4 || n == 6 || n == 8 || n == 12 || n == 16 || n == 24 || n == 32) ? 1 : 0;   // the exact-count kernels
#endif
    a.rinv[0] = 0.0;
    for (int m = 1; m <= P2S_MAX_CAMS; ++m) a.rinv[m] = 1.0 / (double)m;
    a.out_Q = L.out_Q; a.out_err = L.out_err; a.out_nexcl = L.out_nexcl; a.out_mask = L.out_mask;
    a.list = L.deep_list; a.cap = L.deep_cap; a.count = L.tile_counter + 3;
    // P2S_DEEP_CLUSTER CTAs (one thread-block cluster, portable size) per parked unit; the grid stays a multiple of it
    constexpr unsigned KC = P2S_DEEP_CLUSTER;
    const unsigned grid = ((unsigned)(P2S_DEEP_GRID_MULT * L.sm_count) + KC - 1u) / KC * KC;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3((unsigned)kDeepThreads); cfg.dynamicSmemBytes = 0; cfg.stream = L.stream;
    cudaLaunchAttribute attr[2];
    unsigned na = 0;
#ifndef P2S_NO_PDL
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
#endif
    if (KC > 1u) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = KC; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, deep_search_kernel<(int)KC>, cams, a);
}

// The wide-spread / arrival-flag kernel behind the main one.  Always launched: that a likelihood threshold >= 1 / 256
// rules wide units out rests on likelihoods being <= 1, which is a convention of pose estimators, not a contract of this
// interface; an empty launch costs ~2 us behind the search kernel and keeps the accuracy guarantee unconditional.
static cudaError_t launch_fixup(const TriLaunch &L, int main_grid) {
    CamParams<P2S_MAX_CAMS> cams;
    for (int c = 0; c < P2S_MAX_CAMS; ++c)
        for (int j = 0; j < 12; ++j) cams.P[c][j] = (c < L.n_cams) ? L.P[c * 12 + j] : 0.0;
    FixArgs a;
    a.obs = (const float4 *)L.obs; a.px = L.px; a.py = L.py; a.pl = L.pl;
    a.gate = (L.lik_thr == L.lik_thr) && !(L.lik_thr == -INFINITY);
    {
        float tf = (float)L.lik_thr;
        if (a.gate && (double)tf < L.lik_thr) tf = nextafterf(tf, INFINITY);
        a.lik_thr_f = tf;
    }
    a.n_units = L.n_units; a.n_cams = L.n_cams; a.min_cams = L.min_cams; a.thr = L.thr; a.band_eps = L.band_eps;
    a.cand_masks = L.cand_masks; a.max_table_level = L.max_table_level;
    a.out_Q = L.out_Q; a.out_err = L.out_err; a.out_nexcl = L.out_nexcl; a.out_mask = L.out_mask;
    a.stats = L.stats; a.tile_counter = L.tile_counter; a.done_flag = L.done_flag; a.done_value = L.done_value;
    const long long n_tiles = (L.n_units + 31) / 32;
    long long grid = (long long)L.sm_count * 4;
    if (grid > (n_tiles + 3) / 4) grid = (n_tiles + 3) / 4;
    if (grid < 1) grid = 1;
    (void)main_grid;
    const size_t slab = (size_t)4 * L.n_cams * 32 * (sizeof(float2) + sizeof(float));
    if (L.lens) {
        LensSet<P2S_MAX_CAMS> lens;
        std::memset(&lens, 0, sizeof lens);
        for (int c = 0; c < L.n_cams; ++c) {
            const p2s_camera_model &m = L.lens[c];
            LensParams &o = lens.cam[c];
            for (int j = 0; j < 9; ++j) o.R[j] = m.R[j];
            for (int j = 0; j < 3; ++j) o.T[j] = m.T[j];
            o.fx = m.K[0]; o.fy = m.K[4]; o.cx = m.K[2]; o.cy = m.K[5];
            for (int j = 0; j < 8; ++j) o.k[j] = m.dist[j];
        }
        const size_t smem = P2S_MAX_CAMS * 12 * sizeof(double) + sizeof(LensSet<P2S_MAX_CAMS>) + slab;
        cudaError_t e = cudaFuncSetAttribute(wide_fixup_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return launch_dependent(wide_fixup_kernel<true>, (unsigned)grid, smem, L.stream, cams, lens, a);
    } else {
        LensSet<1> none;
        std::memset(&none, 0, sizeof none);
        const size_t smem = P2S_MAX_CAMS * 12 * sizeof(double) + slab;
        cudaError_t e = cudaFuncSetAttribute(wide_fixup_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return launch_dependent(wide_fixup_kernel<false>, (unsigned)grid, smem, L.stream, cams, none, a);
    }
}

cudaError_t launch_triangulate(const TriLaunch &L, int *grid_out) {
    int grid = 0;
    cudaError_t e = launch_main(L, &grid);
    if (grid_out) *grid_out = grid;
    if (e != cudaSuccess) return e;
    L.kernels = 2;
    if (deep_applies(L)) {
        e = launch_deep(L);
        if (e != cudaSuccess) return e;
        L.kernels = 3;
    }
#ifdef P2S_NO_FIXUP_LAUNCH                                     /* A/B switch, tools/kernel_ab.py */
    return e;
#else
    return launch_fixup(L, grid);
#endif
}

cudaError_t launch_stage(const float *x, const float *y, const float *lik, long long n_units, int n_cams,
                         double lik_thr, const p2s_camera_model *lens, void *out, int sm_count, cudaStream_t stream) {
    const int gate = (lik_thr == lik_thr) && !(lik_thr == -INFINITY);
    const size_t smem = (size_t)3 * 256 * (n_cams + 1) * sizeof(float);
    long long blocks = (n_units + 255) / 256;
    long long grid = (long long)sm_count * 8;
    if (grid > blocks) grid = blocks;
    if (grid < 1) grid = 1;
    static UndistortSet set;                                  // zero-initialised; filled per call when used
    cudaError_t e;
    if (lens) {
        UndistortSet u;
        std::memset(&u, 0, sizeof u);
        for (int c = 0; c < n_cams; ++c) {
            const p2s_camera_model &m = lens[c];
            UndistortCam &o = u.cam[c];
            o.cx = m.K[2]; o.cy = m.K[5]; o.ifx = 1.0 / m.K[0]; o.ify = 1.0 / m.K[4];
            for (int j = 0; j < 8; ++j) o.k[j] = m.dist[j];
            for (int j = 0; j < 9; ++j) o.nk[j] = m.newK[j];
        }
        e = cudaFuncSetAttribute(stage_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        stage_kernel<true><<<(unsigned)grid, 256, smem, stream>>>(x, y, lik, n_units, n_cams, lik_thr, gate, u, (float4 *)out);
    } else {
        e = cudaFuncSetAttribute(stage_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        stage_kernel<false><<<(unsigned)grid, 256, smem, stream>>>(x, y, lik, n_units, n_cams, lik_thr, gate, set, (float4 *)out);
    }
    return cudaGetLastError();
}

cudaError_t launch_collect(const unsigned int *arrive, unsigned int *const *ack, int n, unsigned int value,
                           unsigned int ack_value, unsigned int *err_word, cudaStream_t stream) {
    CollectArgs a;
    std::memset(&a, 0, sizeof a);
    a.arrive = arrive; a.n = n; a.value = value; a.ack_value = ack_value; a.err_word = err_word;
    for (int i = 0; i < n && i < P2S_MAX_PEERS; ++i) a.ack[i] = ack ? ack[i] : nullptr;
    collect_kernel<<<1, 32, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_fp64_peak(double *out, int blocks, int iters, cudaStream_t stream) {
    fp64_peak_kernel<<<blocks, 256, 0, stream>>>(out, iters, 1.0);
    return cudaGetLastError();
}

}  // namespace p2s
