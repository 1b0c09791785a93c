// Single-person association search: which detected person of each camera (and which cameras)
// triangulate the tracked keypoint best.
//
// Replaces, per frame, `persons_combinations` + `best_persons_and_cameras_combination` +
// `triangulate_comb` (Pose2Sim/personAssociation.py:67-99, :154-257, :102-151), restated in
// SURVEY.md §8(a) `associate_frame`:
//
//   rows = itertools.product(range(max(n_c, 1)) for c)       last camera fastest            (:95-97)
//   k = 0; err_last = inf; best = inf
//   while err_last > thr and C - (n_missing + k) >= min_cams:                                (:194)
//       for row in rows (product order):                                                     (:196)
//           camera off when n_c == 0, likelihood < lik_thr, or likelihood == 0               (:215-216)
//           skip if fewer than min_cams active                                               (:219-221)
//           candidates = lexicographic k-subsets of the ACTIVE cameras                       (:222-225)
//           solve each; skip row if all NaN; err_last = nanmin                               (:229-238)
//           best updated on strict '<'; first row with err_last < thr ends the row loop      (:242-248)
//       k += 1
//
// Mapping: one warp per frame (frames are plentiful), one LANE PER ROW in ordered chunks of 32 rows.
//   * per frame the warp first builds, camera-person-parallel, the 10-entry normal-matrix block
//     a a^T + b b^T of EVERY detection (camera c, person p) in shared memory; a row's normal matrix is
//     then the sum of its active cameras' blocks (10 FP64 adds per camera instead of 64 flops), and a
//     candidate that drops cameras subtracts their blocks;
//   * a lane keeps its row's person indices as packed 4-bit digits and steps them by 32 rows per chunk
//     with a small mixed-radix addition (multiply-shift by the reciprocal radix) — no 64-bit division;
//   * the lexicographic k-subsets of the active cameras are the subsequence of the all-camera table
//     whose members are subsets of `active`, so the triangulation kernel's mask table is reused;
//   * the ordered early exit is a ballot: the first lane whose row is under the threshold wins and later
//     lanes of the chunk are discarded; the running best is a shuffle arg-min over (error, row).
#include "p2s_math.cuh"
#include "p2s_internal.h"

namespace p2s {

struct AssocArgs {
    const float4 *obs;            // [n_frames][n_cams][max_persons]
    const int32_t *count;         // [n_frames][n_cams]
    long long n_frames;
    int n_cams, max_persons, min_cams;
    double thr, lik_thr;
    const uint32_t *cand_masks;
    uint32_t level_off[P2S_MAX_CAMS + 2];
    int max_table_level;
    double *out_err;
    int8_t *out_comb;
    double *out_Q;
    uint32_t *out_stats;
    unsigned int *tile_counter;    // [0] frame dispenser of the main pass, [1] of the wide pass, [2] frames flagged wide
    uint8_t *wide_flags;           // [n_frames]: 1 = the frame's active likelihoods span more than P2S_WIDE_SPREAD
    // row filter (see associate_kernel): |P_c[2] . (q, 1)| <= bnd_k1 |q| + bnd_k2 for every camera
    double bnd_k1, bnd_k2;
    int search_mode;               // 0 = rows filtered by their lower bound, 1 = every row solved and re-projected
};

// Blocks of camera PAIRS: entry (j, d0, d1) = block of (camera 2j, person d0) + block of (camera 2j + 1, person d1), each only
// if that detection is active — a row's normal matrix is then four 10-entry additions instead of eight with a test per
// camera.  Kept when the table is small (np^2 entries per pair), else the per-camera blocks are summed.
__host__ __device__ inline int assoc_pair_entries(int cmax, int np) {
    return ((size_t)(cmax / 2) * np * np * 10 * sizeof(double) <= 12288) ? np * np : 0;
}
__host__ __device__ inline size_t assoc_slab_bytes(int cmax, int np) {
    return (size_t)cmax * np * (sizeof(float4) + 10 * sizeof(double)) + 4 * (size_t)cmax * sizeof(uint32_t) +
           (size_t)(cmax / 2) * 256 +        // per camera PAIR: byte of two packed digits -> active / NaN bits of the pair
           (size_t)(cmax / 2) * assoc_pair_entries(cmax, np) * 10 * sizeof(double);
}
// row queue of a team: 2 x its threads entries of {row index, candidate mask, packed person digits}
__host__ __device__ inline size_t assoc_queue_bytes(int cmax, int nw) {
    return (size_t)2 * 32 * nw * (2 * sizeof(unsigned long long) + (size_t)((cmax + 7) / 8) * sizeof(uint32_t));
}

template <int CMAX>
__device__ __forceinline__ uint32_t digit_of(const uint32_t (&dig)[4], int c) { return (dig[c >> 3] >> ((c & 7) * 4)) & 15u; }

// same for a camera index only known at run time (select chain instead of a local-memory array index)
__device__ __forceinline__ uint32_t digit_rt(const uint32_t (&dig)[4], int c) {
    const uint32_t w = c < 8 ? dig[0] : c < 16 ? dig[1] : c < 24 ? dig[2] : dig[3];
    return (w >> ((c & 7) * 4)) & 15u;
}

// digits += v in the mixed radix n[c] (last camera fastest, itertools.product order); n <= 16:
// floor(t / n) = (t * ceil(65536 / n)) >> 16 is exact for t < 65536 / n, i.e. for any v <= 4000
template <int CMAX>
__device__ __forceinline__ void radix_add(uint32_t (&dig)[4], uint32_t v, const uint32_t *s_n, const uint32_t *s_inv, int C) {
    uint32_t carry = v;
#pragma unroll
    for (int c = CMAX - 1; c >= 0; --c) {
        if (c < C && carry != 0u) {                             // the carry of a small step dies out after a few cameras
            const uint32_t n = s_n[c] ? s_n[c] : 1u;
            const uint32_t t = digit_of<CMAX>(dig, c) + carry;
            const uint32_t q = (t * s_inv[c]) >> 16;
            const uint32_t d = t - q * n;
            dig[c >> 3] = (dig[c >> 3] & ~(15u << ((c & 7) * 4))) | (d << ((c & 7) * 4));
            carry = q;
        }
    }
}

// ---- the same step for rigs of at most 8 cameras, all cameras at once --------------------------------------------------
// The row's digits live in ONE word with the last camera (the fastest digit) in the lowest nibble, each digit stored with
// the bias 16 - n_c: a nibble then overflows exactly when its digit reaches n_c, so the binary carry of an ordinary
// 32-bit addition IS the mixed-radix carry.  Nibbles that overflowed come out unbiased and get their bias back.
__device__ __forceinline__ uint32_t nibble_reverse(uint32_t x) {
    x = ((x & 0x0f0f0f0fu) << 4) | ((x >> 4) & 0x0f0f0f0fu);
    return __byte_perm(x, 0u, 0x0123);
}
__device__ __forceinline__ uint32_t swar_step(uint32_t rw, uint32_t stepw, uint32_t biasw) {
    const uint32_t sum = rw + stepw;
    const uint32_t cout = (rw & stepw) | ((rw | stepw) & ~sum);      // bit i: carry out of bit i
    const uint32_t ov = (cout >> 3) & 0x11111111u;                    // nibbles that overflowed
    return sum + (biasw & (ov * 15u));
}

// Team = NW warps that share one frame.  NW = 1: a warp per frame (4 frames per CTA) — best when there are
// many frames with few rows.  NW = 8: a 256-thread CTA per frame, 256 consecutive rows per step — for frames
// with many rows (cfg4: 6^8 rows per frame) or few frames; the ordered early exit is resolved across the
// team's warps through shared memory with two CTA barriers per step.
template <int NW>
struct alignas(16) TeamScratch {  // per team, in shared memory (only used when NW > 1)
    uint32_t eval[NW], hit[NW], any[NW];
    unsigned long long key[NW];
    double last[NW];
    double q[NW][3];
    uint32_t valid[NW];
    uint32_t dig[NW][4];
    unsigned int frame;
    unsigned int rows, cands;
    uint32_t wcount[2][NW];       // row filter: survivors per warp of a scan step (two buffers, alternating: one barrier per step)
    unsigned long long lastrow[NW];   // row index of T.last[w]
    unsigned long long deadrow[NW];   // level end: last row of warp w that counts as evaluated but was dropped unsolved
    unsigned int again;           // the frame has to be searched again without the filter
};

template <int NW>
__device__ __forceinline__ void team_sync() {
    if (NW == 1) __syncwarp(); else __syncthreads();
}

// WIDE = false: the main pass; it FLAGS the frames whose active likelihoods span more than P2S_WIDE_SPREAD (only possible
// with a likelihood threshold near 0) and leaves them alone.  WIDE = true: the pass behind it, which returns at once when
// nothing was flagged and otherwise searches the flagged frames with every candidate solved from a factorisation of A
// itself (Givens QR + one-sided Jacobi, p2s_math.cuh) like the reference's SVD, instead of the normal matrix.
// ROW FILTER (search_mode 0, not in the WIDE pass).  The reference solves and re-projects every row it visits until the
// first row whose error is under the threshold (:196-248); when such a row exists it IS the result (every row before it
// is >= the threshold, and the running best is replaced on a strict '<').  So a candidate only needs its exact error if
// that error can be under the threshold.  Iteration 0 of the eigen-solve (one 3x3 factorisation) yields
// q0 = -A^-1 b and s = min_q R(q), R(q) = (q,1)^T M (q,1) = s + (q - q0)^T A (q - q0) >= s + (|q| - |q0|)^2 / tau with
// tau = trace(A^-1) >= 1 / (smallest eigenvalue of A).  For ANY q, R(q) = sum_c (w_c d_c e_c)^2 (e_c: pixel distance of
// camera c, d_c = P_c[2].(q,1), common.py:344-345 / :357-375 are the same rows) and |w_c d_c| <= whi (k1 |q| + k2), so the
// mean error over the m cameras obeys E >= sqrt(sum e_c^2) / m >= sqrt(R(q)) / (m whi (k1 |q| + k2)); minimising the right
// side over |q| gives a bound that needs no knowledge of where the solver ends:
//     E^2 >= s / ((k1 |q0| + k2)^2 + k1^2 s tau) / (m whi)^2
// (p2s_math.cuh, SecularBound).  The team SCANS the rows in product order, computing only that bound for every candidate
// of a row, and queues the rows that have candidates whose bound is not above the threshold (x 1.001 + 1e-5 px) together
// with the mask of those candidates; full batches of queued rows then get the exact evaluation of their marked
// candidates, in row order, with the reference's first-hit rule.  A frame whose search ends WITHOUT a hit needs the exact
// minimum over all its rows instead (and the exact error of every level's last row): it is searched again with the
// filter off.  Results are those of the exhaustive search (tests/test_gpu_associate.py compares the two bit for bit).
template <int CMAX, int NW, bool WIDE>
#ifndef P2S_ASSOC_MIN_BLOCKS
#define P2S_ASSOC_MIN_BLOCKS 4      /* resident 4-warp CTAs per SM of the warp-per-frame variant (A/B: tools/kernel_ab.py) */
#endif
#ifndef P2S_ASSOC_MIN_BLOCKS8
#define P2S_ASSOC_MIN_BLOCKS8 2     /* resident 8-warp CTAs per SM of the CTA-per-frame variant */
#endif
__global__ void __launch_bounds__(NW == 1 ? 128 : 32 * NW, NW == 1 ? P2S_ASSOC_MIN_BLOCKS : NW == 8 ? P2S_ASSOC_MIN_BLOCKS8 : 1)
associate_kernel(const CamParams<CMAX> cams, const AssocArgs a) {
    if (WIDE && *reinterpret_cast<const volatile unsigned int *>(a.tile_counter + 2) == 0u) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TEAMS = (NW == 1) ? 4 : 1;
    constexpr int TEAM_THREADS = 32 * NW;
    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    const int team = (NW == 1) ? warp_in_cta : 0;
    const int warp = (NW == 1) ? 0 : warp_in_cta;           // warp index inside the team
    const int ttid = warp * 32 + lane;                        // thread index inside the team
    const int C = a.n_cams, NP = a.max_persons;
    double *sP = reinterpret_cast<double *>(smem_raw);
    // per-team slab: obs float4 [CMAX][NP], blocks double [CMAX * NP][10], n_c, ok masks, reciprocal radices
    constexpr int DW = (CMAX + 7) / 8;                        // words of packed person digits per row
    constexpr int QCAP = 2 * TEAM_THREADS;
    const size_t team_bytes = assoc_slab_bytes(CMAX, NP) + sizeof(TeamScratch<NW>) + assoc_queue_bytes(CMAX, NW);
    unsigned char *base = smem_raw + CMAX * 12 * sizeof(double) + team_bytes * team;
    float4 *sobs = reinterpret_cast<float4 *>(base);
    double *sblk = reinterpret_cast<double *>(base + (size_t)CMAX * NP * sizeof(float4));
    uint32_t *s_n = reinterpret_cast<uint32_t *>(base + (size_t)CMAX * NP * (sizeof(float4) + 10 * sizeof(double)));
    uint32_t *s_ok = s_n + CMAX;
    uint32_t *s_inv = s_ok + CMAX;
    uint32_t *s_nan = s_inv + CMAX;                   // per camera: persons whose x, y or likelihood is NaN
    // s_pair[j][b]: for the camera pair (2j, 2j + 1) and the byte b = digit(2j) | digit(2j + 1) << 4 of a row's packed
    // digits: bit 0 / 1 = camera 2j / 2j + 1 active, bit 2 / 3 = its chosen detection holds a NaN (one LDS per pair and row
    // instead of a shift-and-test chain per camera: that chain was 10 % of the kernel's stall samples)
    unsigned char *s_pair = reinterpret_cast<unsigned char *>(s_nan + CMAX);
    const int NPP = assoc_pair_entries(CMAX, NP);             // 0: no pair blocks
    double *spb = reinterpret_cast<double *>(s_pair + (CMAX / 2) * 256);                          // [CMAX / 2][NPP][10]
    TeamScratch<NW> &T = *reinterpret_cast<TeamScratch<NW> *>(base + assoc_slab_bytes(CMAX, NP));
    unsigned long long *q_row = reinterpret_cast<unsigned long long *>(base + assoc_slab_bytes(CMAX, NP) + sizeof(TeamScratch<NW>));
    unsigned long long *q_mask = q_row + QCAP;                // candidates of the row that need their exact error
    uint32_t *q_dig = reinterpret_cast<uint32_t *>(q_mask + QCAP);                                 // [QCAP][DW]

    for (int i = threadIdx.x; i < CMAX * 12; i += blockDim.x) sP[i] = (&cams.P[0][0])[i];
    __syncthreads();

    for (;;) {
        unsigned int f = 0;
        if (NW == 1) {
            if (lane == 0) f = atomicAdd(a.tile_counter + (WIDE ? 1 : 0), 1u);
            f = __shfl_sync(P2S_FULL, f, 0);
        } else {
            if (ttid == 0) T.frame = atomicAdd(a.tile_counter + (WIDE ? 1 : 0), 1u);
            __syncthreads();
            f = T.frame;
        }
        if ((long long)f >= a.n_frames) break;
        if (WIDE && !a.wide_flags[f]) { team_sync<NW>(); continue; }   // team-uniform; the barrier keeps T.frame's readers ahead of its next writer

        // ---- stage the frame, gate the detections, build every detection's block ----------------------
        const float4 *fobs = a.obs + (long long)f * C * NP;
        for (int i = ttid; i < C * NP; i += TEAM_THREADS) {
            const float4 o = __ldg(fobs + i);
            sobs[i] = o;
            double b[10];
            // A detection with a NaN coordinate / likelihood stays ACTIVE in the reference (NaN < thr and NaN == 0 are
            // both false, personAssociation.py:215-216) and poisons every subset that keeps its camera (Q = NaN, all
            // distances +inf) until the subset search drops that camera.  Its block is stored as zero so that the
            // row's normal matrix and the block subtractions stay finite; the poisoning is applied per candidate below.
            const bool nanobs = (o.x != o.x) || (o.y != o.y) || (o.z != o.z);
            camera_block(sP + (i / NP) * 12, nanobs ? 0.0 : (double)o.x, nanobs ? 0.0 : (double)o.y, nanobs ? 0.0 : (double)o.z, b);
            double2 *dst = reinterpret_cast<double2 *>(sblk + (size_t)i * 10);
#pragma unroll
            for (int e = 0; e < 5; ++e) dst[e] = make_double2(b[2 * e], b[2 * e + 1]);
        }
        if (NW > 1 && ttid == 0) { T.rows = 0; T.cands = 0; }
        team_sync<NW>();
        if (ttid < C) {
            int n = a.count[(long long)f * C + ttid];
            n = max(0, min(n, NP));
            uint32_t ok = 0, nanbits = 0;
            for (int p = 0; p < n; ++p) {
                const float4 o = sobs[ttid * NP + p];
                const double l = (double)o.z;
                // gate (:215-216): likelihood < thr -> 0 -> off; likelihood == 0 -> off; NaN stays on
                if (!(l < a.lik_thr) && !(l == 0.0)) ok |= 1u << p;
                if ((o.x != o.x) || (o.y != o.y) || (o.z != o.z)) nanbits |= 1u << p;
            }
            s_n[ttid] = (uint32_t)n;
            s_ok[ttid] = ok;
            s_nan[ttid] = nanbits;
            const uint32_t nn = n ? (uint32_t)n : 1u;
            s_inv[ttid] = (65536u + nn - 1u) / nn;
            if (!WIDE) {                                      // smallest / largest active |likelihood| of the camera
                float wlo = __int_as_float(0x7f800000), whi = 0.f;
                for (int p = 0; p < n; ++p)
                    if ((ok >> p) & 1u) {
                        const float la = fabsf(sobs[ttid * NP + p].z);
                        wlo = fminf(wlo, la); whi = fmaxf(whi, la);
                    }
                reinterpret_cast<float *>(s_pair)[2 * ttid] = wlo;         // s_pair is rebuilt below
                reinterpret_cast<float *>(s_pair)[2 * ttid + 1] = whi;
            }
        }
        team_sync<NW>();
        float whi_frame = __int_as_float(0x7f800000);        // largest active |likelihood| of the frame (bounds the row weights)
        if (!WIDE) {
            float wlo = __int_as_float(0x7f800000), whi = 0.f;
            for (int c = 0; c < C; ++c) {
                wlo = fminf(wlo, reinterpret_cast<const float *>(s_pair)[2 * c]);
                whi = fmaxf(whi, reinterpret_cast<const float *>(s_pair)[2 * c + 1]);
            }
            whi_frame = whi;
            const bool wide_frame = whi > P2S_WIDE_SPREAD * wlo;
            team_sync<NW>();
            if (wide_frame) {                                 // team-uniform: left to the WIDE pass
                if (ttid == 0) { a.wide_flags[f] = 1; atomicAdd(a.tile_counter + 2, 1u); }
                continue;
            }
            if (ttid == 0) a.wide_flags[f] = 0;
        }
        for (int i = ttid; i < (CMAX / 2) * 256; i += TEAM_THREADS) {
            const int j = i >> 8, c0 = 2 * j, c1 = c0 + 1;
            const uint32_t d0 = (uint32_t)i & 15u, d1 = ((uint32_t)i >> 4) & 15u;
            uint32_t t = 0;
            if (c0 < C && s_n[c0]) t |= ((s_ok[c0] >> d0) & 1u) | (((s_ok[c0] >> d0) & (s_nan[c0] >> d0) & 1u) << 2);
            if (c1 < C && s_n[c1]) t |= (((s_ok[c1] >> d1) & 1u) << 1) | (((s_ok[c1] >> d1) & (s_nan[c1] >> d1) & 1u) << 3);
            s_pair[i] = (unsigned char)t;
        }
        for (int i = ttid; i < (CMAX / 2) * NPP; i += TEAM_THREADS) {
            const int j = i / NPP, rem = i - j * NPP, d1 = rem / NP, d0 = rem - d1 * NP, c0 = 2 * j, c1 = c0 + 1;
            const bool a0 = c0 < C && s_n[c0] && ((s_ok[c0] >> d0) & 1u), a1 = c1 < C && s_n[c1] && ((s_ok[c1] >> d1) & 1u);
            const double *b0 = sblk + (size_t)(c0 * NP + d0) * 10, *b1 = sblk + (size_t)(c1 * NP + d1) * 10;
#pragma unroll
            for (int e = 0; e < 10; ++e) spb[(size_t)i * 10 + e] = (a0 ? b0[e] : 0.0) + (a1 ? b1[e] : 0.0);
        }
        team_sync<NW>();
        // a row's normal matrix: the sum of its active detections' blocks (ascending camera order)
        auto row_matrix = [&](Sym4 &Mrow, const uint32_t (&dg)[4], uint32_t active) {
            sym4_zero(Mrow);
            if (NPP) {
#pragma unroll
                for (int j = 0; j < CMAX / 2; ++j) {
                    const uint32_t byte = (dg[j >> 2] >> ((j & 3) * 8)) & 255u;
                    const double2 *src = reinterpret_cast<const double2 *>(spb + (size_t)(j * NPP + (int)((byte & 15u) + (uint32_t)NP * (byte >> 4))) * 10);
                    const double2 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4];
                    Mrow.m00 += v0.x; Mrow.m01 += v0.y; Mrow.m02 += v1.x; Mrow.m03 += v1.y; Mrow.m11 += v2.x;
                    Mrow.m12 += v2.y; Mrow.m13 += v3.x; Mrow.m22 += v3.y; Mrow.m23 += v4.x; Mrow.m33 += v4.y;
                }
            } else {
#pragma unroll
                for (int c = 0; c < CMAX; ++c) {
                    if ((active >> c) & 1u) {
                        const double2 *src = reinterpret_cast<const double2 *>(sblk + (size_t)(c * NP + digit_of<CMAX>(dg, c)) * 10);
                        const double2 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4];
                        Mrow.m00 += v0.x; Mrow.m01 += v0.y; Mrow.m02 += v1.x; Mrow.m03 += v1.y; Mrow.m11 += v2.x;
                        Mrow.m12 += v2.y; Mrow.m13 += v3.x; Mrow.m22 += v3.y; Mrow.m23 += v4.x; Mrow.m33 += v4.y;
                    }
                }
            }
        };
        uint32_t present = 0;
        unsigned long long total_rows = 1;
        bool overflow = false;
        for (int c = 0; c < C; ++c) {
            const uint32_t n = s_n[c];
            if (n) present |= 1u << c;
            const unsigned long long nn = n ? n : 1u;
            if (total_rows > (0x7fffffffffffffffULL / nn)) overflow = true; else total_rows *= nn;
        }
        const int n_missing = C - __popc(present);

        // digit stepping for <= 8 cameras (swar_step): bias word and the digits of one step, last camera in the lowest nibble
        uint32_t biasw = 0u, stepw = 0u;
        if (CMAX <= 8) {
            uint32_t sd[4] = {0, 0, 0, 0};
            radix_add<CMAX>(sd, (uint32_t)TEAM_THREADS, s_n, s_inv, C);
            stepw = nibble_reverse(sd[0] << (4 * (8 - C)));
            for (int c = 0; c < C; ++c) biasw |= (16u - (s_n[c] ? s_n[c] : 1u)) << (4 * (C - 1 - c));
        }
        // replicated in every thread of the team (all threads apply the same updates)
        double err_last = inf64();
        // key of the global best.  The reference starts from error = +inf and replaces it on a strict '<' (:242-245), so a
        // row whose best error is +inf (single-camera subsets, subsets poisoned by a NaN observation) never becomes it.
        const unsigned long long kInfKey = 0x7ff0000000000000ULL;
        const unsigned long long kNoRow = 0xffffffffffffffffULL;
        unsigned long long best_key = kInfKey;
        double bqx = nan64(), bqy = bqx, bqz = bqx;
        uint32_t b_valid = 0;                                 // cameras used by the best candidate
        uint32_t b_dig[4] = {0, 0, 0, 0};                     // person digits of the best row
        unsigned int st_rows = 0, st_cands = 0;               // per warp / per lane counters
        const double whi_d = (double)whi_frame;
        const double thr_b = fma(a.thr, 1.001, 1e-5);
        const double thr2 = thr_b * thr_b;                    // a bound above this cannot belong to a hit (NaN / inf threshold: never)
        const double k1sq = a.bnd_k1 * a.bnd_k1;

        for (int pass = 0; pass < 2; ++pass) {                // pass 1 only for a filtered search that ended without a hit
        const bool filter = !WIDE && a.search_mode == 0 && pass == 0;
        bool redo = false, hit_found = false;                 // redo: a candidate was dropped, or an exact error equalled the threshold
        err_last = inf64(); best_key = kInfKey; bqx = bqy = bqz = nan64(); b_valid = 0;
        b_dig[0] = b_dig[1] = b_dig[2] = b_dig[3] = 0u;
        st_rows = 0; st_cands = 0;

        for (int k = 0; !overflow && err_last > a.thr && C - (n_missing + k) >= a.min_cams; ++k) {
            const bool tabled = k <= a.max_table_level;
            const uint32_t ncand_all = (k == 0) ? 1u : tabled ? (a.level_off[k + 1] - a.level_off[k]) : binom_u32(C, k);
            const uint32_t *table = a.cand_masks + a.level_off[tabled ? k : 0];
            const bool masked = ncand_all <= 64u;             // a queued row carries the mask of its candidates to evaluate
            bool hit = false;
            uint32_t dig[4] = {0, 0, 0, 0};                   // 4 bits per camera: the person indices of the row I scan
            radix_add<CMAX>(dig, (uint32_t)ttid, s_n, s_inv, C);
            uint32_t rw = (CMAX <= 8) ? nibble_reverse(dig[0] << (4 * (8 - C))) + biasw : 0u;   // the same, biased, reversed
            unsigned long long rbase = 0;                     // next row block to scan
            unsigned int qhead = 0, qn = 0;                   // the team's row queue (ring): team-uniform
            unsigned long long my_dead = kNoRow;              // my last row that counts as evaluated but was dropped unsolved
            unsigned long long surv_row = kNoRow;             // last row that went through the exact evaluation ...
            double surv_err = inf64();                        // ... and its error (team-uniform)
            unsigned long long hit_row = 0;

            while (!hit) {
                // ---- scan rows into the queue until a full batch is waiting (or the rows are used up) -----------------
                while (qn < (unsigned)TEAM_THREADS && rbase < total_rows) {
                    const unsigned long long r = rbase + ttid;
                    unsigned long long smask = 0ULL;          // candidates of my row that need their exact error
                    if (r < total_rows) {
                        uint32_t active = 0, nanact = 0;
#pragma unroll
                        for (int j = 0; j < CMAX / 2; ++j) {
                            const uint32_t t = s_pair[j * 256 + ((dig[j >> 2] >> ((j & 3) * 8)) & 255u)];
                            active |= (t & 3u) << (2 * j);
                            nanact |= ((t >> 2) & 3u) << (2 * j);
                        }
                        const int na = __popc(active);
                        if (na >= a.min_cams && k <= na) {
                            if (!filter) {
                                smask = ~0ULL;
                            } else {
                                Sym4 Mrow;
                                row_matrix(Mrow, dig, active);
                                bool dead_eval = false;       // some candidate has a definite error (so the row is an evaluated one)
                                // can the candidate with normal matrix M and m cameras be under the threshold?  (0 no, 1 maybe, 2 NaN)
                                auto maybe_hit = [&](const Sym4 &M, int m) -> int {
                                    double lam, x0, x1, x2;
                                    SecularBound B;
                                    if (secular_first(M, lam, x0, x1, x2, B) == 2) return 2;  // NaN candidate: ignored like in the exact pass
                                    // E^2 >= s / ((k1 |q0| + k2)^2 + k1^2 s tau) / (m whi)^2, see the comment above the kernel
                                    const double s_lo = fmax(fma(-2e-9, M.m33, B.s), 0.0);   // rounding of c - b.A^-1 b (|b.A^-1 b| <= c)
                                    const double d0 = fma(a.bnd_k1, sqrt_fast(B.g - 1.0), a.bnd_k2);
                                    const double mw = whi_d * (double)m;
                                    const double den = fma(d0, d0, k1sq * s_lo * B.tau) * (mw * mw);
                                    return (s_lo > thr2 * den) ? 0 : 1;                       // NaN compares false: then it is solved
                                };
                                if (k == 0) {                 // one candidate: all active cameras
                                    if (na < 2) dead_eval = na == 1;                          // +inf (na == 0: NaN, ignored)
                                    else if (nanact) dead_eval = true;                        // +inf
                                    else {
                                        const int h = maybe_hit(Mrow, na);
                                        if (h == 0) { dead_eval = true; redo = true; }
                                        else if (h == 1) smask = 1ULL;
                                    }
                                } else
                                for (uint32_t ci = 0; ci < ncand_all; ++ci) {
                                    const uint32_t cm = tabled ? __ldg(table + ci) : unrank_subset(C, k, ci);
                                    if (cm & ~active) continue;
                                    const uint32_t valid = active & ~cm;
                                    const int m = __popc(valid);
                                    if (m < 2) { dead_eval |= (m == 1); continue; }          // +inf (m == 0: NaN, ignored)
                                    if (valid & nanact) { dead_eval = true; continue; }       // +inf
                                    Sym4 M = Mrow;
                                    uint32_t bits = cm;
                                    while (bits) {
                                        const int c = __ffs(bits) - 1;
                                        bits &= bits - 1;
                                        const double2 *src = reinterpret_cast<const double2 *>(sblk + (size_t)(c * NP + digit_rt(dig, c)) * 10);
                                        const double2 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4];
                                        M.m00 -= v0.x; M.m01 -= v0.y; M.m02 -= v1.x; M.m03 -= v1.y; M.m11 -= v2.x;
                                        M.m12 -= v2.y; M.m13 -= v3.x; M.m22 -= v3.y; M.m23 -= v4.x; M.m33 -= v4.y;
                                    }
                                    const int h = maybe_hit(M, m);
                                    if (h == 0) { dead_eval = true; redo = true; }
                                    else if (h == 1) smask |= masked ? (1ULL << ci) : ~0ULL;  // needs its exact error
                                }
                                if (smask == 0ULL && dead_eval) my_dead = r;
                            }
                        }
                    }
                    // append the surviving rows in row order
                    const bool survive = smask != 0ULL;
                    const uint32_t sb = __ballot_sync(P2S_FULL, survive);
                    unsigned int off = 0, tot = (unsigned)__popc(sb);
                    if (NW > 1) {
                        uint32_t *wc = T.wcount[(unsigned)(rbase / (unsigned)TEAM_THREADS) & 1u];
                        if (lane == 0) wc[warp] = tot;
                        __syncthreads();
                        uint32_t inc = lane < NW ? wc[lane] : 0u;               // inclusive scan over the team's warps
#pragma unroll
                        for (int d = 1; d < NW; d <<= 1) {
                            const uint32_t o = __shfl_up_sync(P2S_FULL, inc, d);
                            if (lane >= d) inc += o;
                        }
                        tot = __shfl_sync(P2S_FULL, inc, NW - 1);
                        off = warp ? __shfl_sync(P2S_FULL, inc, warp - 1) : 0u;
                    }
                    if (survive) {
                        const unsigned int slot = (qhead + qn + off + (unsigned)__popc(sb & ((1u << lane) - 1u))) % (unsigned)QCAP;
                        q_row[slot] = r;
                        q_mask[slot] = smask;
#pragma unroll
                        for (int w = 0; w < DW; ++w) q_dig[slot * DW + w] = dig[w];
                    }
                    qn += tot;
                    rbase += TEAM_THREADS;
                    if (CMAX <= 8) {
                        rw = swar_step(rw, stepw, biasw);
                        dig[0] = nibble_reverse(rw - biasw) >> (4 * (8 - C));
                    } else {
                        radix_add<CMAX>(dig, (uint32_t)TEAM_THREADS, s_n, s_inv, C);
                    }
                }
                if (qn == 0) break;                           // every row of the level has been looked at
                team_sync<NW>();                              // the queued rows are visible to their evaluators

                // ---- exact evaluation of the first batch of queued rows, in row order -------------------------------
                const unsigned int nb = qn < (unsigned)TEAM_THREADS ? qn : (unsigned)TEAM_THREADS;
                const bool row_ok = (unsigned)ttid < nb;
                const unsigned int eslot = (qhead + (unsigned)ttid) % (unsigned)QCAP;
                const unsigned long long er = row_ok ? q_row[eslot] : 0ULL;
                const unsigned long long emask = row_ok ? q_mask[eslot] : 0ULL;
                uint32_t edig[4] = {0, 0, 0, 0};
#pragma unroll
                for (int w = 0; w < DW; ++w) edig[w] = row_ok ? q_dig[eslot * DW + w] : 0u;
                uint32_t active = 0, nanact = 0;              // nanact: active cameras whose chosen detection holds a NaN
#pragma unroll
                for (int j = 0; j < CMAX / 2; ++j) {
                    const uint32_t t = s_pair[j * 256 + ((edig[j >> 2] >> ((j & 3) * 8)) & 255u)];
                    active |= (t & 3u) << (2 * j);
                    nanact |= ((t >> 2) & 3u) << (2 * j);
                }
                const int na = __popc(active);
                unsigned long long rkey = P2S_KEY_EMPTY;
                double rqx = nan64(), rqy = rqx, rqz = rqx;
                uint32_t rvalid = 0;
                if (row_ok && na >= a.min_cams && k <= na) {
                    // the row's normal matrix: sum of its active cameras' blocks, ascending camera order
                    Sym4 Mrow;
                    row_matrix(Mrow, edig, active);
                    for (uint32_t ci = 0; ci < ncand_all; ++ci) {
                        if (masked && !((emask >> ci) & 1ULL)) continue;      // its bound is above the threshold: not a hit
                        const uint32_t cm = (k == 0) ? 0u : tabled ? __ldg(table + ci) : unrank_subset(C, k, ci);
                        if (cm & ~active) continue;           // not a subset of the active cameras
                        const uint32_t valid = active & ~cm;
                        const int m = __popc(valid);
                        double cqx, cqy, cqz, e;
                        if (m < 2) {
                            cqx = cqy = cqz = nan64();
                            e = (m == 0) ? nan64() : inf64();
                        } else if (valid & nanact) {           // a kept camera observes NaN: Q = NaN, every distance +inf
                            cqx = cqy = cqz = nan64();
                            e = inf64();
                        } else {
                            Sym4 M = Mrow;
                            uint32_t bits = cm;
                            while (bits) {
                                const int c = __ffs(bits) - 1;
                                bits &= bits - 1;
                                const double2 *src = reinterpret_cast<const double2 *>(sblk + (size_t)(c * NP + digit_rt(edig, c)) * 10);
                                const double2 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4];
                                M.m00 -= v0.x; M.m01 -= v0.y; M.m02 -= v1.x; M.m03 -= v1.y; M.m11 -= v2.x;
                                M.m12 -= v2.y; M.m13 -= v3.x; M.m22 -= v3.y; M.m23 -= v4.x; M.m33 -= v4.y;
                            }
                            if (WIDE) {
                                Tri4 R4;
                                tri4_zero(R4);
#pragma unroll 1
                                for (int c = 0; c < C; ++c) {
                                    if (!((valid >> c) & 1u)) continue;
                                    const float4 o = sobs[c * NP + digit_rt(edig, c)];
                                    givens_add_camera(R4, sP + c * 12, (double)o.x, (double)o.y, (double)o.z);
                                }
                                smallest_singvec_jacobi(R4, cqx, cqy, cqz);
                            } else {
                                smallest_eigvec_secular(M, cqx, cqy, cqz);
                            }
                            double sum = 0.0;
#pragma unroll
                            for (int c = 0; c < CMAX; ++c) {
                                const float4 o = sobs[c * NP + digit_of<CMAX>(edig, c)];
                                const double dist = reproj_distance(cams.P[c], cqx, cqy, cqz, (double)o.x, (double)o.y);
                                sum += ((valid >> c) & 1u) ? dist : 0.0;
                            }
                            e = div_small(sum, (double)m);
                        }
                        ++st_cands;
                        const unsigned long long key = err_key(e);
                        if (key < rkey) { rkey = key; rqx = cqx; rqy = cqy; rqz = cqz; rvalid = valid; }
                    }
                }
                // all-NaN rows are skipped (:235-236); rows without candidates too
                const bool evaluated = rkey < P2S_KEY_NAN;
                const double rerr = key_err(rkey);
                if (filter && evaluated && rerr == a.thr) redo = true;   // '>' against the threshold decides the next level (:194)
                const uint32_t evalmask = __ballot_sync(P2S_FULL, evaluated);
                const uint32_t hitmask = __ballot_sync(P2S_FULL, evaluated && rerr < a.thr);
                // ---- which rows of this batch count: everything up to the first row under the threshold ------
                int wh = NW;                                  // first warp of the team with a hit
                if (NW == 1) {
                    if (hitmask) wh = 0;
                } else {
                    if (lane == 0) { T.eval[warp] = evalmask; T.hit[warp] = hitmask; }
                    __syncthreads();
#pragma unroll
                    for (int w = NW - 1; w >= 0; --w) if (T.hit[w]) wh = w;
                }
                uint32_t upto = 0xffffffffu;                  // my warp's lanes at or before the first hit
                if (warp > wh) upto = 0u;
                else if (warp == wh) { const int fh = __ffs(hitmask) - 1; upto = (fh == 31) ? 0xffffffffu : ((2u << fh) - 1u); }
                if (wh < NW) hit = true;
                const uint32_t considered = evalmask & upto;
                // warp arg-min over (key, lane) among the considered lanes: min of the high words (one redux.sync), the low
                // words only when several lanes hold it; the lowest lane among the holders is the first row in visiting order
                unsigned long long ck = ((considered >> lane) & 1u) ? rkey : P2S_KEY_EMPTY;
                int cl;
                {
                    const uint32_t hi = (uint32_t)(ck >> 32), lo = (uint32_t)ck;
                    const uint32_t mh = __reduce_min_sync(P2S_FULL, hi);
                    uint32_t holders = __ballot_sync(P2S_FULL, hi == mh);
                    uint32_t ml = lo;
                    if (__popc(holders) > 1) {                 // warp-uniform
                        ml = __reduce_min_sync(P2S_FULL, hi == mh ? lo : 0xffffffffu);
                        holders = __ballot_sync(P2S_FULL, hi == mh && lo == ml);
                    }
                    cl = __ffs(holders) - 1;
                    ck = ((unsigned long long)mh << 32) | __shfl_sync(P2S_FULL, ml, cl);
                }
                if (NW == 1) {
                    if (considered) {
                        // error and index of the last evaluated row in visiting order
                        const int ll = 31 - __clz(considered);
                        surv_err = __shfl_sync(P2S_FULL, rerr, ll);
                        surv_row = __shfl_sync(P2S_FULL, er, ll);
                        if (ck < best_key) {                  // strict '<' (:242)
                            best_key = ck;
                            bqx = __shfl_sync(P2S_FULL, rqx, cl);
                            bqy = __shfl_sync(P2S_FULL, rqy, cl);
                            bqz = __shfl_sync(P2S_FULL, rqz, cl);
                            b_valid = __shfl_sync(P2S_FULL, rvalid, cl);
#pragma unroll
                            for (int w = 0; w < 4; ++w) b_dig[w] = __shfl_sync(P2S_FULL, edig[w], cl);
                        }
                    }
                } else {
                    if (lane == 0) { T.any[warp] = considered; T.key[warp] = ck; }
                    if (considered && lane == 31 - __clz(considered)) { T.last[warp] = rerr; T.lastrow[warp] = er; }
                    if (considered && lane == cl) {
                        T.q[warp][0] = rqx; T.q[warp][1] = rqy; T.q[warp][2] = rqz;
                        T.valid[warp] = rvalid;
#pragma unroll
                        for (int w = 0; w < 4; ++w) T.dig[warp][w] = edig[w];
                    }
                    __syncthreads();
#pragma unroll
                    for (int w = 0; w < NW; ++w) {            // visiting order: ascending warp, first wins ties
                        if (T.any[w]) {
                            surv_err = T.last[w]; surv_row = T.lastrow[w];
                            if (T.key[w] < best_key) {
                                best_key = T.key[w];
                                bqx = T.q[w][0]; bqy = T.q[w][1]; bqz = T.q[w][2];
                                b_valid = T.valid[w];
#pragma unroll
                                for (int j = 0; j < 4; ++j) b_dig[j] = T.dig[w][j];
                            }
                        }
                    }
                    __syncthreads();                          // scratch is rewritten by the next batch
                }
                if (hit) hit_row = surv_row;                  // the last considered row IS the first hit
                qhead = (qhead + nb) % (unsigned)QCAP;
                qn -= nb;
            }

            // ---- end of the level: rows visited, and the error of the LAST evaluated row (:194 tests it against the threshold)
            if (hit) {
                err_last = surv_err;
                hit_found = true;
                if (ttid == 0) st_rows += (unsigned)(hit_row + 1ULL);
            } else {
                if (ttid == 0) st_rows += (unsigned)total_rows;
                // a dropped row's error is above the threshold, whatever it is: +inf stands for it
                unsigned long long dead = my_dead;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(P2S_FULL, dead, off);
                    if (dead == kNoRow || (o != kNoRow && o > dead)) dead = o;
                }
                if (NW > 1) {
                    if (lane == 0) T.deadrow[warp] = dead;
                    __syncthreads();
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        const unsigned long long o = T.deadrow[w];
                        if (dead == kNoRow || (o != kNoRow && o > dead)) dead = o;
                    }
                    __syncthreads();
                }
                if (dead != kNoRow && (surv_row == kNoRow || dead > surv_row)) err_last = inf64();
                else if (surv_row != kNoRow) err_last = surv_err;
            }
        }
        // a filtered search that found no row under the threshold: the result is the minimum over ALL rows (:242-245) — again, unfiltered
        bool again = false;
        if (filter && !hit_found) {
            if (NW == 1) {
                again = __any_sync(P2S_FULL, redo);
            } else {
                if (ttid == 0) T.again = 0u;
                __syncthreads();
                if (redo) T.again = 1u;
                __syncthreads();
                again = T.again != 0u;
                __syncthreads();
            }
        }
        if (!again) break;
        }

        // ---- write the frame's result -----------------------------------------------------------------
        if (a.out_stats) {
            unsigned int sc = st_cands;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sc += __shfl_xor_sync(P2S_FULL, sc, off);
            if (NW == 1) {
                if (lane == 0) { a.out_stats[(long long)f * 2] = st_rows; a.out_stats[(long long)f * 2 + 1] = sc; }
            } else {
                if (lane == 0) { atomicAdd(&T.rows, ttid == 0 ? st_rows : 0u); atomicAdd(&T.cands, sc); }
                __syncthreads();
                if (ttid == 0) { a.out_stats[(long long)f * 2] = T.rows; a.out_stats[(long long)f * 2 + 1] = T.cands; }
            }
        }
        if (ttid == 0) {
            const bool any = best_key != kInfKey;
            a.out_err[f] = any ? key_err(best_key) : inf64();
            double *q = a.out_Q + (long long)f * 3;
            q[0] = any ? bqx : nan64(); q[1] = any ? bqy : nan64(); q[2] = any ? bqz : nan64();
        }
        if (ttid < C) {
            int8_t v = -1;
            if (best_key != kInfKey && ((b_valid >> ttid) & 1u)) {
                const uint32_t w = ttid < 8 ? b_dig[0] : ttid < 16 ? b_dig[1] : ttid < 24 ? b_dig[2] : b_dig[3];
                v = (int8_t)((w >> ((ttid & 7) * 4)) & 15u);
            }
            a.out_comb[(long long)f * C + ttid] = v;
        }
        team_sync<NW>();
    }
}

template <int CMAX, int NW>
static cudaError_t launch_assoc_nw(const AssocLaunch &L, const AssocArgs &a0, int *grid_out) {
    CamParams<CMAX> cams;
    for (int c = 0; c < CMAX; ++c)
        for (int j = 0; j < 12; ++j) cams.P[c][j] = (c < L.n_cams) ? L.P[c * 12 + j] : 0.0;
    AssocArgs a = a0;
    constexpr int teams = (NW == 1) ? 4 : 1;
    constexpr int threads = (NW == 1) ? 128 : 32 * NW;
    const size_t smem = (size_t)CMAX * 12 * sizeof(double) +
                        (assoc_slab_bytes(CMAX, L.max_persons) + sizeof(TeamScratch<NW>) + assoc_queue_bytes(CMAX, NW)) * teams;
    auto kern = associate_kernel<CMAX, NW, false>;
    auto kern_wide = associate_kernel<CMAX, NW, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long want = (L.n_frames + teams - 1) / teams;
    long long grid = (long long)L.sm_count * per_sm;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    if (grid_out) *grid_out = (int)grid;
    kern<<<(unsigned)grid, threads, smem, L.stream>>>(cams, a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    kern_wide<<<(unsigned)grid, threads, smem, L.stream>>>(cams, a);       // returns at once unless frames were flagged
    return cudaGetLastError();
}

// A warp per frame when there are enough frames to fill the machine (or the frames are small), else a
// 256-thread CTA per frame.  `mean_rows`: average size of the person-combination product per frame.
template <int CMAX>
static cudaError_t launch_assoc(const AssocLaunch &L, const AssocArgs &a0, int *grid_out) {
    // fewer frames than SMs and many rows per frame (BASELINE configs[3] on a short trial): a 512-thread CTA per frame
    // puts 16 warps on every busy SM instead of 8
    const bool wide16 = L.team == 16 || (L.team == 0 && L.mean_rows >= 4096.0 && L.n_frames <= (long long)L.sm_count);
    const bool wide = L.team == 8 || (L.team == 0 && L.mean_rows >= 512.0 && (double)L.n_frames < 16.0 * L.sm_count * 4);
    if (wide16) return launch_assoc_nw<CMAX, 16>(L, a0, grid_out);
    if (wide) return launch_assoc_nw<CMAX, 8>(L, a0, grid_out);
    return launch_assoc_nw<CMAX, 1>(L, a0, grid_out);
}

cudaError_t launch_associate(const AssocLaunch &L, int *grid_out) {
    AssocArgs a;
    a.obs = (const float4 *)L.obs; a.count = L.count; a.n_frames = L.n_frames;
    a.n_cams = L.n_cams; a.max_persons = L.max_persons; a.min_cams = L.min_cams;
    a.thr = L.thr; a.lik_thr = L.lik_thr; a.cand_masks = L.cand_masks;
    for (int i = 0; i < P2S_MAX_CAMS + 2; ++i) a.level_off[i] = L.level_off[i];
    a.max_table_level = L.max_table_level;
    a.out_err = L.out_err; a.out_comb = L.out_comb; a.out_Q = L.out_Q; a.out_stats = L.out_stats;
    a.tile_counter = L.tile_counter;
    a.wide_flags = L.wide_flags;
    a.bnd_k1 = 0.0; a.bnd_k2 = 0.0;
    for (int c = 0; c < L.n_cams; ++c) {
        const double *p2 = L.P + c * 12 + 8;
        a.bnd_k1 = std::fmax(a.bnd_k1, std::sqrt(p2[0] * p2[0] + p2[1] * p2[1] + p2[2] * p2[2]));
        a.bnd_k2 = std::fmax(a.bnd_k2, std::fabs(p2[3]));
    }
    a.bnd_k1 *= 1.0 + 1e-12; a.bnd_k2 *= 1.0 + 1e-12;
    if (!(a.bnd_k1 < INFINITY) || !(a.bnd_k2 < INFINITY)) a.bnd_k1 = a.bnd_k2 = INFINITY;   // NaN / inf matrices: bound 0, no row dropped
    a.search_mode = L.search_mode;
    if (L.n_cams <= 4) return launch_assoc<4>(L, a, grid_out);
    if (L.n_cams <= 8) return launch_assoc<8>(L, a, grid_out);
    if (L.n_cams <= 16) return launch_assoc<16>(L, a, grid_out);
    return launch_assoc<32>(L, a, grid_out);
}

}  // namespace p2s
