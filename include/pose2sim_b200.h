/*
 * pose2sim_b200 — C ABI of the B200 (sm_100a) triangulation / person-association hot path.
 *
 * The reference (sakagawa-star/pose2sim) is pure Python and has no FFI; the seam this library
 * replaces is the per-unit Python function boundary, batched:
 *
 *   p2s_triangulate_*   replaces  Pose2Sim/triangulation.py:363  triangulation_from_best_cameras
 *                       (called once per frame x person x keypoint at triangulation.py:840), with the
 *                       leaf math of Pose2Sim/common.py:327 weighted_triangulation, :357 reprojection,
 *                       :378 euclidean_distance;
 *   p2s_stage_*         replaces  the likelihood gate at triangulation.py:817-821 plus the
 *                       (3, n_cams) slicing at triangulation.py:837-838;
 *   p2s_associate_*     replaces  Pose2Sim/personAssociation.py:154  best_persons_and_cameras_combination
 *                       (+ :67 persons_combinations, :102 triangulate_comb), called per frame at :774;
 *   p2s_associate_multi_* replaces  personAssociation.py:793-801 (compute_affinity :347, matchSVT :450,
 *                       the arg-max half of person_index_per_cam :512), the multi_person branch.
 *
 * Conventions
 *   - plain C, no torch / C++ types; all sizes are explicit;
 *   - `*_device` entry points take DEVICE pointers, enqueue work on `stream` (a cudaStream_t passed
 *     as void*, NULL = default stream) and return immediately;
 *   - `*_host` entry points take HOST pointers, do host<->device copies on internal streams and
 *     return when the results are in the caller's buffers;
 *   - return value: 0 = ok, otherwise a P2S_E* code (p2s_status_string gives text).  Argument errors
 *     are reported that way; numerical failure of a unit is NaN in its outputs, as in the reference
 *     (triangulation.py:600-602);
 *   - no global state: everything lives in the p2s_handle (one per GPU per host thread).
 *   - There is NO CPU fallback: without a CUDA device p2s_create fails with P2S_ENODEVICE.
 */
#ifndef POSE2SIM_B200_H
#define POSE2SIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2S_MAX_CAMS 32          /* one uint32 exclusion mask per unit */
#define P2S_MAX_PERSONS 16       /* persons per camera in the association search */
#define P2S_MAX_PEERS 16         /* GPUs of one node that can push into one consumer */
#define P2S_IPC_HANDLE_BYTES 64  /* sizeof(cudaIpcMemHandle_t) */
#define P2S_MAX_DETECTIONS 64    /* detections per frame (all cameras together) in the multi-person matching */

enum {
    P2S_OK = 0,
    P2S_EINVAL = 1,              /* bad argument (n_cams outside 2..32, min_cams < 1, null pointer ...) */
    P2S_ENODEVICE = 2,           /* no usable CUDA device / wrong architecture (needs sm_100) */
    P2S_ECUDA = 3,               /* a CUDA call failed; p2s_last_cuda_error() has the text */
    P2S_ENOMEM = 4,
    P2S_ETOODEEP = 5             /* search level needs more candidates than the engine enumerates */
};

typedef struct p2s_handle p2s_handle;

/* Layout of the statistics block written by the triangulation kernels (uint64 each).           */
enum {
    P2S_STAT_LEVEL0 = 0,         /* [0..32]  units whose LAST EVALUATED exclusion level was k     */
    P2S_STAT_NOT_EVALUATED = 33, /* units for which no level could be evaluated (nexcl = n_cams)  */
    P2S_STAT_FAILED = 34,        /* units returned as NaN (error above threshold at the end)      */
    P2S_STAT_CANDIDATES = 35,    /* candidate camera subsets solved (DLT + reprojection)          */
    P2S_STAT_CAM_SOLVES = 36,    /* sum over solved candidates of their number of valid cameras   */
    P2S_STAT_BAND_THRESHOLD = 37,/* units with |error_min - threshold| < eps at an evaluated level */
    P2S_STAT_BAND_ARGMIN = 38,   /* units whose best and second-best DISTINCT candidate errors at
                                    an evaluated level differ by < eps                            */
    P2S_STAT_NEWTON_STEPS = 39,  /* eigen-solver iterations summed over candidates                */
    P2S_STAT_SOLVED = 40,        /* candidates with >= 2 valid cameras (eigen-solve + reprojection)  */
    P2S_STAT_DIRECT_CAMS = 41,   /* level 0: valid cameras accumulated straight into the normal matrix */
    P2S_STAT_BLOCKS = 42,        /* levels >= 1: per-camera 4x4 blocks built (valid cameras x unit x level) */
    P2S_STAT_ENTRY_ADDS = 43,    /* levels >= 1: FP64 additions forming M_all and M_all -/+ blocks     */
    P2S_STAT_WIDE_UNITS = 44,    /* units whose valid likelihoods span more than 256x: solved from a QR
                                    factorisation of A (like the reference's SVD) by the fix-up kernel   */
    P2S_STAT_COUNT = 48
};

/* Lens model of one camera for `[triangulation] undistort_points = true` (calibration TOML values;
 * common.py:254-288 retrieve_calib_params).  dist = k1 k2 p1 p2 k3 k4 k5 k6 in OpenCV order, unused
 * coefficients 0; newK = cv2.getOptimalNewCameraMatrix(K, dist, size, 1, size), the matrix the
 * reference builds P from (common.py:310-313) and re-projects the undistorted points with.        */
typedef struct p2s_camera_model {
    double K[9];
    double dist[8];
    double R[9];
    double T[3];
    double newK[9];
} p2s_camera_model;

typedef struct p2s_device_info {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    int clock_khz;               /* SM clock reported by the runtime */
    size_t total_mem;
    char name[128];
} p2s_device_info;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* A handle is used by ONE host thread at a time.  Every launch draws its tile / arrival counters from a ring of 256
 * entries owned by the handle: at most 256 launches of the *_device entry points may be in flight (enqueued on user
 * streams and not yet finished) per handle; a caller that pipelines deeper than that creates a second handle.   */
int p2s_create(int device, p2s_handle **out);
int p2s_destroy(p2s_handle *h);
const char *p2s_status_string(int status);
const char *p2s_last_cuda_error(const p2s_handle *h);
int p2s_get_device_info(const p2s_handle *h, p2s_device_info *info);
/* eps of the two decision bands counted in the statistics block (default 1e-6 px). */
int p2s_set_band_eps(p2s_handle *h, double eps_px);
/* eigen-solver selection: 0 = safeguarded secular Newton (default), 1 = cyclic Jacobi sweeps
 * (the north-star's nominal solver; kept for A/B evidence). */
int p2s_set_solver(p2s_handle *h, int solver);

/* how the single-person association search (p2s_associate_*) visits its rows: 0 (default) = candidates whose lower bound
 * (one factorisation of the eigen-solve) is above the threshold are not solved and re-projected; a frame whose search
 * finds no row under the threshold is searched again exhaustively, so the result is the exhaustive search's;
 * 1 = every row is solved and re-projected (personAssociation.py:196-248 as written).                               */
int p2s_set_search_mode(p2s_handle *h, int mode);

/* deep levels of the exclusion search (triangulation.py:408-411: level k enumerates ALL C(n_cams, k) camera subsets).
 * A unit that is pending at a level of at least `min_candidates` subsets is not walked by the one warp that holds its
 * tile: it is parked, and a second kernel behind the search kernel gives every parked unit a cluster of two 512-thread CTAs (same
 * arithmetic, same outputs bit for bit).  Default 2048 — C(16, 5) = 4368, C(32, 3) = 4960: rigs of 14 cameras and up;
 * 0 = never park (the single-kernel search); small values exercise the path on small rigs (tests).  Launches that ask
 * for the statistics block, the Jacobi solver or a lens model never park.                                            */
int p2s_set_deep_search(p2s_handle *h, long long min_candidates);

/* how the triangulation kernels write a full 32-unit tile of outputs: 0 (default) = 16-byte vector stores from the
 * warp's staging area, 1 = four cp.async.bulk (TMA) stores per tile issued by one lane (A/B knob for the push path),
 * 2 = the pooled kernel where it applies (p2s_triangulate_planes_device without statistics, 4 or 8 cameras): level-1
 * passes are shared by the pending units of several tiles, whose outputs are overwritten unit by unit after the tile's
 * vector stores; same results bit for bit, measured equal in time to mode 0, kept as an A/B knob     */
int p2s_set_output_mode(p2s_handle *h, int mode);

/* how p2s_triangulate_host moves its data: 0 (default) = zero-copy when every buffer is pinned host memory (one
 * kernel reads its tiles from and writes its results to host memory over PCIe), else the chunked copy pipeline;
 * 1 = always the copy pipeline; 2 = zero-copy or P2S_EINVAL */
int p2s_set_host_mode(p2s_handle *h, int mode);

/* units per H2D -> search -> D2H pipeline chunk of the *_host entry points (4 chunks in flight);
 * 0 (default) = automatic, about a quarter of the call's units clamped to [2^16, 2^20] */
int p2s_set_chunk_units(p2s_handle *h, long long units);

/* association search: warps that share one frame.  0 (default) = automatic: a warp per frame when frames are plentiful,
 * a 256-thread CTA per frame when the person-combination product is large and frames are few, a 512-thread CTA per
 * frame when there are fewer frames than SMs; 1, 8 or 16 force one of the three (tests, A/B). */
int p2s_set_assoc_team(p2s_handle *h, int warps_per_frame);

/* ---- staging (triangulation.py:817-821 + layout) -------------------------------------------- *
 * x, y, lik: [n_units][n_cams] float32, row-major (unit = frame x person x keypoint).
 * obs_out  : float4 [n_cams][n_units] = {x, y, likelihood, 0}; likelihood < lik_thr (and not NaN)
 *            turns the whole triple into NaN, exactly like the reference's gate.
 *            Pass lik_thr = -INFINITY (or NaN) to stage without gating.                          */
size_t p2s_obs_bytes(long long n_units, int n_cams);
int p2s_stage_observations_device(p2s_handle *h, const float *x, const float *y, const float *lik,
                                  long long n_units, int n_cams, double lik_thr,
                                  void *obs_out, void *stream);

/* ---- triangulation with camera-exclusion search --------------------------------------------- *
 * obs        : staged float4 [n_cams][n_units] (device)
 * P          : HOST pointer, n_cams x 12 float64 (row-major 3x4 per camera, common.py:291 computeP)
 * out_Q      : [n_units][3] float64, NaN when the unit failed
 * out_err    : [n_units]    float64 reprojection error (px), NaN when failed
 * out_nexcl  : [n_units]    uint8   nb_cams_excluded (NaN-or-zero likelihood count of the chosen subset)
 * out_mask   : [n_units]    uint32  bit c set <=> camera c in id_excluded_cams
 * stats      : device pointer to P2S_STAT_COUNT uint64 (accumulated, caller zeroes) or NULL      */
int p2s_triangulate_device(p2s_handle *h, const void *obs, const double *P,
                           long long n_units, int n_cams, double reproj_thr, int min_cams,
                           double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                           unsigned long long *stats, void *stream);

/* Same search straight from the raw planes x, y, lik [n_units][n_cams] (DEVICE pointers, 16-byte
 * aligned): the likelihood gate and the float4 SoA staging happen inside the kernel's tile load, so no
 * staged buffer is written to or read from HBM.                                                  */
int p2s_triangulate_planes_device(p2s_handle *h, const float *x, const float *y, const float *lik,
                                  const double *P, long long n_units, int n_cams, double lik_thr,
                                  double reproj_thr, int min_cams,
                                  double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                                  unsigned long long *stats, void *stream);

/* Whole job from host buffers.  x/y/lik as for p2s_stage_observations_device but HOST pointers; outputs HOST
 * pointers; stats: HOST pointer to P2S_STAT_COUNT uint64 (overwritten) or NULL.  Pinned (page-locked) buffers are
 * read and written by the kernel directly (zero-copy, one launch); pageable buffers go through a chunked
 * H2D -> search -> D2H pipeline on internal streams (p2s_set_host_mode).                                */
int p2s_triangulate_host(p2s_handle *h, const float *x, const float *y, const float *lik,
                         const double *P, long long n_units, int n_cams,
                         double lik_thr, double reproj_thr, int min_cams,
                         double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                         unsigned long long *stats);

/* `undistort_points = true` variants (triangulation.py:808-813, :472-476).  `lens`: HOST pointer to
 * n_cams models.  Staging additionally undistorts x, y like cv2.undistortPoints(float32 points, K, dist,
 * None, newK) before the gate; the search measures the error between the undistorted observation and the
 * DISTORTED re-projection (cv2.projectPoints with K, dist), exactly as the reference does.  P must be
 * the projection built on newK.                                                                  */
int p2s_stage_undistort_device(p2s_handle *h, const float *x, const float *y, const float *lik,
                               long long n_units, int n_cams, double lik_thr, const p2s_camera_model *lens,
                               void *obs_out, void *stream);
int p2s_triangulate_distorted_device(p2s_handle *h, const void *obs, const double *P, const p2s_camera_model *lens,
                                     long long n_units, int n_cams, double reproj_thr, int min_cams,
                                     double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                                     unsigned long long *stats, void *stream);
int p2s_triangulate_undistort_host(p2s_handle *h, const float *x, const float *y, const float *lik,
                                   const double *P, const p2s_camera_model *lens, long long n_units, int n_cams,
                                   double lik_thr, double reproj_thr, int min_cams,
                                   double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                                   unsigned long long *stats);

/* `handle_LR_swap = true` (triangulation.py:509-579): the same search with the left/right swapped evaluation
 * after every level whose error is still above the threshold — as the reference EXECUTES it (its sub-configuration
 * arrays alias, :518-526, so each candidate gets ONE swapped evaluation: the first n_cams - nb_cams_off_tot valid
 * cameras take the partner keypoint's coordinates and only they enter the error; nb_cams_excluded keeps the
 * un-swapped winner's count, :574-577).
 * obs        : staged float4 [n_cams][n_units] (device), units ordered (frame, person, keypoint), so the partner
 *              of unit u is u - u % n_keypoints + partner[u % n_keypoints]  (triangulation.py:838)
 * partner    : DEVICE int32 [n_keypoints], 0 <= partner[k] < n_keypoints (keypoints_idx_swapped, :742-745);
 *              read back and checked (P2S_EINVAL), which synchronises `stream` once per call
 * lens       : HOST pointer to n_cams models (undistort_points) or NULL                                        */
int p2s_triangulate_lrswap_device(p2s_handle *h, const void *obs, const int32_t *partner, int n_keypoints,
                                  const double *P, const p2s_camera_model *lens, long long n_units, int n_cams,
                                  double reproj_thr, int min_cams,
                                  double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                                  void *stream);

/* ---- single-person association search (personAssociation.py:154) ---------------------------- *
 * obs        : float4 [n_frames][n_cams][max_persons] = {x, y, likelihood, 0} of the tracked keypoint
 *              (device for *_device, host for *_host); entries >= count are ignored
 * count      : int32 [n_frames][n_cams] persons detected per camera (0 = none)
 * out_err    : [n_frames] float64 best error (+inf when nothing was evaluable)
 * out_comb   : [n_frames][n_cams] int8 chosen person index per camera, -1 = camera off (NaN)
 * out_Q      : [n_frames][3] float64
 * out_stats  : [n_frames][2] uint32 {combination rows visited, candidates solved} or NULL        */
int p2s_associate_device(p2s_handle *h, const void *obs, const int32_t *count, const double *P,
                         long long n_frames, int n_cams, int max_persons,
                         double reproj_thr, double lik_thr, int min_cams,
                         double *out_err, int8_t *out_comb, double *out_Q, uint32_t *out_stats,
                         void *stream);
int p2s_associate_host(p2s_handle *h, const float *obs, const int32_t *count, const double *P,
                       long long n_frames, int n_cams, int max_persons,
                       double reproj_thr, double lik_thr, int min_cams,
                       double *out_err, int8_t *out_comb, double *out_Q, uint32_t *out_stats);

/* ---- multi-person cross-view association (personAssociation.py:783-801) ---------------------- *
 * Per frame: Plücker-ray affinity between every two detections of different cameras (compute_rays :277,
 * compute_affinity :347), the one-person-per-view constraint (:411), matchSVT (:450, max_iter 20,
 * w_rank 50, tol 1e-4, w_sparse 0.1 as at the call site :799), the min_affinity cut (:800) and the first
 * half of person_index_per_cam (:526-531): for every detection (row) the arg-max detection of each view.
 * The integer bookkeeping that follows in the reference (unique rows, ordering, duplicate and
 * min-cameras filters, :535-547) is left to the caller.
 * obs          : float32 [n_frames][n_cams][max_persons][3 n_joints] = pose_keypoints_2d of each detection
 *                (x, y, likelihood per joint; NaN allowed); entries >= count are ignored
 * count        : int32 [n_frames][n_cams] detections per camera
 * cams         : HOST pointer, n_cams models; K, R (world -> camera rotation matrix) and T are used
 * n_max        : row capacity per frame, >= the largest per-frame sum of count, <= P2S_MAX_DETECTIONS
 * out_rows     : int8 [n_frames][n_max][n_cams]; row r < sum(count[f]) holds per view the index (within
 *                that view) of the best-matching detection, -1 when none has affinity > 0; rows beyond
 *                the frame's detections are left untouched
 * out_affinity : float64 [n_frames][n_max][n_max] matched and thresholded affinity (top-left N x N used) or NULL
 * out_iters    : int32 [n_frames] matchSVT iterations used, or NULL                                  */
int p2s_associate_multi_device(p2s_handle *h, const float *obs, const int32_t *count, const p2s_camera_model *cams,
                               long long n_frames, int n_cams, int max_persons, int n_joints, int n_max,
                               double reconstruction_error_threshold, double min_affinity,
                               int8_t *out_rows, double *out_affinity, int32_t *out_iters, void *stream);
int p2s_associate_multi_host(p2s_handle *h, const float *obs, const int32_t *count, const p2s_camera_model *cams,
                             long long n_frames, int n_cams, int max_persons, int n_joints, int n_max,
                             double reconstruction_error_threshold, double min_affinity,
                             int8_t *out_rows, double *out_affinity, int32_t *out_iters);

/* ---- multi-GPU: the final gather fused into the search kernel (one node, NVLink / NVSwitch) ---- *
 * Units shard over GPUs without any data-path exchange (triangulation.py:831-845 keeps no cross-unit state);
 * the only exchange is the gather of the 37 bytes per unit to the rank that writes the TRC.  Instead of a
 * collective after the kernel, the producer's kernel stores its outputs straight into the CONSUMER's memory
 * (peer mapping, full 16-byte vectors per 32-unit tile) while it computes, and its last CTA raises an arrival
 * flag there; the consumer releases a buffer for reuse by writing an acknowledgement flag back.
 *
 * p2s_peer_alloc   : device buffer that other processes of this node can map; `handle` is what they need
 *                    (exchange it with any host-side channel, e.g. torch.distributed / MPI / a pipe)
 * p2s_peer_open    : map a buffer exported by another process (peer access is enabled on demand)
 * p2s_peer_close / p2s_peer_free : undo the above
 * p2s_triangulate_planes_push_device : p2s_triangulate_planes_device whose out_* may point into a mapped peer
 *                    buffer; before the first store the kernel waits until *wait_flag >= wait_value (LOCAL flag,
 *                    NULL = no wait; wrap-around safe); after the last store it sets *done_flag = done_value
 *                    (local or peer flag, NULL = none) with system-scope release ordering
 * p2s_peer_collect_device : consumer side, one tiny kernel on `stream`: waits until arrive[i] >= value for all
 *                    i < n (LOCAL flags), then writes `value` to every non-NULL ack[i] (local or peer flags)
 * p2s_peer_error   : bit 0: a producer's wait timed out (2 s), bit 1: the consumer's; cleared by the call      */
int p2s_peer_alloc(p2s_handle *h, size_t bytes, void **dptr, unsigned char handle[P2S_IPC_HANDLE_BYTES]);
int p2s_peer_open(p2s_handle *h, const unsigned char handle[P2S_IPC_HANDLE_BYTES], void **dptr);
int p2s_peer_close(p2s_handle *h, void *dptr);
int p2s_peer_free(p2s_handle *h, void *dptr);
int p2s_triangulate_planes_push_device(p2s_handle *h, const float *x, const float *y, const float *lik,
                                       const double *P, long long n_units, int n_cams, double lik_thr,
                                       double reproj_thr, int min_cams,
                                       double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                                       unsigned long long *stats,
                                       const unsigned int *wait_flag, unsigned int wait_value,
                                       unsigned int *done_flag, unsigned int done_value, void *stream);
int p2s_peer_collect_device(p2s_handle *h, const unsigned int *arrive, int n, unsigned int value,
                            unsigned int *const *ack, void *stream);
int p2s_peer_error(p2s_handle *h, unsigned int *bits);

/* ---- host staging: OpenPose JSON -> observation planes (no GPU involved) --------------------- *
 * Replaces the file handling of triangulation.py:607-653 extract_files_frame_f (+ :77-90
 * count_persons_in_json): every file is parsed ONCE, on `n_threads` host threads (<= 0: all cores).
 * paths      : n_frames x n_cams C strings, frame-major ("" or unreadable / unparsable => NaN)
 * keypoint_ids: n_keypoints OpenPose ids (skeleton pre-order); value = pose_keypoints_2d[3 id : 3 id + 3]
 * x, y, lik  : float32 [n_frames][n_persons][n_keypoints][n_cams] — the unit-major, camera-fastest
 *              layout p2s_triangulate_host takes; NaN wherever the reference's lookup would raise
 * n_people   : int32 [n_frames][n_cams] len(people) of each file (-1: unparsable), or NULL
 * status     : uint8 [n_frames][n_cams] 1 = parsed, or NULL
 * n_inexact  : number of finite values that float32 cannot represent exactly, or NULL            */
int p2s_read_pose_files(const char *const *paths, long long n_frames, int n_cams,
                        const int32_t *keypoint_ids, int n_keypoints, int n_persons,
                        float *x, float *y, float *lik, int32_t *n_people, uint8_t *status,
                        long long *n_inexact, int n_threads);

/* Directory index (triangulation.py:752-803, common.py:568-583): the camera folders listed, *.json kept and sorted by the
 * last number in the name, and the frame -> file table built with the reference's rule (all files of a camera whose last
 * number is the frame, or none; flattened; first n_cams entries).  p2s_index_open fails with P2S_EINVAL when a folder
 * cannot be listed, p2s_index_build_table when a name holds no number (the reference raises there).  The table is
 * [f1 - f0][n_cams] absolute paths ("" = no file), valid until the next build / close; it is what p2s_read_pose_files
 * takes.  p2s_index_signature: 128-bit digest of every table entry's path, mtime (ns) and size (staging cache key).   */
typedef struct p2s_dir_index p2s_dir_index;
int p2s_index_open(const char *const *dirs, int n_cams, p2s_dir_index **out);
void p2s_index_close(p2s_dir_index *ix);
long long p2s_index_file_count(const p2s_dir_index *ix, int cam);
const char *p2s_index_file_name(const p2s_dir_index *ix, int cam, long long i);
int p2s_index_build_table(p2s_dir_index *ix, long long f0, long long f1);
const char *const *p2s_index_table_paths(const p2s_dir_index *ix);
/* the same table as one buffer: the n_frames x n_cams paths back to back, each NUL-terminated ("" = none); *len = bytes */
const char *p2s_index_table_arena(const p2s_dir_index *ix, long long *len);
int p2s_index_signature(const p2s_dir_index *ix, unsigned long long sig[2], int n_threads);

/* (mtime in ns, size in bytes) of n files, -1 / -1 where stat fails: the signature the staging cache is keyed on
 * (pose2sim_b200/staging.py: parsed trials are kept as memory-mapped float32 planes and reused while no file changed). */
int p2s_stat_files(const char *const *paths, long long n, long long *mtime_ns, long long *size, int n_threads);

/* Association stage, input side (personAssociation.py:67-99 `persons_combinations`, :260-274 `read_json`, re-read there
 * once per person combination): every file parsed once.  paths [n_frames][n_cams] ("" = no file).  Per file:
 * count_named = people whose x values are not all NaN (the index space of the combinations), count_listed = keypoint
 * lists with >= 3 values (the index space the observations are read from), list_len = their common length (0: none,
 * -1: they differ); obs [n_frames][n_cams][max_persons][n_values] = values [value_offset, value_offset + n_values) of
 * the listed people in order, NaN where the list is shorter (n_values = 0: counts only, obs may be NULL).
 * status: 0 unreadable / not JSON, 1 ok, 2 irregular content (left to the host language's own parser, which mirrors
 * the reference's exception handling), 3 more people than max_persons.                                             */
int p2s_read_people_files(const char *const *paths, long long n_frames, int n_cams, int value_offset, int n_values,
                          int max_persons, float *obs, int32_t *count_named, int32_t *count_listed, int32_t *list_len,
                          uint8_t *status, long long *n_inexact, int n_threads);
/* Association stage, output side (personAssociation.py:552-580 `rewrite_json_files`): per (frame, camera) the source
 * document re-emitted exactly as Python's json.dumps(json.load(f)) writes it, with `people` replaced by the chosen
 * person of every proposal of the frame ({} where the camera is off).  comb: one row of n_cams indices per proposal
 * (-1 = off); the proposals of frame f are rows prop_offset[f] .. prop_offset[f + 1].  No readable source or an index
 * past the people list -> no file (a stale one is removed).  status: 1 written, 0 no file, 2 left to the host
 * language (strings with escapes, duplicate keys, no `people` list).                                               */
int p2s_rewrite_people_files(const char *const *src, const char *const *dst, long long n_frames, int n_cams,
                             const int32_t *prop_offset, const int32_t *comb, uint8_t *status, int n_threads);

/* TRC body (triangulation.py:206-213 `Q.to_csv(sep='\t', index=True, header=None)`): APPENDS n_rows lines
 * "frame \t time \t v0 \t ... \n" to `path` (the caller has written the 5 header lines); values is
 * [n_rows][n_cols] float64, NaN = empty field, numbers formatted like Python's repr(float) (what pandas
 * writes), so the file is byte-identical to the reference's for equal values.                      */
int p2s_write_trc_rows(const char *path, const long long *frames, const double *time_s,
                       const double *values, long long n_rows, int n_cols);
/* The same rows formatted into `buf` (cap >= n_rows * (22 + 25 * (n_cols + 1)) bytes always suffices); *len = bytes
 * written.  For frame-block sharded writers: every rank formats its rows, the byte counts are exchanged and each rank
 * writes its range of the one file (triangulation.write_outputs_sharded).                                        */
int p2s_format_trc_rows(const long long *frames, const double *time_s, const double *values, long long n_rows,
                        int n_cols, char *buf, size_t cap, size_t *len);

/* ---- workload generation (benchmarks / tests; produces INPUTS only) ---------------------------- */
/* Synthetic observations of units [unit0, unit0 + n_units) generated on the device as a pure function of
 * (seed, unit, camera): Philox4x32-10 counters + IEEE add / multiply / divide in a fixed order, so that the NumPy
 * twin (pose2sim_b200/synth_philox.py) regenerates any subsample bit for bit for the CPU oracle (SURVEY.md 8(d)/(e):
 * BASELINE configs[4], 10 M frames x up to 32 cameras, is generated per shard instead of being pushed over PCIe).
 * Units are (frame, keypoint), keypoint fastest.  P: host n_cams x 12.  kp_offsets [n_keypoints][3], circle [600][2],
 * dirs [256][2]: DEVICE float64 tables (synth_philox.tables()).  x, y, lik: device float32 [n_units][n_cams], the
 * likelihood gate NOT applied; truth: device float64 [n_units][3] or NULL.  Asynchronous on `stream`.          */
int p2s_synth_observations_device(p2s_handle *h, const double *P, int n_cams, int n_keypoints, unsigned int seed,
                                  long long unit0, long long n_units, double sigma, double p_out, double p_low,
                                  const double *kp_offsets, const double *circle, const double *dirs,
                                  float *x, float *y, float *lik, double *truth, void *stream);

/* ---- measurement helpers -------------------------------------------------------------------- */
/* Dependent-chain FP64 FMA microbenchmark on the handle's device: achieved DFMA TFLOP/s
 * (2 flops per FMA) — the FP64 roofline denominator MEASURED_PEAKS.json does not carry.         */
int p2s_measure_fp64_peak(p2s_handle *h, double *tflops, double *ms);
/* Number of kernels this library launched through the handle since creation.                    */
long long p2s_launch_count(const p2s_handle *h);
/* Grid size (CTAs) of the last triangulation / association kernel launched through the handle:
 * SM count x resident CTAs per SM for a persistent launch (occupancy evidence for the profiles). */
int p2s_last_grid(const p2s_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* POSE2SIM_B200_H */
