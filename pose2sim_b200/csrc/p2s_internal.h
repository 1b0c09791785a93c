// Internal launch descriptors shared by the kernel translation units and the C-ABI layer.
#pragma once
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/pose2sim_b200.h"

namespace p2s {

struct TriLaunch {
    const void *obs;              // float4 [n_cams][n_units], device (null: raw planes below)
    const float *px = nullptr, *py = nullptr, *pl = nullptr;   // raw planes [n_units][n_cams], device
    double lik_thr = -INFINITY;   // gate of the raw-plane path
    const double *P;              // host, n_cams x 12
    const p2s_camera_model *lens; // host, n_cams lens models (undistort_points) or null
    long long n_units;
    int n_cams, min_cams, solver, sm_count;
    double thr, band_eps;
    const uint32_t *cand_masks;   // device
    uint32_t level_off[P2S_MAX_CAMS + 2];
    int max_table_level;
    double *out_Q, *out_err;
    uint8_t *out_nexcl;
    uint32_t *out_mask;
    unsigned long long *stats;    // device or null
    unsigned int *tile_counter;   // device, four words, zeroed on the same stream before the launch
    const unsigned int *wait_flag = nullptr;   // push path (see TriArgs)
    unsigned int wait_value = 0;
    unsigned int *done_flag = nullptr;
    unsigned int done_value = 0;
    unsigned int *err_word = nullptr;
    int bulk_out = 0;             // full tiles leave by TMA bulk stores (p2s_set_output_mode)
    bool allow_pool = false;      // outputs live in this device's memory: the pooled kernel may serve the launch
    bool pool = false;            // ... and is asked for (p2s_set_output_mode(h, 2))
    // deep levels (deep_search_kernel): list of parked (unit, level) records of THIS launch, its capacity, and the
    // candidate count from which a level is parked (0 / null: never); tile_counter[3] counts the parked units
    unsigned long long *deep_list = nullptr;
    unsigned int deep_cap = 0;
    uint32_t deep_min = 0;
    mutable int kernels = 2;      // kernels the launch enqueued (search + fix-up, + deep search when it applies)
    cudaStream_t stream;
};

struct AssocLaunch {
    const void *obs;              // float4 [n_frames][n_cams][max_persons], device
    const int32_t *count;         // [n_frames][n_cams], device
    const double *P;              // host
    long long n_frames;
    int n_cams, max_persons, min_cams, sm_count;
    int team = 0;                 // 0 auto, 1 or 8 warps per frame
    double mean_rows;             // average rows (product of persons per camera) per frame, picks the team width
    double thr, lik_thr;
    const uint32_t *cand_masks;   // same lexicographic all-camera table as the triangulation kernel
    uint32_t level_off[P2S_MAX_CAMS + 2];
    int max_table_level;
    double *out_err;
    int8_t *out_comb;
    double *out_Q;
    uint32_t *out_stats;
    unsigned int *tile_counter;   // four words, zeroed on the stream before the launch
    uint8_t *wide_flags;          // device, n_frames bytes of scratch
    int search_mode = 0;          // 0 rows filtered by their lower bound, 1 exhaustive (p2s_set_search_mode)
    cudaStream_t stream;
};

struct MpLaunch {
    const float *obs;             // [n_frames][n_cams][max_persons][3 n_joints] float32 {x, y, likelihood}, device
    const int32_t *count;         // [n_frames][n_cams], device
    const p2s_camera_model *cams; // host: K, R, T are used
    long long n_frames;
    int n_cams, max_persons, n_joints, n_max, sm_count;
    size_t smem_per_sm;           // cudaDeviceProp::sharedMemPerMultiprocessor: decides one or two resident frames per SM
    double d_max, min_affinity;
    int8_t *out_rows;             // [n_frames][n_max][n_cams]
    double *out_affinity;         // [n_frames][n_max][n_max] or null
    int32_t *out_iters;           // [n_frames] or null
    unsigned int *tile_counter;
    cudaStream_t stream;
};

struct SwapLaunch {               // handle_LR_swap search (p2s_lrswap.cu)
    const void *obs;              // staged float4 [n_cams][n_units], device
    const int32_t *partner;       // device, [n_keypoints]
    int n_keypoints;
    const double *P;              // host
    const p2s_camera_model *lens; // host or null
    long long n_units;
    int n_cams, min_cams, sm_count;
    double thr;
    const uint32_t *cand_masks;
    uint32_t level_off[P2S_MAX_CAMS + 2];
    int max_table_level;
    double *out_Q, *out_err;
    uint8_t *out_nexcl;
    uint32_t *out_mask;
    cudaStream_t stream;
};

cudaError_t launch_triangulate(const TriLaunch &L, int *grid_out);
cudaError_t launch_lrswap(const SwapLaunch &L);
cudaError_t launch_mp_associate(const MpLaunch &L, int *grid_out);
size_t mp_smem_bytes(int n_max, int n_joints);
cudaError_t launch_stage(const float *x, const float *y, const float *lik, long long n_units, int n_cams,
                         double lik_thr, const p2s_camera_model *lens, void *out, int sm_count, cudaStream_t stream);
cudaError_t launch_collect(const unsigned int *arrive, unsigned int *const *ack, int n, unsigned int value,
                           unsigned int ack_value, unsigned int *err_word, cudaStream_t stream);
cudaError_t launch_fp64_peak(double *out, int blocks, int iters, cudaStream_t stream);
cudaError_t launch_associate(const AssocLaunch &L, int *grid_out);
cudaError_t launch_synth(const double *P, int n_cams, int n_keypoints, unsigned int seed, long long unit0, long long n_units,
                         double sigma, double p_out, double p_low, const double *kp_off, const double *circle,
                         const double *dirs, float *x, float *y, float *lik, double *truth, int sm_count, cudaStream_t stream);

}  // namespace p2s
