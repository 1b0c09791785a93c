// C-ABI layer (include/pose2sim_b200.h): handle, subset tables, stream pipeline for host buffers.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "p2s_internal.h"

namespace {

constexpr int kSlots = 4;                         // H2D / compute / D2H overlap
constexpr long long kChunkUnitsDefault = 0;       // units per pipeline chunk; 0 = automatic: about four chunks per call,
constexpr unsigned int kDeepMinDefault = 2048;   // candidates per level from which a pending unit is parked for deep_search_kernel
                                                  // between 2^16 and 2^20 units (measured, tools/e2e_sweep.py: below ~1e5
                                                  // units per chunk the host-side cost of the 7 copies + launch per chunk
                                                  // shows, above it the call is PCIe-bound at ~47 + 18 GB/s)
constexpr long long kChunkFrames = 1 << 14;       // association frames per chunk
constexpr unsigned long long kMaxTableEntries = 1ULL << 22;

struct SubsetTable {
    uint32_t *d_masks = nullptr;
    uint32_t level_off[P2S_MAX_CAMS + 2] = {0};
    int max_level = -1;
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Slot {
    cudaStream_t stream = nullptr;
    DevBuf x, y, lik, obs, Q, err, nexcl, mask, count, comb, astats, rows, aff, iters, wflags;
};

}  // namespace

struct p2s_handle {
    int device = 0;
    cudaDeviceProp prop;
    double band_eps = 1e-6;
    int solver = 0;
    int assoc_team = 0;
    long long chunk_units = kChunkUnitsDefault;
    int bulk_out = 0;
    int search_mode = 0;                               // p2s_set_search_mode
    int host_mode = 0;                                 // 0 auto, 1 pipeline, 2 zero-copy (p2s_set_host_mode)
    long long launches = 0;
    int last_grid = 0;
    std::string last_error;
    SubsetTable tables[P2S_MAX_CAMS + 1];
    Slot slots[kSlots];
    unsigned int *d_counters = nullptr;           // ring of tile counters
    int counter_next = 0;
    unsigned long long *d_stats = nullptr;
    double *d_peak = nullptr;
    DevBuf wflags;                                // association: per-frame wide-likelihood flags of the *_device entry point
    // deep levels of the triangulation search (p2s_set_deep_search): one list of parked (unit, level) records per
    // counter-ring entry, allocated on the first launch that can park
    unsigned int deep_min = kDeepMinDefault;
    unsigned long long *d_deep = nullptr;
};

namespace {

constexpr int kDeepCap = 16384;                   // parked units per launch; beyond it units are searched by their warp as before
constexpr int kCounterRing = 256;                 // {tile dispenser, fix-up CTAs finished, wide tiles, parked units} per launch; one extra word = error bits

int cuda_fail(p2s_handle *h, cudaError_t e, const char *what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    if (h) h->last_error = buf;
    return P2S_ECUDA;
}

#define P2S_CUDA(h, call)                                          \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail((h), e__, #call); \
    } while (0)

int ensure(p2s_handle *h, DevBuf &b, size_t bytes) {
    if (b.cap >= bytes) return P2S_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { cuda_fail(h, e, "cudaMalloc"); return P2S_ENOMEM; }
    b.cap = want;
    return P2S_OK;
}

// Lexicographic k-subsets of {0..n-1} as bit masks, levels 0..max_level concatenated
// (itertools.combinations order, triangulation.py:411).
int build_table(p2s_handle *h, int n) {
    SubsetTable &t = h->tables[n];
    if (t.d_masks) return P2S_OK;
    std::vector<uint32_t> masks;
    unsigned long long total = 0;
    int level = 0;
    t.level_off[0] = 0;
    for (; level <= n; ++level) {
        // C(n, level)
        unsigned long long cnt = 1;
        for (int i = 1; i <= level; ++i) cnt = cnt * (unsigned)(n - level + i) / (unsigned)i;
        if (total + cnt > kMaxTableEntries) break;
        std::vector<int> idx(level);
        for (int i = 0; i < level; ++i) idx[i] = i;
        for (;;) {
            uint32_t m = 0;
            for (int i = 0; i < level; ++i) m |= 1u << idx[i];
            masks.push_back(m);
            int i = level - 1;
            while (i >= 0 && idx[i] == n - level + i) --i;
            if (i < 0) break;
            ++idx[i];
            for (int j = i + 1; j < level; ++j) idx[j] = idx[j - 1] + 1;
        }
        total += cnt;
        t.level_off[level + 1] = (uint32_t)total;
    }
    t.max_level = level - 1;
    for (int l = level + 1; l < P2S_MAX_CAMS + 2; ++l) t.level_off[l] = (uint32_t)total;
    P2S_CUDA(h, cudaMalloc((void **)&t.d_masks, masks.size() * sizeof(uint32_t)));
    P2S_CUDA(h, cudaMemcpy(t.d_masks, masks.data(), masks.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    return P2S_OK;
}

unsigned int *next_counter(p2s_handle *h) {
    unsigned int *c = h->d_counters + 4 * h->counter_next;
    h->counter_next = (h->counter_next + 1) % kCounterRing;
    return c;
}

unsigned int *error_word(p2s_handle *h) { return h->d_counters + 4 * kCounterRing; }

// Any early return out of a chunk loop (P2S_CUDA, a non-zero rc) must not leave asynchronous copies into the CALLER's
// buffers in flight: this guard drains the slot streams on the way out.
struct DrainSlots {
    p2s_handle *h;
    explicit DrainSlots(p2s_handle *hh) : h(hh) {}
    ~DrainSlots() { for (int k = 0; k < kSlots; ++k) cudaStreamSynchronize(h->slots[k].stream); }
};
struct EventList {                                 // trace events are destroyed on every path
    std::vector<cudaEvent_t> ev;
    cudaEvent_t begin = nullptr;
    ~EventList() {
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        if (begin) cudaEventDestroy(begin);
    }
};

struct PushFlags {
    const unsigned int *wait_flag = nullptr;
    unsigned int wait_value = 0;
    unsigned int *done_flag = nullptr;
    unsigned int done_value = 0;
};

int check_tri_args(int n_cams, int min_cams, long long n_units) {
    if (n_cams < 2 || n_cams > P2S_MAX_CAMS) return P2S_EINVAL;
    if (min_cams < 1) return P2S_EINVAL;
    if (n_units < 0 || n_units > 0x7fffffffLL * 32) return P2S_EINVAL;
    return P2S_OK;
}

struct Planes {                                    // raw-plane input of the fused path (device pointers)
    const float *x = nullptr, *y = nullptr, *lik = nullptr;
    double lik_thr = -INFINITY;
};

int enqueue_triangulate(p2s_handle *h, const void *obs, const double *P, const p2s_camera_model *lens, long long n_units, int n_cams,
                        double thr, int min_cams, double *Q, double *err, uint8_t *nexcl, uint32_t *mask,
                        unsigned long long *stats, cudaStream_t stream, const Planes *planes = nullptr,
                        const PushFlags *push = nullptr, bool device_outputs = false) {
    int rc = build_table(h, n_cams);
    if (rc) return rc;
    if (n_units == 0) return P2S_OK;
    p2s::TriLaunch L;
    if (planes) { L.px = planes->x; L.py = planes->y; L.pl = planes->lik; L.lik_thr = planes->lik_thr; }
    L.obs = obs; L.P = P; L.lens = lens; L.n_units = n_units; L.n_cams = n_cams; L.min_cams = min_cams;
    L.solver = h->solver; L.sm_count = h->prop.multiProcessorCount;
    L.thr = thr; L.band_eps = h->band_eps;
    const SubsetTable &t = h->tables[n_cams];
    L.cand_masks = t.d_masks;
    std::memcpy(L.level_off, t.level_off, sizeof L.level_off);
    L.max_table_level = t.max_level;
    L.out_Q = Q; L.out_err = err; L.out_nexcl = nexcl; L.out_mask = mask; L.stats = stats;
    L.tile_counter = next_counter(h);
    if (h->deep_min != 0 && n_cams - min_cams >= 1) {
        // only rigs whose search can reach a level of deep_min candidates need the list (14 cameras and up by default)
        unsigned long long most = 1, r = 1;
        for (int k = 1; k <= n_cams - min_cams; ++k) { r = r * (unsigned)(n_cams - k + 1) / (unsigned)k; if (r > most) most = r; }
        if (most >= h->deep_min) {
            if (!h->d_deep) {
                cudaError_t e = cudaMalloc((void **)&h->d_deep, (size_t)kCounterRing * kDeepCap * sizeof(unsigned long long));
                if (e != cudaSuccess) { cuda_fail(h, e, "cudaMalloc(deep list)"); return P2S_ENOMEM; }
            }
            L.deep_list = h->d_deep + (size_t)((L.tile_counter - h->d_counters) / 4) * kDeepCap;
            L.deep_cap = kDeepCap;
            L.deep_min = h->deep_min;
        }
    }
    L.stream = stream;
    L.err_word = error_word(h);
    L.bulk_out = h->bulk_out == 1;
    L.allow_pool = device_outputs;
    L.pool = h->bulk_out == 2;
    if (push) { L.wait_flag = push->wait_flag; L.wait_value = push->wait_value; L.done_flag = push->done_flag; L.done_value = push->done_value; }
    P2S_CUDA(h, cudaMemsetAsync(L.tile_counter, 0, 4 * sizeof(unsigned int), stream));
    P2S_CUDA(h, p2s::launch_triangulate(L, &h->last_grid));
    h->launches += L.kernels;                           // search kernel + wide-spread / arrival-flag kernel (+ deep-level kernel)
    return P2S_OK;
}

int enqueue_associate(p2s_handle *h, const void *obs, const int32_t *count, const double *P, long long n_frames,
                      int n_cams, int max_persons, double thr, double lik_thr, int min_cams,
                      double *err, int8_t *comb, double *Q, uint32_t *stats, cudaStream_t stream, double mean_rows,
                      DevBuf *wflags = nullptr) {
    int rc = build_table(h, n_cams);
    if (rc) return rc;
    if (n_frames == 0) return P2S_OK;
    if (!wflags) wflags = &h->wflags;             // one launch in flight per handle on this path (header: lifetime notes)
    if ((rc = ensure(h, *wflags, (size_t)n_frames))) return rc;
    p2s::AssocLaunch L;
    L.obs = obs; L.count = count; L.P = P; L.n_frames = n_frames; L.n_cams = n_cams;
    L.max_persons = max_persons; L.min_cams = min_cams; L.sm_count = h->prop.multiProcessorCount;
    L.mean_rows = mean_rows;
    L.team = h->assoc_team;
    L.thr = thr; L.lik_thr = lik_thr;
    const SubsetTable &t = h->tables[n_cams];
    L.cand_masks = t.d_masks;
    std::memcpy(L.level_off, t.level_off, sizeof L.level_off);
    L.max_table_level = t.max_level;
    L.out_err = err; L.out_comb = comb; L.out_Q = Q; L.out_stats = stats;
    L.tile_counter = next_counter(h);
    L.wide_flags = (uint8_t *)wflags->p;
    L.search_mode = h->search_mode;
    L.stream = stream;
    P2S_CUDA(h, cudaMemsetAsync(L.tile_counter, 0, 4 * sizeof(unsigned int), stream));
    P2S_CUDA(h, p2s::launch_associate(L, &h->last_grid));
    h->launches += 2;                                   // main pass + wide-likelihood pass
    return P2S_OK;
}

}  // namespace

extern "C" {

const char *p2s_status_string(int s) {
    switch (s) {
        case P2S_OK: return "ok";
        case P2S_EINVAL: return "invalid argument";
        case P2S_ENODEVICE: return "no usable CUDA device (sm_100 required); there is no CPU fallback";
        case P2S_ECUDA: return "CUDA error";
        case P2S_ENOMEM: return "out of device memory";
        case P2S_ETOODEEP: return "search too deep to enumerate";
        default: return "unknown status";
    }
}

int p2s_create(int device, p2s_handle **out) {
    if (!out) return P2S_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return P2S_ENODEVICE;
    p2s_handle *h = new (std::nothrow) p2s_handle();
    if (!h) return P2S_ENOMEM;
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&h->prop, device) != cudaSuccess) {
        delete h;
        return P2S_ENODEVICE;
    }
    if (h->prop.major != 10) {                     // the fatbin holds sm_100a code only
        delete h;
        return P2S_ENODEVICE;
    }
    for (int i = 0; i < kSlots; ++i)
        if (cudaStreamCreateWithFlags(&h->slots[i].stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return P2S_ECUDA; }
    if (cudaMalloc((void **)&h->d_counters, (4 * kCounterRing + 4) * sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(h->d_counters, 0, (4 * kCounterRing + 4) * sizeof(unsigned int)) != cudaSuccess ||
        cudaMalloc((void **)&h->d_stats, P2S_STAT_COUNT * sizeof(unsigned long long)) != cudaSuccess) {
        delete h;
        return P2S_ENOMEM;
    }
    *out = h;
    return P2S_OK;
}

int p2s_destroy(p2s_handle *h) {
    if (!h) return P2S_EINVAL;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto &t : h->tables) if (t.d_masks) cudaFree(t.d_masks);
    for (auto &s : h->slots) {
        for (DevBuf *b : {&s.x, &s.y, &s.lik, &s.obs, &s.Q, &s.err, &s.nexcl, &s.mask, &s.count, &s.comb, &s.astats, &s.rows, &s.aff, &s.iters, &s.wflags})
            if (b->p) cudaFree(b->p);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (h->d_counters) cudaFree(h->d_counters);
    if (h->d_stats) cudaFree(h->d_stats);
    if (h->d_peak) cudaFree(h->d_peak);
    if (h->wflags.p) cudaFree(h->wflags.p);
    if (h->d_deep) cudaFree(h->d_deep);
    delete h;
    return P2S_OK;
}

const char *p2s_last_cuda_error(const p2s_handle *h) { return h ? h->last_error.c_str() : ""; }

int p2s_get_device_info(const p2s_handle *h, p2s_device_info *info) {
    if (!h || !info) return P2S_EINVAL;
    std::memset(info, 0, sizeof *info);
    info->device = h->device;
    info->sm_count = h->prop.multiProcessorCount;
    info->cc_major = h->prop.major;
    info->cc_minor = h->prop.minor;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
    info->clock_khz = khz;
    info->total_mem = h->prop.totalGlobalMem;
    std::strncpy(info->name, h->prop.name, sizeof info->name - 1);
    return P2S_OK;
}

int p2s_set_band_eps(p2s_handle *h, double eps) {
    if (!h || !(eps >= 0.0)) return P2S_EINVAL;
    h->band_eps = eps;
    return P2S_OK;
}

int p2s_set_assoc_team(p2s_handle *h, int warps_per_frame) {
    if (!h || (warps_per_frame != 0 && warps_per_frame != 1 && warps_per_frame != 8 && warps_per_frame != 16)) return P2S_EINVAL;
    h->assoc_team = warps_per_frame;
    return P2S_OK;
}

int p2s_set_host_mode(p2s_handle *h, int mode) {
    if (!h || mode < 0 || mode > 2) return P2S_EINVAL;
    h->host_mode = mode;
    return P2S_OK;
}

int p2s_set_output_mode(p2s_handle *h, int mode) {
    if (!h || (mode != 0 && mode != 1 && mode != 2)) return P2S_EINVAL;
    h->bulk_out = mode;
    return P2S_OK;
}

int p2s_set_chunk_units(p2s_handle *h, long long units) {
    if (!h || (units != 0 && (units < 32 || units > (1LL << 26)))) return P2S_EINVAL;
    h->chunk_units = (units + 31) & ~31LL;            // whole tiles, keeps the chunk's planes 16-byte aligned
    return P2S_OK;
}

int p2s_set_search_mode(p2s_handle *h, int mode) {
    if (!h || (mode != 0 && mode != 1)) return P2S_EINVAL;
    h->search_mode = mode;
    return P2S_OK;
}

int p2s_set_deep_search(p2s_handle *h, long long min_candidates) {
    if (!h || min_candidates < 0 || min_candidates > 0xffffffffLL) return P2S_EINVAL;
    h->deep_min = (unsigned int)min_candidates;
    return P2S_OK;
}

int p2s_set_solver(p2s_handle *h, int solver) {
    if (!h || (solver != 0 && solver != 1)) return P2S_EINVAL;
    h->solver = solver;
    return P2S_OK;
}

size_t p2s_obs_bytes(long long n_units, int n_cams) {
    if (n_units < 0 || n_cams < 0) return 0;
    return (size_t)n_units * (size_t)n_cams * 16u;
}

static int stage_device(p2s_handle *h, const float *x, const float *y, const float *lik, long long n_units, int n_cams,
                        double lik_thr, const p2s_camera_model *lens, void *obs_out, void *stream) {
    if (!h || !x || !y || !lik || !obs_out) return P2S_EINVAL;
    if (n_cams < 1 || n_cams > P2S_MAX_CAMS || n_units < 0) return P2S_EINVAL;
    if (n_units == 0) return P2S_OK;
    P2S_CUDA(h, cudaSetDevice(h->device));
    P2S_CUDA(h, p2s::launch_stage(x, y, lik, n_units, n_cams, lik_thr, lens, obs_out, h->prop.multiProcessorCount, (cudaStream_t)stream));
    h->launches += 1;
    return P2S_OK;
}

int p2s_stage_observations_device(p2s_handle *h, const float *x, const float *y, const float *lik,
                                  long long n_units, int n_cams, double lik_thr, void *obs_out, void *stream) {
    return stage_device(h, x, y, lik, n_units, n_cams, lik_thr, nullptr, obs_out, stream);
}

int p2s_stage_undistort_device(p2s_handle *h, const float *x, const float *y, const float *lik,
                               long long n_units, int n_cams, double lik_thr, const p2s_camera_model *lens,
                               void *obs_out, void *stream) {
    if (!lens) return P2S_EINVAL;
    return stage_device(h, x, y, lik, n_units, n_cams, lik_thr, lens, obs_out, stream);
}

static int triangulate_device(p2s_handle *h, const void *obs, const double *P, const p2s_camera_model *lens,
                              long long n_units, int n_cams, double reproj_thr, int min_cams, double *out_Q,
                              double *out_err, uint8_t *out_nexcl, uint32_t *out_mask, unsigned long long *stats,
                              void *stream) {
    if (!h || !P || (n_units > 0 && (!obs || !out_Q || !out_err || !out_nexcl || !out_mask))) return P2S_EINVAL;
    int rc = check_tri_args(n_cams, min_cams, n_units);
    if (rc) return rc;
    P2S_CUDA(h, cudaSetDevice(h->device));
    return enqueue_triangulate(h, obs, P, lens, n_units, n_cams, reproj_thr, min_cams, out_Q, out_err, out_nexcl, out_mask,
                               stats, (cudaStream_t)stream);
}

int p2s_triangulate_device(p2s_handle *h, const void *obs, const double *P, long long n_units, int n_cams,
                           double reproj_thr, int min_cams, double *out_Q, double *out_err, uint8_t *out_nexcl,
                           uint32_t *out_mask, unsigned long long *stats, void *stream) {
    return triangulate_device(h, obs, P, nullptr, n_units, n_cams, reproj_thr, min_cams, out_Q, out_err, out_nexcl,
                              out_mask, stats, stream);
}

int p2s_triangulate_lrswap_device(p2s_handle *h, const void *obs, const int32_t *partner, int n_keypoints,
                                  const double *P, const p2s_camera_model *lens, long long n_units, int n_cams,
                                  double reproj_thr, int min_cams, double *out_Q, double *out_err, uint8_t *out_nexcl,
                                  uint32_t *out_mask, void *stream) {
    if (!h || !P || n_keypoints < 1 || (n_units > 0 && (!obs || !partner || !out_Q || !out_err || !out_nexcl || !out_mask)))
        return P2S_EINVAL;
    if (n_units % n_keypoints != 0) return P2S_EINVAL;         // units are (frame, person) blocks of n_keypoints
    int rc = check_tri_args(n_cams, min_cams, n_units);
    if (rc) return rc;
    P2S_CUDA(h, cudaSetDevice(h->device));
    rc = build_table(h, n_cams);
    if (rc) return rc;
    if (n_units == 0) return P2S_OK;
    {   // the partner map indexes the observation buffer: check it here rather than trust the caller (K ints, not a hot path)
        std::vector<int32_t> part((size_t)n_keypoints);
        P2S_CUDA(h, cudaMemcpyAsync(part.data(), partner, part.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        P2S_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
        for (int32_t v : part)
            if (v < 0 || v >= n_keypoints) return P2S_EINVAL;
    }
    p2s::SwapLaunch L;
    L.obs = obs; L.partner = partner; L.n_keypoints = n_keypoints; L.P = P; L.lens = lens;
    L.n_units = n_units; L.n_cams = n_cams; L.min_cams = min_cams; L.sm_count = h->prop.multiProcessorCount;
    L.thr = reproj_thr;
    const SubsetTable &t = h->tables[n_cams];
    L.cand_masks = t.d_masks;
    std::memcpy(L.level_off, t.level_off, sizeof L.level_off);
    L.max_table_level = t.max_level;
    L.out_Q = out_Q; L.out_err = out_err; L.out_nexcl = out_nexcl; L.out_mask = out_mask;
    L.stream = (cudaStream_t)stream;
    P2S_CUDA(h, p2s::launch_lrswap(L));
    h->launches += 1;
    return P2S_OK;
}

int p2s_triangulate_planes_device(p2s_handle *h, const float *x, const float *y, const float *lik,
                                  const double *P, long long n_units, int n_cams, double lik_thr,
                                  double reproj_thr, int min_cams,
                                  double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                                  unsigned long long *stats, void *stream) {
    if (!h || !P || (n_units > 0 && (!x || !y || !lik || !out_Q || !out_err || !out_nexcl || !out_mask))) return P2S_EINVAL;
    if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)lik) & 15u) return P2S_EINVAL;
    int rc = check_tri_args(n_cams, min_cams, n_units);
    if (rc) return rc;
    P2S_CUDA(h, cudaSetDevice(h->device));
    Planes pl;
    pl.x = x; pl.y = y; pl.lik = lik; pl.lik_thr = lik_thr;
    return enqueue_triangulate(h, nullptr, P, nullptr, n_units, n_cams, reproj_thr, min_cams, out_Q, out_err, out_nexcl,
                               out_mask, stats, (cudaStream_t)stream, &pl, nullptr, true);
}

int p2s_triangulate_planes_push_device(p2s_handle *h, const float *x, const float *y, const float *lik,
                                       const double *P, long long n_units, int n_cams, double lik_thr,
                                       double reproj_thr, int min_cams,
                                       double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                                       unsigned long long *stats,
                                       const unsigned int *wait_flag, unsigned int wait_value,
                                       unsigned int *done_flag, unsigned int done_value, void *stream) {
    if (!h || !P || (n_units > 0 && (!x || !y || !lik || !out_Q || !out_err || !out_nexcl || !out_mask))) return P2S_EINVAL;
    if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)lik) & 15u) return P2S_EINVAL;
    int rc = check_tri_args(n_cams, min_cams, n_units);
    if (rc) return rc;
    P2S_CUDA(h, cudaSetDevice(h->device));
    if (n_units == 0) {                                         // nothing to compute: still hand the flag on
        if (done_flag) {                                        // error_word + 1 is a word that stays 0
            P2S_CUDA(h, p2s::launch_collect(wait_flag ? wait_flag : error_word(h) + 1, &done_flag, 1, wait_flag ? wait_value : 0,
                                            done_value, error_word(h), (cudaStream_t)stream));
            h->launches += 1;
        }
        return P2S_OK;
    }
    Planes pl;
    pl.x = x; pl.y = y; pl.lik = lik; pl.lik_thr = lik_thr;
    PushFlags pf;
    pf.wait_flag = wait_flag; pf.wait_value = wait_value; pf.done_flag = done_flag; pf.done_value = done_value;
    return enqueue_triangulate(h, nullptr, P, nullptr, n_units, n_cams, reproj_thr, min_cams, out_Q, out_err, out_nexcl,
                               out_mask, stats, (cudaStream_t)stream, &pl, &pf);
}

int p2s_peer_alloc(p2s_handle *h, size_t bytes, void **dptr, unsigned char handle[P2S_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == P2S_IPC_HANDLE_BYTES, "IPC handle size");
    if (!h || !dptr || !handle || bytes == 0) return P2S_EINVAL;
    P2S_CUDA(h, cudaSetDevice(h->device));
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { cuda_fail(h, e, "cudaMalloc"); return P2S_ENOMEM; }
    cudaIpcMemHandle_t ih;
    e = cudaIpcGetMemHandle(&ih, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(h, e, "cudaIpcGetMemHandle"); }
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(h, e, "cudaMemset"); }
    std::memcpy(handle, &ih, sizeof ih);
    *dptr = p;
    return P2S_OK;
}

int p2s_peer_open(p2s_handle *h, const unsigned char handle[P2S_IPC_HANDLE_BYTES], void **dptr) {
    if (!h || !dptr || !handle) return P2S_EINVAL;
    P2S_CUDA(h, cudaSetDevice(h->device));
    cudaIpcMemHandle_t ih;
    std::memcpy(&ih, handle, sizeof ih);
    void *p = nullptr;
    P2S_CUDA(h, cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess));
    *dptr = p;
    return P2S_OK;
}

int p2s_peer_close(p2s_handle *h, void *dptr) {
    if (!h || !dptr) return P2S_EINVAL;
    P2S_CUDA(h, cudaSetDevice(h->device));
    P2S_CUDA(h, cudaIpcCloseMemHandle(dptr));
    return P2S_OK;
}

int p2s_peer_free(p2s_handle *h, void *dptr) {
    if (!h || !dptr) return P2S_EINVAL;
    P2S_CUDA(h, cudaSetDevice(h->device));
    P2S_CUDA(h, cudaFree(dptr));
    return P2S_OK;
}

int p2s_peer_collect_device(p2s_handle *h, const unsigned int *arrive, int n, unsigned int value,
                            unsigned int *const *ack, void *stream) {
    if (!h || !arrive || n < 1 || n > P2S_MAX_PEERS) return P2S_EINVAL;
    P2S_CUDA(h, cudaSetDevice(h->device));
    P2S_CUDA(h, p2s::launch_collect(arrive, ack, n, value, value, error_word(h), (cudaStream_t)stream));
    h->launches += 1;
    return P2S_OK;
}

int p2s_peer_error(p2s_handle *h, unsigned int *bits) {
    if (!h || !bits) return P2S_EINVAL;
    P2S_CUDA(h, cudaSetDevice(h->device));
    P2S_CUDA(h, cudaMemcpy(bits, error_word(h), sizeof *bits, cudaMemcpyDeviceToHost));
    P2S_CUDA(h, cudaMemset(error_word(h), 0, sizeof(unsigned int)));
    return P2S_OK;
}

int p2s_triangulate_distorted_device(p2s_handle *h, const void *obs, const double *P, const p2s_camera_model *lens,
                                     long long n_units, int n_cams, double reproj_thr, int min_cams,
                                     double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                                     unsigned long long *stats, void *stream) {
    if (!lens) return P2S_EINVAL;
    return triangulate_device(h, obs, P, lens, n_units, n_cams, reproj_thr, min_cams, out_Q, out_err, out_nexcl,
                              out_mask, stats, stream);
}

static int triangulate_host(p2s_handle *h, const float *x, const float *y, const float *lik, const double *P,
                            const p2s_camera_model *lens, long long n_units, int n_cams, double lik_thr,
                            double reproj_thr, int min_cams, double *out_Q, double *out_err, uint8_t *out_nexcl,
                            uint32_t *out_mask, unsigned long long *stats) {
    if (!h || !P || (n_units > 0 && (!x || !y || !lik || !out_Q || !out_err || !out_nexcl || !out_mask))) return P2S_EINVAL;
    int rc = check_tri_args(n_cams, min_cams, n_units);
    if (rc) return rc;
    P2S_CUDA(h, cudaSetDevice(h->device));
    if (stats) P2S_CUDA(h, cudaMemsetAsync(h->d_stats, 0, P2S_STAT_COUNT * sizeof(unsigned long long), h->slots[0].stream));
    if (stats) P2S_CUDA(h, cudaStreamSynchronize(h->slots[0].stream));
    const size_t C = (size_t)n_cams;
    // ---- zero-copy: pinned host buffers are device-accessible (UVA), so ONE kernel TMA-reads its tiles straight from
    // host memory and writes its outputs straight back — the PCIe transfers are pipelined per 32-unit tile by the
    // kernel's own prefetch, with no chunk boundaries, no staging buffers and no copy set-up (measured 5.07 vs 5.24 ms
    // on cfg2, tools/zero_copy_probe.py).  Pageable or misaligned buffers (and the undistort path, which has a
    // separate stage kernel) take the chunked H2D -> kernel -> D2H pipeline below.
    if (!lens && h->host_mode != 1 && n_units > 0) {
        auto dev = [](const void *p) -> void * {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            return (at.type == cudaMemoryTypeHost) ? at.devicePointer : nullptr;
        };
        void *dx = dev(x), *dy = dev(y), *dl = dev(lik), *dQ = dev(out_Q), *de = dev(out_err), *dn = dev(out_nexcl), *dm = dev(out_mask);
        const bool mapped = dx && dy && dl && dQ && de && dn && dm &&
                            ((((uintptr_t)dx | (uintptr_t)dy | (uintptr_t)dl) & 15u) == 0);
        if (mapped) {
            cudaStream_t st = h->slots[0].stream;
            Planes pl;
            pl.x = (const float *)dx; pl.y = (const float *)dy; pl.lik = (const float *)dl; pl.lik_thr = lik_thr;
            rc = enqueue_triangulate(h, nullptr, P, nullptr, n_units, n_cams, reproj_thr, min_cams, (double *)dQ, (double *)de,
                                     (uint8_t *)dn, (uint32_t *)dm, stats ? h->d_stats : nullptr, st, &pl);
            if (rc) return rc;
            P2S_CUDA(h, cudaStreamSynchronize(st));
            if (stats) P2S_CUDA(h, cudaMemcpy(stats, h->d_stats, P2S_STAT_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            return P2S_OK;
        }
        if (h->host_mode == 2) return P2S_EINVAL;              // zero-copy demanded but the buffers are not pinned / aligned
    }
    const bool automatic = h->chunk_units <= 0;
    long long chunk = h->chunk_units;
    if (automatic) chunk = std::min<long long>(1LL << 20, std::max<long long>(1LL << 16, ((n_units / 4) + 31) & ~31LL));
    chunk = std::min<long long>(chunk, std::max<long long>(n_units, 1));
    int i = 0;
    long long nu = 0;
    // P2S_TRACE=1: device-side timeline of the pipeline (events after each chunk's H2D, kernel and D2H) on stderr
    const bool trace = std::getenv("P2S_TRACE") != nullptr;
    EventList events;
    std::vector<cudaEvent_t> &tev = events.ev;
    std::vector<long long> tunits;
    cudaEvent_t &t_begin = events.begin;
    if (trace) {
        cudaEventCreate(&t_begin);
        cudaEventRecord(t_begin, h->slots[0].stream);
        for (int k = 1; k < kSlots; ++k) cudaStreamWaitEvent(h->slots[k].stream, t_begin, 0);
    }
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tev.push_back(e);
    };
    DrainSlots drain(h);
    for (long long u0 = 0; u0 < n_units; u0 += nu, ++i) {
        const long long rem = n_units - u0;
        nu = std::min(chunk, rem);
        // Automatic mode tapers the tail: once the rest fits one chunk it is halved down to 2^17 units, so that what
        // is left exposed after the last H2D copy (one kernel + one D2H) is a few MB instead of ~25 MB.  Not finer:
        // every copy costs ~13 us of set-up (P2S_TRACE timeline), which shows below ~10 MB per copy.
        if (automatic && rem <= chunk && rem > (1LL << 17)) nu = std::max<long long>(1LL << 17, ((rem / 2) + 31) & ~31LL);
        Slot &s = h->slots[i % kSlots];
        if ((rc = ensure(h, s.x, nu * C * 4)) || (rc = ensure(h, s.y, nu * C * 4)) || (rc = ensure(h, s.lik, nu * C * 4)) ||
            (lens && (rc = ensure(h, s.obs, nu * C * 16))) || (rc = ensure(h, s.Q, nu * 24)) || (rc = ensure(h, s.err, nu * 8)) ||
            (rc = ensure(h, s.nexcl, nu)) || (rc = ensure(h, s.mask, nu * 4)))
            return rc;
        P2S_CUDA(h, cudaMemcpyAsync(s.x.p, x + u0 * C, nu * C * 4, cudaMemcpyHostToDevice, s.stream));
        P2S_CUDA(h, cudaMemcpyAsync(s.y.p, y + u0 * C, nu * C * 4, cudaMemcpyHostToDevice, s.stream));
        P2S_CUDA(h, cudaMemcpyAsync(s.lik.p, lik + u0 * C, nu * C * 4, cudaMemcpyHostToDevice, s.stream));
        mark(s.stream);
        tunits.push_back(nu);
        if (lens) {                                               // undistort: separate stage kernel (iterative lens inversion)
            P2S_CUDA(h, p2s::launch_stage((const float *)s.x.p, (const float *)s.y.p, (const float *)s.lik.p, nu, n_cams,
                                         lik_thr, lens, s.obs.p, h->prop.multiProcessorCount, s.stream));
            h->launches += 1;
            rc = enqueue_triangulate(h, s.obs.p, P, lens, nu, n_cams, reproj_thr, min_cams, (double *)s.Q.p, (double *)s.err.p,
                                     (uint8_t *)s.nexcl.p, (uint32_t *)s.mask.p, stats ? h->d_stats : nullptr, s.stream);
        } else {                                                  // gate + staging fused into the search kernel
            Planes pl;
            pl.x = (const float *)s.x.p; pl.y = (const float *)s.y.p; pl.lik = (const float *)s.lik.p; pl.lik_thr = lik_thr;
            rc = enqueue_triangulate(h, nullptr, P, nullptr, nu, n_cams, reproj_thr, min_cams, (double *)s.Q.p, (double *)s.err.p,
                                     (uint8_t *)s.nexcl.p, (uint32_t *)s.mask.p, stats ? h->d_stats : nullptr, s.stream, &pl);
        }
        if (rc) return rc;
        mark(s.stream);
        P2S_CUDA(h, cudaMemcpyAsync(out_Q + u0 * 3, s.Q.p, nu * 24, cudaMemcpyDeviceToHost, s.stream));
        P2S_CUDA(h, cudaMemcpyAsync(out_err + u0, s.err.p, nu * 8, cudaMemcpyDeviceToHost, s.stream));
        P2S_CUDA(h, cudaMemcpyAsync(out_nexcl + u0, s.nexcl.p, nu, cudaMemcpyDeviceToHost, s.stream));
        P2S_CUDA(h, cudaMemcpyAsync(out_mask + u0, s.mask.p, nu * 4, cudaMemcpyDeviceToHost, s.stream));
        mark(s.stream);
    }
    for (int k = 0; k < kSlots; ++k) P2S_CUDA(h, cudaStreamSynchronize(h->slots[k].stream));
    if (trace) {
        for (size_t c = 0; c < tunits.size(); ++c) {
            float a = 0, b = 0, d = 0;
            cudaEventElapsedTime(&a, t_begin, tev[3 * c]);
            cudaEventElapsedTime(&b, t_begin, tev[3 * c + 1]);
            cudaEventElapsedTime(&d, t_begin, tev[3 * c + 2]);
            fprintf(stderr, "p2s trace chunk %2zu units %8lld  h2d done %7.3f  kernel done %7.3f  d2h done %7.3f ms\n", c, tunits[c], a, b, d);
        }
    }
    if (stats) P2S_CUDA(h, cudaMemcpy(stats, h->d_stats, P2S_STAT_COUNT * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return P2S_OK;
}

int p2s_triangulate_host(p2s_handle *h, const float *x, const float *y, const float *lik, const double *P,
                         long long n_units, int n_cams, double lik_thr, double reproj_thr, int min_cams,
                         double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                         unsigned long long *stats) {
    return triangulate_host(h, x, y, lik, P, nullptr, n_units, n_cams, lik_thr, reproj_thr, min_cams, out_Q, out_err,
                            out_nexcl, out_mask, stats);
}

int p2s_triangulate_undistort_host(p2s_handle *h, const float *x, const float *y, const float *lik,
                                   const double *P, const p2s_camera_model *lens, long long n_units, int n_cams,
                                   double lik_thr, double reproj_thr, int min_cams,
                                   double *out_Q, double *out_err, uint8_t *out_nexcl, uint32_t *out_mask,
                                   unsigned long long *stats) {
    if (!lens) return P2S_EINVAL;
    return triangulate_host(h, x, y, lik, P, lens, n_units, n_cams, lik_thr, reproj_thr, min_cams, out_Q, out_err,
                            out_nexcl, out_mask, stats);
}

int p2s_associate_device(p2s_handle *h, const void *obs, const int32_t *count, const double *P, long long n_frames,
                         int n_cams, int max_persons, double reproj_thr, double lik_thr, int min_cams,
                         double *out_err, int8_t *out_comb, double *out_Q, uint32_t *out_stats, void *stream) {
    if (!h || !P || (n_frames > 0 && (!obs || !count || !out_err || !out_comb || !out_Q))) return P2S_EINVAL;
    if (n_cams < 2 || n_cams > P2S_MAX_CAMS || min_cams < 1 || n_frames < 0 || n_frames > 0xfffffff0LL) return P2S_EINVAL;
    if (max_persons < 1 || max_persons > P2S_MAX_PERSONS) return P2S_EINVAL;
    P2S_CUDA(h, cudaSetDevice(h->device));
    // the per-camera counts live on the device here: size the team for the largest possible product
    const double rows_bound = std::pow((double)max_persons, (double)n_cams);
    return enqueue_associate(h, obs, count, P, n_frames, n_cams, max_persons, reproj_thr, lik_thr, min_cams, out_err,
                             out_comb, out_Q, out_stats, (cudaStream_t)stream, rows_bound);
}

int p2s_associate_host(p2s_handle *h, const float *obs, const int32_t *count, const double *P, long long n_frames,
                       int n_cams, int max_persons, double reproj_thr, double lik_thr, int min_cams,
                       double *out_err, int8_t *out_comb, double *out_Q, uint32_t *out_stats) {
    if (!h || !P || (n_frames > 0 && (!obs || !count || !out_err || !out_comb || !out_Q))) return P2S_EINVAL;
    if (n_cams < 2 || n_cams > P2S_MAX_CAMS || min_cams < 1 || n_frames < 0 || n_frames > 0xfffffff0LL) return P2S_EINVAL;
    if (max_persons < 1 || max_persons > P2S_MAX_PERSONS) return P2S_EINVAL;
    P2S_CUDA(h, cudaSetDevice(h->device));
    const size_t C = (size_t)n_cams, NP = (size_t)max_persons;
    int rc, i = 0;
    const long long chunk = kChunkFrames;
    DrainSlots drain(h);
    for (long long f0 = 0; f0 < n_frames; f0 += chunk, ++i) {
        const long long nf = std::min(chunk, n_frames - f0);
        Slot &s = h->slots[i % kSlots];
        if ((rc = ensure(h, s.obs, nf * C * NP * 16)) || (rc = ensure(h, s.count, nf * C * 4)) || (rc = ensure(h, s.err, nf * 8)) ||
            (rc = ensure(h, s.comb, nf * C)) || (rc = ensure(h, s.Q, nf * 24)) || (rc = ensure(h, s.astats, nf * 8)))
            return rc;
        P2S_CUDA(h, cudaMemcpyAsync(s.obs.p, obs + f0 * C * NP * 4, nf * C * NP * 16, cudaMemcpyHostToDevice, s.stream));
        P2S_CUDA(h, cudaMemcpyAsync(s.count.p, count + f0 * C, nf * C * 4, cudaMemcpyHostToDevice, s.stream));
        double rows = 0.0;                                        // mean size of the person-combination product
        for (long long f = f0; f < f0 + nf; ++f) {
            double r = 1.0;
            for (size_t c = 0; c < C; ++c) { const int n = count[f * C + c]; r *= (n > 1) ? (double)n : 1.0; }
            rows += r;
        }
        rc = enqueue_associate(h, s.obs.p, (const int32_t *)s.count.p, P, nf, n_cams, max_persons, reproj_thr, lik_thr,
                               min_cams, (double *)s.err.p, (int8_t *)s.comb.p, (double *)s.Q.p,
                               out_stats ? (uint32_t *)s.astats.p : nullptr, s.stream, rows / (double)nf, &s.wflags);
        if (rc) return rc;
        P2S_CUDA(h, cudaMemcpyAsync(out_err + f0, s.err.p, nf * 8, cudaMemcpyDeviceToHost, s.stream));
        P2S_CUDA(h, cudaMemcpyAsync(out_comb + f0 * C, s.comb.p, nf * C, cudaMemcpyDeviceToHost, s.stream));
        P2S_CUDA(h, cudaMemcpyAsync(out_Q + f0 * 3, s.Q.p, nf * 24, cudaMemcpyDeviceToHost, s.stream));
        if (out_stats) P2S_CUDA(h, cudaMemcpyAsync(out_stats + f0 * 2, s.astats.p, nf * 8, cudaMemcpyDeviceToHost, s.stream));
    }
    for (int k = 0; k < kSlots; ++k) P2S_CUDA(h, cudaStreamSynchronize(h->slots[k].stream));
    return P2S_OK;
}

static int check_mp_args(const p2s_handle *h, const void *obs, const void *count, const void *cams, long long n_frames,
                         int n_cams, int max_persons, int n_joints, int n_max, double d_max, const void *rows) {
    if (!h || !cams || (n_frames > 0 && (!obs || !count || !rows))) return P2S_EINVAL;
    if (n_cams < 1 || n_cams > P2S_MAX_CAMS || n_frames < 0 || n_frames > 0x7ffffff0LL) return P2S_EINVAL;
    if (max_persons < 1 || max_persons > P2S_MAX_DETECTIONS || n_joints < 1 || n_joints > 1024) return P2S_EINVAL;
    if (n_max < 1 || n_max > P2S_MAX_DETECTIONS || !(d_max > 0.0)) return P2S_EINVAL;
    if (p2s::mp_smem_bytes(n_max, n_joints) > (size_t)h->prop.sharedMemPerBlockOptin) return P2S_EINVAL;
    return P2S_OK;
}

static int enqueue_mp(p2s_handle *h, const float *obs, const int32_t *count, const p2s_camera_model *cams, long long n_frames,
                      int n_cams, int max_persons, int n_joints, int n_max, double d_max, double min_affinity,
                      int8_t *rows, double *aff, int32_t *iters, cudaStream_t stream) {
    if (n_frames == 0) return P2S_OK;
    p2s::MpLaunch L;
    L.obs = obs; L.count = count; L.cams = cams; L.n_frames = n_frames; L.n_cams = n_cams; L.max_persons = max_persons;
    L.n_joints = n_joints; L.n_max = n_max; L.sm_count = h->prop.multiProcessorCount; L.d_max = d_max;
    L.smem_per_sm = h->prop.sharedMemPerMultiprocessor;
    L.min_affinity = min_affinity; L.out_rows = rows; L.out_affinity = aff; L.out_iters = iters;
    L.tile_counter = next_counter(h);
    L.stream = stream;
    P2S_CUDA(h, cudaMemsetAsync(L.tile_counter, 0, sizeof(unsigned int), stream));
    P2S_CUDA(h, p2s::launch_mp_associate(L, &h->last_grid));
    h->launches += 1;
    return P2S_OK;
}

int p2s_associate_multi_device(p2s_handle *h, const float *obs, const int32_t *count, const p2s_camera_model *cams,
                               long long n_frames, int n_cams, int max_persons, int n_joints, int n_max,
                               double d_max, double min_affinity, int8_t *out_rows, double *out_affinity,
                               int32_t *out_iters, void *stream) {
    int rc = check_mp_args(h, obs, count, cams, n_frames, n_cams, max_persons, n_joints, n_max, d_max, out_rows);
    if (rc) return rc;
    P2S_CUDA(h, cudaSetDevice(h->device));
    return enqueue_mp(h, obs, count, cams, n_frames, n_cams, max_persons, n_joints, n_max, d_max, min_affinity, out_rows,
                      out_affinity, out_iters, (cudaStream_t)stream);
}

int p2s_associate_multi_host(p2s_handle *h, const float *obs, const int32_t *count, const p2s_camera_model *cams,
                             long long n_frames, int n_cams, int max_persons, int n_joints, int n_max,
                             double d_max, double min_affinity, int8_t *out_rows, double *out_affinity,
                             int32_t *out_iters) {
    int rc = check_mp_args(h, obs, count, cams, n_frames, n_cams, max_persons, n_joints, n_max, d_max, out_rows);
    if (rc) return rc;
    P2S_CUDA(h, cudaSetDevice(h->device));
    // the counts are host memory here: a frame with more detections than n_max is an argument error, not something to
    // truncate silently (the device entry point clamps in the kernel instead)
    for (long long f = 0; f < n_frames; ++f) {
        long long tot = 0;
        for (int c = 0; c < n_cams; ++c) tot += std::max(0, std::min(count[f * n_cams + c], max_persons));
        if (tot > n_max) return P2S_EINVAL;
    }
    const size_t C = (size_t)n_cams, per_frame = C * (size_t)max_persons * 3u * (size_t)n_joints, NM = (size_t)n_max;
    long long chunk = std::max<long long>(1, std::min<long long>(kChunkFrames, (long long)((64u << 20) / (per_frame * 4u + 1))));
    int i = 0;
    DrainSlots drain(h);
    for (long long f0 = 0; f0 < n_frames; f0 += chunk, ++i) {
        const long long nf = std::min(chunk, n_frames - f0);
        Slot &s = h->slots[i % kSlots];
        if ((rc = ensure(h, s.obs, nf * per_frame * 4)) || (rc = ensure(h, s.count, nf * C * 4)) ||
            (rc = ensure(h, s.rows, nf * NM * C)) || (rc = ensure(h, s.iters, nf * 4)) ||
            (out_affinity && (rc = ensure(h, s.aff, nf * NM * NM * 8))))
            return rc;
        P2S_CUDA(h, cudaMemcpyAsync(s.obs.p, obs + f0 * per_frame, nf * per_frame * 4, cudaMemcpyHostToDevice, s.stream));
        P2S_CUDA(h, cudaMemcpyAsync(s.count.p, count + f0 * C, nf * C * 4, cudaMemcpyHostToDevice, s.stream));
        P2S_CUDA(h, cudaMemsetAsync(s.rows.p, 0xff, nf * NM * C, s.stream));
        if (out_affinity) P2S_CUDA(h, cudaMemsetAsync(s.aff.p, 0, nf * NM * NM * 8, s.stream));
        rc = enqueue_mp(h, (const float *)s.obs.p, (const int32_t *)s.count.p, cams, nf, n_cams, max_persons, n_joints, n_max,
                        d_max, min_affinity, (int8_t *)s.rows.p, out_affinity ? (double *)s.aff.p : nullptr,
                        (int32_t *)s.iters.p, s.stream);
        if (rc) return rc;
        P2S_CUDA(h, cudaMemcpyAsync(out_rows + f0 * NM * C, s.rows.p, nf * NM * C, cudaMemcpyDeviceToHost, s.stream));
        if (out_affinity)
            P2S_CUDA(h, cudaMemcpyAsync(out_affinity + f0 * NM * NM, s.aff.p, nf * NM * NM * 8, cudaMemcpyDeviceToHost, s.stream));
        if (out_iters) P2S_CUDA(h, cudaMemcpyAsync(out_iters + f0, s.iters.p, nf * 4, cudaMemcpyDeviceToHost, s.stream));
    }
    for (int k = 0; k < kSlots; ++k) P2S_CUDA(h, cudaStreamSynchronize(h->slots[k].stream));
    return P2S_OK;
}

int p2s_measure_fp64_peak(p2s_handle *h, double *tflops, double *ms_out) {
    if (!h || !tflops) return P2S_EINVAL;
    P2S_CUDA(h, cudaSetDevice(h->device));
    const int blocks = h->prop.multiProcessorCount * 8, iters = 4096;
    if (!h->d_peak) P2S_CUDA(h, cudaMalloc((void **)&h->d_peak, (size_t)blocks * 256 * sizeof(double)));
    cudaStream_t st = h->slots[0].stream;
    cudaEvent_t e0, e1;
    P2S_CUDA(h, cudaEventCreate(&e0));
    P2S_CUDA(h, cudaEventCreate(&e1));
    double best = 0.0, best_ms = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        P2S_CUDA(h, cudaEventRecord(e0, st));
        P2S_CUDA(h, p2s::launch_fp64_peak(h->d_peak, blocks, iters, st));
        P2S_CUDA(h, cudaEventRecord(e1, st));
        P2S_CUDA(h, cudaEventSynchronize(e1));
        h->launches += 1;
        float ms = 0.f;
        P2S_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * 256.0;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) { best = tf; best_ms = ms; }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    if (ms_out) *ms_out = best_ms;
    return P2S_OK;
}

int p2s_synth_observations_device(p2s_handle *h, const double *P, int n_cams, int n_keypoints, unsigned int seed,
                                  long long unit0, long long n_units, double sigma, double p_out, double p_low,
                                  const double *kp_offsets, const double *circle, const double *dirs,
                                  float *x, float *y, float *lik, double *truth, void *stream) {
    if (!h || !P || !kp_offsets || !circle || !dirs || (n_units > 0 && (!x || !y || !lik))) return P2S_EINVAL;
    if (n_cams < 1 || n_cams > P2S_MAX_CAMS || n_keypoints < 1 || unit0 < 0 || n_units < 0) return P2S_EINVAL;
    if (n_units == 0) return P2S_OK;
    P2S_CUDA(h, cudaSetDevice(h->device));
    P2S_CUDA(h, p2s::launch_synth(P, n_cams, n_keypoints, seed, unit0, n_units, sigma, p_out, p_low, kp_offsets, circle, dirs,
                                  x, y, lik, truth, h->prop.multiProcessorCount, (cudaStream_t)stream));
    h->launches += 1;
    return P2S_OK;
}

long long p2s_launch_count(const p2s_handle *h) { return h ? h->launches : 0; }

int p2s_last_grid(const p2s_handle *h) { return h ? h->last_grid : 0; }

}  // extern "C"
