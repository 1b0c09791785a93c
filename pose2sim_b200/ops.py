"""Batched device operators behind the reference's per-unit function seam.

`Engine.triangulate*`  <->  Pose2Sim/triangulation.py:363 `triangulation_from_best_cameras`
`Engine.associate*`    <->  Pose2Sim/personAssociation.py:154 `best_persons_and_cameras_combination`

PyTorch is used for device memory and streams only; all arithmetic runs in the hand-written sm_100a
kernels of `libp2s_b200.so`, reached through its C ABI with ctypes.  No CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib

_ENGINES = {}


def _torch():
    import torch
    return torch


def get_engine(device=0):
    """One Engine (C-ABI handle) per GPU per process."""
    if device not in _ENGINES:
        _ENGINES[device] = Engine(device)
    return _ENGINES[device]


def _as_P(P, n_cams):
    P = np.ascontiguousarray(np.asarray(P, dtype=np.float64).reshape(n_cams, 12))
    return P


def lens_array(models):
    """List of dicts {K[3,3], dist[<=8], R[3,3], T[3], newK[3,3]} -> ctypes array of p2s_camera_model
    (dist and newK optional: the multi-person matching only uses K, R, T)."""
    arr = (_lib.CameraModel * len(models))()
    for m, o in zip(models, arr):
        o.K[:] = np.asarray(m["K"], float).reshape(9)
        d = np.zeros(8)
        dist = np.asarray(m.get("dist", []), float).reshape(-1)
        if dist.size > 8:
            raise ValueError("at most 8 distortion coefficients (k1 k2 p1 p2 k3 k4 k5 k6) are supported")
        d[:dist.size] = dist
        o.dist[:] = d
        o.R[:] = np.asarray(m["R"], float).reshape(9)
        o.T[:] = np.asarray(m["T"], float).reshape(3)
        o.newK[:] = np.asarray(m.get("newK", m["K"]), float).reshape(9)
    return arr


def _ptr(a):
    """Raw pointer of a numpy array or torch tensor (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def stats_dict(stats):
    s = [int(v) for v in stats]
    levels = s[_lib.STAT_LEVEL0:_lib.STAT_LEVEL0 + 33]
    while len(levels) > 1 and levels[-1] == 0:
        levels.pop()
    return {"level_hist": levels, "not_evaluated": s[_lib.STAT_NOT_EVALUATED], "failed": s[_lib.STAT_FAILED],
            "candidates": s[_lib.STAT_CANDIDATES], "cam_solves": s[_lib.STAT_CAM_SOLVES],
            "band_threshold": s[_lib.STAT_BAND_THRESHOLD], "band_argmin": s[_lib.STAT_BAND_ARGMIN],
            "solver_steps": s[_lib.STAT_NEWTON_STEPS], "solved": s[_lib.STAT_SOLVED],
            "direct_cams": s[_lib.STAT_DIRECT_CAMS], "blocks": s[_lib.STAT_BLOCKS], "entry_adds": s[_lib.STAT_ENTRY_ADDS],
            "wide_units": s[_lib.STAT_WIDE_UNITS]}


class Engine:
    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(None, self.lib.p2s_create(int(device), C.byref(h)))
        self.h = h
        self.device = int(device)
        info = _lib.DeviceInfo()
        _lib.check(self.h, self.lib.p2s_get_device_info(self.h, C.byref(info)))
        self.info = {"name": info.name.decode(), "sm_count": info.sm_count, "cc": (info.cc_major, info.cc_minor),
                     "clock_khz": info.clock_khz, "total_mem": info.total_mem}
        self._partner_maps = {}          # keypoint partner maps resident on the device (triangulate_lr_swap)

    def close(self):
        if getattr(self, "h", None):
            self.lib.p2s_destroy(self.h)
            self.h = None

    # ---- knobs ---------------------------------------------------------------------------------
    def set_band_eps(self, eps_px):
        _lib.check(self.h, self.lib.p2s_set_band_eps(self.h, float(eps_px)))

    def set_solver(self, solver):
        _lib.check(self.h, self.lib.p2s_set_solver(self.h, {"secular": 0, "jacobi": 1}.get(solver, solver)))

    def set_assoc_team(self, warps_per_frame):
        _lib.check(self.h, self.lib.p2s_set_assoc_team(self.h, int(warps_per_frame)))

    def set_host_mode(self, mode):
        """How `triangulate_host` moves data: "auto" (zero-copy for pinned buffers, else copy pipeline), "pipeline",
        "zero_copy" (raises when the buffers are not pinned)."""
        _lib.check(self.h, self.lib.p2s_set_host_mode(self.h, {"auto": 0, "pipeline": 1, "zero_copy": 2}.get(mode, mode)))

    def set_search_mode(self, mode):
        """Single-person association search: "filtered" (default: candidates whose lower bound is above the threshold are
        not solved; same results) or "exhaustive" (every row solved and re-projected)."""
        _lib.check(self.h, self.lib.p2s_set_search_mode(self.h, {"filtered": 0, "exhaustive": 1}.get(mode, mode)))

    def set_deep_search(self, min_candidates):
        """Triangulation search: a unit pending at a level of >= `min_candidates` camera subsets is parked and searched by
        a cluster of two 512-thread CTAs of the deep-level kernel instead of by the warp that holds its tile (default 2048; 0 = never).
        Same outputs bit for bit."""
        _lib.check(self.h, self.lib.p2s_set_deep_search(self.h, int(min_candidates)))

    def set_output_mode(self, mode):
        """0 = vector stores (default), 1 = TMA bulk stores of whole tile records (`bulk`), 2 = the pooled kernel
        (`pooled`: level-1 passes shared across tiles, device-resident calls without statistics at 4 / 8 cameras)."""
        _lib.check(self.h, self.lib.p2s_set_output_mode(self.h, {"vector": 0, "bulk": 1, "pooled": 2}.get(mode, mode)))

    def set_chunk_units(self, units):
        _lib.check(self.h, self.lib.p2s_set_chunk_units(self.h, int(units)))

    def launch_count(self):
        return int(self.lib.p2s_launch_count(self.h))

    def last_grid(self):
        return int(self.lib.p2s_last_grid(self.h))

    def fp64_peak(self):
        tf, ms = C.c_double(), C.c_double()
        _lib.check(self.h, self.lib.p2s_measure_fp64_peak(self.h, C.byref(tf), C.byref(ms)))
        return tf.value

    def _stream(self):
        return _torch().cuda.current_stream(self.device).cuda_stream

    # ---- workload generation (benchmarks / tests) ------------------------------------------------------
    def synth_observations(self, P, unit0, n_units, n_keypoints=26, seed=500, sigma=2.0, p_out=0.05, p_low=0.05,
                           want_truth=False, out=None):
        """Observations of units [unit0, unit0 + n_units) generated ON THE DEVICE as a pure function of
        (seed, unit, camera) — the device side of `synth_philox.observations` (bit-identical float32 values).
        Returns dict of CUDA tensors x, y, lik [n_units, C] float32 (ungated) and, on request, truth [n_units, 3]."""
        torch = _torch()
        from . import synth_philox
        dev = torch.device("cuda", self.device)
        Pm = _as_P(P, np.asarray(P).reshape(-1, 12).shape[0])
        Cn = Pm.shape[0]
        key = ("synth_tables", int(n_keypoints))
        if key not in self.__dict__.setdefault("_cache", {}):
            self._cache[key] = tuple(torch.from_numpy(t).to(dev) for t in synth_philox.tables(int(n_keypoints)))
        off, circle, dirs = self._cache[key]
        if out is None:
            out = {k: torch.empty((n_units, Cn), dtype=torch.float32, device=dev) for k in ("x", "y", "lik")}
        truth = torch.empty((n_units, 3), dtype=torch.float64, device=dev) if want_truth else None
        _lib.check(self.h, self.lib.p2s_synth_observations_device(
            self.h, Pm.ctypes.data, Cn, int(n_keypoints), int(seed) & 0xFFFFFFFF, int(unit0), int(n_units), float(sigma),
            float(p_out), float(p_low), _ptr(off), _ptr(circle), _ptr(dirs), _ptr(out["x"]), _ptr(out["y"]), _ptr(out["lik"]),
            _ptr(truth), self._stream()))
        if want_truth:
            out["truth"] = truth
        return out

    # ---- device-resident path ------------------------------------------------------------------------
    def stage_observations(self, x, y, lik, lik_thr=None, out=None, lens=None):
        """x, y, lik: CUDA float32 tensors [U, C] -> staged float4 tensor [C, U, 4] with the
        likelihood gate of triangulation.py:817-821 (lik_thr=None: no gate).  lens: list of camera
        models -> the points are undistorted first (undistort_points, triangulation.py:808-813)."""
        torch = _torch()
        U, Cn = x.shape
        for t in (x, y, lik):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (U, Cn)
        if out is None:
            out = torch.empty((Cn, U, 4), dtype=torch.float32, device=x.device)
        thr = float("-inf") if lik_thr is None else float(lik_thr)
        if lens is not None:
            arr = lens_array(lens)
            _lib.check(self.h, self.lib.p2s_stage_undistort_device(self.h, _ptr(x), _ptr(y), _ptr(lik), U, Cn, thr,
                                                                   C.cast(arr, C.c_void_p), _ptr(out), self._stream()))
            return out
        _lib.check(self.h, self.lib.p2s_stage_observations_device(self.h, _ptr(x), _ptr(y), _ptr(lik), U, Cn, thr,
                                                                  _ptr(out), self._stream()))
        return out

    def triangulate(self, obs, P, reproj_thr, min_cams, out=None, stats=None, lens=None):
        """obs: staged CUDA tensor [C, U, 4]; returns dict of CUDA tensors Q[U,3] f64, err[U] f64,
        nexcl[U] u8, mask[U] i32 (bit pattern of the uint32 mask).  Asynchronous on the current stream."""
        torch = _torch()
        Cn, U, four = obs.shape
        assert four == 4 and obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous()
        dev = obs.device
        if out is None:
            out = {"Q": torch.empty((U, 3), dtype=torch.float64, device=dev),
                   "err": torch.empty((U,), dtype=torch.float64, device=dev),
                   "nexcl": torch.empty((U,), dtype=torch.uint8, device=dev),
                   "mask": torch.empty((U,), dtype=torch.int32, device=dev)}
        Pm = _as_P(P, Cn)
        if lens is not None:
            arr = lens_array(lens)
            _lib.check(self.h, self.lib.p2s_triangulate_distorted_device(
                self.h, _ptr(obs), Pm.ctypes.data, C.cast(arr, C.c_void_p), U, Cn, float(reproj_thr), int(min_cams),
                _ptr(out["Q"]), _ptr(out["err"]), _ptr(out["nexcl"]), _ptr(out["mask"]), _ptr(stats), self._stream()))
            return out
        _lib.check(self.h, self.lib.p2s_triangulate_device(
            self.h, _ptr(obs), Pm.ctypes.data, U, Cn, float(reproj_thr), int(min_cams),
            _ptr(out["Q"]), _ptr(out["err"]), _ptr(out["nexcl"]), _ptr(out["mask"]), _ptr(stats), self._stream()))
        return out

    def triangulate_lr_swap(self, obs, partner, P, reproj_thr, min_cams, lens=None):
        """`handle_LR_swap = true` (triangulation.py:509-579).  obs: staged CUDA tensor [C, U, 4] with the units ordered
        (frame, person, keypoint); partner: K integers, the keypoint index of each keypoint's left/right partner
        (`keypoints_idx_swapped`, :742-745).  Same outputs as `triangulate`."""
        torch = _torch()
        Cn, U, four = obs.shape
        assert four == 4 and obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous()
        part = np.ascontiguousarray(partner, dtype=np.int32).ravel()
        K = int(part.size)
        if K < 1 or U % K or part.min(initial=0) < 0 or part.max(initial=0) >= K:
            raise ValueError(f"partner must hold K indices in [0, K) with K dividing the {U} units")
        dev = obs.device
        part_d = self._partner_maps.get(part.tobytes())           # the launch is asynchronous: the map must outlive it
        if part_d is None or part_d.device != dev:
            part_d = self._partner_maps[part.tobytes()] = torch.from_numpy(part).to(dev)
        out = {"Q": torch.empty((U, 3), dtype=torch.float64, device=dev),
               "err": torch.empty((U,), dtype=torch.float64, device=dev),
               "nexcl": torch.empty((U,), dtype=torch.uint8, device=dev),
               "mask": torch.empty((U,), dtype=torch.int32, device=dev)}
        Pm = _as_P(P, Cn)
        arr = lens_array(lens) if lens is not None else None
        _lib.check(self.h, self.lib.p2s_triangulate_lrswap_device(
            self.h, _ptr(obs), _ptr(part_d), K, Pm.ctypes.data, C.cast(arr, C.c_void_p) if arr is not None else None,
            U, Cn, float(reproj_thr), int(min_cams),
            _ptr(out["Q"]), _ptr(out["err"]), _ptr(out["nexcl"]), _ptr(out["mask"]), self._stream()))
        return out

    def triangulate_planes(self, x, y, lik, P, lik_thr, reproj_thr, min_cams, out=None, stats=None):
        """x, y, lik: CUDA float32 tensors [U, C] (raw planes).  Gate + float4 staging are fused into the
        search kernel's tile load: ONE kernel, no staged buffer in HBM.  Returns the same dict as `triangulate`."""
        torch = _torch()
        U, Cn = x.shape
        for t in (x, y, lik):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (U, Cn)
        if out is None:
            dev = x.device
            out = {"Q": torch.empty((U, 3), dtype=torch.float64, device=dev),
                   "err": torch.empty((U,), dtype=torch.float64, device=dev),
                   "nexcl": torch.empty((U,), dtype=torch.uint8, device=dev),
                   "mask": torch.empty((U,), dtype=torch.int32, device=dev)}
        Pm = _as_P(P, Cn)
        thr = float("-inf") if lik_thr is None else float(lik_thr)
        _lib.check(self.h, self.lib.p2s_triangulate_planes_device(
            self.h, _ptr(x), _ptr(y), _ptr(lik), Pm.ctypes.data, U, Cn, thr, float(reproj_thr), int(min_cams),
            _ptr(out["Q"]), _ptr(out["err"]), _ptr(out["nexcl"]), _ptr(out["mask"]), _ptr(stats), self._stream()))
        return out

    def new_stats(self):
        torch = _torch()
        return torch.zeros(_lib.P2S_STAT_COUNT, dtype=torch.int64, device=f"cuda:{self.device}")

    # ---- host-buffer path (what triangulate_all calls) ---------------------------------------------
    def triangulate_host(self, x, y, lik, P, lik_thr, reproj_thr, min_cams, out=None, want_stats=True, lens=None):
        """x, y, lik: host float32 arrays [U, C] (numpy, or pinned torch CPU tensors).  Returns dict of
        numpy arrays Q[U,3], err[U], nexcl[U] (uint8), mask[U] (uint32) and `stats`."""
        U, Cn = x.shape
        for t in (x, y, lik):
            assert tuple(t.shape) == (U, Cn)
        xs, ys, ls = (np.ascontiguousarray(t, dtype=np.float32) if isinstance(t, np.ndarray) else t for t in (x, y, lik))
        if out is None:
            out = {"Q": np.empty((U, 3), np.float64), "err": np.empty(U, np.float64),
                   "nexcl": np.empty(U, np.uint8), "mask": np.empty(U, np.uint32)}
        stats = np.zeros(_lib.P2S_STAT_COUNT, np.uint64) if want_stats else None
        Pm = _as_P(P, Cn)
        thr = float("-inf") if lik_thr is None else float(lik_thr)
        if lens is not None:
            arr = lens_array(lens)
            _lib.check(self.h, self.lib.p2s_triangulate_undistort_host(
                self.h, _ptr(xs), _ptr(ys), _ptr(ls), Pm.ctypes.data, C.cast(arr, C.c_void_p), U, Cn, thr,
                float(reproj_thr), int(min_cams), _ptr(out["Q"]), _ptr(out["err"]), _ptr(out["nexcl"]),
                _ptr(out["mask"]), _ptr(stats)))
        else:
            _lib.check(self.h, self.lib.p2s_triangulate_host(
                self.h, _ptr(xs), _ptr(ys), _ptr(ls), Pm.ctypes.data, U, Cn, thr, float(reproj_thr), int(min_cams),
                _ptr(out["Q"]), _ptr(out["err"]), _ptr(out["nexcl"]), _ptr(out["mask"]), _ptr(stats)))
        if want_stats:
            out["stats"] = stats_dict(stats)
        return out

    # ---- association ----------------------------------------------------------------------------------
    def associate(self, obs, count, P, reproj_thr, lik_thr, min_cams, want_stats=False):
        """obs: CUDA float32 [F, C, NP, 4]; count: CUDA int32 [F, C].  Returns dict of CUDA tensors."""
        torch = _torch()
        F, Cn, NP, four = obs.shape
        assert four == 4 and obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous()
        assert count.is_cuda and count.dtype == torch.int32 and tuple(count.shape) == (F, Cn) and count.is_contiguous()
        dev = obs.device
        out = {"err": torch.empty((F,), dtype=torch.float64, device=dev),
               "comb": torch.empty((F, Cn), dtype=torch.int8, device=dev),
               "Q": torch.empty((F, 3), dtype=torch.float64, device=dev)}
        st = torch.zeros((F, 2), dtype=torch.int32, device=dev) if want_stats else None
        Pm = _as_P(P, Cn)
        _lib.check(self.h, self.lib.p2s_associate_device(
            self.h, _ptr(obs), _ptr(count), Pm.ctypes.data, F, Cn, NP, float(reproj_thr), float(lik_thr), int(min_cams),
            _ptr(out["err"]), _ptr(out["comb"]), _ptr(out["Q"]), _ptr(st), self._stream()))
        if want_stats:
            out["stats"] = st
        return out

    def associate_host(self, obs, count, P, reproj_thr, lik_thr, min_cams, want_stats=False):
        """obs: host float32 [F, C, NP, 3 or 4] (x, y, likelihood[, pad]); count: int32 [F, C]."""
        obs = np.asarray(obs, dtype=np.float32)
        F, Cn, NP = obs.shape[:3]
        if obs.shape[3] == 3:
            o4 = np.zeros((F, Cn, NP, 4), np.float32)
            o4[..., :3] = obs
            obs = o4
        obs = np.ascontiguousarray(obs)
        count = np.ascontiguousarray(count, dtype=np.int32)
        out = {"err": np.empty(F, np.float64), "comb": np.empty((F, Cn), np.int8), "Q": np.empty((F, 3), np.float64)}
        st = np.zeros((F, 2), np.uint32) if want_stats else None
        Pm = _as_P(P, Cn)
        _lib.check(self.h, self.lib.p2s_associate_host(
            self.h, obs.ctypes.data, count.ctypes.data, Pm.ctypes.data, F, Cn, NP, float(reproj_thr), float(lik_thr),
            int(min_cams), _ptr(out["err"]), _ptr(out["comb"]), _ptr(out["Q"]), _ptr(st)))
        if want_stats:
            out["stats"] = st
        return out

    # ---- multi-person association -----------------------------------------------------------------------
    def associate_multi_host(self, obs, count, models, max_distance, min_affinity, n_max=None, want_affinity=False):
        """personAssociation.py:793-801 for all frames.  obs: host float32 [F, C, NP, 3 J] (pose_keypoints_2d of
        every detection); count: int32 [F, C]; models: list of {K, R, T} per camera.  Returns `rows`
        int8 [F, n_max, C] (per detection the arg-max detection of each view, -1 = none), `iters` [F] and, on
        request, the matched affinity [F, n_max, n_max]."""
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        F, Cn, NP, L = obs.shape
        assert L % 3 == 0
        count = np.ascontiguousarray(count, dtype=np.int32)
        assert tuple(count.shape) == (F, Cn) and len(models) == Cn
        if n_max is None:
            n_max = max(1, int(count.sum(axis=1).max(initial=0)))
        out = {"rows": np.full((F, n_max, Cn), -1, np.int8), "iters": np.zeros(F, np.int32)}
        aff = np.zeros((F, n_max, n_max), np.float64) if want_affinity else None
        arr = lens_array(models)
        _lib.check(self.h, self.lib.p2s_associate_multi_host(
            self.h, obs.ctypes.data, count.ctypes.data, C.cast(arr, C.c_void_p), F, Cn, NP, L // 3, int(n_max),
            float(max_distance), float(min_affinity), _ptr(out["rows"]), _ptr(aff), _ptr(out["iters"])))
        if want_affinity:
            out["affinity"] = aff
        return out

    def associate_multi(self, obs, count, models, max_distance, min_affinity, n_max, want_affinity=False):
        """Device-resident variant: obs CUDA float32 [F, C, NP, 3 J], count CUDA int32 [F, C]; asynchronous on the
        current stream.  Returns CUDA tensors."""
        torch = _torch()
        F, Cn, NP, L = obs.shape
        assert obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and L % 3 == 0
        assert count.is_cuda and count.dtype == torch.int32 and tuple(count.shape) == (F, Cn) and count.is_contiguous()
        dev = obs.device
        out = {"rows": torch.full((F, n_max, Cn), -1, dtype=torch.int8, device=dev),
               "iters": torch.zeros((F,), dtype=torch.int32, device=dev)}
        aff = torch.zeros((F, n_max, n_max), dtype=torch.float64, device=dev) if want_affinity else None
        arr = lens_array(models)
        _lib.check(self.h, self.lib.p2s_associate_multi_device(
            self.h, _ptr(obs), _ptr(count), C.cast(arr, C.c_void_p), F, Cn, NP, L // 3, int(n_max),
            float(max_distance), float(min_affinity), _ptr(out["rows"]), _ptr(aff), _ptr(out["iters"]), self._stream()))
        if want_affinity:
            out["affinity"] = aff
        return out

    # ---- multi-GPU push path (include/pose2sim_b200.h "the final gather fused into the search kernel") ------------
    def peer_alloc(self, nbytes):
        """Device buffer other processes of this node can map.  Returns (device address, 64-byte handle)."""
        p = C.c_void_p()
        hb = (C.c_ubyte * _lib.P2S_IPC_HANDLE_BYTES)()
        _lib.check(self.h, self.lib.p2s_peer_alloc(self.h, int(nbytes), C.byref(p), C.cast(hb, C.c_void_p)))
        return int(p.value), bytes(hb)

    def peer_open(self, handle):
        p = C.c_void_p()
        hb = (C.c_ubyte * _lib.P2S_IPC_HANDLE_BYTES).from_buffer_copy(handle)
        _lib.check(self.h, self.lib.p2s_peer_open(self.h, C.cast(hb, C.c_void_p), C.byref(p)))
        return int(p.value)

    def peer_close(self, ptr):
        _lib.check(self.h, self.lib.p2s_peer_close(self.h, C.c_void_p(ptr)))

    def peer_free(self, ptr):
        _lib.check(self.h, self.lib.p2s_peer_free(self.h, C.c_void_p(ptr)))

    def peer_error(self):
        bits = C.c_uint()
        _lib.check(self.h, self.lib.p2s_peer_error(self.h, C.byref(bits)))
        return int(bits.value)

    def peer_collect(self, arrive_ptr, n, value, ack_ptrs, stream=None):
        """Consumer side: wait (on the device, on `stream`) until arrive[i] >= value for i < n, then write `value`
        to every ack pointer (0 = skip)."""
        arr = (C.c_void_p * n)(*[C.c_void_p(p or None) for p in ack_ptrs])
        st = self._stream() if stream is None else stream
        _lib.check(self.h, self.lib.p2s_peer_collect_device(self.h, C.c_void_p(arrive_ptr), int(n), int(value) & 0xffffffff,
                                                            C.cast(arr, C.c_void_p), st))

    def triangulate_planes_push(self, x, y, lik, P, lik_thr, reproj_thr, min_cams, out_ptrs, wait_flag=0, wait_value=0,
                                done_flag=0, done_value=0, stats=None):
        """`triangulate_planes` whose outputs are raw device addresses {"Q", "err", "nexcl", "mask"} — typically a
        slot in the CONSUMER GPU's memory (sharding.PeerGather) — bracketed by the flag protocol of the push path."""
        torch = _torch()
        U, Cn = x.shape
        for t in (x, y, lik):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (U, Cn)
        Pm = _as_P(P, Cn)
        thr = float("-inf") if lik_thr is None else float(lik_thr)
        _lib.check(self.h, self.lib.p2s_triangulate_planes_push_device(
            self.h, _ptr(x), _ptr(y), _ptr(lik), Pm.ctypes.data, U, Cn, thr, float(reproj_thr), int(min_cams),
            C.c_void_p(out_ptrs["Q"]), C.c_void_p(out_ptrs["err"]), C.c_void_p(out_ptrs["nexcl"]), C.c_void_p(out_ptrs["mask"]),
            _ptr(stats), C.c_void_p(wait_flag or None), int(wait_value) & 0xffffffff, C.c_void_p(done_flag or None),
            int(done_value) & 0xffffffff, self._stream()))


def bind_host_threads_to_gpu(device=0):
    """Pin this process to the CPU cores NVML reports as local to `device` (same NUMA node / PCIe root), so that
    the pinned staging buffers allocated afterwards are first-touched next to the GPU.  Returns the previous
    affinity (restore with os.sched_setaffinity(0, prev)) or None when NVML or the call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        prev = os.sched_getaffinity(0)
        want = cpus & prev
        if not want:
            return None
        os.sched_setaffinity(0, want)
        return prev
    except Exception:
        return None
