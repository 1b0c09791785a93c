"""ctypes binding of the C-ABI library `libp2s_b200.so` (include/pose2sim_b200.h).

There is no CPU fallback: if the library is missing, or no sm_100 GPU is present, the product path
raises.  `load()` only dlopens (works on a CPU box, used by the "exports every symbol" test);
creating a handle needs a B200.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# P2S_LIB: another build of the same library (kernel A/B experiments, tools/kernel_ab.py)
LIB_PATH = os.environ.get("P2S_LIB") or os.path.join(_HERE, "libp2s_b200.so")

P2S_MAX_CAMS = 32
P2S_MAX_PERSONS = 16
P2S_MAX_DETECTIONS = 64
P2S_MAX_PEERS = 16
P2S_IPC_HANDLE_BYTES = 64
P2S_STAT_COUNT = 48
STAT_LEVEL0 = 0
STAT_NOT_EVALUATED = 33
STAT_FAILED = 34
STAT_CANDIDATES = 35
STAT_CAM_SOLVES = 36
STAT_BAND_THRESHOLD = 37
STAT_BAND_ARGMIN = 38
STAT_NEWTON_STEPS = 39
STAT_SOLVED = 40
STAT_DIRECT_CAMS = 41
STAT_BLOCKS = 42
STAT_ENTRY_ADDS = 43
STAT_WIDE_UNITS = 44

STATUS = {0: "P2S_OK", 1: "P2S_EINVAL", 2: "P2S_ENODEVICE", 3: "P2S_ECUDA", 4: "P2S_ENOMEM", 5: "P2S_ETOODEEP"}


class DeviceInfo(C.Structure):
    _fields_ = [("device", C.c_int), ("sm_count", C.c_int), ("cc_major", C.c_int), ("cc_minor", C.c_int),
                ("clock_khz", C.c_int), ("total_mem", C.c_size_t), ("name", C.c_char * 128)]


class CameraModel(C.Structure):
    """p2s_camera_model (include/pose2sim_b200.h): lens model for `undistort_points = true`."""
    _fields_ = [("K", C.c_double * 9), ("dist", C.c_double * 8), ("R", C.c_double * 9), ("T", C.c_double * 3),
                ("newK", C.c_double * 9)]


class P2SError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        super().__init__(f"{STATUS.get(status, status)}: {detail}")


_vp, _i, _ll, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_double

# name -> (restype, argtypes); must list every function include/pose2sim_b200.h declares
SIGNATURES = {
    "p2s_create": (_i, [_i, C.POINTER(_vp)]),
    "p2s_destroy": (_i, [_vp]),
    "p2s_status_string": (C.c_char_p, [_i]),
    "p2s_last_cuda_error": (C.c_char_p, [_vp]),
    "p2s_get_device_info": (_i, [_vp, C.POINTER(DeviceInfo)]),
    "p2s_set_band_eps": (_i, [_vp, _d]),
    "p2s_set_solver": (_i, [_vp, _i]),
    "p2s_set_assoc_team": (_i, [_vp, _i]),
    "p2s_set_chunk_units": (_i, [_vp, _ll]),
    "p2s_set_output_mode": (_i, [_vp, _i]),
    "p2s_set_search_mode": (_i, [_vp, _i]),
    "p2s_set_deep_search": (_i, [_vp, _ll]),
    "p2s_set_host_mode": (_i, [_vp, _i]),
    "p2s_obs_bytes": (C.c_size_t, [_ll, _i]),
    "p2s_stage_observations_device": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _d, _vp, _vp]),
    "p2s_triangulate_device": (_i, [_vp, _vp, _vp, _ll, _i, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "p2s_triangulate_planes_device": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "p2s_triangulate_host": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _vp]),
    "p2s_stage_undistort_device": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _d, _vp, _vp, _vp]),
    "p2s_triangulate_distorted_device": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "p2s_triangulate_lrswap_device": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _ll, _i, _d, _i, _vp, _vp, _vp, _vp, _vp]),
    "p2s_triangulate_undistort_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _vp]),
    "p2s_associate_device": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _vp]),
    "p2s_associate_host": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _d, _d, _i, _vp, _vp, _vp, _vp]),
    "p2s_associate_multi_device": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _d, _d, _vp, _vp, _vp, _vp]),
    "p2s_associate_multi_host": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _d, _d, _vp, _vp, _vp]),
    "p2s_peer_alloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp), _vp]),
    "p2s_peer_open": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "p2s_peer_close": (_i, [_vp, _vp]),
    "p2s_peer_free": (_i, [_vp, _vp]),
    "p2s_triangulate_planes_push_device": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _d, _d, _i, _vp, _vp, _vp, _vp, _vp,
                                                _vp, C.c_uint, _vp, C.c_uint, _vp]),
    "p2s_peer_collect_device": (_i, [_vp, _vp, _i, C.c_uint, _vp, _vp]),
    "p2s_peer_error": (_i, [_vp, C.POINTER(C.c_uint)]),
    "p2s_read_pose_files": (_i, [_vp, _ll, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "p2s_index_open": (_i, [_vp, _i, C.POINTER(_vp)]),
    "p2s_index_close": (None, [_vp]),
    "p2s_index_file_count": (_ll, [_vp, _i]),
    "p2s_index_file_name": (C.c_char_p, [_vp, _i, _ll]),
    "p2s_index_build_table": (_i, [_vp, _ll, _ll]),
    "p2s_index_table_paths": (_vp, [_vp]),
    "p2s_index_table_arena": (_vp, [_vp, _vp]),
    "p2s_index_signature": (_i, [_vp, _vp, _i]),
    "p2s_stat_files": (_i, [_vp, _ll, _vp, _vp, _i]),
    "p2s_read_people_files": (_i, [_vp, _ll, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i]),
    "p2s_rewrite_people_files": (_i, [_vp, _vp, _ll, _i, _vp, _vp, _vp, _i]),
    "p2s_write_trc_rows": (_i, [C.c_char_p, _vp, _vp, _vp, _ll, _i]),
    "p2s_format_trc_rows": (_i, [_vp, _vp, _vp, _ll, _i, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "p2s_synth_observations_device": (_i, [_vp, _vp, _i, _i, C.c_uint, _ll, _ll, _d, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "p2s_measure_fp64_peak": (_i, [_vp, C.POINTER(_d), C.POINTER(_d)]),
    "p2s_launch_count": (_ll, [_vp]),
    "p2s_last_grid": (_i, [_vp]),
}

_lib = None


def load():
    """dlopen the in-tree library and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise P2SError(2, f"{LIB_PATH} not built — run `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"or `make -C pose2sim_b200/csrc`; there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if "P2S_LIB" in os.environ and not hasattr(lib, name):
            continue                                  # an A/B build of an older checkout (tools/kernel_ab.py) may lack new entry points
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(handle, status):
    if status != 0:
        lib = load()
        detail = lib.p2s_status_string(status).decode()
        if handle:
            cuda = lib.p2s_last_cuda_error(handle).decode()
            if cuda:
                detail += f" [{cuda}]"
        raise P2SError(status, detail)
