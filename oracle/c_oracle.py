"""TEST INFRASTRUCTURE — ctypes loader of the plain-C oracle (oracle/p2s_oracle.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libp2s_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_PATH)
        _lib.p2s_oracle_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(load().p2s_oracle_max_threads())


def triangulate_units(x, y, w, P, thr, min_cams):
    """x, y, w: [U, C] float32.  Returns Q[U,3], err[U], nexcl[U] u8, mask[U] u32, level[U] i32, n_candidates."""
    lib = load()
    x, y, w = (np.ascontiguousarray(a, np.float32) for a in (x, y, w))
    U, Cn = x.shape
    P = np.ascontiguousarray(np.asarray(P, np.float64).reshape(Cn, 12))
    Q = np.empty((U, 3)); err = np.empty(U); nexcl = np.empty(U, np.uint8); mask = np.empty(U, np.uint32)
    level = np.empty(U, np.int32); nc = C.c_longlong(0)
    lib.p2s_oracle_triangulate(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_void_p(w.ctypes.data),
                               C.c_void_p(P.ctypes.data), C.c_longlong(U), C.c_int(Cn), C.c_double(thr), C.c_int(min_cams),
                               C.c_void_p(Q.ctypes.data), C.c_void_p(err.ctypes.data), C.c_void_p(nexcl.ctypes.data),
                               C.c_void_p(mask.ctypes.data), C.c_void_p(level.ctypes.data), C.byref(nc))
    return Q, err, nexcl, mask, level, nc.value


def triangulate_units_lr_swap(x, y, w, partner, P, thr, min_cams):
    """`handle_LR_swap = true`: x, y, w [U, C] float32 with the units ordered (.., keypoint), partner = K keypoint
    indices.  Returns Q[U,3], err[U], nexcl[U] u8, mask[U] u32."""
    lib = load()
    x, y, w = (np.ascontiguousarray(a, np.float32) for a in (x, y, w))
    U, Cn = x.shape
    part = np.ascontiguousarray(partner, np.int32).ravel()
    assert U % part.size == 0 and part.min() >= 0 and part.max() < part.size
    P = np.ascontiguousarray(np.asarray(P, np.float64).reshape(Cn, 12))
    Q = np.empty((U, 3)); err = np.empty(U); nexcl = np.empty(U, np.uint8); mask = np.empty(U, np.uint32)
    lib.p2s_oracle_triangulate_lrswap(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_void_p(w.ctypes.data),
                                      C.c_void_p(part.ctypes.data), C.c_int(part.size), C.c_void_p(P.ctypes.data),
                                      C.c_longlong(U), C.c_int(Cn), C.c_double(thr), C.c_int(min_cams),
                                      C.c_void_p(Q.ctypes.data), C.c_void_p(err.ctypes.data), C.c_void_p(nexcl.ctypes.data),
                                      C.c_void_p(mask.ctypes.data))
    return Q, err, nexcl, mask


def associate_frames(obs, count, P, thr, lik_thr, min_cams):
    """obs: [F, C, NP, 3|4] float32; count [F, C] int32.  Returns err[F], comb[F,C] int8, Q[F,3]."""
    lib = load()
    obs = np.asarray(obs, np.float32)
    F, Cn, NP = obs.shape[:3]
    if obs.shape[3] == 3:
        o4 = np.zeros((F, Cn, NP, 4), np.float32)
        o4[..., :3] = obs
        obs = o4
    obs = np.ascontiguousarray(obs)
    count = np.ascontiguousarray(count, np.int32)
    P = np.ascontiguousarray(np.asarray(P, np.float64).reshape(Cn, 12))
    err = np.empty(F); comb = np.empty((F, Cn), np.int8); Q = np.empty((F, 3))
    lib.p2s_oracle_associate(C.c_void_p(obs.ctypes.data), C.c_void_p(count.ctypes.data), C.c_void_p(P.ctypes.data),
                             C.c_longlong(F), C.c_int(Cn), C.c_int(NP), C.c_double(thr), C.c_double(lik_thr),
                             C.c_int(min_cams), C.c_void_p(err.ctypes.data), C.c_void_p(comb.ctypes.data),
                             C.c_void_p(Q.ctypes.data))
    return err, comb, Q
