"""Counter-based synthetic multi-view keypoint streams: a PURE FUNCTION of (seed, unit, camera).

SURVEY.md §8(d) asks for a generator that a device kernel can evaluate per shard (BASELINE configs[4] — 10 M
frames x 32 cameras would be 100 GB of host-made planes) and that a NumPy twin can regenerate for any subsample so
the oracle sees bit-identical float32 inputs.  This module is the twin; `csrc/p2s_synth.cu::synth_kernel` is the
device side (`ops.Engine.synth_observations`).  Both evaluate

    Philox4x32-10( counter = (unit_lo, unit_hi, camera, stream), key = (seed, 0x5032534D) )

and use only IEEE add / multiply / divide in a fixed order (no fused multiply-add, no transcendental function), so
the two sides agree bit for bit (`tests/test_synth_philox.py` on the CPU, `tests/test_gpu_synth.py` on the device).

Scene = SURVEY §8(d): ring cameras (`synth.ring_cameras`), per-keypoint offsets, the person walking a 2 m circle
with period 600 frames (cos / sin from a 600-entry host table handed to the kernel), 2 cm jitter; observations =
projection + noise of standard deviation `sigma` (sum of four 16-bit uniforms, i.e. Irwin-Hall(4): bell-shaped with
exact integer arithmetic), outliers of 50-300 px in one of 256 tabulated directions with probability p_out,
likelihoods U(0.5, 1) / U(0.3, 0.7) for outliers / U(0, 0.3) with probability p_low.  The likelihood gate is NOT
applied (the kernel's fused gate does that).  Units are (frame, keypoint) in that order, one person.
"""
import numpy as np

from . import synth

KEY1 = 0x5032534D                     # "P2SM"
M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
CAM_UNIT = 0xFFFF                     # "camera" index of the per-unit (camera-independent) draws
PERIOD = 600
N_DIR = 256
SQRT3 = 1.7320508075688772            # Irwin-Hall(4) of variance 1/3 -> unit variance


def philox4x32(c0, c1, c2, c3, k0, k1=KEY1, rounds=10):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  Inputs broadcast; returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & np.uint64(0xFFFFFFFF) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint64(int(k0) & 0xFFFFFFFF)
    k1 = np.uint64(int(k1) & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _u24(r):
    """uint32 -> double in [0, 1) with 24 random bits (exact)."""
    return (r >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)


def _ih4(ra, rb):
    """Two uint32 -> unit-variance bell-shaped double: (sum of four 16-bit uniforms - 2) * sqrt(3), exact integer sum."""
    s = ((ra & np.uint32(0xFFFF)).astype(np.int64) + (ra >> np.uint32(16)).astype(np.int64) +
         (rb & np.uint32(0xFFFF)).astype(np.int64) + (rb >> np.uint32(16)).astype(np.int64))
    return (s - 131070).astype(np.float64) * (SQRT3 / 65536.0)      # 4 * 65535 / 2 = 131070: centred exactly


def tables(K):
    """Host tables the device kernel receives verbatim: keypoint offsets [K, 3] (the same as synth.truth_points),
    the walk circle [600, 2] and the outlier directions [256, 2]."""
    g = synth._rng(1)
    off = np.stack([g.uniform(-0.3, 0.3, K), g.uniform(-0.3, 0.3, K), g.uniform(0.0, 1.8, K)], axis=1)
    a = 2.0 * np.pi * np.arange(PERIOD) / PERIOD
    circle = np.stack([2.0 * np.cos(a), 2.0 * np.sin(a)], axis=1)
    d = 2.0 * np.pi * np.arange(N_DIR) / N_DIR
    dirs = np.stack([np.cos(d), np.sin(d)], axis=1)
    return np.ascontiguousarray(off), np.ascontiguousarray(circle), np.ascontiguousarray(dirs)


def truth(units, K, seed, off, circle):
    """3D truth [len(units), 3] of the given unit indices (frame = unit // K, keypoint = unit % K)."""
    units = np.asarray(units, np.int64)
    f, k = units // K, units % K
    r = philox4x32(units & 0xFFFFFFFF, units >> 32, CAM_UNIT, 0, seed)
    r2 = philox4x32(units & 0xFFFFFFFF, units >> 32, CAM_UNIT, 1, seed)
    jx, jy, jz = _ih4(r[0], r[1]) * 0.02, _ih4(r[2], r[3]) * 0.02, _ih4(r2[0], r2[1]) * 0.02
    c = circle[f % PERIOD]
    X = (c[:, 0] + off[k, 0]) + jx
    Y = (c[:, 1] + off[k, 1]) + jy
    Z = off[k, 2] + jz
    return np.stack([X, Y, Z], axis=1)


def observations(units, P, K, seed, sigma=2.0, p_out=0.05, p_low=0.05, tabs=None):
    """x, y, lik float32 [len(units), C] and the truth [len(units), 3] for the given unit indices."""
    units = np.asarray(units, np.int64)
    C = P.shape[0]
    off, circle, dirs = tabs if tabs is not None else tables(K)
    Q = truth(units, K, seed, off, circle)
    Pm = np.asarray(P, np.float64).reshape(C, 12)
    X, Y, Z = (Q[:, i][:, None] for i in range(3))
    # (((P0 X + P1 Y) + P2 Z) + P3): separate multiplies and adds, the order the kernel uses
    hu = ((Pm[None, :, 0] * X + Pm[None, :, 1] * Y) + Pm[None, :, 2] * Z) + Pm[None, :, 3]
    hv = ((Pm[None, :, 4] * X + Pm[None, :, 5] * Y) + Pm[None, :, 6] * Z) + Pm[None, :, 7]
    hd = ((Pm[None, :, 8] * X + Pm[None, :, 9] * Y) + Pm[None, :, 10] * Z) + Pm[None, :, 11]
    cam = np.arange(C, dtype=np.int64)[None, :]
    ulo, uhi = (units & 0xFFFFFFFF)[:, None], (units >> 32)[:, None]
    a = philox4x32(ulo, uhi, cam, 0, seed)
    b = philox4x32(ulo, uhi, cam, 1, seed)
    x = hu / hd + _ih4(a[0], a[1]) * sigma
    y = hv / hd + _ih4(a[2], a[3]) * sigma
    is_out = _u24(b[0]) < p_out
    mag = 50.0 + 250.0 * _u24(b[1])
    d = dirs[(b[2] & np.uint32(0xFF)).astype(np.int64)]
    x = np.where(is_out, x + mag * d[..., 0], x)
    y = np.where(is_out, y + mag * d[..., 1], y)
    u_l = _u24(b[3])
    lik = np.where(is_out, 0.3 + 0.4 * u_l, 0.5 + 0.5 * u_l)
    is_low = _u24(b[2]) < p_low                                           # bits 8..31 of the third word (0..7: direction)
    lik = np.where(is_low, 0.3 * u_l, lik)
    return x.astype(np.float32), y.astype(np.float32), lik.astype(np.float32), Q


def make_workload(C, unit0, n_units, K=26, seed=500, sigma=2.0, p_out=0.05, p_low=0.05, P=None):
    """NumPy twin of one shard [unit0, unit0 + n_units): dict with P, x, y, lik [n_units, C] float32 (ungated), truth."""
    if P is None:
        P = synth.ring_cameras(C)[0]
    units = np.arange(unit0, unit0 + n_units, dtype=np.int64)
    x, y, lik, Q = observations(units, P, K, seed, sigma, p_out, p_low)
    return {"P": P, "x": x, "y": y, "lik": lik, "truth": Q, "C": C, "K": K}
