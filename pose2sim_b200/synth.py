"""Deterministic synthetic multi-view keypoint streams.

Used by bench.py, the tests and the golden-vector generator; it only produces INPUTS (cameras and
2D observations), never results.  Everything is a pure function of (seed, shapes): a counter-based
Philox stream is consumed in a fixed order, so the same call reproduces the same float32 values on
any box.

Workloads (BASELINE.json `configs`):
  cfg1  C=4 (shipped Qualisys demo cameras), F=100, N=1           seed 101
  cfg2  C=8,  F=1e5, N=1, thr=15 px, min_cams=2                    seed 202   (headline metric)
  cfg3  C=16, F=1e6, N=1, min_cams=3                               seed 303
  cfg4  C=8,  6 persons per camera, association + triangulation    seed 404
  cfg5  C in 4..32, F=1e7, min_cams=max(2, C-4)                    seed 500+C
"""
import numpy as np

HALPE_26_COUNT = 26

# Demo_SinglePerson/calibration/Calib.qca.txt converted to (K, Rodrigues vector, translation[m]) with
# the reference's own conversion (calibration.py:70-190), done once in the build container by
# oracle/make_golden.py; values are stored in tests/golden/demo_calib.npz and the TOML fixture.


def ring_cameras(C, radius=5.0, fx=1600.0, cx=540.0, cy=960.0):
    """C pinhole cameras on a ring looking at the origin, Z-up world (SURVEY.md §8(d) 'Cameras').

    Returns (P[C,3,4] float64, K[C,3,3], R[C,3,3], t[C,3]); P = K [R | t]."""
    P = np.zeros((C, 3, 4))
    Ks = np.zeros((C, 3, 3))
    Rs = np.zeros((C, 3, 3))
    ts = np.zeros((C, 3))
    for c in range(C):
        az = 2.0 * np.pi * c / C + 0.1 * c
        pos = np.array([radius * np.cos(az), radius * np.sin(az), 1.5 + 0.2 * (c % 8)])
        fwd = -pos / np.linalg.norm(pos)               # optical axis: towards the origin
        right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd])               # world -> camera
        t = -R @ pos
        K = np.array([[fx, 0.0, cx], [0.0, fx, cy], [0.0, 0.0, 1.0]])
        P[c] = K @ np.concatenate([R, t[:, None]], axis=1)
        Ks[c], Rs[c], ts[c] = K, R, t
    return P, Ks, Rs, ts


def _rng(seed):
    return np.random.Generator(np.random.Philox(key=int(seed)))


def truth_points(F, N, K, seed, frame0=0):
    """3D ground truth [F, N, K, 3] (SURVEY.md §8(d) '3D truth')."""
    g = _rng(1)                                        # keypoint offsets: seed 1, shared by all configs
    off = np.stack([g.uniform(-0.3, 0.3, K), g.uniform(-0.3, 0.3, K), g.uniform(0.0, 1.8, K)], axis=1)
    f = np.arange(frame0, frame0 + F)
    ang = 2.0 * np.pi * f / 600.0
    centre = np.stack([2.0 * np.cos(ang), 2.0 * np.sin(ang), np.zeros(F)], axis=1)   # 2 m circle
    pers = np.stack([np.array([(1.2 * p) % 3.6, -0.8 * p, 0.0]) for p in range(N)])
    jitter = _rng(seed * 7919 + 13).normal(0.0, 0.02, (F, N, K, 3))
    return centre[:, None, None, :] + pers[None, :, None, :] + off[None, None, :, :] + jitter


def observe(Q, P, seed, sigma=2.0, p_out=0.05, p_low=0.05):
    """Project Q[..., 3] with P[C,3,4]; add noise, outliers and low likelihoods.

    Returns float32 arrays x, y, lik of shape Q.shape[:-1] + (C,) — ALREADY float32-rounded, which is
    what both the GPU path and the oracle must consume (SURVEY.md 'FP32 staging').
    """
    C = P.shape[0]
    g = _rng(seed)
    Qh = np.concatenate([Q, np.ones(Q.shape[:-1] + (1,))], axis=-1)
    proj = np.einsum("cij,...j->...ci", P, Qh)
    x = proj[..., 0] / proj[..., 2]
    y = proj[..., 1] / proj[..., 2]
    shp = x.shape
    x = x + g.normal(0.0, sigma, shp)
    y = y + g.normal(0.0, sigma, shp)
    is_out = g.random(shp) < p_out
    mag = g.uniform(50.0, 300.0, shp)
    ang = g.uniform(0.0, 2.0 * np.pi, shp)
    x = np.where(is_out, x + mag * np.cos(ang), x)
    y = np.where(is_out, y + mag * np.sin(ang), y)
    lik = np.where(is_out, g.uniform(0.3, 0.7, shp), g.uniform(0.5, 1.0, shp))
    is_low = g.random(shp) < p_low
    lik = np.where(is_low, g.uniform(0.0, 0.3, shp), lik)
    assert x.shape[-1] == C
    return x.astype(np.float32), y.astype(np.float32), lik.astype(np.float32)


def gate_likelihood(x, y, lik, lik_thr):
    """Pose2Sim/triangulation.py:817-821: likelihood below threshold => x, y, likelihood = NaN."""
    with np.errstate(invalid="ignore"):
        low = lik.astype(np.float64) < float(lik_thr)          # the reference compares in float64
    x = np.where(low, np.nan, x).astype(x.dtype)
    y = np.where(low, np.nan, y).astype(y.dtype)
    lik = np.where(low, np.nan, lik).astype(lik.dtype)
    return x, y, lik


def make_triangulation_workload(C, F, N=1, K=HALPE_26_COUNT, seed=202, lik_thr=0.3, frame0=0, P=None,
                                sigma=2.0, p_out=0.05, p_low=0.05):
    """Units in (frame, person, keypoint) order: returns dict with P[C,3,4] and x,y,lik [U, C] float32,
    likelihood gate (lik < lik_thr -> NaN) already applied like triangulate_all does before the search
    (lik_thr=None: raw likelihoods, gate left to the device)."""
    if P is None:
        P = ring_cameras(C)[0]
    Q = truth_points(F, N, K, seed, frame0)
    x, y, lik = observe(Q, P, seed, sigma, p_out, p_low)
    if lik_thr is not None:                                   # None: leave the gate to the device stage kernel
        x, y, lik = gate_likelihood(x, y, lik, lik_thr)
    U = F * N * K
    return {"P": P, "x": x.reshape(U, C), "y": y.reshape(U, C), "lik": lik.reshape(U, C),
            "truth": Q.reshape(U, 3), "F": F, "N": N, "K": K, "C": C}


def make_association_workload(C, F, n_persons, seed=404, K=HALPE_26_COUNT, tracked=13, P=None,
                              sigma=2.0, p_out=0.05, p_low=0.05, p_missing=0.0):
    """Person-association frames: per (frame, camera) the persons appear in a random permutation.

    Returns dict with P, obs[F, C, n_persons, 3] float32 (x, y, likelihood of the tracked keypoint,
    persons in per-camera detection order), count[F, C] (persons detected), perm[F, C, n_persons]
    (perm[f,c,j] = true person shown at detection slot j)."""
    if P is None:
        P = ring_cameras(C)[0]
    Q = truth_points(F, n_persons, K, seed)[:, :, tracked, :]           # [F, Np, 3]
    x, y, lik = observe(Q, P, seed, sigma, p_out, p_low)                # [F, Np, C]
    g = _rng(seed * 31 + 7)
    perm = np.stack([np.stack([g.permutation(n_persons) for _ in range(C)]) for _ in range(F)])
    obs = np.empty((F, C, n_persons, 3), np.float32)
    fidx = np.arange(F)[:, None, None]
    cidx = np.arange(C)[None, :, None]
    obs[..., 0] = x.transpose(0, 2, 1)[fidx, cidx, perm]
    obs[..., 1] = y.transpose(0, 2, 1)[fidx, cidx, perm]
    obs[..., 2] = lik.transpose(0, 2, 1)[fidx, cidx, perm]
    count = np.full((F, C), n_persons, np.int32)
    if p_missing > 0:
        drop = g.random((F, C)) < p_missing
        count = np.where(drop, g.integers(0, n_persons, (F, C)), count).astype(np.int32)
    return {"P": P, "obs": obs, "count": count, "perm": perm, "truth": Q, "F": F, "C": C, "Np": n_persons}


def make_multi_person_workload(C, F, n_persons, seed=404, K=HALPE_26_COUNT, sigma=2.0, p_out=0.05, p_low=0.05,
                               p_missing=0.1, p_nan=0.02):
    """Multi-person association frames (cfg4 shape): every camera lists its detected persons in a random
    order, each with ALL K keypoints; with probability p_missing a (frame, camera) loses some persons, and
    with probability p_nan a keypoint triple is NaN (an undetected joint).

    Returns dict with models (list of {K, R, T}), P, obs[F, C, n_persons, 3 K] float32 (pose_keypoints_2d
    layout x, y, likelihood per joint), count[F, C], perm[F, C, n_persons]."""
    P, Ks, Rs, ts = ring_cameras(C)
    Q = truth_points(F, n_persons, K, seed)                             # [F, Np, K, 3]
    x, y, lik = observe(Q, P, seed, sigma, p_out, p_low)                # [F, Np, K, C]
    g = _rng(seed * 31 + 11)
    perm = np.stack([np.stack([g.permutation(n_persons) for _ in range(C)]) for _ in range(F)])
    kp = np.stack([x, y, lik], axis=-1).transpose(0, 3, 1, 2, 4)        # [F, C, Np, K, 3]
    fidx = np.arange(F)[:, None, None]
    cidx = np.arange(C)[None, :, None]
    kp = kp[fidx, cidx, perm]                                           # detection order per camera
    nan = g.random((F, C, n_persons, K)) < p_nan
    kp = np.where(nan[..., None], np.float32(np.nan), kp)
    obs = np.ascontiguousarray(kp.reshape(F, C, n_persons, 3 * K), dtype=np.float32)
    count = np.full((F, C), n_persons, np.int32)
    if p_missing > 0:
        drop = g.random((F, C)) < p_missing
        count = np.where(drop, g.integers(0, n_persons, (F, C)), count).astype(np.int32)
    models = [{"K": Ks[c], "R": Rs[c], "T": ts[c]} for c in range(C)]
    return {"P": P, "models": models, "obs": obs, "count": count, "perm": perm, "F": F, "C": C, "Np": n_persons}
