"""TEST INFRASTRUCTURE — NumPy restatement of the reference's multi-person cross-view association
(the `multi_person = true` branch of `associate_all`, Pose2Sim/personAssociation.py:783-801): Plücker-ray
affinity between every pair of detections of different cameras (:277-408), the one-person-per-view
constraint (:411-428), low-rank matching by singular-value thresholding (:431-509), proposal extraction
(:512-549).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product
path (pose2sim_b200/multi_person.py) runs the arithmetic in `mp_associate_kernel` on the GPU.

Pinned by tests/golden/e2e_assoc_multi.npz (JSON written by the live reference on a 4-camera, 3-person
trial, oracle/make_golden_e2e.py): the proposals of this restatement give the same files.
"""
import numpy as np


def camera_ray_params(models):
    """Per camera: inverse intrinsics, world-from-camera rotation, translation, optical centre
    (common.py:254-288 `retrieve_calib_params`: inv_K, R_mat = Rodrigues(rotation), T).
    models: list of {"K": 3x3, "R": 3x3 rotation matrix, "T": 3}."""
    out = []
    for cam in models:
        K = np.array(cam["K"], dtype=np.float64).reshape(3, 3)
        R = np.array(cam["R"], dtype=np.float64).reshape(3, 3)
        T = np.array(cam["T"], dtype=np.float64).reshape(3)
        out.append({"inv_K": np.linalg.inv(K), "Rt": R.T, "T": T, "centre": -R.T @ T})
    return out


def person_rays(keypoints, cam):
    """personAssociation.py:277-318 `compute_rays` for one detection: keypoints = flat [x, y, lik, ...].
    Returns [J, 7]: unit direction of the camera->keypoint ray, its moment about the origin, likelihood;
    a joint with any NaN becomes seven zeros (zero weight)."""
    kp = np.asarray(keypoints, dtype=np.float64)
    x, y, lik = kp[0::3], kp[1::3], kp[2::3]
    pix = np.stack([x, y, np.ones_like(x)], axis=1)                         # [J, 3]
    with np.errstate(invalid="ignore", divide="ignore"):
        world = (cam["Rt"] @ ((cam["inv_K"] @ pix.T) - cam["T"][:, None])).T   # R^T (K^-1 q - T)
        line = world - cam["centre"]
        line = line / np.linalg.norm(line, axis=1, keepdims=True)
        moment = np.cross(np.broadcast_to(cam["centre"], line.shape), line)
    rays = np.concatenate([line, moment, lik[:, None]], axis=1)
    rays[np.isnan(rays).any(axis=1)] = 0.0
    return rays


def ray_affinity(detections, cams, cum, max_distance):
    """personAssociation.py:347-408 `compute_affinity`.  detections[c] = list of flat keypoint lists of
    camera c; cum = cumulative person counts.  Affinity = 1 - d / max_distance with d the
    likelihood-weighted mean |reciprocal product| of the two detections' joint rays, clamped at
    max_distance; pairs of the same camera (and cameras without detection) keep d = 2 max_distance."""
    rays = [np.array([person_rays(p, cams[c]) for p in det]) for c, det in enumerate(detections)]
    n = cum[-1]
    dist = np.zeros((n, n)) + 2 * max_distance
    C = len(detections)
    for c0 in range(C):
        for c1 in range(c0 + 1, C):
            if cum[c0] == cum[c0 + 1] or cum[c1] == cum[c1 + 1]:
                continue
            a, b = rays[c0][:, None], rays[c1][None, :]                    # [n0, 1, J, 7], [1, n1, J, 7]
            recip = np.abs(np.sum(a[..., :3] * b[..., 3:6], axis=-1) + np.sum(b[..., :3] * a[..., 3:6], axis=-1))
            w = np.sqrt(a[..., -1] * b[..., -1])
            d = np.sum(recip * w, axis=-1) / (1e-5 + w.sum(axis=-1))
            dist[cum[c0]:cum[c0 + 1], cum[c1]:cum[c1 + 1]] = d
            dist[cum[c1]:cum[c1 + 1], cum[c0]:cum[c0 + 1]] = d.T
    dist[dist > max_distance] = max_distance
    return 1 - dist / max_distance


def view_constraint(cum):
    """personAssociation.py:411-428: 1 on the diagonal and between detections of different views."""
    n = cum[-1]
    view = np.repeat(np.arange(len(cum) - 1), np.diff(cum))
    return ((view[:, None] != view[None, :]) | np.eye(n, dtype=bool)).astype(np.float64)


def shrink_singular_values(matrix, tau):
    """personAssociation.py:431-447 `SVT`."""
    U, s, Vt = np.linalg.svd(matrix)
    return U @ np.diag(np.maximum(s - tau, 0)) @ Vt


def match_svt(affinity, cum, constraint, max_iter=20, w_rank=50, tol=1e-4, w_sparse=0.1):
    """personAssociation.py:450-509 `matchSVT`: ADMM-style alternation between a low-rank step
    (singular-value shrinkage by w_rank / mu) and the projection on [0, 1] with zero same-view blocks,
    unit diagonal and symmetry; mu doubles / halves on the ratio of primal and dual residuals."""
    X = affinity.copy()
    n = X.shape[0]
    di = np.arange(n)
    X[di, di] = 0.0
    Y = np.zeros_like(X)
    W = w_sparse - X
    mu = 64
    for _ in range(max_iter):
        X_prev = X.copy()
        Q = shrink_singular_values(X + Y * 1.0 / mu, w_rank / mu)
        X = Q - (W + Y) / mu
        for i in range(len(cum) - 1):
            X[cum[i]:cum[i + 1], cum[i]:cum[i + 1]] = 0
        X[di, di] = 1.0
        X[X < 0] = 0
        X[X > 1] = 1
        X = X * constraint
        X = (X + X.T) / 2
        Y = Y + mu * (X - Q)
        primal = np.linalg.norm(X - Q) / n
        dual = mu * np.linalg.norm(X - X_prev) / n
        if primal < tol and dual < tol:
            break
        if primal > 10 * dual:
            mu = 2 * mu
        elif dual > 10 * primal:
            mu = mu / 2
    return X


def argmax_rows(affinity, cum):
    """personAssociation.py:526-531: per row the arg-max detection of every view, -1 when none is above 0
    (what the device kernel returns per frame)."""
    n_views = len(cum) - 1
    rows = np.full((affinity.shape[0], n_views), -1, dtype=np.int8)
    for r in range(affinity.shape[0]):
        for v in range(n_views):
            seg = affinity[r, cum[v]:cum[v + 1]]
            if len(seg) > 0 and max(seg) > 0:
                rows[r, v] = np.argmax(seg)
    return rows


def frame_affinity(detections, cams, max_distance, min_affinity):
    """personAssociation.py:789-800: matched, thresholded affinity of one frame and the cumulative counts."""
    cum = np.cumsum([0] + [len(d) for d in detections])
    affinity = ray_affinity(detections, cams, cum, max_distance)
    constraint = view_constraint(cum)
    affinity = affinity * constraint
    affinity = match_svt(affinity, cum, constraint)
    affinity[affinity < min_affinity] = 0
    return affinity, cum


def proposals_from_affinity(affinity, cum, min_cams):
    """personAssociation.py:512-549 `person_index_per_cam`: per row the arg-max detection of every view
    (-1 when the view has none above 0), unique rows ordered by multiplicity, rows that reuse a
    detection of an earlier row dropped, rows seen by fewer than min_cams views dropped."""
    n_views = len(cum) - 1
    rows = []
    for r in range(affinity.shape[0]):
        row = []
        for v in range(n_views):
            seg = affinity[r, cum[v]:cum[v + 1]]
            row.append(np.argmax(seg) if (len(seg) > 0 and max(seg) > 0) else -1)
        rows.append(row)
    prop = np.array(rows, dtype=float)
    prop, counts = np.unique(prop, axis=0, return_counts=True)
    prop = prop[np.argsort(counts)[::-1]]
    prop[prop == -1] = np.nan
    keep = np.ones(prop.shape[0], dtype=bool)
    for i in range(1, len(prop)):
        keep[i] = ~np.any(prop[i] == prop[:i], axis=0).any()
    prop = prop[keep]
    seen = [np.count_nonzero(~np.isnan(p)) for p in prop]
    return np.array([p for n, p in zip(seen, prop) if n >= min_cams])


def associate_frame(detections, cams, max_distance, min_affinity, min_cams):
    """One frame of the multi-person branch (personAssociation.py:783-801).  Returns proposals
    [n_persons, n_cams] (detection index per camera, NaN = not seen)."""
    cum = np.cumsum([0] + [len(d) for d in detections])
    affinity = ray_affinity(detections, cams, cum, max_distance)
    constraint = view_constraint(cum)
    affinity = affinity * constraint
    affinity = match_svt(affinity, cum, constraint)
    affinity[affinity < min_affinity] = 0
    return proposals_from_affinity(affinity, cum, min_cams)
