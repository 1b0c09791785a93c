"""TEST INFRASTRUCTURE — CPU restatement (NumPy) of the reference's triangulation hot path.

This module is the *checker*: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it.  The product (`pose2sim_b200/`) never does; it fails loudly
when the CUDA library is missing.

Parity pinning: the reference's own tests hold no numeric vectors for this path, so
this restatement is pinned against outputs of the UNMODIFIED reference run in the build container
(`oracle/make_golden.py` -> `tests/golden/*.npz`, checked by `tests/test_oracle_golden.py`).

Every function cites the reference file:line it follows (paths relative to the reference root).
It is written per unit / per candidate on purpose: that is the reference's cost structure
(a Python call per (frame, person, keypoint), a NumPy SVD per candidate camera subset), which makes it
the stand-in for "the reference's CPU path" when bench.py times a CPU baseline on a box where the
reference itself is absent.
"""
import itertools
import math

import numpy as np

__all__ = [
    "weighted_dlt", "reproject", "pixel_distance", "solve_subset", "undistort_points", "project_distorted",
    "triangulate_unit", "triangulate_units", "person_id_rows", "associate_frame",
]


# ---------------------------------------------------------------------------------------------
# leaf math
# ---------------------------------------------------------------------------------------------
def weighted_dlt(P_sub, x, y, w):
    """Likelihood-weighted DLT.  Pose2Sim/common.py:327-354 (`weighted_triangulation`).

    Two rows per camera, `(P[0]-x*P[2])*w` and `(P[1]-y*P[2])*w` (:344-345); with >= 4 rows the
    solution is the right singular vector of the smallest singular value, de-homogenised
    (:347-350); otherwise NaN (:351-352).  The reference calls `cv2.SVDecomp`; any SVD gives the same
    vector up to sign/rounding because the result is `V[:3,3]/V[3,3]`.
    """
    m = len(x)
    if 2 * m < 4:
        return np.array([np.nan, np.nan, np.nan, 1.0])
    A = np.empty((2 * m, 4))
    if not (np.all(np.isfinite(x)) and np.all(np.isfinite(y)) and np.all(np.isfinite(w))):
        # cv2.SVDecomp does not raise on NaN input, it returns NaN (np.linalg.svd would raise): a NaN observation
        # that is still "active" (personAssociation.py:215-216 only switches likelihood 0 off) yields Q = NaN
        return np.array([np.nan, np.nan, np.nan, 1.0])
    for c in range(m):
        Pc = P_sub[c]
        A[2 * c] = (Pc[0] - x[c] * Pc[2]) * w[c]
        A[2 * c + 1] = (Pc[1] - y[c] * Pc[2]) * w[c]
    v = np.linalg.svd(A)[2][3]
    with np.errstate(all="ignore"):
        return np.array([v[0] / v[3], v[1] / v[3], v[2] / v[3], 1.0])


def reproject(P_sub, Q):
    """Pose2Sim/common.py:357-375 (`reprojection`): x = P[0].Q / P[2].Q, y = P[1].Q / P[2].Q."""
    xs, ys = [], []
    with np.errstate(all="ignore"):
        for Pc in P_sub:
            den = Pc[2] @ Q
            xs.append(Pc[0] @ Q / den)
            ys.append(Pc[1] @ Q / den)
    return xs, ys


def pixel_distance(q1, q2):
    """Pose2Sim/common.py:378-403 (`euclidean_distance`): sqrt(nansum(d^2)); if every component of
    the difference is NaN the distance is +inf (:394-396)."""
    d = np.asarray(q2, float) - np.asarray(q1, float)
    if np.isnan(d).all():
        return np.inf
    return float(np.sqrt(np.nansum(d * d)))


def undistort_points(x, y, K, dist, new_K):
    """Pose2Sim/triangulation.py:808-813: `cv2.undistortPoints(points.astype('float32'), K, dist, None,
    optim_K)` — OpenCV's algorithm restated: 5 fixed-point iterations of the inverse radial/tangential
    model in double, re-projection with the new camera matrix, float32 result.  (cv2 itself is absent
    from this restatement on purpose; tests compare it with cv2 where cv2 is importable.)"""
    k = np.zeros(8)
    k[:len(dist)] = np.asarray(dist, float).reshape(-1)
    u = np.asarray(x, np.float32).astype(np.float64)
    v = np.asarray(y, np.float32).astype(np.float64)
    x0 = (u - K[0][2]) * (1.0 / K[0][0])
    y0 = (v - K[1][2]) * (1.0 / K[1][1])
    xx, yy = x0.copy(), y0.copy()
    with np.errstate(all="ignore"):
        for _ in range(5):
            r2 = xx * xx + yy * yy
            icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
            dx = 2 * k[2] * xx * yy + k[3] * (r2 + 2 * xx * xx)
            dy = k[2] * (r2 + 2 * yy * yy) + 2 * k[3] * xx * yy
            xx, yy = (x0 - dx) * icdist, (y0 - dy) * icdist
        nk = np.asarray(new_K, float)
        ww = 1.0 / (nk[2, 0] * xx + nk[2, 1] * yy + nk[2, 2])
        ox = (nk[0, 0] * xx + nk[0, 1] * yy + nk[0, 2]) * ww
        oy = (nk[1, 0] * xx + nk[1, 1] * yy + nk[1, 2]) * ww
    return ox.astype(np.float32), oy.astype(np.float32)


def project_distorted(lens, Q):
    """Pose2Sim/triangulation.py:472-476: `cv2.projectPoints(Q, R, T, K, dist)` — pinhole projection
    followed by OpenCV's radial/tangential model (k1 k2 p1 p2 k3 k4 k5 k6)."""
    k = np.zeros(8)
    k[:len(lens["dist"])] = np.asarray(lens["dist"], float).reshape(-1)
    K = lens["K"]
    with np.errstate(all="ignore"):
        X = np.asarray(lens["R"], float) @ np.asarray(Q[:3], float) + np.asarray(lens["T"], float)
        x, y = X[0] / X[2], X[1] / X[2]
        r2 = x * x + y * y
        r4, r6 = r2 * r2, r2 * r2 * r2
        a1, a2, a3 = 2 * x * y, r2 + 2 * x * x, r2 + 2 * y * y
        s = (1 + k[0] * r2 + k[1] * r4 + k[4] * r6) / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6)
        xd = x * s + k[2] * a1 + k[3] * a2
        yd = y * s + k[2] * a3 + k[3] * a1
        return xd * K[0][0] + K[0][2], yd * K[1][1] + K[1][2]


def solve_subset(P, x, y, w, cams, lens=None):
    """DLT + reprojection + mean pixel error over the cameras `cams` (ascending indices).

    Pose2Sim/triangulation.py:469 (DLT), :478 (reprojection), :485-489 (mean distance).
    Zero cameras -> mean of an empty list = NaN; one camera -> NaN point -> distance inf.
    `lens` (undistort_points): the re-projection goes through the lens model (:472-476).
    """
    P_sub = [P[c] for c in cams]
    xs = [x[c] for c in cams]
    ys = [y[c] for c in cams]
    ws = [w[c] for c in cams]
    Q = weighted_dlt(P_sub, xs, ys, ws)
    if lens is not None:
        proj = [project_distorted(lens[c], Q) for c in cams]
        xc, yc = [p[0] for p in proj], [p[1] for p in proj]
    else:
        xc, yc = reproject(P_sub, Q)
    d = [pixel_distance((xs[i], ys[i]), (xc[i], yc[i])) for i in range(len(cams))]
    err = float(np.mean(d)) if len(d) else float("nan")
    return Q, err


# ---------------------------------------------------------------------------------------------
# triangulation exclusion search  (Pose2Sim/triangulation.py:363-604; `lens` turns undistort_points on
# (:472-476), `swapped` / `partner` turn handle_LR_swap on (:509-579) — both off in every shipped config)
# ---------------------------------------------------------------------------------------------
def swapped_pass(P, x, y, w, xs, ys, cands, counts, C, lens=None):
    """`handle_LR_swap` branch of one level (triangulation.py:509-579), as the reference EXECUTES it.

    The reference builds `[[x] * n for x in x_files_filt]` (:518-519): every "sub-configuration" of a
    candidate is the SAME array object (the candidate's compacted valid-camera array), so the assignments
    at :525-526 accumulate in place: after the loops the first n_cams - nb_cams_off_tot compacted positions
    of every candidate hold the swapped coordinates, for every sub-configuration and for every later
    `n_cams_swapped` (the arrays stay mutated, re-assigning the same values).  Hence one evaluation per
    candidate: DLT over ALL its valid cameras (first T' = n_cams - nb_cams_off_tot positions swapped, the
    rest original, original likelihoods :529), error = mean distance over the first T' positions ONLY
    (:557-559).  Returns (min error, first candidate index with it, its Q) — np.min / argmin (:565-566) —
    or None when the swap loop's condition (:513) never holds.
    """
    T = max(counts)
    n_first = C - T
    if not 1 < n_first / 2:                     # n_cams_swapped = 1 < (n_cams - nb_cams_off_tot) / 2   (:513)
        return None
    errs, Qs = [], []
    for cand in cands:
        wl = w.copy()
        wl[list(cand)] = np.nan
        cams = [c for c in range(C) if not np.isnan(wl[c]) and wl[c] != 0.0]
        xm = [xs[c] if i < n_first else x[c] for i, c in enumerate(cams)]
        ym = [ys[c] if i < n_first else y[c] for i, c in enumerate(cams)]
        P_sub = [P[c] for c in cams]
        Q = weighted_dlt(P_sub, xm, ym, [w[c] for c in cams])
        if lens is not None:
            proj = [project_distorted(lens[c], Q) for c in cams]
            xc, yc = [p[0] for p in proj], [p[1] for p in proj]
        else:
            xc, yc = reproject(P_sub, Q)
        d = [pixel_distance((xm[i], ym[i]), (xc[i], yc[i])) for i in range(n_first)]
        errs.append(float(np.mean(d)))
        Qs.append(Q)
    errs = np.array(errs)
    b = int(errs.argmin())
    return float(errs.min()), b, Qs[b]


def triangulate_unit(x, y, w, P, thr, min_cams, lens=None, swapped=None):
    """One (frame, person, keypoint) unit.  Returns (Q[3], err, nb_cams_excluded, id_excluded_cams).
    `swapped` = (x, y) of the left/right partner keypoint turns `handle_LR_swap` on (see swapped_pass):
    when a level's swapped evaluation beats its error, error / Q / id list become the swapped winner's while
    `nb_cams_excluded` stays the un-swapped winner's (:574-577 do not touch it).

    Follows `triangulation_from_best_cameras`:
      * level loop condition `error_min > thr and n_cams - k >= min_cams`            (:408)
      * candidates = lexicographic k-subsets of ALL camera indices                   (:411)
      * excluded cameras become NaN                                                   (:426-432)
      * per candidate: list of NaN cameras (:435), count of NaN-or-zero cameras       (:436)
      * the level is abandoned, results untouched, when the WORST candidate excludes
        more than n_cams - min_cams cameras                                           (:437-441)
      * valid cameras = likelihood neither NaN nor 0                                  (:450-465)
      * error_min = nanmin, best = nanargmin (first index on ties)                    (:500-505)
      * after the loop: ids of the last evaluated level's best candidate, or all
        cameras when no level was evaluated                                           (:588-596)
      * failure -> error NaN, Q NaN                                                   (:600-602)
    """
    x = np.asarray(x, float)
    y = np.asarray(y, float)
    w = np.asarray(w, float)
    C = len(w)
    err_min = math.inf
    Q = np.array([np.nan, np.nan, np.nan])
    nexcl, ids = None, None
    k = 0
    while err_min > thr and C - k >= min_cams:
        cands = list(itertools.combinations(range(C), k))
        nan_sets, counts = [], []
        for cand in cands:
            wl = w.copy()
            wl[list(cand)] = np.nan
            nan_sets.append(np.flatnonzero(np.isnan(wl)))
            counts.append(int(np.count_nonzero(np.nan_to_num(wl) == 0)))
        if max(counts) > C - min_cams:
            break
        errs, Qs = [], []
        for cand in cands:
            wl = w.copy()
            wl[list(cand)] = np.nan
            cams = [c for c in range(C) if not np.isnan(wl[c]) and wl[c] != 0.0]
            Qc, ec = solve_subset(P, x, y, w, cams, lens)
            Qs.append(Qc)
            errs.append(ec)
        errs = np.array(errs)
        if np.all(np.isnan(errs)):
            # the reference would raise in np.nanargmin here (unreachable for sane inputs);
            # defined behaviour for the restatement: keep the first candidate, NaN error.
            best = 0
            err_min = float("nan")
        else:
            best = int(np.nanargmin(errs))
            err_min = float(errs[best])
        nexcl = counts[best]
        ids = [int(i) for i in nan_sets[best]]
        Q = Qs[best][:3].copy()
        if swapped is not None and err_min > thr:
            sw = swapped_pass(P, x, y, w, np.asarray(swapped[0], float), np.asarray(swapped[1], float), cands, counts, C, lens)
            if sw is not None and sw[0] < err_min:
                err_min = sw[0]
                ids = [int(i) for i in nan_sets[sw[1]]]
                Q = sw[2][:3].copy()
        k += 1
    if ids is None:
        ids = list(range(C))
        nexcl = C
    err = err_min
    if err_min > thr:
        err = float("nan")
        Q = np.array([np.nan, np.nan, np.nan])
    return Q, err, nexcl, ids


def triangulate_units(x, y, w, P, thr, min_cams, lens=None, partner=None):
    """Batched convenience wrapper: x, y, w are [U, C]; returns (Q[U,3], err[U], nexcl[U], mask[U])
    with mask bit c set iff camera c is in `id_excluded_cams`.  `partner` (K keypoint indices, units ordered
    (.., keypoint)) turns handle_LR_swap on: unit u's swapped coordinates are those of unit
    u - u % K + partner[u % K]  (triangulation.py:838)."""
    U = x.shape[0]
    Q = np.empty((U, 3))
    err = np.empty(U)
    nexcl = np.empty(U, np.int32)
    mask = np.zeros(U, np.uint32)
    for u in range(U):
        swapped = None
        if partner is not None:
            K = len(partner)
            up = u - u % K + int(partner[u % K])
            swapped = (x[up], y[up])
        q, e, n, ids = triangulate_unit(x[u], y[u], w[u], P, thr, min_cams, lens, swapped)
        Q[u], err[u], nexcl[u] = q, e, n
        m = 0
        for c in ids:
            m |= 1 << c
        mask[u] = m
    return Q, err, nexcl, mask


# ---------------------------------------------------------------------------------------------
# single-person association search  (Pose2Sim/personAssociation.py:67-257)
# ---------------------------------------------------------------------------------------------
def person_id_rows(n_per_cam):
    """Pose2Sim/personAssociation.py:67-99 (`persons_combinations`): cartesian product of the
    per-camera person ranges in camera order; cameras without a detection get NaN."""
    ranges = [range(n if n != 0 else 1) for n in n_per_cam]
    rows = np.array(list(itertools.product(*ranges)), float)
    rows[:, [c for c, n in enumerate(n_per_cam) if n == 0]] = np.nan
    return rows


def associate_frame(obs, n_per_cam, P, thr, lik_thr, min_cams):
    """One frame of `best_persons_and_cameras_combination` (personAssociation.py:154-257).

    obs[c][p] = (x, y, likelihood) of the tracked keypoint of person p in camera c.
    Returns (best_error, comb[C] (person index or NaN per camera), Q[3]).

      * cameras NaN in every row are "missing"                                       (:187-188)
      * level loop `error_min > thr and C - (missing + k) >= min_cams`               (:194)
      * rows visited in product order; a camera whose likelihood is below lik_thr
        becomes 0 and the row entry is set to NaN IN PLACE (persists)                (:215-216)
      * row skipped when fewer than min_cams cameras stay active                     (:219-221)
      * candidates = k-subsets of the ACTIVE cameras                                 (:222-225)
      * row skipped when every candidate error is NaN                                (:235-236)
      * error_min = nanmin, chosen candidate = np.argmin (first NaN wins if any)     (:238-240)
      * global best updated on strict '<'                                            (:242-245)
      * first row whose error_min < thr ends the level's row loop                    (:247-248)
      * nothing evaluable -> (inf, NaN.., NaN)                                       (:252-253)
    """
    C = len(n_per_cam)
    rows = person_id_rows(n_per_cam)
    n_missing = int(np.all(np.isnan(rows), axis=0).sum())
    err_min = math.inf
    best_err, best_comb, best_Q = math.inf, None, None
    k = 0
    while err_min > thr and C - (n_missing + k) >= min_cams:
        for row in rows:
            coords = np.full((C, 3), np.nan)
            for c in range(C):
                if not np.isnan(row[c]):
                    p = int(row[c])
                    if p < len(obs[c]):
                        coords[c] = obs[c][p]
            with np.errstate(invalid="ignore"):
                coords[coords[:, 2] < lik_thr, 2] = 0.0
            row[coords[:, 2] == 0.0] = np.nan
            active = np.flatnonzero(~np.isnan(row))
            if len(active) < min_cams:
                continue
            cands = list(itertools.combinations(active, k))
            errs, combs, Qs = [], [], []
            for cand in cands:
                comb = row.copy()
                comb[list(cand)] = np.nan
                cams = [c for c in range(C) if not np.isnan(comb[c])]
                Qc, ec = solve_subset(P, coords[:, 0], coords[:, 1], coords[:, 2], cams)
                errs.append(ec)
                combs.append(comb)
                Qs.append(Qc)
            if len(errs) == 0 or np.all(np.isnan(errs)):
                continue
            err_min = float(np.nanmin(errs))
            b = int(np.argmin(errs))
            if err_min < best_err:
                best_err, best_comb, best_Q = err_min, combs[b], Qs[b][:3].copy()
            if err_min < thr:
                break
        k += 1
    if best_comb is None:
        return math.inf, np.full(C, np.nan), np.full(3, np.nan)
    return best_err, best_comb, best_Q
