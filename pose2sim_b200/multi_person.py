"""Multi-person cross-view association — the `multi_person = true` branch of `associate_all`
(Pose2Sim/personAssociation.py:783-801), SURVEY §8(f) row 3.

    stage_detections()      host     :783-790  `read_json` lists -> obs[F, C, NP, 3 J] float32, count[F, C]
    Engine.associate_multi  DEVICE   :793-800 + :526-531  ray affinity, view constraint, matchSVT, min_affinity
                                     cut, per-row / per-view arg-max — `mp_associate_kernel`, one CTA per frame
    proposals_from_rows()   host     :535-547  integer bookkeeping on <= 64 small rows per frame (unique rows,
                                     ordering by multiplicity, duplicate and min-cameras filters)

All floating-point work is on the GPU; nothing here computes an affinity or an SVD, and there is no CPU
fallback (the NumPy restatement lives in oracle/p2s_oracle_mp.py, test infrastructure).
"""
import numpy as np

from . import _lib
from . import staging as _stg


def stage_detections(parsed):
    """parsed[f][c] = decoded JSON (or None).  Returns obs [F, C, NP, 3 J] float32, count [F, C] int32 and the
    number of values float32 cannot hold exactly.  Detections are `read_json`'s lists (people with >= 3
    values, personAssociation.py:260-274).  All keypoint lists must have the same length: the reference
    stacks them with np.array (:378) and broadcasts across cameras (:390-392), which raises otherwise."""
    F = len(parsed)
    C = len(parsed[0]) if F else 0
    people = [[_stg.read_people(js) if js is not None else [] for js in row] for row in parsed]
    lengths = {len(p) for row in people for cam in row for p in cam}
    if len(lengths) > 1:
        raise ValueError(f"pose_keypoints_2d lists of different lengths {sorted(lengths)}: the reference cannot "
                         f"stack them either (personAssociation.py:378)")
    L = lengths.pop() if lengths else 3
    if L % 3:
        raise ValueError(f"pose_keypoints_2d has {L} values, not a multiple of 3")
    NP = max([len(cam) for row in people for cam in row] + [1])
    count = np.zeros((F, C), np.int32)
    obs64 = np.full((F, C, NP, L), np.nan, np.float64)
    for f, row in enumerate(people):
        for c, cam in enumerate(row):
            count[f, c] = len(cam)
            for p, kp in enumerate(cam):
                obs64[f, c, p] = [np.nan if v is None else v for v in kp]
    n_max = int(count.sum(axis=1).max(initial=0))
    if n_max > _lib.P2S_MAX_DETECTIONS:
        raise ValueError(f"{n_max} detections in one frame: the device matching handles at most "
                         f"{_lib.P2S_MAX_DETECTIONS} per frame")
    return obs64.astype(np.float32), count, _stg.float32_inexact(obs64)


def proposals_from_rows(rows, min_cams):
    """personAssociation.py:532-547 on the per-row arg-max table of one frame (rows [N, C] integers, -1 = no
    detection): unique rows ordered by multiplicity, rows reusing a detection of an earlier row dropped,
    rows seen by fewer than min_cams views dropped.  Returns [n_persons, C] float (NaN = not seen).
    The calls are the reference's own (np.unique / argsort), so ties order identically."""
    prop = np.array(rows, dtype=float)
    if prop.ndim != 2:
        prop = prop.reshape(0, 0)
    prop, counts = np.unique(prop, axis=0, return_counts=True)
    prop = prop[np.argsort(counts)[::-1]]
    prop[prop == -1] = np.nan
    keep = np.ones(prop.shape[0], dtype=bool)
    for i in range(1, len(prop)):
        keep[i] = ~np.any(prop[i] == prop[:i], axis=0).any()
    prop = prop[keep]
    seen = [np.count_nonzero(~np.isnan(p)) for p in prop]
    return np.array([p for n, p in zip(seen, prop) if n >= min_cams])


def associate_frames(engine, obs, count, models, max_distance, min_affinity, min_cams):
    """All frames in ONE device call; returns the list of proposals per frame."""
    n_max = max(1, int(count.sum(axis=1).max(initial=0)))
    out = engine.associate_multi_host(obs, count, models, max_distance, min_affinity, n_max=n_max)
    n = count.sum(axis=1)
    return [proposals_from_rows(out["rows"][f, :n[f]], min_cams) for f in range(len(count))]
