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


def stage_detections(parsed, check_total=True, want_length=False):
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
    if check_total and n_max > _lib.P2S_MAX_DETECTIONS:
        raise ValueError(f"{n_max} detections in one frame: the device matching handles at most "
                         f"{_lib.P2S_MAX_DETECTIONS} per frame")
    if want_length:                                             # block of a larger trial (merge_staged)
        return obs64.astype(np.float32), count, _stg.float32_inexact(obs64), (L if count.any() else None)
    return obs64.astype(np.float32), count, _stg.float32_inexact(obs64)


def merge_staged(parts):
    """Blocks of `stage_detections(..., check_total=False, want_length=True)` of consecutive frame ranges -> what one
    call over all frames returns (same checks, same padding)."""
    lengths = sorted({p[3] for p in parts if p[3] is not None})
    if len(lengths) > 1:
        raise ValueError(f"pose_keypoints_2d lists of different lengths {lengths}: the reference cannot "
                         f"stack them either (personAssociation.py:378)")
    L = lengths[0] if lengths else 3
    NP = max([p[0].shape[2] for p in parts if p[3] is not None] + [1])
    blocks = []
    for obs, count, _, n_values in parts:
        F, C = count.shape
        full = np.full((F, C, NP, L), np.nan, np.float32)
        if n_values is not None:
            full[:, :, :obs.shape[2], :] = obs
        blocks.append(full)
    obs = np.concatenate(blocks) if blocks else np.zeros((0, 0, NP, L), np.float32)
    count = np.concatenate([p[1] for p in parts]) if parts else np.zeros((0, 0), np.int32)
    n_max = int(count.sum(axis=1).max(initial=0))
    if n_max > _lib.P2S_MAX_DETECTIONS:
        raise ValueError(f"{n_max} detections in one frame: the device matching handles at most "
                         f"{_lib.P2S_MAX_DETECTIONS} per frame")
    return obs, count, int(sum(p[2] for p in parts))


def proposals_from_rows(rows, min_cams):
    """personAssociation.py:532-547 on the per-row arg-max table of one frame (rows [N, C] integers, -1 = no
    detection): unique rows ordered by multiplicity, rows reusing a detection of an earlier row dropped,
    rows seen by fewer than min_cams views dropped.  Returns [n_persons, C] float (NaN = not seen).

    Same results as the reference's calls, including tie orders: `np.unique(axis=0)` sorts rows
    lexicographically, which for small integers is the order of their base-b digit strings, so the unique
    rows / counts come from a 1-D `np.unique` on int64 keys (half the time of the axis=0 form); `np.argsort(counts)[::-1]` is the
    reference's own call on the same counts array."""
    rows = np.asarray(rows)
    if rows.ndim != 2 or rows.shape[0] == 0:
        return np.array([], dtype=float)
    n, C = rows.shape
    base = int(rows.max()) + 2                                  # digits 0 .. base-1 for values -1 .. max
    if base ** C < 2 ** 62:
        weights = base ** np.arange(C - 1, -1, -1, dtype=np.int64)
        keys = (rows.astype(np.int64) + 1) @ weights
        _, first, counts = np.unique(keys, return_index=True, return_counts=True)
        prop = rows[first].astype(float)
    else:                                                       # too wide for one int64 key: the reference's call itself
        prop, counts = np.unique(rows.astype(float), axis=0, return_counts=True)
    prop = prop[np.argsort(counts)[::-1]]
    prop[prop == -1] = np.nan
    # a row is dropped when any of its entries equals the entry of ANY earlier row in the same column (:541-542)
    same = (prop[:, None, :] == prop[None, :, :]).any(axis=2)
    keep = ~np.tril(same, -1).any(axis=1)
    prop = prop[keep]
    prop = prop[(~np.isnan(prop)).sum(axis=1) >= min_cams]
    return prop if len(prop) else np.array([], dtype=float)


def proposals_from_rows_batch(rows, n, min_cams):
    """`proposals_from_rows` for all frames at once: rows [F, NM, C] integers (-1 = no detection), n[f] = detections of
    frame f.  Same results, frame by frame (tests/test_dropin_host.py compares them on random tables): the keys, the stable
    sort that yields `np.unique`'s sorted unique rows / first occurrences / counts, and the duplicate and min-cameras
    filters run on padded [F, ...] arrays; only `np.argsort(counts)[::-1]` — whose tie order is NumPy's own — stays a call
    per frame on exactly the counts array the per-frame function passes."""
    rows = np.asarray(rows)
    F, NM, C = rows.shape
    n = np.asarray(n, dtype=np.int64)
    empty = np.array([], dtype=float)
    if F == 0:
        return []
    valid = np.arange(NM)[None, :] < n[:, None]
    r64 = rows.astype(np.int64)
    base = int(r64[valid].max(initial=-1)) + 2                  # any base > max + 1 keeps the lexicographic order of the rows
    if NM == 0 or float(base) ** C >= 2.0 ** 62:
        return [proposals_from_rows(rows[f, :n[f]], min_cams) for f in range(F)]
    weights = base ** np.arange(C - 1, -1, -1, dtype=np.int64)
    sentinel = np.iinfo(np.int64).max
    keys = np.where(valid, (r64 + 1) @ weights, sentinel)
    order = np.argsort(keys, axis=1, kind="stable")             # equal keys keep their index order: first = first occurrence
    sk = np.take_along_axis(keys, order, axis=1)
    start = np.ones((F, NM), bool)
    start[:, 1:] = sk[:, 1:] != sk[:, :-1]
    start &= sk != sentinel                                      # run starts of the sorted valid keys = the unique rows, ascending
    n_unique = start.sum(axis=1)
    U = int(n_unique.max(initial=0))
    if U == 0:
        return [empty for _ in range(F)]
    fi, pos = np.nonzero(start)                                  # row-major: frame by frame, ascending key
    first_of_frame = np.concatenate([[0], np.cumsum(n_unique)[:-1]])
    slot = np.arange(len(fi)) - first_of_frame[fi]
    run_end = np.empty(len(fi), np.int64)                        # position after the run: next start of the frame, or n[f]
    run_end[:-1] = pos[1:]
    run_end[-1] = 0
    last = np.ones(len(fi), bool)
    last[:-1] = fi[1:] != fi[:-1]
    run_end[last] = n[fi[last]]
    counts_flat = run_end - pos
    first_flat = order[fi, pos]                                  # np.unique's return_index
    # order by multiplicity: NumPy's own argsort on each frame's counts (ties are its business), then padded [F, U, C]
    bounds = np.concatenate([first_of_frame, [len(fi)]]).tolist()
    sel = np.empty(len(fi), np.int64)
    for f in np.flatnonzero(n_unique).tolist():
        a, b = bounds[f], bounds[f + 1]
        sel[a:b] = first_flat[a:b][counts_flat[a:b].argsort()[::-1]]      # = np.argsort(counts)[::-1], the reference's call
    prop = np.full((F, U, C), -1, np.int64)
    filled = np.zeros((F, U), bool)
    prop[fi, slot] = r64[fi, sel]
    filled[fi, slot] = True
    # a row is dropped when any of its entries equals the entry of ANY earlier row in the same column (:541-542; -1 never matches)
    # — i.e. when one of its entries is not the FIRST occurrence of that detection in its column: one pass per detection index
    # over [F, U, C] instead of the [F, U, U, C] comparison table
    seen = prop >= 0
    upos = np.arange(U)[None, :, None]
    repeat = np.zeros((F, U, C), bool)
    for v in range(int(prop.max(initial=-1)) + 1):
        m = prop == v
        repeat |= m & (upos > m.argmax(axis=1)[:, None, :])     # argmax of booleans = first True of the column
    keep = filled & ~repeat.any(axis=2) & (seen.sum(axis=2) >= min_cams)
    propf = np.where(seen, prop, 0).astype(float)
    propf[~seen] = np.nan
    kept = keep.sum(axis=1)
    out = np.split(propf[keep], np.cumsum(kept)[:-1])           # row-major: frame by frame, in proposal order
    for f in np.flatnonzero(kept == 0).tolist():
        out[f] = empty                                          # what the per-frame function returns for "nobody"
    return out


def associate_frames(engine, obs, count, models, max_distance, min_affinity, min_cams):
    """All frames in ONE device call; returns the list of proposals per frame."""
    n_max = max(1, int(count.sum(axis=1).max(initial=0)))
    out = engine.associate_multi_host(obs, count, models, max_distance, min_affinity, n_max=n_max)
    return proposals_from_rows_batch(out["rows"], count.sum(axis=1), min_cams)
