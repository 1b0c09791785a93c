"""Drop-in `triangulate_all(config_dict)` — same entry point, inputs and TRC output as
Pose2Sim/triangulation.py:656-959, with the per-(frame, person, keypoint) Python loop (:831-845,
`triangulation_from_best_cameras` :363-604) replaced by ONE batched call into the sm_100a library
(`p2s_triangulate_host`, include/pose2sim_b200.h).

    stage_project()     host   :678-821   config, calibration TOML, skeleton order, JSON -> [F, N, K, C] arrays
    solve_units()       DEVICE :831-845   likelihood gate + weighted DLT + camera-exclusion search, all units
    reidentify()        host   :847-865   multi-person re-ID, sequential over frames (common.py:1037-1136)
    write_outputs()     host   :877-959   interpolation, trimming, fill, TRC (+C3D), exclusion recap

There is no CPU implementation of the search in this package: `solve_units` needs the built CUDA
library and a B200 and raises otherwise.  `[triangulation] undistort_points = true` runs on the device
too (points undistorted by the stage kernel, distorted re-projection in the search kernel);
`handle_LR_swap = true` (off in every shipped config) runs on the device as well, in its own kernel
(`p2s_lrswap.cu`), reproducing what the reference executes for that flag; the two flags combine.
"""
import glob
import logging
import os

import numpy as np

from . import calib as _calib
from . import skeletons as _skel
from . import staging as _stg


# ---------------------------------------------------------------------------------------------------
# settings
# ---------------------------------------------------------------------------------------------------
def read_settings(config_dict):
    """The config keys the stage reads (triangulation.py:678-696, :389-393)."""
    prj, tri = config_dict.get("project"), config_dict.get("triangulation")
    s = {
        "project_dir": prj.get("project_dir"),
        "multi_person": prj.get("multi_person"),
        "frame_range": prj.get("frame_range"),
        "frame_rate": prj.get("frame_rate"),
        "pose_model": config_dict.get("pose").get("pose_model"),
        "vid_img_extension": config_dict.get("pose").get("vid_img_extension"),
        "reproj_thr": tri.get("reproj_error_threshold_triangulation"),
        "lik_thr": tri.get("likelihood_threshold_triangulation"),
        "min_cams": tri.get("min_cameras_for_triangulation"),
        "interpolation": tri.get("interpolation"),
        "interp_gap": tri.get("interp_if_gap_smaller_than"),
        "max_distance_m": tri.get("max_distance_m", None),
        "remove_incomplete_frames": tri.get("remove_incomplete_frames", False),
        "sections_to_keep": tri.get("sections_to_keep"),
        "min_chunk_size": tri.get("min_chunk_size", 10),
        "fill_large_gaps_with": tri.get("fill_large_gaps_with"),
        "show_interp_indices": tri.get("show_interp_indices"),
        "undistort_points": tri.get("undistort_points"),
        "handle_LR_swap": tri.get("handle_LR_swap"),
        "make_c3d": tri.get("make_c3d"),
    }
    if s["min_chunk_size"] is None:
        s["min_chunk_size"] = 10
    return s


def swapped_keypoint_indices(keypoints_names):
    """`keypoints_idx_swapped` (triangulation.py:741-749); see `skeletons.swapped_indices`."""
    return _skel.swapped_indices(keypoints_names)


# ---------------------------------------------------------------------------------------------------
# staging
# ---------------------------------------------------------------------------------------------------
class StagedProject:
    """Everything `solve_units` / `write_outputs` need; arrays are unit-major, camera fastest."""
    __slots__ = ("settings", "calib_file", "P", "lens", "keypoints_ids", "keypoints_names", "cam_dirs", "input_dir",
                 "f_range", "n_cams", "n_persons", "x", "y", "lik", "inexact")


def _world():
    """(rank, world_size, local_rank) of an initialised torch.distributed job, else (0, 1, 0)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", dist.get_rank()))
    except ImportError:
        pass
    return 0, 1, 0


def stage_project(config_dict, rank=0, world=1):
    """Host staging.  With world > 1 (one process per GPU under torchrun) every rank discovers the whole
    trial but parses only the JSON of its own contiguous frame block (sharding.frame_block)."""
    s = read_settings(config_dict)
    session_dir = _calib.session_dir_of(s["project_dir"])
    calib_file = _calib.find_calibration_file(session_dir)
    P = _calib.compute_P(calib_file, undistort=bool(s["undistort_points"]))
    lens = _calib.camera_models(calib_file) if s["undistort_points"] else None
    ids, names = _skel.keypoints(s["pose_model"], config_dict)

    dirs = _stg.PoseDirs(s["project_dir"])
    cam_dirs = dirs.camera_dirs()
    index, files = None, None
    if os.environ.get("P2S_NATIVE_IO", "1") != "0":
        input_dir, index = dirs.index_for_triangulation(cam_dirs)      # listing, order and frame table in native code
        counts = index.counts()
    else:
        input_dir, files = dirs.files_for_triangulation(cam_dirs)
        counts = [len(j) for j in files]
    n_cams = len(cam_dirs)
    fr = s["frame_range"]
    f_range = [0, min(counts)] if fr in ("all", "auto", []) else fr
    if n_cams != len(P):
        raise Exception(f"Error: The number of cameras is not consistent: Found {len(P)} cameras in the calibration "
                        f"file, and {n_cams} cameras based on the number of pose folders.")
    if s["multi_person"]:
        if files is None:
            files = index.names()
        # every rank counts its stride of the files; a file that cannot be parsed raises (like the reference, :88-89) — but
        # only AFTER the all-reduce, so that the ranks whose files are fine are not left waiting in it: the failure is
        # part of what is reduced and every rank raises
        failure = None
        try:
            n_persons = _stg.count_persons(input_dir, cam_dirs, [f[rank::world] for f in files])
        except Exception as e:                                  # noqa: BLE001 — re-raised below, on every rank
            failure, n_persons = e, 0
        if world > 1:
            import torch
            import torch.distributed as dist
            dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))) if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor([n_persons, 0 if failure is None else 1], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            n_persons = int(t[0].item())
            if failure is None and int(t[1].item()):
                failure = RuntimeError("another rank could not parse one of its pose files (see its traceback)")
        if failure is not None:
            raise failure
    else:
        n_persons = 1

    from . import sharding
    n_frames = len(range(*f_range))
    b0, b1 = sharding.frame_block(n_frames, rank, world)
    my_range = [f_range[0] + b0, f_range[0] + b1]
    staged = _stg.stage_triangulation_indexed(index, my_range, ids, n_persons) if index is not None else None
    if staged is None:
        if files is None:
            files = index.names()
        staged = _stg.stage_triangulation(input_dir, cam_dirs, files, my_range, ids, n_persons)
    x, y, lik, inexact = staged
    if index is not None:
        index.close()
    st = StagedProject()
    st.settings, st.calib_file, st.P, st.lens = s, calib_file, np.asarray(P, dtype=np.float64), lens
    st.keypoints_ids, st.keypoints_names = ids, names
    st.cam_dirs, st.input_dir, st.f_range, st.n_cams, st.n_persons = cam_dirs, input_dir, list(f_range), n_cams, n_persons
    st.inexact = inexact
    st.x, st.y, st.lik = x, y, lik
    if st.inexact:
        logging.warning(f"{st.inexact} 2D values are not exactly representable in float32 and were rounded for the "
                        f"device staging layout (Pose2Sim's own pose stage writes float32 values).")
    return st


# ---------------------------------------------------------------------------------------------------
# device
# ---------------------------------------------------------------------------------------------------
def solve_units(st, engine=None, device=0):
    """ONE device call for all F*N*K units: likelihood gate (:817-821) + exclusion search (:363-604).
    Returns Q[F,N,K,3], err[F,N,K], nexcl[F,N,K] (int), mask[F,N,K] (uint32 id_excluded_cams bit sets)."""
    from . import ops
    eng = engine if engine is not None else ops.get_engine(device)
    F, N, K, C = st.x.shape
    U = F * N * K
    s = st.settings
    if U == 0:
        # nothing staged (a reversed or out-of-range frame_range): no device call; write_outputs raises the reference's
        # "No persons have been triangulated" (triangulation.py:955-956)
        return {"Q": np.empty((F, N, K, 3)), "err": np.empty((F, N, K)), "nexcl": np.empty((F, N, K), np.int64),
                "mask": np.empty((F, N, K), np.uint32), "stats": None}
    if s["handle_LR_swap"]:
        import torch
        dev = torch.device("cuda", eng.device)
        planes = [torch.from_numpy(np.ascontiguousarray(a.reshape(U, C), dtype=np.float32)).to(dev) for a in (st.x, st.y, st.lik)]
        obs = eng.stage_observations(*planes, s["lik_thr"], lens=st.lens)
        res = eng.triangulate_lr_swap(obs, swapped_keypoint_indices(st.keypoints_names), st.P, s["reproj_thr"], s["min_cams"],
                                      lens=st.lens)
        return {"Q": res["Q"].cpu().numpy().reshape(F, N, K, 3), "err": res["err"].cpu().numpy().reshape(F, N, K),
                "nexcl": res["nexcl"].cpu().numpy().reshape(F, N, K).astype(np.int64),
                "mask": res["mask"].cpu().numpy().view(np.uint32).reshape(F, N, K), "stats": None}
    out = eng.triangulate_host(st.x.reshape(U, C), st.y.reshape(U, C), st.lik.reshape(U, C), st.P,
                               s["lik_thr"], s["reproj_thr"], s["min_cams"], lens=st.lens)
    return {"Q": out["Q"].reshape(F, N, K, 3), "err": out["err"].reshape(F, N, K),
            "nexcl": out["nexcl"].reshape(F, N, K).astype(np.int64), "mask": out["mask"].reshape(F, N, K),
            "stats": out.get("stats")}


# ---------------------------------------------------------------------------------------------------
# multi-person re-identification across frames (host, sequential)
# ---------------------------------------------------------------------------------------------------
def match_to_previous(prev, cur, max_dist=None):
    """common.py:1037-1136 `sort_people_sports2d` without scores: Hungarian assignment on the mean
    per-keypoint distance, `max_dist` gate, unmatched detections appended as new persons.
    Returns (carried_prev, sorted_cur, ids) with ids[i] = index into `cur` or -1."""
    from scipy.optimize import linear_sum_assignment
    n_prev, n_cur = len(prev), len(cur)
    with np.errstate(invalid="ignore"):
        d = np.sqrt(np.nansum((cur[None, :, :, :] - prev[:, None, :, :]) ** 2, axis=3))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            D = np.nanmean(d, axis=2)
    D = np.nan_to_num(D, nan=1e10, posinf=1e10)
    rows, cols = linear_sum_assignment(D)
    pairs = [(r, c) for r, c in zip(rows, cols) if max_dist is None or D[r, c] <= max_dist]
    taken = {c for _, c in pairs}
    fresh = [i for i in range(n_cur) if i not in taken]
    n_total = n_prev + len(fresh)
    sorted_cur = np.full((n_total,) + cur.shape[1:], np.nan)
    ids = np.full(n_total, -1)
    for r, c in pairs:
        sorted_cur[r], ids[r] = cur[c], c
    for j, c in enumerate(fresh):
        sorted_cur[n_prev + j], ids[n_prev + j] = cur[c], c
    padded = prev if n_prev >= n_total else np.concatenate([prev, np.full((n_total - n_prev,) + prev.shape[1:], np.nan)])
    carried = np.where(np.isnan(sorted_cur) & ~np.isnan(padded), padded, sorted_cur)
    return carried, sorted_cur, ids


def reidentify(res, f_range, n_cams, max_distance_m):
    """triangulation.py:822-865 for all frames: reorder persons frame by frame.  Only the first
    N rows survive each frame (:868), persons not seen get NaN / n_cams / all-camera ids (:860-863)."""
    Q, err, nexcl, mask = res["Q"], res["err"], res["nexcl"], res["mask"]
    F, N, K, _ = Q.shape
    all_cams = np.uint32((1 << n_cams) - 1) if n_cams < 32 else np.uint32(0xFFFFFFFF)
    oQ, oe, on, om = np.empty_like(Q), np.empty_like(err), np.empty_like(nexcl), np.empty_like(mask)
    last = np.full((N, K, 3), np.nan)          # the previous frame's (sorted) persons
    memory = np.full((N, K, 3), np.nan)        # last known position of every tracked person
    for fi, f in enumerate(range(*f_range)):
        memory = np.where(np.isnan(last), memory, last)
        cur = Q[fi]
        if f != 0:
            if N == 0:                          # nobody in any file: sort_people_sports2d unpacks an empty array (:852)
                raise ValueError("not enough values to unpack (expected 3, got 2)")
            memory, cur_sorted, ids = match_to_previous(memory, cur, max_distance_m)
            for n in range(N):
                j = ids[n]
                if j >= 0:
                    oe[fi, n], on[fi, n], om[fi, n] = err[fi, j], nexcl[fi, j], mask[fi, j]
                else:
                    oe[fi, n], on[fi, n], om[fi, n] = np.nan, n_cams, all_cams
            cur = cur_sorted
        else:
            oe[fi], on[fi], om[fi] = err[fi], nexcl[fi], mask[fi]
        oQ[fi] = cur[:N]
        last = cur
    return {"Q": oQ, "err": oe, "nexcl": on, "mask": om}


# ---------------------------------------------------------------------------------------------------
# post-processing
# ---------------------------------------------------------------------------------------------------
def fill_small_gaps(col, index, max_gap, kind):
    """common.py:669-712 `interpolate_zeros_nans` on one coordinate column (values at labels `index`)."""
    good = ~(np.isnan(col) | (col == 0))
    if np.count_nonzero(good) <= 4:
        return col
    bad = np.flatnonzero(~good)
    out = col.copy()
    if bad.size:
        # = np.where(good, col, f(index)) with f = interp1d(index[good], col[good], kind, fill_value="extrapolate"),
        # evaluated only where it is used
        xg, yg = index[good], col[good]
        if kind == "linear" and np.all(np.diff(xg) > 0):
            # scipy's own statements for kind="linear" with extrapolation (interp1d._call_linear), without building an
            # interpolator object per column: the segment is found by searchsorted, end segments extrapolate
            xn = index[bad]
            hi = np.clip(np.searchsorted(xg, xn), 1, len(xg) - 1).astype(int)
            lo = hi - 1
            x_lo, x_hi, y_lo, y_hi = xg[lo], xg[hi], yg[lo], yg[hi]
            out[bad] = ((xn - x_lo) / (x_hi - x_lo)) * y_hi + ((x_hi - xn) / (x_hi - x_lo)) * y_lo
        else:
            from scipy import interpolate
            f = interpolate.interp1d(xg, yg, kind=kind, fill_value="extrapolate", bounds_error=False)
            out[bad] = f(index[bad])
        for seq in np.split(bad, np.flatnonzero(np.diff(index[bad]) > 1) + 1):
            if len(seq) > max_gap:
                out[seq] = np.nan
    return out


def valid_chunk(series, min_chunk_size=10, method="largest"):
    """triangulation.py:93-148: (start, end) positions of the runs of >= min_chunk_size non-NaN values
    selected by `method` ('largest' | 'all' | 'first' | 'last'); (0, 0) when there is none."""
    ok = ~np.isnan(np.asarray(series, dtype=np.float64))
    edges = np.flatnonzero(np.diff(np.concatenate([[0], ok.astype(np.int8), [0]])))
    runs = [(int(a), int(b)) for a, b in zip(edges[::2], edges[1::2]) if b - a >= min_chunk_size]
    if not runs:
        return 0, 0
    if method not in ("largest", "all", "first", "last"):
        method = "all"
    if method == "largest":
        return max(runs, key=lambda r: r[1] - r[0])       # first of the longest, like a stable sort
    if method == "all":
        return runs[0][0], runs[-1][1]
    return runs[0] if method == "first" else runs[-1]


def _ffill_bfill(a):
    """DataFrame.ffill().bfill() down the rows (NaN only); only the columns that hold a NaN are touched."""
    def ffill(v):
        idx = np.where(~np.isnan(v), np.arange(v.shape[0])[:, None], 0)
        np.maximum.accumulate(idx, axis=0, out=idx)
        return np.take_along_axis(v, idx, axis=0)
    cols = np.flatnonzero(np.isnan(a).any(axis=0))
    if cols.size == 0:
        return a
    sub = ffill(a[:, cols])
    out = a.copy()
    out[:, cols] = ffill(sub[::-1])[::-1]
    return out


def frame_rate_of(s):
    """triangulation.py:171-185."""
    rate = s["frame_rate"]
    if rate == "auto":
        try:
            import cv2
            vids = glob.glob(os.path.join(s["project_dir"], "videos", "*" + s["vid_img_extension"]))
            cap = cv2.VideoCapture(vids[0])
            cap.read()
            if cap.read()[0] is False:
                raise RuntimeError
            rate = round(cap.get(cv2.CAP_PROP_FPS))
        except Exception:
            logging.warning("Cannot read video. Frame rate will be set to 60 fps.")
            rate = 30                                      # sic: the reference logs 60 and uses 30
    return rate


def trc_header(s, n_rows, first_frame, last_frame, keypoints_names, id_person=-1):
    """triangulation.py:151-205: (path, rate, the five header lines) of the TRC of `n_rows` frames."""
    project_dir = s["project_dir"]
    base = os.path.basename(os.path.realpath(project_dir))
    seq = f"{base}_P{id_person}" if s["multi_person"] else base
    out_dir = os.path.join(project_dir, "pose-3d")
    rate = frame_rate_of(s)
    name = f"{seq}_{first_frame}-{last_frame}.trc"
    K = len(keypoints_names)
    header = ["PathFileType\t4\t(X/Y/Z)\t" + name,
              "DataRate\tCameraRate\tNumFrames\tNumMarkers\tUnits\tOrigDataRate\tOrigDataStartFrame\tOrigNumFrames",
              "\t".join(map(str, [rate, rate, n_rows, K, "m", rate, first_frame, n_rows])),
              "Frame#\tTime\t" + "\t\t\t".join(keypoints_names) + "\t\t\t",
              "\t\t" + "\t".join(f"X{i + 1}\tY{i + 1}\tZ{i + 1}" for i in range(K)) + "\t"]
    return os.path.realpath(os.path.join(out_dir, name)), out_dir, rate, "\n".join(header) + "\n"


def trc_row_arrays(Q, frames, rate, K):
    """common.py:596-612 `zup2yup` (X, Y, Z <- Y, Z, X) + the frame / time columns, as the native writer takes them."""
    yup = np.ascontiguousarray(Q.reshape(len(Q), K, 3)[:, :, [1, 2, 0]].reshape(len(Q), 3 * K), dtype=np.float64)
    fr = np.ascontiguousarray(frames, dtype=np.int64)
    t = np.ascontiguousarray(fr / rate, dtype=np.float64)
    return fr, t, yup


def write_trc(s, Q, frames, keypoints_names, id_person=-1):
    """triangulation.py:151-215 `make_trc`: Q [n, 3K] Z-up, `frames` their labels.  Returns the path.
    The body is written by the native writer (`p2s_write_trc_rows`, Python-repr number formatting =
    what `DataFrame.to_csv` emits) instead of pandas' per-value string conversion."""
    from . import _lib
    K = len(keypoints_names)
    path, out_dir, rate, header = trc_header(s, len(Q), frames[0], frames[-1], keypoints_names, id_person)
    fr, t, yup = trc_row_arrays(Q, frames, rate, K)
    if not os.path.exists(out_dir):
        os.mkdir(out_dir)
    with open(path, "w") as f:
        f.write(header)
    _lib.check(None, _lib.load().p2s_write_trc_rows(path.encode(), fr.ctypes.data, t.ctypes.data, yup.ctypes.data,
                                                    len(fr), 3 * K))
    return path


def write_c3d(trc_path):
    """common.py:615-665 `convert_to_c3d`, when the optional `c3d` package is importable."""
    import c3d
    with open(trc_path) as f:
        names = f.readlines()[3].strip().split("\t")[2::3]
    data = np.genfromtxt(trc_path, skip_header=5, delimiter="\t")[:, 1:]
    t = data[:, 0]
    rate = round((len(t) - 1) / (t[-1] - t[0]))
    w = c3d.Writer(point_rate=rate, analog_rate=0, point_scale=1.0, point_units="mm", gen_scale=-1.0)
    w.set_point_labels(names)
    w.set_screen_axis(X="+Z", Y="+Y")
    for row in data:
        pts = np.hstack([row[1:].reshape(-1, 3) * 1000, np.zeros((len(names), 2))])
        w.add_frames([(pts, np.array([]))])
    w.set_start_frame(0)
    w._set_last_frame(len(data) - 1)
    path = trc_path.replace(".trc", ".c3d")
    with open(path, "wb") as h:
        w.write(h)
    return path


def _gap_strings(positions, max_gap):
    seqs = np.split(positions, np.flatnonzero(np.diff(positions) > 1) + 1)
    short = [f"{q[0]}:{q[-1]}" for q in seqs if 0 < len(q) <= max_gap]
    long_ = [f"{q[0]}:{q[-1]}" for q in seqs if len(q) > max_gap]
    return short, long_


def write_outputs(st, res):
    """triangulation.py:877-959.  `res`: Q[F,N,K,3], err[F,N,K], nexcl[F,N,K], mask[F,N,K].
    Returns a dict with the TRC paths and the recap statistics (also logged)."""
    s = st.settings
    F, N, K, _ = res["Q"].shape
    frames = np.arange(*st.f_range)
    n_cams = st.n_cams
    min_chunk = s["min_chunk_size"]
    trimmed, trc_paths, cam_excluded, interp_frames, non_interp_frames = [], [], [], [], []
    err_kept, nexcl_kept = [], []
    for n in range(N):
        Qn = res["Q"][:, n].reshape(F, 3 * K).astype(np.float64).copy()
        en = res["err"][:, n].astype(np.float64)
        xn = res["nexcl"][:, n].astype(np.float64)
        mn = res["mask"][:, n]
        if s["interpolation"] != "none":
            try:
                filled = np.empty_like(Qn)                    # assigned as a whole: a failing column leaves Qn as it was
                for j in range(3 * K):
                    filled[:, j] = fill_small_gaps(Qn[:, j], frames, s["interp_gap"], s["interpolation"])
                Qn = filled
            except Exception:
                logging.warning(f"Interpolation was not possible for person {n}. This means that not enough points "
                                f"are available, which is often due to a bad calibration.")
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            e_mean = en.mean(axis=1) if s["remove_incomplete_frames"] else np.nanmean(en, axis=1)
        x_mean = xn.mean(axis=1)
        a, b = valid_chunk(e_mean, min_chunk, s["sections_to_keep"])
        trimmed.append([a, b])
        if b - a <= min_chunk:
            cam_excluded.append({})
            interp_frames.append([])
            non_interp_frames.append([])
            trc_paths.append("")
            err_kept.append(None)
            nexcl_kept.append(None)
            logging.info(f"\nPerson {n}: Less than {min_chunk} valid frames in a row. Deleting person.")
            continue
        Qn, en, xn, mn, fr = Qn[a:b], en[a:b], xn[a:b], mn[a:b], frames[a:b]
        err_kept.append(np.concatenate([en, e_mean[a:b, None]], axis=1))
        nexcl_kept.append(np.concatenate([xn, x_mean[a:b, None]], axis=1))
        xs = Qn[:, ::3].T
        kpt_i, pos = np.where((xs == 0) | ~np.isfinite(xs))
        bad_per_kpt = [pos[kpt_i == k] for k in range(K)]
        bad_per_kpt = [z[(a < z) & (b > z)] for z in bad_per_kpt]            # sic (:906): positions vs. bounds

        if s["fill_large_gaps_with"] == "last_value":
            Qn = _ffill_bfill(Qn)
            Qn[np.isnan(Qn) | (Qn == np.inf)] = 0
        elif s["fill_large_gaps_with"] == "zeros":
            Qn[np.isnan(Qn) | (Qn == np.inf)] = 0

        trc_paths.append(write_trc(s, Qn, fr, st.keypoints_names, id_person=n))
        if s["make_c3d"]:
            try:
                write_c3d(trc_paths[-1])
            except ImportError:
                logging.warning("make_c3d = true but the optional `c3d` package is not installed: no .c3d written.")

        opportunities = len(Qn) * K
        counts = {c: int(np.count_nonzero((mn >> np.uint32(c)) & np.uint32(1))) for c in range(n_cams)}
        cam_excluded.append({c: v / opportunities for c, v in counts.items()})
        if s["show_interp_indices"]:
            pairs = [_gap_strings(bad_per_kpt[k], s["interp_gap"]) for k in range(K)]
            interp_frames.append([p[0] for p in pairs])
            non_interp_frames.append([p[1] for p in pairs])
        else:
            interp_frames.append(None)
            non_interp_frames.append([])

    if np.all(np.diff(np.array(trimmed)) == 0):
        raise Exception("No persons have been triangulated. Please check your calibration and your synchronization, "
                        "or the triangulation parameters in Config.toml.")
    recap = {"trc_paths": trc_paths, "f_range_trimmed": trimmed, "cam_excluded_count": cam_excluded,
             "error": err_kept, "nb_cams_excluded": nexcl_kept, "interp_frames": interp_frames,
             "non_interp_frames": non_interp_frames}
    log_recap(st, recap)
    return recap


def log_recap(st, r):
    """triangulation.py:255-360 `recap_triangulate` (log lines only)."""
    s = st.settings
    names = np.array(_calib.camera_names(st.calib_file))
    names = names[list(r["cam_excluded_count"][0].keys())]
    fm, Dm = _calib.first_camera_scale(st.calib_file)
    kind, gap = s["interpolation"], s["interp_gap"]
    logging.info("")
    N = len(r["error"])
    import warnings
    for n in range(N):
        a, b = r["f_range_trimmed"][n]
        if b - a <= s["min_chunk_size"]:
            continue
        if N > 1:
            logging.info(f"\n\nPARTICIPANT {n}\n")
        err, nex = r["error"][n], r["nb_cams_excluded"][n]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            # np.nanmean(err[:, k]) of every column in one call each: rows of the transposed copy are contiguous, so every
            # mean is the same pairwise sum as the per-column statement's (tests/test_dropin_host.py checks the bits)
            err_means = np.nanmean(np.ascontiguousarray(err.T), axis=1)
            nex_means = np.nanmean(np.ascontiguousarray(nex.T), axis=1)
            for k, name in enumerate(st.keypoints_names):
                e_px = np.around(err_means[k], decimals=1)
                e_m = np.around(e_px * Dm / fm, decimals=3)
                excl = np.around(nex_means[k], decimals=2)
                logging.info(f"Mean reprojection error for {name} is {e_px} px (~ {e_m} m), reached with {excl} excluded cameras. ")
                if s["show_interp_indices"]:
                    if kind != "none":
                        done, skipped = r["interp_frames"][n][k], r["non_interp_frames"][n][k]
                        if not done and not skipped:
                            logging.info("  No frames needed to be interpolated.")
                        if done:
                            logging.info("  Frames " + ", ".join(d.replace(":", " to ") for d in done) + " were interpolated.")
                        if skipped:
                            logging.info("  Frames " + ", ".join(d.replace(":", " to ") for d in skipped) + " were not interpolated.")
                    else:
                        logging.info("  No frames were interpolated because 'interpolation_kind' was set to none. ")
            e_px = np.around(err_means[-1], decimals=1)
            e_mm = np.around(e_px * Dm / fm * 1000, decimals=1)
            excl = np.around(nex_means[-1], decimals=2)
        logging.info(f"\n--> Mean reprojection error for all points on frames {a} to {b} is {e_px} px, which roughly corresponds to {e_mm} mm. ")
        logging.info(f"Cameras were excluded if likelihood was below {s['lik_thr']} and if the reprojection error was above {s['reproj_thr']} px.")
        if kind != "none":
            fill = {"last_value": "the last valid value", "zeros": "zeros"}.get(s["fill_large_gaps_with"], "NaNs")
            logging.info(f"Gaps were interpolated with {kind} method if smaller than {gap} frames. Larger gaps were filled with {fill}.")
        logging.info(f"In average, {excl} cameras had to be excluded to reach these thresholds.")
        if len(range(a, b)) < len(range(*st.f_range)):
            logging.warning(f"\nSome frames could not be correctly triangulated: trial trimmed between frames {[a, b]}.\n"
                            "You might need to tweak the triangulation parameters in Config.toml (for example, try "
                            'increasing "reproj_error_threshold_triangulation").')
        named = dict(zip(names, r["cam_excluded_count"][n].values()))
        named = dict(sorted(named.items(), key=lambda kv: kv[1])[::-1])
        r["cam_excluded_count"][n] = named
        parts = []
        for i, (cam, v) in enumerate(named.items()):
            pct = int(np.round(v * 100))
            if i == 0:
                parts.append(f"Camera {cam} was excluded {pct}% of the time, ")
            elif i == len(named) - 1:
                parts.append(f"and Camera {cam}: {pct}%.")
            else:
                parts.append(f"Camera {cam}: {pct}%, ")
        logging.info("".join(parts))
        logging.info(f"3D coordinates are stored at {r['trc_paths'][n]}.")
    logging.info("\n\n")
    if s["make_c3d"]:
        logging.info("All trc files have been converted to c3d.")
    logging.info(f"Limb swapping was {'handled' if s['handle_LR_swap'] else 'not handled'}.")
    logging.info(f"Lens distortions were {'taken into account' if s['undistort_points'] else 'not taken into account'}.")


# ---------------------------------------------------------------------------------------------------
# frame-block sharded post-processing (N > 1 ranks, results stay rank-local)
# ---------------------------------------------------------------------------------------------------
def rank_local_supported(st):
    """The sharded writer covers what shards by frame block with a small halo: one person (the multi-person re-ID,
    triangulation.py:847-865, is sequential over frames) and the interpolation kinds whose value depends on the two
    neighbouring good samples only ('linear', the shipped default, and 'none'); a spline over the whole column does not."""
    s = st.settings
    return (not s["multi_person"]) and st.n_persons == 1 and s["interpolation"] in ("linear", "none")


def _halo_points(gathered, rank, j, side):
    """Up to two nearest good (label, value) samples of column j held by the ranks before (`side` = -1, nearest last)
    or after (`side` = +1, nearest first) this rank."""
    pts = []
    ranks = range(rank - 1, -1, -1) if side < 0 else range(rank + 1, len(gathered))
    for r in ranks:
        lab, val = gathered[r]["tail" if side < 0 else "head"]
        seq = range(1, -1, -1) if side < 0 else range(2)
        for i in seq:
            if not np.isnan(lab[j, i]):
                pts.append((lab[j, i], val[j, i]))
                if len(pts) == 2:
                    return pts[::-1] if side < 0 else pts
    return pts[::-1] if side < 0 else pts


def write_outputs_sharded(st, res, rank, world):
    """`write_outputs` (triangulation.py:877-959) with the results left where they were computed: every rank
    post-processes its own frame block and writes its own byte range of the TRC; what crosses ranks is per COLUMN a
    handful of boundary samples (interpolation and fill halos), per FRAME one mean error (the trimming decision), and
    per rank its text length and its recap sums — a few hundred bytes per keypoint instead of 37 bytes per unit.
    The file is byte-identical to the single-process one (tests/test_sharding_gloo.py)."""
    import warnings
    import torch.distributed as dist
    from scipy import interpolate
    from . import _lib, sharding

    def gather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    s = st.settings
    frames_all = np.arange(*st.f_range)
    F_tot = len(frames_all)
    b0, b1 = sharding.frame_block(F_tot, rank, world)
    Fl, K = b1 - b0, res["Q"].shape[2]
    frames = frames_all[b0:b1]
    n_cams, min_chunk = st.n_cams, s["min_chunk_size"]
    Qn = res["Q"][:, 0].reshape(Fl, 3 * K).astype(np.float64).copy()
    en = res["err"][:, 0].astype(np.float64)
    xn = res["nexcl"][:, 0].astype(np.float64)
    mn = res["mask"][:, 0]

    # ---- interpolation of small gaps (common.py:669-712), halo = two good samples per column and side ------------
    if s["interpolation"] == "linear":
        good = ~(np.isnan(Qn) | (Qn == 0))
        head_l, head_v = np.full((3 * K, 2), np.nan), np.full((3 * K, 2), np.nan)
        tail_l, tail_v = np.full((3 * K, 2), np.nan), np.full((3 * K, 2), np.nan)
        for j in range(3 * K):
            idx = np.flatnonzero(good[:, j])
            h, t = idx[:2], idx[-2:]
            head_l[j, :len(h)], head_v[j, :len(h)] = frames[h], Qn[h, j]
            tail_l[j, 2 - len(t):], tail_v[j, 2 - len(t):] = frames[t], Qn[t, j]
        G = gather({"count": good.sum(axis=0), "head": (head_l, head_v), "tail": (tail_l, tail_v)})
        total = np.sum([g["count"] for g in G], axis=0)
        for j in range(3 * K):
            if total[j] <= 4 or Fl == 0:                              # :683: columns with <= 4 good samples stay as they are
                continue
            bad = np.flatnonzero(~good[:, j])
            if bad.size == 0:
                continue
            before, after = _halo_points(G, rank, j, -1), _halo_points(G, rank, j, +1)
            xs = np.concatenate([[p[0] for p in before], frames[good[:, j]], [p[0] for p in after]]).astype(np.float64)
            ys = np.concatenate([[p[1] for p in before], Qn[good[:, j], j], [p[1] for p in after]])
            f = interpolate.interp1d(xs, ys, kind="linear", fill_value="extrapolate", bounds_error=False)
            col = Qn[:, j].copy()
            col[bad] = f(frames[bad])
            first_label = (before[-1][0] + 1) if before else frames_all[0]      # where a gap touching my first row began
            last_label = (after[0][0] - 1) if after else frames_all[-1]        # ... and where one touching my last row ends
            for seq in np.split(bad, np.flatnonzero(np.diff(bad) > 1) + 1):
                g0 = first_label if seq[0] == 0 else frames[seq[0]]
                g1 = last_label if seq[-1] == Fl - 1 else frames[seq[-1]]
                if g1 - g0 + 1 > s["interp_gap"]:
                    col[seq] = np.nan
            Qn[:, j] = col

    # ---- trimming decision on the whole trial: one mean error per frame ------------------------------------------
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        e_mean = (en.mean(axis=1) if s["remove_incomplete_frames"] else np.nanmean(en, axis=1)) if Fl else np.zeros(0)
    x_mean = xn.mean(axis=1) if Fl else np.zeros(0)
    e_all = np.concatenate(gather(e_mean))
    a, b = valid_chunk(e_all, min_chunk, s["sections_to_keep"])
    if b - a <= min_chunk:
        if rank == 0:
            logging.info(f"\nPerson 0: Less than {min_chunk} valid frames in a row. Deleting person.")
        if b - a == 0:
            raise Exception("No persons have been triangulated. Please check your calibration and your synchronization, "
                            "or the triangulation parameters in Config.toml.")
        if rank == 0:
            log_recap(st, {"trc_paths": [""], "f_range_trimmed": [[a, b]], "cam_excluded_count": [{}], "error": [None],
                           "nb_cams_excluded": [None], "interp_frames": [[]], "non_interp_frames": [[]]})
        return None
    lo, hi = min(max(a, b0), b1) - b0, max(min(b, b1), b0) - b0           # my rows of the kept section
    hi = max(hi, lo)
    Qt, et, xt, mt, fr = Qn[lo:hi], en[lo:hi], xn[lo:hi], mn[lo:hi], frames[lo:hi]
    first_pos = lo + b0 - a                                              # position of my first kept row in the section

    # ---- recap inputs (per rank sums; rank 0 adds them up) --------------------------------------------------------
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        err_cols = np.concatenate([et, e_mean[lo:hi, None]], axis=1)
        nex_cols = np.concatenate([xt, x_mean[lo:hi, None]], axis=1)
        sums = {"err_sum": np.nansum(err_cols, axis=0), "err_n": np.count_nonzero(~np.isnan(err_cols), axis=0),
                "nex_sum": np.nansum(nex_cols, axis=0), "nex_n": np.count_nonzero(~np.isnan(nex_cols), axis=0),
                "cams": np.array([np.count_nonzero((mt >> np.uint32(c)) & np.uint32(1)) for c in range(n_cams)], np.int64)}
    xs = Qt[:, ::3].T
    kpt_i, pos = np.where((xs == 0) | ~np.isfinite(xs))
    pos = pos + first_pos
    keep = (a < pos) & (b > pos)                                         # sic (:906): positions against frame bounds
    sums["bad"] = [pos[keep & (kpt_i == k)] for k in range(K)]

    # ---- large gaps (:910-916): forward / backward fill needs the nearest valid sample of the neighbouring ranks ----
    if s["fill_large_gaps_with"] == "last_value":
        ok = ~np.isnan(Qt)
        has = ok.any(axis=0)
        first_v = np.where(has, Qt[np.argmax(ok, axis=0), np.arange(3 * K)], np.nan) if hi > lo else np.full(3 * K, np.nan)
        last_v = np.where(has, Qt[len(Qt) - 1 - np.argmax(ok[::-1], axis=0), np.arange(3 * K)], np.nan) if hi > lo else np.full(3 * K, np.nan)
        H = gather((first_v, last_v))
        if hi > lo:
            prev_last = np.full(3 * K, np.nan)
            for r in range(rank):                                        # nearest previous rank wins
                prev_last = np.where(np.isnan(H[r][1]), prev_last, H[r][1])
            next_first = np.full(3 * K, np.nan)
            for r in range(world - 1, rank, -1):                         # nearest following rank wins
                next_first = np.where(np.isnan(H[r][0]), next_first, H[r][0])
            ext = np.concatenate([prev_last[None], Qt, next_first[None]], axis=0)
            Qt = _ffill_bfill(ext)[1:-1]
            Qt[np.isnan(Qt) | (Qt == np.inf)] = 0
    elif s["fill_large_gaps_with"] == "zeros":
        Qt = Qt.copy()
        Qt[np.isnan(Qt) | (Qt == np.inf)] = 0

    # ---- TRC: rank 0 writes the header, every rank its own rows at its own offset ------------------------------------
    path, out_dir, rate, header = trc_header(s, b - a, frames_all[a], frames_all[b - 1], st.keypoints_names)
    fr_a, t_a, yup = trc_row_arrays(Qt, fr, rate, K)
    cap = max(1, len(fr_a) * (22 + 25 * (3 * K + 1)))
    buf = np.empty(cap, np.uint8)
    import ctypes
    n_bytes = ctypes.c_size_t(0)
    _lib.check(None, _lib.load().p2s_format_trc_rows(fr_a.ctypes.data, t_a.ctypes.data, yup.ctypes.data, len(fr_a), 3 * K,
                                                     buf.ctypes.data, cap, ctypes.byref(n_bytes)))
    sums["bytes"] = int(n_bytes.value)
    S = gather(sums)
    if rank == 0:
        if not os.path.exists(out_dir):
            os.mkdir(out_dir)
        with open(path, "wb") as f:
            f.write(header.encode())
            f.truncate(len(header.encode()) + sum(g["bytes"] for g in S))
    dist.barrier()
    offset = len(header.encode()) + sum(g["bytes"] for g in S[:rank])
    fd = os.open(path, os.O_WRONLY)
    try:
        os.pwrite(fd, memoryview(buf)[:n_bytes.value], offset)
    finally:
        os.close(fd)
    dist.barrier()
    if rank != 0:
        return None

    # ---- recap on rank 0 (triangulation.py:255-360) --------------------------------------------------------------------
    if s["make_c3d"]:
        try:
            write_c3d(path)
        except ImportError:
            logging.warning("make_c3d = true but the optional `c3d` package is not installed: no .c3d written.")
    with np.errstate(invalid="ignore", divide="ignore"):
        err_mean = np.sum([g["err_sum"] for g in S], axis=0) / np.sum([g["err_n"] for g in S], axis=0)
        nex_mean = np.sum([g["nex_sum"] for g in S], axis=0) / np.sum([g["nex_n"] for g in S], axis=0)
    cams = np.sum([g["cams"] for g in S], axis=0)
    opportunities = (b - a) * K
    bad_per_kpt = [np.sort(np.concatenate([g["bad"][k] for g in S])) for k in range(K)]
    if s["show_interp_indices"]:
        pairs = [_gap_strings(bad_per_kpt[k], s["interp_gap"]) for k in range(K)]
        interp_frames, non_interp = [[p[0] for p in pairs]], [[p[1] for p in pairs]]
    else:
        interp_frames, non_interp = [None], [[]]
    recap = {"trc_paths": [path], "f_range_trimmed": [[a, b]],
             "cam_excluded_count": [{c: int(cams[c]) / opportunities for c in range(n_cams)}],
             "error": [err_mean[None, :]], "nb_cams_excluded": [nex_mean[None, :]],     # one row = the column means
             "interp_frames": interp_frames, "non_interp_frames": non_interp}
    log_recap(st, recap)
    return recap


# ---------------------------------------------------------------------------------------------------
def gather_units(res, n_frames, rank, world):
    """The ONE collective of the path: every rank's packed per-unit outputs (37 B/unit) to rank 0
    (NCCL over NVLink on a GPU box, gloo in the CPU tests).  Returns the full result on rank 0, None
    elsewhere."""
    import torch
    import torch.distributed as dist
    from . import sharding
    F, N, K, _ = res["Q"].shape
    per_frame = N * K
    units = [(b - a) * per_frame for a, b in sharding.frame_blocks(n_frames, world)]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.empty(sharding.PACK_BYTES * units[rank], dtype=torch.uint8, device=dev)
    v = sharding.packed_views(buf, units[rank])
    v["Q"].copy_(torch.from_numpy(np.ascontiguousarray(res["Q"].reshape(-1, 3))))
    v["err"].copy_(torch.from_numpy(np.ascontiguousarray(res["err"].reshape(-1))))
    v["mask"].copy_(torch.from_numpy(np.ascontiguousarray(res["mask"].reshape(-1)).view(np.int32)))
    v["nexcl"].copy_(torch.from_numpy(np.ascontiguousarray(res["nexcl"].reshape(-1)).astype(np.uint8)))
    bufs, _ = sharding.gather_packed(buf, units, dst=0)
    if rank != 0:
        return None
    out = sharding.unpack_concat(bufs, units)
    return {"Q": out["Q"].reshape(n_frames, N, K, 3), "err": out["err"].reshape(n_frames, N, K),
            "nexcl": out["nexcl"].reshape(n_frames, N, K).astype(np.int64), "mask": out["mask"].reshape(n_frames, N, K)}


def raise_together(failure, world, where):
    """One small object collective: every rank learns whether any rank failed in its rank-local work.  The failing rank
    raises its own exception, the others a RuntimeError naming it; with no failure it is a no-op."""
    import torch.distributed as dist
    reports = [None] * world
    dist.all_gather_object(reports, None if failure is None else f"{type(failure).__name__}: {failure}")
    if failure is not None:
        raise failure
    failed = [(r, msg) for r, msg in enumerate(reports) if msg is not None]
    if failed:
        raise RuntimeError(f"rank {failed[0][0]} failed in {where}: {failed[0][1]}")


def triangulate_all(config_dict):
    """Same contract as Pose2Sim/triangulation.py:656: reads the calibration TOML and the per-camera
    OpenPose JSON of the trial, writes `pose-3d/*.trc`, logs the recap.  Returns None.

    Under torchrun (torch.distributed initialised, one process per GPU) the frames are sharded in
    contiguous blocks over the ranks and each rank stages and solves its block on its own GPU.  For a single
    person with linear (or no) interpolation the results then STAY rank-local: every rank post-processes its block
    and writes its byte range of the TRC (`write_outputs_sharded`).  Otherwise (multi-person re-ID is sequential over
    frames, spline interpolation spans the whole column) the packed outputs are gathered once on rank 0, which alone
    writes the files (`P2S_GATHER=rank0` forces that path)."""
    rank, world, local = _world()
    if world > 1:
        import torch.distributed as dist
        if dist.get_backend() == "nccl":
            import torch
            torch.cuda.set_device(local)                    # the collectives below put their tensors on the current device
        # rank-local work (staging of the own frame block, the device call) first; then the ranks tell each other whether
        # it worked, so that a rank-local failure raises everywhere instead of leaving the others in the next collective
        failure, st, res = None, None, None
        try:
            st = stage_project(config_dict, rank, world)
            res = solve_units(st, device=local)
        except Exception as e:                              # noqa: BLE001 — re-raised by raise_together
            failure = e
        raise_together(failure, world, "triangulate_all")
        if rank_local_supported(st) and os.environ.get("P2S_GATHER", "local") != "rank0":
            write_outputs_sharded(st, res, rank, world)     # results stay rank-local: no gather of the 37 B/unit outputs
            return
        res = gather_units(res, len(range(*st.f_range)), rank, world)
        if rank != 0:
            return
        if st.settings["multi_person"]:
            res = reidentify(res, st.f_range, st.n_cams, st.settings["max_distance_m"])
        write_outputs(st, res)
        return
    st = stage_project(config_dict)
    res = solve_units(st)
    if st.settings["multi_person"]:
        res = reidentify(res, st.f_range, st.n_cams, st.settings["max_distance_m"])
    write_outputs(st, res)
