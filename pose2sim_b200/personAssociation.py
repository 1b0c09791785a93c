"""Drop-in `associate_all(config_dict)` — same entry point, inputs and `pose-associated/` output as
Pose2Sim/personAssociation.py:642-809 (single-person mode), with the per-frame search
(`persons_combinations` :67, `best_persons_and_cameras_combination` :154, `triangulate_comb` :102)
replaced by ONE batched call into the sm_100a library (`p2s_associate_host`).

    stage_project()    host    :663-772   config, calibration, tracked keypoint, JSON -> obs[F,C,P,4], count[F,C]
    solve_frames()     DEVICE  :771-774   ordered person-combination x camera-subset search, all frames
    write_outputs()    host    :776-808   rewrite JSON with the chosen person per camera, recap

Multi-person mode (Plücker-ray affinity + SVT matching, :277-549) is a different
kernel family — small dense linear algebra per frame — and also runs on the device (`multi_person.py`,
`p2s_associate_multi_host`, one CTA per frame).  Neither search has a CPU implementation in this package.
"""
import json
import logging
import os

import numpy as np

from . import _lib
from . import calib as _calib
from . import skeletons as _skel
from . import staging as _stg


def read_settings(config_dict):
    """personAssociation.py:663-675, :177-180."""
    prj, pa, tri = config_dict.get("project"), config_dict.get("personAssociation"), config_dict.get("triangulation")
    sp = pa.get("single_person")
    mp = pa.get("multi_person") or {}
    return {
        "project_dir": prj.get("project_dir"),
        "multi_person": prj.get("multi_person"),
        "frame_range": prj.get("frame_range"),
        "pose_model": config_dict.get("pose").get("pose_model"),
        "tracked_keypoint": sp.get("tracked_keypoint"),
        "reproj_thr": sp.get("reproj_error_threshold_association"),
        "lik_thr": pa.get("likelihood_threshold_association"),            # sic: read from the parent table (:180)
        "lik_thr_recap": sp.get("likelihood_threshold_association", 0.3),  # the recap reads the other one (:609)
        "min_cams": tri.get("min_cameras_for_triangulation"),
        "undistort_points": tri.get("undistort_points"),
        "reconstruction_error_threshold": mp.get("reconstruction_error_threshold"),
        "min_affinity": mp.get("min_affinity"),
    }


class StagedAssociation:
    __slots__ = ("settings", "calib_file", "P", "cam_dirs", "dirs", "table", "f_range", "n_cams",
                 "tracked_keypoint_id", "obs", "count", "parsed", "inexact", "workers", "mp_staged", "native", "world",
                 "src_c_paths", "index")


# ---- file I/O on a process pool -------------------------------------------------------------------------------------
# Reading 8 cameras x N frames of OpenPose JSON with Python's `json` and writing them back is what the stage costs once
# the search runs on the GPU (1 s per 1000 frames against ~1 ms of device time).  Frames are independent, so large
# trials are cut into contiguous frame blocks that worker processes (fork) parse into the staging arrays and, after the
# device call, re-read and rewrite.  The workers never touch CUDA.  P2S_HOST_WORKERS=<n> forces the count (1 = off).
def host_workers(n_files, world=1):
    forced = os.environ.get("P2S_HOST_WORKERS")
    if forced is not None:
        n = int(forced)
    elif world > 1:
        n = 1                                       # a torchrun job already spreads the frames over its ranks
    else:
        n = min(len(os.sched_getaffinity(0)), n_files // 512)
    return n if n >= 2 else 0


def _blocks(n_rows, n_workers):
    n_blocks = min(n_rows, 2 * n_workers)
    edges = [n_rows * i // n_blocks for i in range(n_blocks + 1)]
    return [(a, b) for a, b in zip(edges, edges[1:]) if b > a]


def _pool_map(fn, jobs, n_workers):
    import multiprocessing
    with multiprocessing.get_context("fork").Pool(n_workers) as pool:
        return pool.map(fn, jobs, chunksize=1)


def _stage_block(job):
    pose_dir, cam_dirs, rows, kid, multi = job
    if multi:
        from . import multi_person as mp
        parsed = [[_stg.load_json(os.path.join(pose_dir, cam_dirs[c], names[c])) for c in range(len(cam_dirs))] for names in rows]
        obs, count, inexact, n_values = mp.stage_detections(parsed, check_total=False, want_length=True)
        return obs, count, inexact, n_values
    obs, count, _ = _stg.stage_association(pose_dir, cam_dirs, rows, kid, _lib.P2S_MAX_PERSONS)
    return obs, count


def _rewrite_block(job):
    pose_dir, tracked_dir, cam_dirs, rows, proposals = job
    for names, prop in zip(rows, proposals):
        source = [_stg.load_json(os.path.join(pose_dir, cam_dirs[c], names[c])) for c in range(len(cam_dirs))]
        tracked = [os.path.join(tracked_dir, cam_dirs[c], names[c]) for c in range(len(cam_dirs))]
        rewrite_frame(tracked, source, prop)
    return len(rows)


def _rewrite_native(st, proposals):
    """`rewrite_json_files` (:552-580) for all frames through the native writer; the few files it leaves alone (status 2:
    strings with escapes, duplicate keys, no `people` list) are rewritten frame by frame by the Python statements."""
    C = st.n_cams
    src = getattr(st, "src_c_paths", None)                 # the C array the native reader was given (same table)
    if src is None:
        src = _stg.frame_paths(st.dirs.pose_dir, st.cam_dirs, st.table)
    dst = _stg.frame_paths(st.dirs.tracked_dir, st.cam_dirs, st.table)
    # 'none' entries: the reference opens <tracked>/<cam>/none for writing, fails on the source and removes it again
    status = _stg.rewrite_people_files(src, dst, proposals, _io_threads(st.world))
    for fi in np.flatnonzero((status == 2).any(axis=1)):
        names = st.table[fi]
        source = [_stg.load_json(os.path.join(st.dirs.pose_dir, st.cam_dirs[c], names[c])) for c in range(C)]
        tracked = [os.path.join(st.dirs.tracked_dir, st.cam_dirs[c], names[c]) for c in range(C)]
        rewrite_frame(tracked, source, proposals[fi])


def _rewrite_parallel(st, proposals):
    if st.native or os.environ.get("P2S_NATIVE_IO", "1") != "0":
        return _rewrite_native(st, proposals)
    jobs = [(st.dirs.pose_dir, st.dirs.tracked_dir, st.cam_dirs, st.table[a:b], proposals[a:b]) for a, b in _blocks(len(st.table), st.workers)]
    _pool_map(_rewrite_block, jobs, st.workers)


def stage_project(config_dict, rank=0, world=1):
    """Host staging.  With world > 1 (one process per GPU under torchrun) every rank discovers the whole trial
    but stages only its own contiguous frame block (sharding.frame_block): frames are independent
    (personAssociation.py:758-804 keeps no cross-frame state) and every rank writes its own output files."""
    s = read_settings(config_dict)
    if s["undistort_points"]:
        raise NotImplementedError("[triangulation] undistort_points = true is not available in the B200 path")
    session_dir = _calib.session_dir_of(s["project_dir"])
    calib_file = _calib.find_calibration_file(session_dir)
    P = _calib.compute_P(calib_file, undistort=False)
    _skel.model_nodes(s["pose_model"], config_dict)          # NameError when the model is unknown (:689-707)

    dirs = _stg.PoseDirs(s["project_dir"])
    cam_dirs = dirs.camera_dirs_for_association()
    # Listing, file order and the frame table in native code (staging.NativeIndex) when the files are listed where they are
    # read — pose/ — i.e. when there is no pose-sync/ (the reference lists pose-sync/ but reads pose/, :724-731 / :762-766:
    # that case keeps the Python statements)
    index, files = None, None
    if os.environ.get("P2S_NATIVE_IO", "1") != "0" and not os.path.isdir(dirs.sync_dir):
        try:
            index = _stg.NativeIndex(dirs.pose_dir, cam_dirs)
            counts = index.counts()
        except OSError:
            index = None
    if index is None:
        _, files = dirs.files_for_association(cam_dirs)
        counts = [len(j) for j in files]
    mkdir_failure = None
    if rank == 0:
        try:
            if not os.path.exists(dirs.tracked_dir):
                os.mkdir(dirs.tracked_dir)
            for d in cam_dirs:
                try:
                    os.mkdir(os.path.join(dirs.tracked_dir, d))
                except OSError:
                    break                                     # the reference stops at the first failure (:735-736)
        except OSError as e:
            mkdir_failure = e                                 # raised after the barrier: nobody is left waiting in it
    if world > 1:
        import torch.distributed as dist
        dist.barrier()                                        # the output directories exist before anyone writes
    if mkdir_failure is not None:
        raise mkdir_failure
    fr = s["frame_range"]
    f_range = [0, max(counts)] if fr in ("all", "auto", []) else fr
    n_cams = len(cam_dirs)
    if n_cams != len(P):
        raise Exception(f"Error: The number of cameras is not consistent: Found {len(P)} cameras in the calibration "
                        f"file, and {n_cams} cameras based on the number of pose folders.")
    kid = 0
    if s["multi_person"]:
        logging.info("\nMulti-person analysis selected.")
    else:
        logging.info("\nSingle-person analysis selected.")
        kid, fallback = _skel.tracked_keypoint_id(s["pose_model"], s["tracked_keypoint"], config_dict)
        if fallback is not None:
            logging.warning(f"{s['tracked_keypoint']} not found in {s['pose_model']}, consider editing "
                            f"tracked_keypoint in Config.toml. Tracking {fallback} instead.")

    st = StagedAssociation()
    st.world, st.src_c_paths = world, None
    st.settings, st.calib_file, st.P = s, calib_file, np.asarray(P, dtype=np.float64)
    st.cam_dirs, st.dirs, st.f_range, st.n_cams, st.tracked_keypoint_id = cam_dirs, dirs, list(f_range), n_cams, kid
    st.index = None
    if index is not None and not all(type(v) is int for v in f_range):
        files = index.names()                                 # range(*f_range) raises on these like the reference's does
        index.close()
        index = None
    if index is not None:
        # the table of this rank's frame block only (frames are tabulated independently of each other)
        fr_all = range(*f_range)
        a, b = 0, len(fr_all)
        if world > 1:
            from . import sharding
            a, b = sharding.frame_block(len(fr_all), rank, world)
        step_ok = fr_all.step == 1
        if step_ok and index.build_table([fr_all.start + a, fr_all.start + b] if len(fr_all) else [0, 0]):
            st.table, st.index = index.table_names(), index
        else:                                                 # a name without a number (the statements below raise the
            files = index.names()                             # reference's IndexError) or a stepped frame range
            index.close()
            index = None
    if index is None:
        st.table = _stg.frame_file_table(files, f_range)
        if world > 1:
            from . import sharding
            a, b = sharding.frame_block(len(st.table), rank, world)
            st.table = st.table[a:b]
    st.workers, st.mp_staged, st.native = host_workers(len(st.table) * n_cams, world), None, False
    # the reference always READS from pose/ (`os.path.exist` typo, :762-766)
    if os.environ.get("P2S_NATIVE_IO", "1") != "0" and _stage_native(st, world):
        return st
    if st.workers:
        jobs = [(dirs.pose_dir, cam_dirs, st.table[a:b], kid, bool(s["multi_person"])) for a, b in _blocks(len(st.table), st.workers)]
        parts = _pool_map(_stage_block, jobs, st.workers)
        st.parsed = None                                      # the rewrite re-reads its sources in the workers
        if s["multi_person"]:
            from . import multi_person as mp
            st.mp_staged = mp.merge_staged(parts)
            st.obs, st.count, st.inexact = None, None, 0
            return st
        obs = np.concatenate([p[0] for p in parts]) if parts else np.zeros((0, n_cams, _lib.P2S_MAX_PERSONS, 4))
        st.count = np.concatenate([p[1] for p in parts]) if parts else np.zeros((0, n_cams), np.int32)
    elif s["multi_person"]:
        st.parsed = [[_stg.load_json(os.path.join(dirs.pose_dir, cam_dirs[c], names[c])) for c in range(n_cams)]
                     for names in st.table]
        st.obs, st.count, st.inexact = None, None, 0
        return st
    else:
        obs, st.count, st.parsed = _stg.stage_association(dirs.pose_dir, cam_dirs, st.table, kid, _lib.P2S_MAX_PERSONS)
    st.inexact = _stg.float32_inexact(obs[..., :3])
    if st.inexact:
        logging.warning(f"{st.inexact} 2D values are not exactly representable in float32 and were rounded for the "
                        f"device staging layout.")
    st.obs = obs.astype(np.float32)
    return st


def _io_threads(world):
    return max(1, len(os.sched_getaffinity(0)) // max(1, world))


def _stage_native(st, world):
    """Staging through the native reader (`p2s_read_people_files`: every file parsed once, all host cores).  Returns
    False — nothing staged — when any file is irregular (status 2: the Python path below mirrors the reference's
    exception handling statement by statement) or shows more people than the device search takes (status 3: the
    Python path raises the documented ValueError)."""
    s, F, C = st.settings, len(st.table), st.n_cams
    if F == 0:
        return False
    if st.index is not None:                                # the native index owns the [F][C] path array
        paths, c_arr = (F, C), st.index.table_c_array()
    else:
        paths = _stg.frame_paths(st.dirs.pose_dir, st.cam_dirs, st.table)
        c_arr = _stg._c_paths(paths)                        # one ctypes array for every native pass over this table
    st.src_c_paths = c_arr
    nt = _io_threads(world)
    if not s["multi_person"]:
        NP = _lib.P2S_MAX_PERSONS
        o3, named, listed, _, status, inexact = _stg.read_people_files(paths, 3 * st.tracked_keypoint_id, 3, NP, nt, c_arr)
        if (status >= 2).any():
            return False
        obs = np.full((F, C, NP, 4), np.nan, np.float32)
        obs[..., 3] = 0.0
        have = np.arange(NP)[None, None, :] < np.minimum(named, listed)[:, :, None]     # :199-205: p < min(A, B)
        obs[..., :3] = np.where(have[..., None], o3, np.float32(np.nan))
        st.obs, st.count, st.inexact, st.parsed, st.native = obs, named, inexact, None, True
        if inexact:
            logging.warning(f"{inexact} 2D values are not exactly representable in float32 and were rounded for the "
                            f"device staging layout.")
        return True
    _, _, listed, llen, status, _ = _stg.read_people_files(paths, 0, 0, _lib.P2S_MAX_DETECTIONS, nt, c_arr)
    if (status >= 2).any():
        return False
    lengths = sorted(set(int(v) for v in llen[llen != 0]))
    if len(lengths) > 1 or (lengths and lengths[0] < 0):
        return False                                          # the Python path raises the stacking error with the lengths
    L = lengths[0] if lengths else 3
    if L % 3:
        return False
    NP = max(int(listed.max(initial=0)), 1)
    obs, _, listed2, _, status2, inexact = _stg.read_people_files(paths, 0, L, NP, nt, c_arr)
    if (status2 >= 2).any() or not np.array_equal(listed, listed2):
        return False
    n_max = int(listed.sum(axis=1).max(initial=0))
    if n_max > _lib.P2S_MAX_DETECTIONS:
        return False
    st.mp_staged = (obs, listed, inexact)
    st.obs, st.count, st.inexact, st.parsed, st.native = None, None, 0, None, True
    return True


def solve_frames(st, engine=None, device=0):
    """ONE device call for all frames.  Returns err[F], comb[F, C] (float, NaN = camera off), Q[F, 3]."""
    from . import ops
    eng = engine if engine is not None else ops.get_engine(device)
    s = st.settings
    n_p = max(1, int(st.count.max(initial=0)))
    obs = np.ascontiguousarray(st.obs[:, :, :n_p, :])
    out = eng.associate_host(obs, st.count, st.P, s["reproj_thr"], s["lik_thr"], s["min_cams"])
    comb = out["comb"].astype(np.float64)
    comb[out["comb"] < 0] = np.nan
    return {"err": out["err"], "comb": comb, "Q": out["Q"]}


def stage_multi_person(st):
    """personAssociation.py:783-790 for every frame -> obs[F, C, NP, 3 J], count[F, C], camera models."""
    from . import multi_person as mp
    obs, count, inexact = st.mp_staged if st.mp_staged is not None else mp.stage_detections(st.parsed)
    if inexact:
        logging.warning(f"{inexact} 2D values are not exactly representable in float32 and were rounded for the "
                        f"device staging layout.")
    return obs, count, _calib.camera_models(st.calib_file)


def solve_frames_multi_person(st, engine=None, device=0):
    """DEVICE: personAssociation.py:793-801 for all frames in one call.  Returns the proposals per frame."""
    from . import multi_person as mp
    from . import ops
    eng = engine if engine is not None else ops.get_engine(device)
    s = st.settings
    obs, count, models = stage_multi_person(st)
    return mp.associate_frames(eng, obs, count, models, s["reconstruction_error_threshold"], s["min_affinity"],
                               s["min_cams"])


def rewrite_frame(tracked_paths, source_js, proposals):
    """personAssociation.py:552-580: per camera write the source JSON with `people` replaced by the chosen
    person of every proposal ({} when the camera is off); a camera without a readable source gets no file.
    N.B. like the reference the RAW `people` list is indexed here."""
    for cam, path in enumerate(tracked_paths):
        try:
            with open(path, "w") as out:
                js = source_js[cam]
                if js is None:
                    raise FileNotFoundError
                new = dict(js)
                new["people"] = []
                for comb in proposals:
                    new["people"] += [js["people"][int(comb[cam])]] if not np.isnan(comb[cam]) else [{}]
                out.write(json.dumps(new))
        except Exception:
            os.remove(path)


def write_outputs(st, res, log=True):
    """personAssociation.py:776-808."""
    F = len(st.table)
    err, comb_all = np.asarray(res["err"], dtype=np.float64)[:F], np.asarray(res["comb"], dtype=np.float64)[:F]
    errors = [float(e) for e in err[~np.isinf(err)]]
    cams_off = [int(v) for v in np.count_nonzero(np.isnan(comb_all), axis=1)] if F else []
    if st.parsed is not None:
        for fi, names in enumerate(st.table):
            tracked = [os.path.join(st.dirs.tracked_dir, st.cam_dirs[c], names[c]) for c in range(st.n_cams)]
            rewrite_frame(tracked, st.parsed[fi], [comb_all[fi]])
    elif st.native or os.environ.get("P2S_NATIVE_IO", "1") != "0":
        _rewrite_native(st, comb_all.reshape(F, 1, st.n_cams))
    else:
        _rewrite_parallel(st, [[comb_all[fi]] for fi in range(F)])
    if log:
        log_recap(st, errors, cams_off)
    return {"error": errors, "cameras_off": cams_off}


def write_outputs_multi_person(st, proposals, log=True):
    if st.parsed is None:
        _rewrite_parallel(st, proposals)
    else:
        for fi, names in enumerate(st.table):
            tracked = [os.path.join(st.dirs.tracked_dir, st.cam_dirs[c], names[c]) for c in range(st.n_cams)]
            rewrite_frame(tracked, st.parsed[fi], proposals[fi])
    if not log:
        return
    s = st.settings
    logging.info(f"\n--> A person was reconstructed if the lines from cameras to their keypoints intersected within "
                 f"{s['reconstruction_error_threshold']} m and if the calculated affinity stayed above {s['min_affinity']}.")
    logging.info("--> Beware that people were sorted across cameras, but not across frames. This will be done in the "
                 "triangulation stage.")
    logging.info(f"\nTracked json files are stored in {os.path.realpath(st.dirs.tracked_dir)}.")


def log_recap(st, errors, cams_off):
    """personAssociation.py:583-639 `recap_tracking`, single-person branch."""
    s = st.settings
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        e_px = np.around(np.nanmean(errors) if len(errors) else np.nan, decimals=1)
    fm, Dm = _calib.first_camera_scale(st.calib_file)
    e_mm = np.around(e_px * Dm / fm * 1000, decimals=1)
    off = np.around(np.mean(cams_off), decimals=2)
    logging.info(f"\n--> Mean reprojection error for {s['tracked_keypoint']} point on all frames is {e_px} px, which "
                 f"roughly corresponds to {e_mm} mm. ")
    logging.info(f"--> In average, {off} cameras had to be excluded to reach the demanded {s['reproj_thr']} px error "
                 f"threshold after excluding points with likelihood below {s['lik_thr_recap']}.")
    logging.info(f"\nTracked json files are stored in {os.path.realpath(st.dirs.tracked_dir)}.")


def associate_all(config_dict):
    """Same contract as Pose2Sim/personAssociation.py:642: reads calibration + per-camera JSON, writes
    `pose-associated/<cam>_json/*.json` with one person of interest, logs the recap.  Returns None.

    Under torchrun (torch.distributed initialised, one process per GPU) the frames are sharded in contiguous
    blocks over the ranks; each rank stages, solves on its own GPU and writes its own files — no data-path
    exchange; only the recap's per-frame scalars travel to rank 0, which logs."""
    from .triangulation import _world
    rank, world, local = _world()
    if world == 1:
        st = stage_project(config_dict)
        if st.settings["multi_person"]:
            write_outputs_multi_person(st, solve_frames_multi_person(st))
            return
        write_outputs(st, solve_frames(st))
        return
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        import torch
        torch.cuda.set_device(local)                          # the object collectives below use the current device
    # rank-local work first, ONE object collective after it: a rank that fails (a file it alone reads is broken, its
    # device call errors) reports the failure through the collective instead of leaving the others waiting in it
    failure, part, st = None, None, None
    try:
        st = stage_project(config_dict, rank, world)
        if st.settings["multi_person"]:
            proposals = solve_frames_multi_person(st, device=local) if len(st.table) else []
            write_outputs_multi_person(st, proposals, log=rank == 0)
        else:
            res = solve_frames(st, device=local) if len(st.table) else {"err": np.zeros(0), "comb": np.zeros((0, st.n_cams)), "Q": np.zeros((0, 3))}
            part = write_outputs(st, res, log=False)
    except Exception as e:                                      # noqa: BLE001 — re-raised below
        failure = e
    reports = [None] * world
    dist.all_gather_object(reports, (part, None if failure is None else f"{type(failure).__name__}: {failure}"))
    if failure is not None:
        raise failure
    failed = [(r, msg) for r, (_, msg) in enumerate(reports) if msg is not None]
    if failed:
        raise RuntimeError(f"rank {failed[0][0]} failed in associate_all: {failed[0][1]}")
    if st.settings["multi_person"]:
        return
    parts = [p for p, _ in reports]
    if rank == 0:
        log_recap(st, [e for p in parts for e in p["error"]], [c for p in parts for c in p["cameras_off"]])
