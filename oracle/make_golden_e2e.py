"""TEST INFRASTRUCTURE — end-to-end golden fixtures made by running the UNMODIFIED reference's stage
entry points on synthetic trials (build container only; needs /root/reference):

    python oracle/make_golden_e2e.py

  * Pose2Sim/triangulation.py:656      triangulate_all(config)   single person, shipped demo cameras (cfg1)
  * Pose2Sim/triangulation.py:656      triangulate_all(config)   multi person (re-ID across frames)
  * Pose2Sim/personAssociation.py:642  associate_all(config)     single-person mode

For each trial the INPUT keypoint arrays (float32, enough to rebuild the JSON directories with
pose2sim_b200.synth_project) and the reference's OUTPUT (TRC text / chosen people) are stored in
tests/golden/e2e_*.npz.  tests/test_dropin_*.py rebuild the trial in a temp dir, run the drop-in and
compare.
"""
import contextlib
import glob
import io
import json
import logging
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from pose2sim_b200 import skeletons, synth, synth_project  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
J = 26                                                  # HALPE_26 keypoints per person in the JSON


@contextlib.contextmanager
def in_dir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def run_reference(fn, cfg, project_dir):
    """cwd = the trial directory (single-trial layout, triangulation.py:680-682)."""
    buf = io.StringIO()
    handler = logging.StreamHandler(buf)
    logging.getLogger().addHandler(handler)
    logging.getLogger().setLevel(logging.INFO)
    try:
        with in_dir(project_dir), warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            fn(cfg)
    finally:
        logging.getLogger().removeHandler(handler)
    return buf.getvalue()


def single_person_trial():
    """cfg1-shaped: the 4 shipped Qualisys cameras, 100 frames, HALPE_26; the person is absent in the
    first frames (trimmed), some keypoints vanish for short (interpolated) and long (filled) spans."""
    ids, names = skeletons.keypoints("HALPE_26")
    calib_text = open(os.path.join(GOLDEN, "Calib_demo.toml")).read()
    g = np.load(os.path.join(GOLDEN, "tri_cfg1_demo.npz"))
    P = g["P"]
    F, C, K = 100, 4, 26
    wl = synth.make_triangulation_workload(C, F, 1, K, seed=111, P=P, lik_thr=None, p_out=0.08, p_low=0.10)
    x, y, lik = (wl[k].reshape(F, 1, K, C).transpose(0, 3, 1, 2).copy() for k in ("x", "y", "lik"))   # [F, C, 1, K]
    lik[:4] = 0.05                                       # nobody visible in frames 0..3
    lik[30:36, :, :, 5] = 0.1                            # short gap on one keypoint  -> interpolated
    lik[50:80, :, :, 9] = 0.1                            # long gap on another        -> not interpolated
    lik[60:63] = 0.0                                     # whole person lost for 3 frames
    kp = synth_project.pack_openpose(x, y, lik, ids, J)  # [F, C, 1, 78]
    return calib_text, [f"cam{c + 1:02d}" for c in range(C)], kp, None


def multi_person_trial():
    """3 persons, 4 ring cameras, 70 frames; the people list is permuted identically in all cameras
    from some frames on (what `pose-associated` looks like), one person leaves for a while."""
    ids, names = skeletons.keypoints("HALPE_26")
    C, F, Np, K = 4, 70, 3, 26
    calib_text, cams, P = synth_project.ring_calibration_toml(C)
    Q = synth.truth_points(F, Np, K, 222)
    x, y, lik = synth.observe(Q, P, 222, sigma=1.5, p_out=0.04, p_low=0.05)      # [F, Np, K, C]
    x, y, lik = (a.transpose(0, 3, 1, 2).copy() for a in (x, y, lik))            # [F, C, Np, K]
    order = np.tile(np.arange(Np), (F, 1))
    order[20:45] = [1, 2, 0]
    order[45:] = [2, 0, 1]
    fidx = np.arange(F)[:, None]
    x, y, lik = (a[fidx, :, order].transpose(0, 2, 1, 3) for a in (x, y, lik))   # people permuted per frame
    present = np.ones((F, C, Np), bool)
    lik[30:41, :, 2] = 0.01                               # the person at slot 2 is not seen for 11 frames
    kp = synth_project.pack_openpose(x, y, lik, ids, J)
    return calib_text, cams, kp, present


def association_trial():
    """Single-person association: 4 ring cameras, 2-3 detections per camera in random order,
    outliers, low likelihoods, missing detections."""
    ids, names = skeletons.keypoints("HALPE_26")
    C, F, Np, K = 4, 60, 3, 26
    calib_text, cams, P = synth_project.ring_calibration_toml(C)
    Q = synth.truth_points(F, Np, K, 333)
    x, y, lik = synth.observe(Q, P, 333, sigma=1.5, p_out=0.10, p_low=0.10)
    x, y, lik = (a.transpose(0, 3, 1, 2).copy() for a in (x, y, lik))            # [F, C, Np, K]
    g = np.random.default_rng(333)
    perm = np.stack([np.stack([g.permutation(Np) for _ in range(C)]) for _ in range(F)])   # [F, C, Np]
    fidx, cidx = np.arange(F)[:, None, None], np.arange(C)[None, :, None]
    x, y, lik = (a[fidx, cidx, perm] for a in (x, y, lik))
    present = np.ones((F, C, Np), bool)
    present[:, :, 2] = g.random((F, C)) < 0.7             # the third detection is often missing
    present[g.random((F, C)) < 0.08] = False              # sometimes a camera sees nobody
    kp = synth_project.pack_openpose(x, y, lik, ids, J)
    return calib_text, cams, kp, present


def read_people_arrays(root, cams, frames, n_slots):
    """[F, C, n_slots, 3J] keypoints of the files under root (NaN where a slot is `{}` or missing) and
    exists[F, C]."""
    F, C = len(frames), len(cams)
    out = np.full((F, C, n_slots, 3 * J), np.nan)
    exists = np.zeros((F, C), bool)
    for c, cam in enumerate(cams):
        for fi, f in enumerate(frames):
            path = os.path.join(root, f"{cam}_json", f"{cam}_{f:06d}.json")
            if not os.path.exists(path):
                continue
            exists[fi, c] = True
            people = json.load(open(path))["people"]
            assert len(people) <= n_slots
            for p, person in enumerate(people):
                if person:
                    out[fi, c, p] = person["pose_keypoints_2d"]
    return out, exists


def main():
    ref = ref_shim.load_reference()
    os.makedirs(GOLDEN, exist_ok=True)

    # 1. single person --------------------------------------------------------------------------------
    for tag, trial, multi, extra in (("e2e_tri_single", single_person_trial, False, {}),
                                     ("e2e_tri_multi", multi_person_trial, True, {"interpolation": "cubic", "fill_large_gaps_with": "nan"})):
        calib_text, cams, kp, present = trial()
        with tempfile.TemporaryDirectory() as td:
            proj = synth_project.write_project(os.path.join(td, "trial_demo"), calib_text, cams, kp, present=present)
            cfg = synth_project.base_config(proj, multi_person=multi, **extra)
            log = run_reference(ref.triangulation.triangulate_all, cfg, proj)
            trcs = sorted(glob.glob(os.path.join(proj, "pose-3d", "*.trc")))
            assert trcs, "reference wrote no TRC"
            out = {"calib": np.array(calib_text), "cams": np.array(cams), "kp": kp.astype(np.float32),
                   "present": np.ones(kp.shape[:3], bool) if present is None else present,
                   "multi_person": np.array(multi), "extra": np.array(json.dumps(extra)),
                   "trc_names": np.array([os.path.basename(t) for t in trcs]), "log": np.array(log)}
            for i, t in enumerate(trcs):
                out[f"trc{i}"] = np.array(open(t).read())
            np.savez_compressed(os.path.join(GOLDEN, tag + ".npz"), **out)
            print(tag, [os.path.basename(t) for t in trcs])
            print("   " + "\n   ".join(l for l in log.splitlines() if l.startswith("-->") or "Camera" in l)[:600])

    # 2. association -----------------------------------------------------------------------------------
    calib_text, cams, kp, present = association_trial()
    with tempfile.TemporaryDirectory() as td:
        proj = synth_project.write_project(os.path.join(td, "trial_assoc"), calib_text, cams, kp, present=present)
        cfg = synth_project.base_config(proj)
        log = run_reference(ref.personAssociation.associate_all, cfg, proj)
        chosen, exists = read_people_arrays(os.path.join(proj, "pose-associated"), cams, range(kp.shape[0]), 1)
        np.savez_compressed(os.path.join(GOLDEN, "e2e_assoc_single.npz"), calib=np.array(calib_text), cams=np.array(cams),
                            kp=kp.astype(np.float32), present=present, chosen=chosen[:, :, 0].astype(np.float32),
                            exists=exists, log=np.array(log))
        print("e2e_assoc_single", "cameras off per frame:", np.isnan(chosen[:, :, 0, 0]).sum(1).mean())
        print("   " + "\n   ".join(l for l in log.splitlines() if l.startswith("-->")))


def undistort_trial(swap_frac=0.0):
    """Ring cameras WITH lens distortion (5 coefficients, like the fork's k3-enabled calibration),
    observations produced by cv2.projectPoints on the truth, `[triangulation] undistort_points = true`.
    swap_frac > 0: in that share of the (frame, camera) views every keypoint trades places with its left/right
    partner (what `handle_LR_swap` is for); the random draws of the base trial are unchanged."""
    import cv2
    from pose2sim_b200 import calib
    ids, names = skeletons.keypoints("HALPE_26")
    C, F, K = 4, 60, 26
    P, Ks, Rs, ts = synth.ring_cameras(C)
    dists = [[-0.06, 0.03, 8e-4, -6e-4, 0.01], [-0.04, 0.01, -5e-4, 3e-4, 0.0], [0.03, -0.02, 2e-4, 1e-4, 0.005],
             [-0.08, 0.05, 0.0, 0.0, -0.01]]
    cams = [f"cam{c + 1:02d}" for c in range(C)]
    td = tempfile.mkdtemp()
    path = os.path.join(td, "c.toml")
    calib.write_calibration_toml(path, cams, [(1080.0, 1920.0)] * C, Ks, dists, [calib.rotation_to_rodrigues(R) for R in Rs], ts)
    calib_text = open(path).read()
    Q = synth.truth_points(F, 1, K, 444)[:, 0]                                   # [F, K, 3]
    g = np.random.default_rng(444)
    x = np.empty((F, C, 1, K)); y = np.empty((F, C, 1, K)); lik = np.empty((F, C, 1, K))
    for c in range(C):
        rvec = cv2.Rodrigues(Rs[c])[0]
        uv = cv2.projectPoints(Q.reshape(-1, 3), rvec, ts[c], Ks[c], np.array(dists[c]))[0].reshape(F, K, 2)
        x[:, c, 0], y[:, c, 0] = uv[..., 0], uv[..., 1]
    x += g.normal(0, 1.5, x.shape); y += g.normal(0, 1.5, y.shape)
    out = g.random(x.shape) < 0.06
    x = np.where(out, x + g.uniform(60, 200, x.shape), x)
    lik[:] = g.uniform(0.5, 1.0, lik.shape)
    lik = np.where(g.random(lik.shape) < 0.06, g.uniform(0.0, 0.3, lik.shape), lik)
    if swap_frac > 0:
        part = np.asarray(skeletons.swapped_indices(names))
        sw = np.random.default_rng(445).random((F, C, 1, 1)) < swap_frac
        x, y, lik = (np.where(sw, a[..., part], a) for a in (x, y, lik))
    kp = synth_project.pack_openpose(x.astype(np.float32), y.astype(np.float32), lik.astype(np.float32), ids, J)
    return calib_text, cams, kp, None


def main_undistort(tag="e2e_tri_undistort", extra=None, swap_frac=0.0):
    """`extra` = the [triangulation] overrides of the trial; with handle_LR_swap in it the per-unit reference outputs are
    taken with the partner keypoint's coordinates as `coords_swapped` (triangulation.py:838)."""
    ref = ref_shim.load_reference()
    calib_text, cams, kp, present = undistort_trial(swap_frac)
    with tempfile.TemporaryDirectory() as td:
        proj = synth_project.write_project(os.path.join(td, "trial_demo"), calib_text, cams, kp, present=present)
        extra = dict(extra or {"undistort_points": True})
        cfg = synth_project.base_config(proj, **extra)
        log = run_reference(ref.triangulation.triangulate_all, cfg, proj)
        trcs = sorted(glob.glob(os.path.join(proj, "pose-3d", "*.trc")))
        assert trcs
        out = {"calib": np.array(calib_text), "cams": np.array(cams), "kp": kp.astype(np.float32),
               "present": np.ones(kp.shape[:3], bool), "multi_person": np.array(False), "extra": np.array(json.dumps(extra)),
               "trc_names": np.array([os.path.basename(t) for t in trcs]), "log": np.array(log)}
        for i, t in enumerate(trcs):
            out[f"trc{i}"] = np.array(open(t).read())
        # per-unit outputs of the reference's search on the same (cv2-undistorted, gated) inputs
        import cv2
        calib_file = glob.glob(os.path.join(proj, "calibration", "*.toml"))[0]
        Pm = ref.common.computeP(calib_file, undistort=True)
        cp = ref.common.retrieve_calib_params(calib_file)
        ids, names = skeletons.keypoints("HALPE_26")
        part = skeletons.swapped_indices(names)
        F, C = kp.shape[0], kp.shape[1]
        Kn = len(ids)
        xs = kp[:, :, 0, :][:, :, 3 * np.asarray(ids)].astype(np.float64)                  # [F, C, K]
        ys = kp[:, :, 0, :][:, :, 3 * np.asarray(ids) + 1].astype(np.float64)
        ls = kp[:, :, 0, :][:, :, 3 * np.asarray(ids) + 2].astype(np.float64)
        ux, uy = np.empty_like(xs), np.empty_like(ys)
        for f in range(F):
            for c in range(C):
                pts = np.stack([xs[f, c], ys[f, c]], 1).reshape(-1, 1, 2).astype("float32")
                u = cv2.undistortPoints(pts, cp["K"][c], cp["dist"][c], None, cp["optim_K"][c]).reshape(-1, 2)
                ux[f, c], uy[f, c] = u[:, 0], u[:, 1]
        low = ls < 0.3
        gx, gy, gl = np.where(low, np.nan, ux), np.where(low, np.nan, uy), np.where(low, np.nan, ls)
        cfg_u = {"triangulation": {"reproj_error_threshold_triangulation": extra.get("reproj_error_threshold_triangulation", 15),
                                   "min_cameras_for_triangulation": 2,
                                   "handle_LR_swap": bool(extra.get("handle_LR_swap", False)), "undistort_points": True}}
        U = F * Kn
        Qo, eo, no, mo = np.empty((U, 3)), np.empty(U), np.empty(U, np.int32), np.zeros(U, np.uint32)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for f in range(F):
                for k in range(Kn):
                    coords = np.array([gx[f, :, k], gy[f, :, k], gl[f, :, k]])
                    swapped = np.array([gx[f, :, part[k]], gy[f, :, part[k]], gl[f, :, part[k]]])
                    q, e, n, idl = ref.triangulation.triangulation_from_best_cameras(cfg_u, coords, swapped, Pm, cp)
                    u = f * Kn + k
                    Qo[u], eo[u], no[u] = np.asarray(q, float)[:3], e, n
                    for c in np.asarray(idl).ravel():
                        mo[u] |= np.uint32(1 << int(c))
        out.update(unit_x=xs.transpose(0, 2, 1).reshape(U, C).astype(np.float32), unit_y=ys.transpose(0, 2, 1).reshape(U, C).astype(np.float32),
                   unit_lik=ls.transpose(0, 2, 1).reshape(U, C).astype(np.float32),
                   unit_ux=ux.transpose(0, 2, 1).reshape(U, C).astype(np.float32), unit_uy=uy.transpose(0, 2, 1).reshape(U, C).astype(np.float32),
                   unit_Q=Qo, unit_err=eo, unit_nexcl=no, unit_mask=mo, unit_P=np.array(Pm))
        np.savez_compressed(os.path.join(GOLDEN, tag + ".npz"), **out)
        print(tag, [os.path.basename(t) for t in trcs], "triangulated", np.isfinite(eo).mean(), "mean nexcl", no.mean())


VARIANTS = [
    # (name, config overrides under [triangulation] / [project], files to delete as (camera index, frame))
    ("frame_range", {"frame_range": [10, 80]}, []),
    ("incomplete_largest_zeros", {"remove_incomplete_frames": True, "sections_to_keep": "largest",
                                  "fill_large_gaps_with": "zeros", "min_chunk_size": 5}, []),
    ("no_interp_nan_fill", {"interpolation": "none", "fill_large_gaps_with": "nan", "show_interp_indices": False}, []),
    ("cubic_first_section", {"interpolation": "cubic", "sections_to_keep": "first", "interp_if_gap_smaller_than": 4}, []),
    ("missing_files", {"sections_to_keep": "last"}, [(1, f) for f in range(20, 26)] + [(3, 50), (0, 97), (0, 98), (0, 99)]),
    ("min_cams_3_thr_8", {"min_cameras_for_triangulation": 3, "reproj_error_threshold_triangulation": 8,
                          "likelihood_threshold_triangulation": 0.5}, []),
    ("lr_swap_thr_6", {"handle_LR_swap": True, "reproj_error_threshold_triangulation": 6}, []),
]


# a second batch (own file, so that the first one's bytes stay as they are): the remaining interpolation kinds, the
# empty-list frame range, a threshold tight enough to fail most units (long gaps -> fill rules), a trial whose first
# and last frames are missing in one camera, every option at a non-default value at once
VARIANTS_2 = [
    ("quadratic_gap_50", {"interpolation": "quadratic", "interp_if_gap_smaller_than": 50}, []),
    ("slinear_incomplete_all", {"interpolation": "slinear", "remove_incomplete_frames": True, "sections_to_keep": "all"}, []),
    ("frame_range_empty_list", {"frame_range": []}, []),
    ("thr_3_min_cams_4_last_value", {"reproj_error_threshold_triangulation": 3, "min_cameras_for_triangulation": 4,
                                     "interp_if_gap_smaller_than": 3}, []),
    ("edges_missing_largest", {"sections_to_keep": "largest", "fill_large_gaps_with": "nan", "min_chunk_size": 3},
     [(2, f) for f in range(0, 4)] + [(1, f) for f in range(95, 100)]),
    ("everything_non_default", {"frame_range": [5, 90], "interpolation": "cubic", "interp_if_gap_smaller_than": 7,
                                "sections_to_keep": "last", "min_chunk_size": 4,
                                "fill_large_gaps_with": "zeros", "show_interp_indices": False,
                                "reproj_error_threshold_triangulation": 10, "likelihood_threshold_triangulation": 0.4,
                                "min_cameras_for_triangulation": 3}, [(0, 40), (0, 41), (3, 41)]),
]


# third batch: configuration values at their edges (the list oracle/diff_variants_live.py `edge` runs side by side with
# the live reference) stored as goldens so that the GPU box runs them through the device too.  A variant either writes
# one TRC or raises; the exception's type and message are then the golden.
VARIANTS_3 = [
    ("lik_thr_0", {"likelihood_threshold_triangulation": 0.0}, []),
    ("lik_thr_1", {"likelihood_threshold_triangulation": 1.0}, []),
    ("min_cams_5_of_4", {"min_cameras_for_triangulation": 5}, []),
    ("min_cams_1", {"min_cameras_for_triangulation": 1}, []),
    ("interp_gap_0", {"interp_if_gap_smaller_than": 0}, []),
    ("frame_rate_auto", {"frame_rate": "auto"}, []),
    ("alias_body_with_feet", {"pose_model": "BODY_WITH_FEET"}, []),
    ("unknown_interpolation", {"interpolation": "spline9"}, []),
    ("unknown_fill", {"fill_large_gaps_with": "sevens", "interp_if_gap_smaller_than": 2, "reproj_error_threshold_triangulation": 5}, []),
    ("unknown_sections", {"sections_to_keep": "middle", "reproj_error_threshold_triangulation": 5}, []),
    ("frame_range_reversed", {"frame_range": [40, 10]}, []),
    ("frame_range_partly_past_end", {"frame_range": [90, 130]}, []),
    ("min_chunk_100", {"min_chunk_size": 100, "reproj_error_threshold_triangulation": 5}, []),
    ("frame_rate_59.94", {"frame_rate": 59.94}, []),
    ("frame_rate_float_30", {"frame_rate": 30.0}, []),
    ("frame_rate_120_range", {"frame_rate": 120, "frame_range": [7, 61]}, []),
    ("max_distance_tiny_multi_off", {"max_distance_m": 0.001}, []),
    ("lik_thr_nan_free_0.999", {"likelihood_threshold_triangulation": 0.999, "min_cameras_for_triangulation": 2}, []),
]
PROJECT_KEYS = ("frame_range", "frame_rate")


def variant_config(proj, over):
    """[project] keys, the pose model and [triangulation] keys of a variant's overrides -> config_dict."""
    prj = {k: over[k] for k in over if k in PROJECT_KEYS}
    cfg = synth_project.base_config(proj, **{k: v for k, v in over.items() if k not in prj and k != "pose_model"})
    cfg["project"].update(prj)
    if "pose_model" in over:
        cfg["pose"]["pose_model"] = over["pose_model"]
    return cfg


def main_variants_edge(variants=None, file_name="e2e_tri_variants3.npz"):
    variants = VARIANTS_3 if variants is None else variants
    ref = ref_shim.load_reference()
    calib_text, cams, kp, present = single_person_trial()
    out = {"names": np.array([v[0] for v in variants])}
    for i, (name, over, missing) in enumerate(variants):
        with tempfile.TemporaryDirectory() as td:
            proj = synth_project.write_project(os.path.join(td, "trial_demo"), calib_text, cams, kp, present=present)
            for c, f in missing:
                os.remove(os.path.join(proj, "pose", f"{cams[c]}_json", f"{cams[c]}_{f:06d}.json"))
            cfg = variant_config(proj, over)
            exc = ""
            try:
                run_reference(ref.triangulation.triangulate_all, cfg, proj)
            except Exception as e:                                   # the reference's own failure is the contract
                exc = type(e).__name__ + ": " + str(e)[:60]
            trcs = sorted(glob.glob(os.path.join(proj, "pose-3d", "*.trc")))
            assert len(trcs) <= 1, (name, trcs)
            out[f"v{i}_over"] = np.array(json.dumps(over))
            out[f"v{i}_missing"] = np.array(missing, dtype=np.int64).reshape(-1, 2)
            out[f"v{i}_exc"] = np.array(exc)
            out[f"v{i}_trc_name"] = np.array(os.path.basename(trcs[0]) if trcs else "")
            out[f"v{i}_trc"] = np.array(open(trcs[0]).read() if trcs else "")
            print("edge variant", name, os.path.basename(trcs[0]) if trcs else None, exc)
    np.savez_compressed(os.path.join(GOLDEN, file_name), **out)


def main_variants(variants=None, file_name="e2e_tri_variants.npz"):
    """The single-person trial under other settings: frame ranges, trimming / fill / interpolation modes,
    missing files, other thresholds.  Inputs are those of e2e_tri_single.npz; only the reference's TRC text
    per variant is stored."""
    variants = VARIANTS if variants is None else variants
    ref = ref_shim.load_reference()
    calib_text, cams, kp, present = single_person_trial()
    out = {"names": np.array([v[0] for v in variants])}
    for i, (name, over, missing) in enumerate(variants):
        with tempfile.TemporaryDirectory() as td:
            proj = synth_project.write_project(os.path.join(td, "trial_demo"), calib_text, cams, kp, present=present)
            for c, f in missing:
                os.remove(os.path.join(proj, "pose", f"{cams[c]}_json", f"{cams[c]}_{f:06d}.json"))
            prj = {k: over[k] for k in over if k in ("frame_range",)}
            tri = {k: over[k] for k in over if k not in prj}
            cfg = synth_project.base_config(proj, **tri)
            cfg["project"].update(prj)
            log = run_reference(ref.triangulation.triangulate_all, cfg, proj)
            trcs = sorted(glob.glob(os.path.join(proj, "pose-3d", "*.trc")))
            assert len(trcs) == 1, (name, trcs)
            out[f"v{i}_over"] = np.array(json.dumps(over))
            out[f"v{i}_missing"] = np.array(missing, dtype=np.int64).reshape(-1, 2)
            out[f"v{i}_trc_name"] = np.array(os.path.basename(trcs[0]))
            out[f"v{i}_trc"] = np.array(open(trcs[0]).read())
            out[f"v{i}_log"] = np.array(log)
            print("variant", name, os.path.basename(trcs[0]))
    np.savez_compressed(os.path.join(GOLDEN, file_name), **out)


def multi_association_trial():
    """Multi-person association: 4 ring cameras, 3 persons in random per-camera order, a detection
    missing now and then."""
    calib_text, cams, kp, present = association_trial()
    return calib_text, cams, kp[:40], present[:40]


def main_multi_association():
    ref = ref_shim.load_reference()
    calib_text, cams, kp, present = multi_association_trial()
    with tempfile.TemporaryDirectory() as td:
        proj = synth_project.write_project(os.path.join(td, "trial_massoc"), calib_text, cams, kp, present=present)
        cfg = synth_project.base_config(proj, multi_person=True)
        log = run_reference(ref.personAssociation.associate_all, cfg, proj)
        chosen, exists = read_people_arrays(os.path.join(proj, "pose-associated"), cams, range(kp.shape[0]), 8)
        n_people = np.zeros(exists.shape, np.int32)
        for c, cam in enumerate(cams):
            for f in range(kp.shape[0]):
                path = os.path.join(proj, "pose-associated", f"{cam}_json", f"{cam}_{f:06d}.json")
                if os.path.exists(path):
                    n_people[f, c] = len(json.load(open(path))["people"])
        np.savez_compressed(os.path.join(GOLDEN, "e2e_assoc_multi.npz"), calib=np.array(calib_text), cams=np.array(cams),
                            kp=kp.astype(np.float32), present=present, multi_person=np.array(True),
                            chosen=chosen.astype(np.float32), exists=exists, n_people=n_people, log=np.array(log))
        print("e2e_assoc_multi: persons per frame", np.bincount(n_people[:, 0]))


# both off-by-default flags at once: lens distortion + left/right swapped limbs in 20 % of the views
UNDISTORT_LRSWAP = {"undistort_points": True, "handle_LR_swap": True, "reproj_error_threshold_triangulation": 6}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "multi_assoc":
        main_multi_association()
    elif len(sys.argv) > 1 and sys.argv[1] == "undistort":
        main_undistort()
    elif len(sys.argv) > 1 and sys.argv[1] == "undistort_lrswap":
        main_undistort("e2e_tri_undistort_lrswap", UNDISTORT_LRSWAP, swap_frac=0.2)
    elif len(sys.argv) > 1 and sys.argv[1] == "variants":
        main_variants()
    elif len(sys.argv) > 1 and sys.argv[1] == "variants3":
        main_variants_edge()
    elif len(sys.argv) > 1 and sys.argv[1] == "variants2":
        main_variants(VARIANTS_2, "e2e_tri_variants2.npz")
        main_variants_edge()
    else:
        main()
        main_multi_association()
        main_undistort()
        main_undistort("e2e_tri_undistort_lrswap", UNDISTORT_LRSWAP, swap_frac=0.2)
        main_variants()
        main_variants(VARIANTS_2, "e2e_tri_variants2.npz")
        main_variants_edge()
