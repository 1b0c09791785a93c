#!/usr/bin/env python
"""Error behaviour side by side: broken trials (no calibration, no pose folder, unknown skeleton, camera-count mismatch,
an empty camera folder, a frame range past the end, nobody in any file) through the UNMODIFIED reference's
`triangulate_all` / `associate_all` and this package's host pipelines (oracle in place of the device call; no GPU
needed).  Prints the exception type and message of both.  Build-container tool (needs /root/reference).

    python oracle/diff_errors_live.py 2>&1 | grep -E " SAME | DIFFERENT "
"""
import glob
import os
import shutil
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)

import make_golden_e2e as mg  # noqa: E402
import p2s_oracle as orc  # noqa: E402
import ref_shim  # noqa: E402
import test_dropin_host as tdh  # noqa: E402
from dropin_util import assert_trc_equal  # noqa: E402
from pose2sim_b200 import personAssociation as pa  # noqa: E402
from pose2sim_b200 import synth_project  # noqa: E402
from pose2sim_b200 import triangulation as tri  # noqa: E402


def break_no_calibration(proj, cfg, cams):
    shutil.rmtree(os.path.join(proj, "calibration"))


def break_no_pose(proj, cfg, cams):
    shutil.rmtree(os.path.join(proj, "pose"))


def break_unknown_model(proj, cfg, cams):
    cfg["pose"]["pose_model"] = "NO_SUCH_MODEL"


def break_camera_mismatch(proj, cfg, cams):
    shutil.rmtree(os.path.join(proj, "pose", f"{cams[-1]}_json"))


def break_empty_camera(proj, cfg, cams):
    for f in glob.glob(os.path.join(proj, "pose", f"{cams[1]}_json", "*.json")):
        os.remove(f)


def break_range_past_end(proj, cfg, cams):
    cfg["project"]["frame_range"] = [500, 600]


def break_nobody(proj, cfg, cams):
    for f in glob.glob(os.path.join(proj, "pose", "*", "*.json")):
        with open(f, "w") as out:
            out.write('{"version": 1.3, "people": []}')


def _files(proj, cam):
    return sorted(glob.glob(os.path.join(proj, "pose", f"{cam}_json", "*.json")))


def odd_truncated_json(proj, cfg, cams):
    for path in _files(proj, cams[0])[2:5] + _files(proj, cams[2])[7:8]:
        text = open(path).read()
        open(path, "w").write(text[:len(text) // 2])


def odd_short_and_null_keypoints(proj, cfg, cams):
    import json
    for i, path in enumerate(_files(proj, cams[1])[3:9]):
        js = json.load(open(path))
        if not js["people"]:
            continue
        kp = js["people"][0]["pose_keypoints_2d"]
        js["people"][0]["pose_keypoints_2d"] = kp[:10] if i % 2 else [None if j % 7 == 0 else v for j, v in enumerate(kp)]
        json.dump(js, open(path, "w"))


def odd_people_entries(proj, cfg, cams):
    import json
    variants = [["x"], [{}], [{"pose_keypoints_2d": []}], {"not": "a list"}, [{"pose_keypoints_2d": "abc"}], [None]]
    for path, people in zip(_files(proj, cams[3])[1:7], variants):
        js = json.load(open(path))
        js["people"] = people
        json.dump(js, open(path, "w"))


def odd_extra_dirs(proj, cfg, cams):
    os.mkdir(os.path.join(proj, "pose", f"{cams[0]}_img"))
    open(os.path.join(proj, "pose", f"{cams[0]}_img", "frame_000001.png"), "w").write("x")
    open(os.path.join(proj, "pose", f"{cams[0]}_json", "notes.txt"), "w").write("x")


def odd_renumbered(proj, cfg, cams):
    for cam in cams:
        for path in reversed(_files(proj, cam)):
            d, name = os.path.split(path)
            stem, num = name[:-5].rsplit("_", 1)
            os.rename(path, os.path.join(d, f"{stem}_{int(num) + 100:06d}.json"))


def odd_pose_sync(proj, cfg, cams):
    shutil.copytree(os.path.join(proj, "pose"), os.path.join(proj, "pose-sync"))
    os.remove(sorted(glob.glob(os.path.join(proj, "pose-sync", f"{cams[0]}_json", "*.json")))[3])
    for path in _files(proj, cams[1])[:2]:
        os.remove(path)                                        # pose/ and pose-sync/ now disagree


def _rename_all(proj, cams, fn):
    for ci, cam in enumerate(cams):
        for path in _files(proj, cam):
            d, name = os.path.split(path)
            stem, num = name[:-5].rsplit("_", 1)
            new = fn(ci, stem, int(num))
            if new != name:
                os.rename(path, os.path.join(d, new + "__tmp"))
        for path in glob.glob(os.path.join(proj, "pose", f"{cam}_json", "*__tmp")):
            os.rename(path, path[:-5])


def odd_no_zero_padding(proj, cfg, cams):
    _rename_all(proj, cams, lambda ci, stem, n: f"{stem}_{n}.json")


def odd_extra_numbers_in_names(proj, cfg, cams):
    _rename_all(proj, cams, lambda ci, stem, n: f"take2_{stem}_v3_{n:04d}_keypoints.json" if ci % 2 else f"{stem}_{n:06d}.json")


def odd_one_camera_shifted(proj, cfg, cams):
    _rename_all(proj, cams, lambda ci, stem, n: f"{stem}_{n + (3 if ci == 1 else 0):06d}.json")


def odd_duplicate_frame_number(proj, cfg, cams):
    src = _files(proj, cams[2])[6]
    shutil.copy(src, os.path.join(os.path.dirname(src), "again_5.json"))     # a second file carrying frame number 5


def odd_frame_range_auto(proj, cfg, cams):
    cfg["project"]["frame_range"] = "auto"


BREAKS = [break_no_calibration, break_no_pose, break_unknown_model, break_camera_mismatch, break_empty_camera,
          break_range_past_end, break_nobody, odd_truncated_json, odd_short_and_null_keypoints, odd_people_entries,
          odd_extra_dirs, odd_renumbered, odd_pose_sync, odd_frame_range_auto, odd_no_zero_padding, odd_extra_numbers_in_names,
          odd_one_camera_shifted, odd_duplicate_frame_number]


def ours_triangulate(cfg, proj):
    with mg.in_dir(proj):
        st = tri.stage_project(cfg)
        tri.write_outputs(st, tdh.oracle_units(st))


def ours_associate(cfg, proj):
    with mg.in_dir(proj):
        st = pa.stage_project(cfg)
        F, C = st.count.shape
        err, comb, Q = np.empty(F), np.empty((F, C)), np.empty((F, 3))
        s = st.settings
        for f in range(F):
            ob = [[st.obs[f, c, p, :3].astype(float) for p in range(st.count[f, c])] for c in range(C)]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                err[f], comb[f], Q[f] = orc.associate_frame(ob, list(st.count[f]), st.P, s["reproj_thr"], s["lik_thr"], s["min_cams"])
        pa.write_outputs(st, {"err": err, "comb": comb, "Q": Q}, log=False)


def ours_associate_multi(cfg, proj):
    import p2s_oracle_mp as omp
    from pose2sim_b200 import multi_person as mp
    with mg.in_dir(proj):
        st = pa.stage_project(cfg)
        obs, count, models = pa.stage_multi_person(st)
        s = st.settings
        rays = omp.camera_ray_params(models)
        proposals = []
        for f in range(len(count)):
            det = [[obs[f, c, p].astype(float) for p in range(count[f, c])] for c in range(st.n_cams)]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                aff, cum = omp.frame_affinity(det, rays, s["reconstruction_error_threshold"], s["min_affinity"])
            proposals.append(mp.proposals_from_rows(omp.argmax_rows(aff, cum), s["min_cams"]))
        pa.write_outputs_multi_person(st, proposals, log=False)


def ours_triangulate_multi(cfg, proj):
    with mg.in_dir(proj):
        st = tri.stage_project(cfg)
        res = tri.reidentify(tdh.oracle_units(st), st.f_range, st.n_cams, st.settings["max_distance_m"])
        tri.write_outputs(st, res)


def outputs(proj):
    """What the stage left on disk: relative path -> text (TRC) or re-serialised JSON (NaN-safe comparison)."""
    import json
    out = {}
    for path in sorted(glob.glob(os.path.join(proj, "pose-3d", "*")) + glob.glob(os.path.join(proj, "pose-associated", "*", "*.json"))):
        with open(path) as f:
            out[os.path.relpath(path, proj)] = json.dumps(json.load(f), sort_keys=True) if path.endswith(".json") else f.read()
    return out


def outcome(fn, cfg, proj):
    try:
        fn(cfg, proj)
        return ("no exception", ""), outputs(proj)
    except BaseException as e:                                   # noqa: BLE001 — the point is to see what comes out
        return (type(e).__name__, " ".join(str(e).split())[:110]), outputs(proj)


def main():
    ref = ref_shim.load_reference()
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    run_tri = lambda cfg, proj: mg.run_reference(ref.triangulation.triangulate_all, cfg, proj)          # noqa: E731
    run_assoc = lambda cfg, proj: mg.run_reference(ref.personAssociation.associate_all, cfg, proj)     # noqa: E731
    single = mg.single_person_trial()
    stages = [("triangulate_all", single, False, 20, run_tri, ours_triangulate),
              ("associate_all", single, False, 20, run_assoc, ours_associate),
              ("triangulate_all[multi_person]", mg.multi_person_trial(), True, 24, run_tri, ours_triangulate_multi),
              ("associate_all[multi_person]", mg.multi_association_trial(), True, 16, run_assoc, ours_associate_multi)]
    for stage, (calib_text, cams, kp, present), multi, n_frames, run_ref, run_ours in stages:
        if only and only not in stage:
            continue
        kp, present = kp[:n_frames], (present[:n_frames] if present is not None else None)
        for brk in BREAKS:
            if len(sys.argv) > 2 and sys.argv[2] not in brk.__name__:
                continue
            res = {}
            for who, run in (("ref", run_ref), ("ours", run_ours)):
                with tempfile.TemporaryDirectory() as td:
                    proj = synth_project.write_project(os.path.join(td, "trial_demo"), calib_text, cams, kp, present=present)
                    cfg = synth_project.base_config(proj, multi_person=multi)
                    brk(proj, cfg, cams)
                    res[who] = outcome(run, cfg, proj)
            (r_exc, r_out), (o_exc, o_out) = res["ref"], res["ours"]
            same = r_exc[0] == o_exc[0] and sorted(r_out) == sorted(o_out)
            for k in (r_out if same else ()):
                if k.endswith(".trc"):                          # coordinates within 1e-6 m, everything else identical
                    try:
                        assert_trc_equal(o_out[k], r_out[k], tol=1e-6)
                    except AssertionError:
                        same = False
                else:
                    same = same and r_out[k] == o_out[k]
            print(stage, brk.__name__, "SAME" if same else "DIFFERENT", "| ref:", r_exc, len(r_out), "files | ours:", o_exc, len(o_out), "files")


if __name__ == "__main__":
    main()
