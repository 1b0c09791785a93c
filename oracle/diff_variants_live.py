#!/usr/bin/env python
"""Side-by-side run of the multi-person trial under other settings: the UNMODIFIED reference's `triangulate_all` and
this package's host pipeline (staging -> oracle units -> re-ID -> TRC; no GPU needed) on the same on-disk trial, TRC
files compared.  Build-container tool like make_golden_*.py (needs /root/reference); nothing is stored — a mismatch
here is a host-logic defect to fix, and the case then becomes a golden variant.

    python oracle/diff_variants_live.py [multi|edge|undistort] 2>&1 | grep -E " OK | MISMATCH |TRC mismatch"

`edge`: the single-person trial with configuration values at their edges (thresholds 0 / 1, more cameras required than
exist, unknown option values, reversed and overshooting frame ranges, ...).
"""
import glob
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import ref_shim, make_golden_e2e as mg
from pose2sim_b200 import synth_project, triangulation as tri
import test_dropin_host as tdh
from dropin_util import assert_trc_equal
ref = ref_shim.load_reference()
TRIALS = {"multi": mg.multi_person_trial(), "single": mg.single_person_trial(), "undistort": mg.undistort_trial(0.2)}
UNDISTORT_CASES = [
    # the trial with lens distortion (limbs swapped in 20 % of the views), `undistort_points = true` throughout
    ("undistort_frame_range", {"undistort_points": True, "frame_range": [5, 40]}, []),
    ("undistort_missing_files_thr_8", {"undistort_points": True, "reproj_error_threshold_triangulation": 8}, [(0, 3), (0, 4), (2, 30), (3, 59)]),
    ("undistort_lr_swap_min_cams_3", {"undistort_points": True, "handle_LR_swap": True, "min_cameras_for_triangulation": 3,
                                      "reproj_error_threshold_triangulation": 5}, []),
    ("undistort_off_on_distorted_data", {"undistort_points": False}, []),
]
EDGE_CASES = [
    # configuration values at their edges, single-person trial: (name, [triangulation] / [project] / [pose] overrides, missing)
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
CASES = [
    ("frame_range", {"frame_range": [5, 30]}, []),
    ("missing+largest", {"sections_to_keep": "largest", "fill_large_gaps_with": "nan"}, [(1, f) for f in range(8, 12)] + [(0, 0)]),
    ("thr6_min3_cubic", {"reproj_error_threshold_triangulation": 6, "min_cameras_for_triangulation": 3, "interpolation": "cubic", "max_distance_m": 0.3}, []),
    ("incomplete_first", {"remove_incomplete_frames": True, "sections_to_keep": "first", "min_chunk_size": 3}, []),
    ("lr_swap", {"handle_LR_swap": True, "reproj_error_threshold_triangulation": 5}, []),
]
which = sys.argv[1] if len(sys.argv) > 1 else "multi"
for name, over, missing in {"edge": EDGE_CASES, "undistort": UNDISTORT_CASES}.get(which, CASES):
    calib_text, cams, kp, present = TRIALS[{"edge": "single", "undistort": "undistort"}.get(which, "multi")]
    multi = which not in ("edge", "undistort")
    out = {}
    for who in ("ref", "ours"):
        with tempfile.TemporaryDirectory() as td:
            proj = synth_project.write_project(os.path.join(td, "trial_demo"), calib_text, cams, kp, present=present)
            for c, f in missing:
                os.remove(os.path.join(proj, "pose", f"{cams[c]}_json", f"{cams[c]}_{f:06d}.json"))
            prj = {k: over[k] for k in over if k in ("frame_range", "frame_rate")}
            cfg = synth_project.base_config(proj, multi_person=multi, **{k: v for k, v in over.items() if k not in prj and k != "pose_model"})
            cfg["project"].update(prj)
            if "pose_model" in over:
                cfg["pose"]["pose_model"] = over["pose_model"]
            try:
                if who == "ref":
                    mg.run_reference(ref.triangulation.triangulate_all, cfg, proj)
                else:
                    with tdh.in_dir(proj):
                        st = tri.stage_project(cfg)
                        res = tdh.oracle_units(st)
                        if multi:
                            res = tri.reidentify(res, st.f_range, st.n_cams, st.settings["max_distance_m"])
                        tri.write_outputs(st, res)
                exc = None
            except Exception as e:
                exc = (type(e).__name__, str(e)[:80])
            out[who] = (exc, {os.path.basename(f): open(f).read() for f in glob.glob(os.path.join(proj, "pose-3d", "*"))})
    r, o = out["ref"], out["ours"]
    ok = r[0] == o[0] and sorted(r[1]) == sorted(o[1])
    worst = 0.0
    if ok:
        try:
            for k in r[1]:
                worst = max(worst, assert_trc_equal(o[1][k], r[1][k], tol=1e-6))
        except AssertionError as e:
            ok = False; print("   TRC mismatch", k, str(e)[:200])
    print(name, "OK" if ok else "MISMATCH", r[0], o[0], sorted(r[1]), sorted(o[1]), worst)
