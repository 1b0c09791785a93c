#!/usr/bin/env python
"""Side-by-side run of the single-person association trial under other settings: the UNMODIFIED reference's
`associate_all` and this package's host pipeline (staging -> oracle search -> JSON rewrite; no GPU needed) on the same
on-disk trial; the rewritten pose-associated/ trees are compared file by file (existence, and the decoded JSON).
Build-container tool like make_golden_*.py (needs /root/reference).

    python oracle/diff_assoc_live.py [multi] 2>&1 | grep -E " OK | MISMATCH "
"""
import glob
import json
import os
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
from pose2sim_b200 import personAssociation as pa  # noqa: E402
from pose2sim_b200 import synth_project  # noqa: E402

CASES = [
    # (name, [personAssociation] single_person overrides, [triangulation] overrides, [project] overrides, missing files)
    ("thr_5", {"reproj_error_threshold_association": 5}, {}, {}, []),
    ("lik_0.6_min_cams_3", {"likelihood_threshold_association": 0.6}, {"min_cameras_for_triangulation": 3}, {}, []),
    ("tracked_RHip", {"tracked_keypoint": "RHip"}, {}, {}, []),
    ("tracked_unknown_falls_back", {"tracked_keypoint": "Tail"}, {}, {}, []),
    ("frame_range_missing_files", {}, {}, {"frame_range": [3, 25]}, [(0, 5), (0, 6), (2, 10), (1, 24)]),
]


CASES_MULTI = [
    # (name, [personAssociation] multi_person overrides, [triangulation] overrides, [project] overrides, missing files)
    ("reconstruction_thr_0.03", {"reconstruction_error_threshold": 0.03}, {}, {}, []),
    ("reconstruction_thr_0.5", {"reconstruction_error_threshold": 0.5}, {}, {}, []),
    ("min_affinity_0.7", {"min_affinity": 0.7}, {}, {}, []),
    ("min_cams_3", {}, {"min_cameras_for_triangulation": 3}, {}, []),
    ("min_cams_4_frame_range_missing", {}, {"min_cameras_for_triangulation": 4}, {"frame_range": [2, 30]}, [(0, 5), (3, 5), (1, 12)]),
]


def tree(proj):
    out = {}
    for path in sorted(glob.glob(os.path.join(proj, "pose-associated", "*", "*.json"))):
        with open(path) as f:
            out[os.path.relpath(path, proj)] = json.load(f)
    return out


def same_json(a, b):
    """Equal including NaN positions (json.load yields float('nan') for the NaN literal)."""
    return json.dumps(a, sort_keys=True) == json.dumps(b, sort_keys=True)


def ours_single(cfg, proj):
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


def main():
    import diff_errors_live as de
    ref = ref_shim.load_reference()
    multi = len(sys.argv) > 1 and sys.argv[1] == "multi"
    calib_text, cams, kp, present = mg.multi_association_trial() if multi else mg.association_trial()
    kp, present = kp[:30], present[:30]
    for name, over, tri_over, prj, missing in (CASES_MULTI if multi else CASES):
        out = {}
        for who in ("ref", "ours"):
            with tempfile.TemporaryDirectory() as td:
                proj = synth_project.write_project(os.path.join(td, "trial_assoc"), calib_text, cams, kp, present=present)
                for c, f in missing:
                    path = os.path.join(proj, "pose", f"{cams[c]}_json", f"{cams[c]}_{f:06d}.json")
                    if os.path.exists(path):
                        os.remove(path)
                cfg = synth_project.base_config(proj, multi_person=multi, **tri_over)
                cfg["project"].update(prj)
                cfg["personAssociation"]["multi_person" if multi else "single_person"].update(over)
                if "likelihood_threshold_association" in over:
                    cfg["personAssociation"]["likelihood_threshold_association"] = over["likelihood_threshold_association"]
                try:
                    if who == "ref":
                        mg.run_reference(ref.personAssociation.associate_all, cfg, proj)
                    elif multi:
                        de.ours_associate_multi(cfg, proj)
                    else:
                        ours_single(cfg, proj)
                    exc = None
                except Exception as e:
                    exc = (type(e).__name__, str(e)[:80])
                out[who] = (exc, tree(proj))
        r, o = out["ref"], out["ours"]
        ok = r[0] == o[0] and sorted(r[1]) == sorted(o[1]) and all(same_json(r[1][k], o[1][k]) for k in r[1])
        diff = [k for k in r[1] if k in o[1] and not same_json(r[1][k], o[1][k])][:3]
        print(name, "OK" if ok else "MISMATCH", r[0], o[0], len(r[1]), len(o[1]), diff)


if __name__ == "__main__":
    main()
