#!/usr/bin/env python
"""Other skeletons side by side: a short single-person trial per pose model (shipped tables and a custom one declared in
Config.toml) through `associate_all` then `triangulate_all` of the UNMODIFIED reference and of this package's host
pipelines (oracle in place of the device calls; no GPU needed).  Compares the pose-associated trees and the TRC files
(marker names and order in the header included).  Build-container tool (needs /root/reference).

    python oracle/diff_models_live.py 2>&1 | grep -E " SAME | DIFFERENT "
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)

import diff_errors_live as de  # noqa: E402
import make_golden_e2e as mg  # noqa: E402
import ref_shim  # noqa: E402
from dropin_util import assert_trc_equal  # noqa: E402
from pose2sim_b200 import skeletons, synth, synth_project  # noqa: E402

CUSTOM = {"name": "Hip", "id": 3, "children": [
    {"name": "RKnee", "id": 0, "children": [{"name": "RFoot", "id": 5}]},
    {"name": "LKnee", "id": 1, "children": [{"name": "LFoot", "id": 6}]},
    {"name": "Spine", "id": "None", "children": [{"name": "Neck", "id": 2, "children": [{"name": "Head", "id": 4}]}]}]}

MODELS = ["COCO_17", "BODY_25B", "BLAZEPOSE", "COCO_133_WRIST", "HAND_21", "MPII", "BODY", "CUSTOM"]


def trial(model, cfg_pose):
    ids, names = skeletons.keypoints(model, {"pose": cfg_pose})
    readable = [k for k, i in enumerate(ids) if i < skeletons.UNREADABLE_ID]     # a custom node with a string id has no data
    K, J = len(ids), max(ids[k] for k in readable) + 1
    calib_text = open(os.path.join(mg.GOLDEN, "Calib_demo.toml")).read()
    P = np.load(os.path.join(mg.GOLDEN, "tri_cfg1_demo.npz"))["P"]
    F, C = 30, 4
    wl = synth.make_triangulation_workload(C, F, 1, K, seed=17 + K, P=P, lik_thr=None, p_out=0.08, p_low=0.10)
    x, y, lik = (wl[k].reshape(F, 1, K, C).transpose(0, 3, 1, 2).copy() for k in ("x", "y", "lik"))
    lik[10:13, :, :, K // 2] = 0.1
    x, y, lik = (a[..., readable] for a in (x, y, lik))
    return calib_text, [f"cam{c + 1:02d}" for c in range(C)], synth_project.pack_openpose(x, y, lik, [ids[k] for k in readable], J)


def main():
    ref = ref_shim.load_reference()
    for model in MODELS:
        pose = {"pose_model": model, "vid_img_extension": "mp4"}
        if model == "CUSTOM":
            pose["CUSTOM"] = CUSTOM
        calib_text, cams, kp = trial(model, pose)
        res = {}
        for who in ("ref", "ours"):
            with tempfile.TemporaryDirectory() as td:
                proj = synth_project.write_project(os.path.join(td, "trial_demo"), calib_text, cams, kp)
                cfg = synth_project.base_config(proj)
                cfg["pose"] = dict(pose)
                if who == "ref":
                    a = de.outcome(lambda c, p: mg.run_reference(ref.personAssociation.associate_all, c, p), cfg, proj)[0]
                    t, out = de.outcome(lambda c, p: mg.run_reference(ref.triangulation.triangulate_all, c, p), cfg, proj)
                else:
                    a = de.outcome(de.ours_associate, cfg, proj)[0]
                    t, out = de.outcome(de.ours_triangulate, cfg, proj)
                res[who] = (a, t, out)
        (ra, rt, ro), (oa, ot, oo) = res["ref"], res["ours"]
        same = ra[0] == oa[0] and rt[0] == ot[0] and sorted(ro) == sorted(oo)
        for k in (ro if same else ()):
            if k.endswith(".trc"):
                try:
                    assert_trc_equal(oo[k], ro[k], tol=1e-6)
                except AssertionError as e:
                    same = False
                    print("   ", k, str(e)[:160])
            else:
                same = same and ro[k] == oo[k]
        print(model, "SAME" if same else "DIFFERENT", "| ref:", ra[0], rt[0], len(ro), "files | ours:", oa[0], ot[0], len(oo), "files")


if __name__ == "__main__":
    main()
