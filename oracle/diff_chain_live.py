#!/usr/bin/env python
"""The two stages chained, side by side with the UNMODIFIED reference (no GPU needed): the reference's `associate_all`
writes pose-associated/ (with its `{}` entries for persons a camera does not see), then `triangulate_all` of the
reference and this package's host pipeline (oracle in place of the device call) read that same tree — single- and
multi-person.  TRC files compared.  Build-container tool (needs /root/reference).

    python oracle/diff_chain_live.py 2>&1 | grep -E " SAME | DIFFERENT "
"""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)
import diff_errors_live as de, make_golden_e2e as mg, ref_shim
from dropin_util import assert_trc_equal
from pose2sim_b200 import synth_project
ref = ref_shim.load_reference()
for multi, trial in ((True, mg.multi_association_trial()), (False, mg.association_trial())):
    calib_text, cams, kp, present = trial
    kp, present = kp[:30], present[:30]
    with tempfile.TemporaryDirectory() as td:
        proj = synth_project.write_project(os.path.join(td, "trial_demo"), calib_text, cams, kp, present=present)
        cfg = synth_project.base_config(proj, multi_person=multi)
        mg.run_reference(ref.personAssociation.associate_all, cfg, proj)          # pose-associated/ with {} entries
        res = {}
        for who in ("ref", "ours"):
            shutil.rmtree(os.path.join(proj, "pose-3d"), ignore_errors=True)
            if who == "ref":
                res[who] = de.outcome(lambda c, p: mg.run_reference(ref.triangulation.triangulate_all, c, p), cfg, proj)
            else:
                res[who] = de.outcome(de.ours_triangulate_multi if multi else de.ours_triangulate, cfg, proj)
            res[who] = (res[who][0], {k: v for k, v in res[who][1].items() if k.endswith(".trc")})
        (re_, ro), (oe, oo) = res["ref"], res["ours"]
        same = re_[0] == oe[0] and sorted(ro) == sorted(oo)
        worst = 0.0
        for k in (ro if same else ()):
            try:
                worst = max(worst, assert_trc_equal(oo[k], ro[k], tol=1e-6))
            except AssertionError as e:
                same = False; print("   ", k, str(e)[:100])
        print("chain multi_person=%s" % multi, "SAME" if same else "DIFFERENT", re_, oe, sorted(ro), sorted(oo), worst)
