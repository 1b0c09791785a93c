#!/usr/bin/env python
"""Log output side by side with the UNMODIFIED reference (no GPU needed): every INFO / WARNING line the two stages emit
(recap of errors in px and mm, excluded cameras, interpolated / non-interpolated spans per keypoint, trimming warnings,
paths with the temporary directory masked) for the single- and multi-person trials, reference against this package's
host pipelines with the oracle in place of the device call.  Build-container tool (needs /root/reference).

    python oracle/diff_logs_live.py 2>&1 | grep -E "^(tri|assoc)|^   (REF|OURS)"
"""
import inspect
import io
import logging
import os
import re
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)

import diff_errors_live as de  # noqa: E402
import make_golden_e2e as mg  # noqa: E402
import ref_shim  # noqa: E402
from pose2sim_b200 import synth_project  # noqa: E402

for _name in ("ours_associate", "ours_associate_multi"):          # the same drivers with the recap switched on
    _src = inspect.getsource(getattr(de, _name)).replace("log=False", "log=True").replace(f"def {_name}", f"def {_name}_log")
    exec(compile(_src, _name, "exec"), de.__dict__)


def capture(fn, cfg, proj):
    buf = io.StringIO()
    h = logging.StreamHandler(buf)
    h.setFormatter(logging.Formatter("%(levelname)s|%(message)s"))
    root = logging.getLogger()
    old = root.level
    root.setLevel(logging.INFO)
    root.addHandler(h)
    try:
        fn(cfg, proj)
    finally:
        root.removeHandler(h)
        root.setLevel(old)
    return [re.sub(r"/tmp/\S+", "<path>", line).rstrip() for line in buf.getvalue().splitlines() if line.strip()]


def main():
    ref = ref_shim.load_reference()
    tri_ref = lambda c, p: ref.triangulation.triangulate_all(c)              # noqa: E731
    assoc_ref = lambda c, p: ref.personAssociation.associate_all(c)          # noqa: E731
    for label, trial, multi, ours_fn, ref_fn in (
            ("tri single", mg.single_person_trial(), False, de.ours_triangulate, tri_ref),
            ("tri multi", mg.multi_person_trial(), True, de.ours_triangulate_multi, tri_ref),
            ("assoc single", mg.association_trial(), False, de.ours_associate_log, assoc_ref),
            ("assoc multi", mg.multi_association_trial(), True, de.ours_associate_multi_log, assoc_ref)):
        calib_text, cams, kp, present = trial
        kp, present = kp[:40], (present[:40] if present is not None else None)
        logs = {}
        for who, fn in (("ref", ref_fn), ("ours", ours_fn)):
            with tempfile.TemporaryDirectory() as td:
                proj = synth_project.write_project(os.path.join(td, "trial_demo"), calib_text, cams, kp, present=present)
                cfg = synth_project.base_config(proj, multi_person=multi)
                if who == "ref":
                    with mg.in_dir(proj):
                        logs[who] = capture(fn, cfg, proj)
                else:
                    logs[who] = capture(fn, cfg, proj)
        r, o = logs["ref"], logs["ours"]
        only_r, only_o = [line for line in r if line not in o], [line for line in o if line not in r]
        print(label, "log lines ref / ours:", len(r), "/", len(o), "| only in ref:", len(only_r), "| only in ours:", len(only_o),
              "| same order:", r == o)
        for line in only_r[:8]:
            print("   REF :", line[:170])
        for line in only_o[:8]:
            print("   OURS:", line[:170])


if __name__ == "__main__":
    main()
