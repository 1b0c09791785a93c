#!/usr/bin/env python
"""End-to-end timing of the drop-in `triangulate_all(config)` on synthetic trials in the reference's project
layout (JSON directories in, TRC out) — BASELINE.json configs[0] shape (4 shipped Qualisys cameras x 100
frames x HALPE_26) and a larger 8-camera trial.  Phases: host staging (native JSON reader), device call,
host post-processing + TRC writer.

    python tests/perf/dropin_bench.py [--reference]     # --reference: time the UNMODIFIED reference instead
                                                   #   (build container only; needs /root/reference)
Lines go to stdout and gpurun_out/dropin_bench.jsonl."""
import json
import logging
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def make_trial(td, name, C, F, seed, demo_calib):
    from pose2sim_b200 import skeletons, synth, synth_project
    ids, _ = skeletons.keypoints("HALPE_26")
    if demo_calib:
        calib_text = open(os.path.join(ROOT, "tests", "golden", "Calib_demo.toml")).read()
        P = np.load(os.path.join(ROOT, "tests", "golden", "tri_cfg1_demo.npz"))["P"]
        cams = [f"cam{c + 1:02d}" for c in range(C)]
    else:
        calib_text, cams, P = synth_project.ring_calibration_toml(C)
    wl = synth.make_triangulation_workload(C, F, 1, 26, seed=seed, P=P, lik_thr=None)
    x, y, lik = (wl[k].reshape(F, 1, 26, C).transpose(0, 3, 1, 2) for k in ("x", "y", "lik"))
    kp = synth_project.pack_openpose(x, y, lik, ids, 26)
    proj = synth_project.write_project(os.path.join(td, name), calib_text, cams, kp)
    return proj, synth_project.base_config(proj)


def make_assoc_trial(td, name, C, F, persons, seed):
    """Every camera lists `persons` people in a random order, all 26 keypoints each (JSON id order)."""
    from pose2sim_b200 import synth, synth_project
    w = synth.make_multi_person_workload(C, F, persons, seed=seed, p_missing=0.1, p_nan=0.0)
    calib_text, cams, _ = synth_project.ring_calibration_toml(C)
    present = np.arange(persons)[None, None, :] < w["count"][:, :, None]
    proj = synth_project.write_project(os.path.join(td, name), calib_text, cams, w["obs"], present=present)
    return proj


def bench_association(td, out):
    """`associate_all(config)` wall time, single-person (ordered combination search) and multi-person (ray affinity +
    SVT matching) modes, on an 8-camera x 3-person x 1000-frame trial on disk (8000 JSON files in, 8000 out)."""
    import shutil
    from pose2sim_b200 import ops, personAssociation as pa, synth_project
    C, F, NP = 8, 1000, 3
    proj = make_assoc_trial(td, "assoc_8cams_3persons_1000frames", C, F, NP, 404)
    ops.get_engine(0)
    for multi in (False, True):
        cfg = synth_project.base_config(proj, multi_person=multi)
        os.chdir(proj)
        for rep in range(2):                                   # second pass is the timed one
            shutil.rmtree(os.path.join(proj, "pose-associated"), ignore_errors=True)
            t0 = time.perf_counter()
            st = pa.stage_project(cfg)
            t1 = time.perf_counter()
            if multi:
                res = pa.solve_frames_multi_person(st)
                t2 = time.perf_counter()
                pa.write_outputs_multi_person(st, res)
            else:
                res = pa.solve_frames(st)
                t2 = time.perf_counter()
                pa.write_outputs(st, res)
            t3 = time.perf_counter()
        line = {"bench": "dropin_associate_all", "trial": os.path.basename(proj), "multi_person": multi, "cams": C, "frames": F,
                "persons_per_cam": NP, "json_files": C * F, "impl": "pose2sim_b200", "stage_s": t1 - t0, "device_call_s": t2 - t1,
                "write_s": t3 - t2, "total_s": t3 - t0, "frames_per_s": F / (t3 - t0)}
        print(json.dumps(line), flush=True)
        out.append(line)
        os.chdir(ROOT)


def main():
    use_ref = "--reference" in sys.argv
    logging.getLogger().setLevel(logging.ERROR)
    out = []
    with tempfile.TemporaryDirectory() as td:
        for name, C, F, seed, demo in (("cfg1_demo_4cams_100frames", 4, 100, 101, True), ("ring_8cams_5000frames", 8, 5000, 202, False)):
            if use_ref and F > 1000:
                F = 300                                       # the reference needs ~25 ms per frame and camera
            proj, cfg = make_trial(td, name, C, F, seed, demo)
            os.chdir(proj)
            line = {"bench": "dropin_triangulate_all", "trial": name, "cams": C, "frames": F, "units": F * 26,
                    "json_files": C * F, "cores": len(os.sched_getaffinity(0))}
            if use_ref:
                import contextlib
                import io
                import warnings
                import ref_shim
                ref = ref_shim.load_reference()
                t0 = time.perf_counter()
                with warnings.catch_warnings(), contextlib.redirect_stderr(io.StringIO()):
                    warnings.simplefilter("ignore")
                    ref.triangulation.triangulate_all(cfg)
                line.update(impl="reference", total_s=time.perf_counter() - t0)
            else:
                from pose2sim_b200 import ops, triangulation as tri
                ops.get_engine(0)                             # context creation is not part of the stage
                os.environ["P2S_STAGE_CACHE"] = "0"
                tri.triangulate_all(cfg)                      # warm-up (library load, allocations), staging cache off
                tc = time.perf_counter()
                tri.stage_project(cfg)                        # every JSON parsed (cold staging, page cache warm)
                line["stage_uncached_s"] = time.perf_counter() - tc
                os.environ["P2S_STAGE_CACHE"] = "1"
                tri.stage_project(cfg)                        # fills the staging cache (large trials only)
                t0 = time.perf_counter()
                st = tri.stage_project(cfg)                   # the timed pass: listing + stat signature + mmap of the cache entry
                t1 = time.perf_counter()
                res = tri.solve_units(st)
                t2 = time.perf_counter()
                tri.write_outputs(st, res)
                t3 = time.perf_counter()
                line.update(impl="pose2sim_b200", stage_s=t1 - t0, device_call_s=t2 - t1, post_s=t3 - t2, total_s=t3 - t0,
                            staged_from="cache (mmap)" if isinstance(st.x, np.memmap) else "JSON",
                            total_uncached_s=line["stage_uncached_s"] + (t3 - t1),
                            units_per_s_uncached=line["units"] / (line["stage_uncached_s"] + (t3 - t1)),
                            level_hist=res["stats"]["level_hist"])
            line["units_per_s"] = line["units"] / line["total_s"]
            print(json.dumps(line), flush=True)
            out.append(line)
            os.chdir(ROOT)
        if not use_ref:
            bench_association(td, out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dropin_bench.jsonl"), "a") as f:
        for line in out:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
