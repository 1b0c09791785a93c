"""TEST INFRASTRUCTURE — tests/golden/assoc_wide_rigs.npz: the single-person association search of the UNMODIFIED
reference on 6 / 7 / 8-camera rigs (Pose2Sim/personAssociation.py:67 persons_combinations, :154-257
best_persons_and_cameras_combination), same key layout as assoc_random_frames.npz.

Run in the build container only (needs /root/reference; ~10 minutes — a frame whose search reaches level 2 of an
8-camera x 3-person table costs the reference 6 561 x 28 solves):

    python oracle/make_golden_wide_assoc.py

Why a separate set: assoc_random_frames.npz / assoc_six_persons.npz stop at 5 cameras.  The 8-camera instantiation of
associate_kernel (BASELINE configs[3]'s rig), its row filter and the mixed-radix row stepping over 6-8 digits are
pinned to the oracle by the GPU tests and fuzzers; this set pins the oracle and the CUDA path to the reference there.
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from make_golden import GOLDEN  # noqa: E402
from pose2sim_b200 import synth  # noqa: E402

KPT, N_KPT_JSON = 18, 26          # 'Neck' in HALPE_26 JSON order

# (cameras, persons, min_cameras, threshold px, frames, outlier rate, missing-person rate)
CASES = [
    (8, 3, 2, 20.0, 20, 0.06, 0.10),
    (8, 2, 3, 8.0, 24, 0.15, 0.15),
    (7, 3, 2, 5.0, 14, 0.10, 0.10),
    (6, 4, 4, 10.0, 14, 0.12, 0.10),
]


def reference_frame(ref, cfg, td, f, obs_f, count_f, Plist):
    files = []
    for c in range(len(Plist)):
        people = []
        for p in range(count_f[c]):
            kp = np.zeros(N_KPT_JSON * 3)
            kp[0::3] = 100.0 + p                              # a non-NaN x so that the person is counted (:81-89)
            kp[2::3] = 0.9
            kp[KPT * 3: KPT * 3 + 3] = obs_f[c, p].astype(np.float64)
            people.append({"person_id": [-1], "pose_keypoints_2d": kp.tolist()})
        fn = os.path.join(td, f"cam{c}_{f:05d}.json")
        with open(fn, "w") as js:
            json.dump({"version": 1.3, "people": people}, js)
        files.append(fn)
    rows = ref.personAssociation.persons_combinations(files)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e, comb, q = ref.personAssociation.best_persons_and_cameras_combination(cfg, files, rows, Plist, KPT, None)
    return e, np.asarray(comb[0], float), np.asarray(q[0], float)[:3]


def main():
    ref = ref_shim.load_reference()
    out = {}
    g = np.random.default_rng(808)
    for idx, (C, n_p, min_cams, thr, F, p_out, p_missing) in enumerate(CASES):
        t0 = time.time()
        wl = synth.make_association_workload(C, F, n_p, seed=860 + idx, p_out=p_out, p_low=0.1, p_missing=p_missing)
        obs, count, P = wl["obs"].copy(), wl["count"].copy(), wl["P"]
        lowm = g.random(obs.shape[:3]) < 0.08                 # likelihoods under the 0.3 gate
        obs[..., 2] = np.where(lowm, g.uniform(0.1, 0.3, obs.shape[:3]), obs[..., 2]).astype(np.float32)
        cfg = {"personAssociation": {"single_person": {"reproj_error_threshold_association": thr},
                                     "likelihood_threshold_association": 0.3},
               "triangulation": {"min_cameras_for_triangulation": min_cams, "undistort_points": False}}
        errs, combs, Qs = np.empty(F), np.empty((F, C)), np.empty((F, 3))
        Plist = [P[c] for c in range(C)]
        with tempfile.TemporaryDirectory() as td:
            for f in range(F):
                errs[f], combs[f], Qs[f] = reference_frame(ref, cfg, td, f, obs[f], count[f], Plist)
        pre = f"assoc{idx}_"
        out[pre + "P"], out[pre + "obs"], out[pre + "count"] = P, obs, count
        out[pre + "params"] = np.array([thr, 0.3, min_cams])
        out[pre + "err"], out[pre + "comb"], out[pre + "Q"] = errs, combs, Qs
        print(f"  wide-rig association case {idx}: C={C} persons={n_p} min_cams={min_cams} thr={thr} F={F} "
              f"under thr {np.mean(errs < thr):.2f}, cameras off per frame {np.isnan(combs).sum(1).tolist()}, "
              f"{time.time() - t0:.0f} s", flush=True)
    out["assoc_n"] = np.array(len(CASES))
    np.savez_compressed(os.path.join(GOLDEN, "assoc_wide_rigs.npz"), **out)


if __name__ == "__main__":
    main()
