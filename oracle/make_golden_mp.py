"""TEST INFRASTRUCTURE — tests/golden/mp_random_frames.npz: multi-person matching of random frames by the UNMODIFIED
reference (build container only, needs /root/reference):

    python oracle/make_golden_mp.py

Reference functions exercised, in the order of Pose2Sim/personAssociation.py:789-801:
  compute_affinity (:347, with compute_rays :277), circular_constraint (:411), matchSVT (:450, the call site's
  constants), the min_affinity cut (:800), person_index_per_cam (:512).
Per case the float32-rounded inputs and the reference's thresholded affinity matrix and proposals are stored.
Cases: 3 / 4 / 5 / 8 cameras, 0-5 persons per camera, cameras without detections, undetected (NaN) joints, outlier
joints, low likelihoods, min_cameras 2 and 3, two thresholds.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from pose2sim_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    ref = ref_shim.load_reference()
    pa = ref.personAssociation
    import cv2
    out = {}
    n = 0
    rng = np.random.default_rng(20260)
    for C, NP, F, seed in ((3, 2, 30, 1), (4, 3, 40, 2), (5, 4, 30, 3), (8, 5, 24, 4), (4, 1, 10, 5), (8, 6, 12, 6)):
        w = synth.make_multi_person_workload(C, F, NP, seed=900 + seed, p_out=0.08, p_low=0.08, p_missing=0.3, p_nan=0.05)
        if seed == 2:
            w["count"][0] = 0                                   # nobody anywhere
            w["count"][1] = [2, 0, 0, 0]                        # one view only
            w["obs"][2] = np.nan                                # everybody undetected
        K = np.stack([m["K"] for m in w["models"]]); R = np.stack([m["R"] for m in w["models"]]); T = np.stack([m["T"] for m in w["models"]])
        calib = {"inv_K": [np.linalg.inv(k) for k in K], "R_mat": [r for r in R], "T": [t for t in T]}
        d_max, min_aff = (0.1, 0.2) if seed % 2 else (0.05, 0.1)
        min_cams = 2 if seed % 3 else 3
        for f in range(F):
            det = [[w["obs"][f, c, p].astype(float).tolist() for p in range(w["count"][f, c])] for c in range(C)]
            cum = np.cumsum([0] + [len(d) for d in det])
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                try:
                    aff = pa.compute_affinity(det, calib, cum, reconstruction_error_threshold=d_max)
                    circ = pa.circular_constraint(cum)
                    aff = aff * circ
                    aff = pa.matchSVT(aff, cum, circ, max_iter=20, w_rank=50, tol=1e-4, w_sparse=0.1)
                    aff[aff < min_aff] = 0
                    prop = pa.person_index_per_cam(aff, cum, min_cams)
                    failed = ""
                except Exception as e:                           # e.g. an empty frame: the reference raises
                    aff, prop, failed = np.zeros((0, 0)), np.zeros((0, C)), type(e).__name__
            p = f"m{n}_"
            out[p + "obs"] = w["obs"][f]
            out[p + "count"] = w["count"][f]
            out[p + "K"], out[p + "R"], out[p + "T"] = K, R, T
            out[p + "params"] = np.array([d_max, min_aff, min_cams], dtype=np.float64)
            out[p + "affinity"] = np.asarray(aff, dtype=np.float64)
            out[p + "proposals"] = np.asarray(prop, dtype=np.float64).reshape(-1, C) if np.size(prop) else np.zeros((0, C))
            out[p + "failed"] = np.array(failed)
            n += 1
    out["n"] = np.array(n)
    np.savez_compressed(os.path.join(GOLDEN, "mp_random_frames.npz"), **out)
    fails = sum(1 for i in range(n) if str(out[f"m{i}_failed"]))
    print(f"mp_random_frames.npz: {n} frames, {fails} on which the reference raises")


if __name__ == "__main__":
    main()
