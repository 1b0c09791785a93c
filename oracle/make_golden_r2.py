"""TEST INFRASTRUCTURE — second batch of golden vectors from the UNMODIFIED reference (build container only).

    python oracle/make_golden_r2.py

  * tests/golden/assoc_six_persons.npz — the single-person association search at six persons per camera (the person
    count of BASELINE configs[3]) on 4 and 5 cameras (6^4 = 1 296 / 6^5 = 7 776 combination rows per frame), through
    persons_combinations + best_persons_and_cameras_combination (Pose2Sim/personAssociation.py:67, :154-257).
  * tests/golden/tri_wide_likelihood.npz — triangulation units whose valid likelihoods span 1e-4 ... 1 and 1e-6 ... 1
    (a likelihood threshold of 0 is a legal configuration): the reference takes the SVD of A (common.py:347-350), so a
    formulation through A^T A loses (w_max / w_min)^2 and must switch to a factorisation of A for such units.
  * tests/golden/assoc_wide_likelihood.npz — the same for the association search (likelihood_threshold_association 0).
"""
import contextlib
import io
import json
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
from make_golden import GOLDEN, run_reference_units  # noqa: E402
from pose2sim_b200 import synth  # noqa: E402


def association_six(ref):
    out = {}
    kpt, n_kpt_json = 18, 26
    g = np.random.default_rng(606)
    idx = 0
    for C, n_p, min_cams, thr, F, p_out, p_missing in [(4, 6, 2, 20.0, 36, 0.10, 0.0), (4, 6, 3, 8.0, 24, 0.15, 0.15),
                                                       (5, 6, 2, 20.0, 10, 0.10, 0.1), (4, 5, 2, 3.0, 24, 0.2, 0.1)]:
        wl = synth.make_association_workload(C, F, n_p, seed=640 + idx, p_out=p_out, p_low=0.1, p_missing=p_missing)
        obs, count, P = wl["obs"].copy(), wl["count"].copy(), wl["P"]
        lowm = g.random(obs.shape[:3]) < 0.1
        obs[..., 2] = np.where(lowm, g.uniform(0.1, 0.3, obs.shape[:3]), obs[..., 2]).astype(np.float32)
        cfg = {"personAssociation": {"single_person": {"reproj_error_threshold_association": thr},
                                     "likelihood_threshold_association": 0.3},
               "triangulation": {"min_cameras_for_triangulation": min_cams, "undistort_points": False}}
        errs, combs, Qs = np.empty(F), np.empty((F, C)), np.empty((F, 3))
        Plist = [P[c] for c in range(C)]
        with tempfile.TemporaryDirectory() as td:
            for f in range(F):
                files = []
                for c in range(C):
                    people = []
                    for p in range(count[f, c]):
                        kp = np.zeros(n_kpt_json * 3)
                        kp[0::3] = 100.0 + p
                        kp[2::3] = 0.9
                        kp[kpt * 3: kpt * 3 + 3] = obs[f, c, p].astype(np.float64)
                        people.append({"person_id": [-1], "pose_keypoints_2d": kp.tolist()})
                    fn = os.path.join(td, f"cam{c}_{f:05d}.json")
                    with open(fn, "w") as js:
                        json.dump({"version": 1.3, "people": people}, js)
                    files.append(fn)
                rows = ref.personAssociation.persons_combinations(files)
                with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    e, comb, q = ref.personAssociation.best_persons_and_cameras_combination(cfg, files, rows, Plist, kpt, None)
                errs[f], combs[f], Qs[f] = e, np.asarray(comb[0], float), np.asarray(q[0], float)[:3]
        pre = f"assoc{idx}_"
        out[pre + "P"], out[pre + "obs"], out[pre + "count"] = P, obs, count
        out[pre + "params"] = np.array([thr, 0.3, min_cams])
        out[pre + "err"], out[pre + "comb"], out[pre + "Q"] = errs, combs, Qs
        idx += 1
        print(f"  six-person association case {idx}: C={C} persons={n_p} min_cams={min_cams} thr={thr} F={F} "
              f"under thr {np.mean(errs < thr):.2f} mean cams off {np.isnan(combs).sum(1).mean():.2f}", flush=True)
    out["assoc_n"] = np.array(idx)
    np.savez_compressed(os.path.join(GOLDEN, "assoc_six_persons.npz"), **out)


def association_wide(ref):
    """Association frames whose detections' likelihoods span up to 1e6 (likelihood threshold 0): the search must solve
    from a factorisation of A there as well."""
    import make_golden as mg
    out = {}
    kpt, n_kpt_json = 18, 26
    idx = 0
    for C, n_p, min_cams, thr, F in [(4, 3, 2, 20.0, 60), (5, 2, 3, 10.0, 40)]:
        wl = synth.make_association_workload(C, F, n_p, seed=690 + idx, p_out=0.1, p_low=0.0, p_missing=0.1)
        obs, count, P = wl["obs"].copy(), wl["count"].copy(), wl["P"]
        g = np.random.default_rng(691 + idx)
        obs[..., 2] = (10.0 ** g.uniform(-6.0, 0.0, obs.shape[:3])).astype(np.float32)
        keep = g.random(obs.shape[:3]) < 0.4
        obs[..., 2] = np.where(keep, g.uniform(0.6, 1.0, obs.shape[:3]), obs[..., 2]).astype(np.float32)
        cfg = {"personAssociation": {"single_person": {"reproj_error_threshold_association": thr},
                                     "likelihood_threshold_association": 0.0},
               "triangulation": {"min_cameras_for_triangulation": min_cams, "undistort_points": False}}
        errs, combs, Qs = np.empty(F), np.empty((F, C)), np.empty((F, 3))
        Plist = [P[c] for c in range(C)]
        with tempfile.TemporaryDirectory() as td:
            for f in range(F):
                files = []
                for c in range(C):
                    people = []
                    for p in range(count[f, c]):
                        kp = np.zeros(n_kpt_json * 3)
                        kp[0::3] = 100.0 + p
                        kp[2::3] = 0.9
                        kp[kpt * 3: kpt * 3 + 3] = obs[f, c, p].astype(np.float64)
                        people.append({"person_id": [-1], "pose_keypoints_2d": kp.tolist()})
                    fn = os.path.join(td, f"cam{c}_{f:05d}.json")
                    with open(fn, "w") as js:
                        json.dump({"version": 1.3, "people": people}, js)
                    files.append(fn)
                rows = ref.personAssociation.persons_combinations(files)
                with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    e, comb, q = ref.personAssociation.best_persons_and_cameras_combination(cfg, files, rows, Plist, kpt, None)
                errs[f], combs[f], Qs[f] = e, np.asarray(comb[0], float), np.asarray(q[0], float)[:3]
        pre = f"assoc{idx}_"
        out[pre + "P"], out[pre + "obs"], out[pre + "count"] = P, obs, count
        out[pre + "params"] = np.array([thr, 0.0, min_cams])
        out[pre + "err"], out[pre + "comb"], out[pre + "Q"] = errs, combs, Qs
        idx += 1
        print(f"  wide-likelihood association case {idx}: C={C} persons={n_p} F={F} under thr {np.mean(errs < thr):.2f}", flush=True)
    out["assoc_n"] = np.array(idx)
    np.savez_compressed(os.path.join(GOLDEN, "assoc_wide_likelihood.npz"), **out)


def wide_likelihood(ref):
    """Likelihoods log-uniform in [1e-4, 1]: no gate (threshold 0), so the weights of one unit span up to 1e4."""
    out = {}
    i = 0
    # the last three configurations go down to 1e-6 (a float32 likelihood can be that small): there A^T A is off by
    # 1e-5 m and more, the factorisation of A is not
    for C, mc, thr, U, lo in [(3, 2, 15.0, 150, -4.0), (4, 2, 15.0, 200, -4.0), (5, 3, 10.0, 150, -4.0), (8, 2, 15.0, 150, -4.0),
                              (8, 4, 30.0, 100, -4.0), (12, 8, 15.0, 60, -4.0), (16, 13, 15.0, 40, -4.0),
                              (4, 2, 15.0, 200, -6.0), (8, 2, 15.0, 150, -6.0), (6, 3, 20.0, 120, -6.0)]:
        seed = 7000 + i
        P = synth.ring_cameras(C)[0]
        Q = synth.truth_points(U, 1, 1, seed)[:, 0, 0, :]
        x, y, _ = synth.observe(Q, P, seed, sigma=1.0, p_out=0.06, p_low=0.0)
        g = np.random.default_rng(seed)
        lik = (10.0 ** g.uniform(lo, 0.0, (U, C))).astype(np.float32)
        # a third of the units: ONE or TWO confident cameras, the rest barely above zero (worst case for A^T A)
        hard = g.random(U) < 0.35
        few = g.integers(1, 3, U)
        for u in np.flatnonzero(hard):
            lik[u] = (10.0 ** g.uniform(lo, lo + 1.0, C)).astype(np.float32)
            lik[u, g.choice(C, few[u], replace=False)] = g.uniform(0.6, 1.0, few[u]).astype(np.float32)
        lik[g.random((U, C)) < 0.03] = 0.0
        Qr, err, nexcl, mask = run_reference_units(ref, x, y, lik, P, thr, mc)
        pre = f"r{i}_"
        out[pre + "P"], out[pre + "x"], out[pre + "y"], out[pre + "w"] = P, x, y, lik
        out[pre + "params"] = np.array([thr, mc], float)
        out[pre + "Q"], out[pre + "err"], out[pre + "nexcl"], out[pre + "mask"] = Qr, err, nexcl, mask
        print(f"  wide-likelihood units C={C} min_cams={mc}: {np.isfinite(err).mean() * 100:.0f}% triangulated", flush=True)
        i += 1
    out["n"] = np.array(i)
    np.savez_compressed(os.path.join(GOLDEN, "tri_wide_likelihood.npz"), **out)


if __name__ == "__main__":
    ref = ref_shim.load_reference()
    if "--assoc-wide-only" in sys.argv:
        association_wide(ref)
        sys.exit(0)
    if "--wide-only" not in sys.argv:
        association_six(ref)
        association_wide(ref)
    if "--assoc-only" not in sys.argv:
        wide_likelihood(ref)
