"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

For every case the float32-rounded inputs AND the reference's outputs are stored, so that the
oracle restatements (oracle/p2s_oracle.py, oracle/p2s_oracle.c) and the CUDA path can be checked
against the reference itself on boxes where the reference is absent.

Reference entry points exercised (unmodified, through oracle/ref_shim.py):
  * Pose2Sim/triangulation.py:363  triangulation_from_best_cameras
  * Pose2Sim/personAssociation.py:67, :154  persons_combinations, best_persons_and_cameras_combination
  * Pose2Sim/common.py:291  computeP (demo calibration fixture)
"""
import io
import contextlib
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
from pose2sim_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def tri_config(thr, min_cams):
    return {"triangulation": {"reproj_error_threshold_triangulation": thr,
                              "min_cameras_for_triangulation": min_cams,
                              "handle_LR_swap": False, "undistort_points": False}}


def run_reference_units(ref, x, y, w, P, thr, min_cams):
    """x, y, w: [U, C] float32 (NaN = invalid).  Returns reference outputs."""
    U, C = x.shape
    cfg = tri_config(thr, min_cams)
    Plist = [P[c] for c in range(C)]
    Q = np.empty((U, 3))
    err = np.empty(U)
    nexcl = np.empty(U, np.int32)
    mask = np.zeros(U, np.uint32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for u in range(U):
            coords = np.array([x[u].astype(np.float64), y[u].astype(np.float64), w[u].astype(np.float64)])
            q, e, n, ids = ref.triangulation.triangulation_from_best_cameras(cfg, coords, coords.copy(), Plist, None)
            Q[u], err[u], nexcl[u] = np.asarray(q, float)[:3], e, n
            m = 0
            for c in np.asarray(ids).ravel():
                m |= 1 << int(c)
            mask[u] = m
    return Q, err, nexcl, mask


# ---------------------------------------------------------------------------------------------
def edge_cases():
    """SURVEY.md §8(a) edge-case table: (name, C, min_cams, thr, x, y, w)."""
    out = []
    thr = 15.0
    P4 = synth.ring_cameras(4)[0]
    P8 = synth.ring_cameras(8)[0]
    Qt = np.array([0.2, -0.1, 1.0])

    def base(P, seed):
        g = np.random.default_rng(seed)
        C = P.shape[0]
        h = P @ np.append(Qt, 1.0)
        x = h[:, 0] / h[:, 2] + g.normal(0, 1.0, C)
        y = h[:, 1] / h[:, 2] + g.normal(0, 1.0, C)
        w = g.uniform(0.6, 0.95, C)
        return x, y, w

    def add(name, P, min_cams, nan=(), zero=(), outl=(), outl_off=None, seed=5):
        x, y, w = base(P, seed)
        for i, c in enumerate(outl):
            off = outl_off[i] if outl_off is not None else (120.0 + 40.0 * i, -90.0 + 25.0 * i)
            x[c] += off[0]
            y[c] += off[1]
        for c in nan:
            x[c] = y[c] = w[c] = np.nan
        for c in zero:
            w[c] = 0.0
        out.append((name, P, min_cams, thr, x.astype(np.float32), y.astype(np.float32), w.astype(np.float32)))

    add("clean", P4, 2)
    add("cam3_nan", P4, 2, nan=(3,))
    add("cam3_nan_cam0_outlier", P4, 2, nan=(3,), outl=(0,))
    add("two_valid", P4, 2, nan=(2, 3))
    add("two_valid_outlier", P4, 2, nan=(2, 3), outl=(0,), outl_off=[(400.0, 300.0)])
    add("one_valid", P4, 2, nan=(1, 2, 3))
    add("all_nan", P4, 2, nan=(0, 1, 2, 3))
    add("cam1_zero", P4, 2, zero=(1,))
    add("cam1_zero_cam0_outlier", P4, 2, zero=(1,), outl=(0,))
    add("two_outliers", P4, 2, outl=(0, 1))
    add("three_outliers_same_dir", P4, 2, outl=(0, 1, 2), outl_off=[(150.0, 150.0)] * 3)
    add("min3_one_outlier", P4, 3, outl=(2,))
    add("min3_two_outliers", P4, 3, outl=(1, 2))
    add("min4_one_outlier", P4, 4, outl=(1,))
    add("min1_three_nan", P4, 1, nan=(1, 2, 3))
    add("c8_two_nan_one_outlier", P8, 2, nan=(6, 7), outl=(2,))
    add("c8_clean", P8, 2)
    add("c8_three_outliers", P8, 2, outl=(1, 4, 6))
    return out


def random_units(C, U, seed, p_out=0.15, p_nan=0.15, p_zero=0.04):
    P = synth.ring_cameras(C)[0]
    Q = synth.truth_points(U, 1, 1, seed)[:, 0, 0, :]
    x, y, lik = synth.observe(Q, P, seed, sigma=2.0, p_out=p_out, p_low=0.0)
    g = np.random.default_rng(seed + 1)
    nanm = g.random((U, C)) < p_nan
    zerom = (g.random((U, C)) < p_zero) & ~nanm
    x[nanm] = np.nan
    y[nanm] = np.nan
    lik[nanm] = np.nan
    lik[zerom] = 0.0
    return P, x, y, lik


def demo_calibration(ref):
    """Convert Demo_SinglePerson/calibration/Calib.qca.txt following calibration.py:107-190 / :70-104
    (stdlib xml.etree instead of lxml), write a TOML in the layout of calibration.py:1521-1533 and let
    the reference's computeP (common.py:291) build P from it."""
    import xml.etree.ElementTree as ET
    import cv2
    qca = os.path.join(ref_shim.REFERENCE_ROOT, "Pose2Sim", "Demo_SinglePerson", "calibration", "Calib.qca.txt")
    root = ET.parse(qca).getroot()
    names, S, D, K, R, T = [], [], [], [], [], []
    for cam in root.iter("camera"):
        names.append(cam.attrib.get("serial"))
    # image size, intrinsics, distortion, extrinsics
    for cam in root.iter("camera"):
        fov = cam.find("fov_video")
        w = (float(fov.attrib["right"]) - float(fov.attrib["left"]) + 1)
        h = (float(fov.attrib["bottom"]) - float(fov.attrib["top"]) + 1)
        S.append([w, h])
        intr = cam.find("intrinsic")
        fu = float(intr.attrib["focalLengthU"]) / 64.0
        fv = float(intr.attrib["focalLengthV"]) / 64.0
        cu = float(intr.attrib["centerPointU"]) / 64.0 - float(fov.attrib["left"])
        cv = float(intr.attrib["centerPointV"]) / 64.0 - float(fov.attrib["top"])
        K.append(np.array([[fu, 0.0, cu], [0.0, fv, cv], [0.0, 0.0, 1.0]]))
        D.append([float(intr.attrib["radialDistortion1"]) / 64.0, float(intr.attrib["radialDistortion2"]) / 64.0,
                  float(intr.attrib["tangentalDistortion1"]) / 64.0, float(intr.attrib["tangentalDistortion2"]) / 64.0])
        tr = cam.find("transform")
        T.append(np.array([float(tr.attrib[k]) for k in ("x", "y", "z")]) / 1000.0)
        r = np.array([float(tr.attrib[f"r{i}{j}"]) for i in (1, 2, 3) for j in (1, 2, 3)]).reshape(3, 3).T
        R.append(r)
    # world -> camera, then rotate camera by pi about x: the reference's own helpers
    # (common.py:458 world_to_camera_persp, :482 rotate_cam), as calib_qca_fun does (calibration.py:92-101)
    tables = []
    for c in range(len(names)):
        r_cam, t_cam = ref.common.world_to_camera_persp(R[c], T[c])
        r_cam, t_cam = ref.common.rotate_cam(r_cam, t_cam, ang_x=np.pi, ang_y=0, ang_z=0)
        rvec = np.array(cv2.Rodrigues(r_cam)[0]).flatten()
        tables.append((f"cam{c + 1:02d}", S[c], K[c], D[c], rvec, t_cam))
    lines = []
    for name, s, k, d, rv, t in tables:
        lines.append(f"[{name}]")
        lines.append(f'name = "{name}"')
        lines.append(f"size = [ {s[0]!r}, {s[1]!r}]")
        lines.append("matrix = [ " + ", ".join("[ " + ", ".join(repr(float(v)) for v in row) + "]" for row in k) + "]")
        lines.append("distortions = [ " + ", ".join(repr(float(v)) for v in d) + "]")
        lines.append("rotation = [ " + ", ".join(repr(float(v)) for v in rv) + "]")
        lines.append("translation = [ " + ", ".join(repr(float(v)) for v in t) + "]")
        lines.append("fisheye = false")
        lines.append("")
    lines.append("[metadata]")
    lines.append("adjusted = false")
    lines.append("error = 0.0")
    text = "\n".join(lines) + "\n"
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, "Calib_demo.toml")
    with open(path, "w") as f:
        f.write(text)
    P = np.array(ref.common.computeP(path, undistort=False))
    return path, P


def association_cases(ref, out):
    """Random association frames run through the reference (JSON on disk, as it insists)."""
    g = np.random.default_rng(77)
    kpt = 18                                            # 'Neck' id in HALPE_26 JSON order
    n_kpt_json = 26
    idx = 0
    # the last two configurations carry NaN and zero likelihoods / NaN coordinates of the tracked keypoint: such a
    # detection stays ACTIVE in the reference (:215-216) and poisons every subset that keeps its camera
    for C, min_cams, thr, F in [(3, 2, 20.0, 120), (4, 2, 20.0, 200), (4, 3, 5.0, 120), (5, 2, 20.0, 120),
                                (5, 3, 5.0, 80), (4, 2, 5.0, 120), (4, 2, 20.0, 100), (5, 1, 5.0, 60)]:
        P = synth.ring_cameras(C)[0]
        wl = synth.make_association_workload(C, F, 3, seed=404 + idx, p_out=0.15, p_low=0.0)
        obs = wl["obs"].copy()
        # likelihood range down to 0.1 so that the 0.3 gate fires; 0..3 persons per camera
        lowm = g.random(obs.shape[:3]) < 0.2
        obs[..., 2] = np.where(lowm, g.uniform(0.1, 0.3, obs.shape[:3]), obs[..., 2]).astype(np.float32)
        count = g.integers(0, 4, (F, C)).astype(np.int32)
        if idx >= 6:
            m = g.random(obs.shape[:3])
            obs[..., 2][m < 0.06] = np.nan
            obs[..., 2][(m >= 0.06) & (m < 0.10)] = 0.0
            obs[..., 0][(m >= 0.10) & (m < 0.12)] = np.nan
        cfg = {"personAssociation": {"single_person": {"reproj_error_threshold_association": thr},
                                     "likelihood_threshold_association": 0.3},
               "triangulation": {"min_cameras_for_triangulation": min_cams, "undistort_points": False}}
        errs = np.empty(F)
        combs = np.empty((F, C))
        Qs = np.empty((F, 3))
        Plist = [P[c] for c in range(C)]
        with tempfile.TemporaryDirectory() as td:
            for f in range(F):
                files = []
                for c in range(C):
                    people = []
                    for p in range(count[f, c]):
                        kp = np.zeros(n_kpt_json * 3)
                        kp[0::3] = 100.0 + p                      # non-NaN x so the person is counted
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
                errs[f] = e
                combs[f] = np.asarray(comb[0], float)
                Qs[f] = np.asarray(q[0], float)[:3]
        pre = f"assoc{idx}_"
        out[pre + "P"] = P
        out[pre + "obs"] = obs
        out[pre + "count"] = count
        out[pre + "params"] = np.array([thr, 0.3, min_cams])
        out[pre + "err"] = errs
        out[pre + "comb"] = combs
        out[pre + "Q"] = Qs
        idx += 1
        print(f"  association case {idx}: C={C} min_cams={min_cams} thr={thr} F={F}  "
              f"mean cams off {np.isnan(combs).sum(1).mean():.2f}")
    out["assoc_n"] = np.array(idx)


def main():
    ref = ref_shim.load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    if "--association-only" in sys.argv:
        out = {}
        association_cases(ref, out)
        np.savez_compressed(os.path.join(GOLDEN, "assoc_random_frames.npz"), **out)
        return

    # 1. edge-case table -----------------------------------------------------------------------
    out = {}
    names = []
    for i, (name, P, mc, thr, x, y, w) in enumerate(edge_cases()):
        Q, err, nexcl, mask = run_reference_units(ref, x[None], y[None], w[None], P, thr, mc)
        pre = f"e{i}_"
        out[pre + "P"], out[pre + "x"], out[pre + "y"], out[pre + "w"] = P, x[None], y[None], w[None]
        out[pre + "params"] = np.array([thr, mc], float)
        out[pre + "Q"], out[pre + "err"], out[pre + "nexcl"], out[pre + "mask"] = Q, err, nexcl, mask
        names.append(name)
        print(f"  edge {name:28s} Q={np.round(Q[0], 4)} err={err[0]:.4f} nexcl={nexcl[0]} mask={mask[0]:b}")
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLDEN, "tri_edge_cases.npz"), **out)

    # 2. random units ---------------------------------------------------------------------------
    out = {}
    i = 0
    for C in (3, 4, 5, 6, 8):
        for mc in (2, 3, 4):
            for thr in (5.0, 15.0, 30.0):
                U = 60 if C == 8 else 90
                P, x, y, w = random_units(C, U, seed=1000 + i)
                Q, err, nexcl, mask = run_reference_units(ref, x, y, w, P, thr, mc)
                pre = f"r{i}_"
                out[pre + "P"], out[pre + "x"], out[pre + "y"], out[pre + "w"] = P, x, y, w
                out[pre + "params"] = np.array([thr, mc], float)
                out[pre + "Q"], out[pre + "err"], out[pre + "nexcl"], out[pre + "mask"] = Q, err, nexcl, mask
                i += 1
        print(f"  random units C={C} done")
    out["n"] = np.array(i)
    np.savez_compressed(os.path.join(GOLDEN, "tri_random_units.npz"), **out)

    # 3. cfg1: shipped demo cameras x 100 frames x HALPE_26 -------------------------------------
    calib_path, P = demo_calibration(ref)
    # demo volume is metres around the origin of the Qualisys frame; use the same truth generator
    wl = synth.make_triangulation_workload(4, 100, 1, 26, seed=101, P=P, p_out=0.08, p_low=0.12)
    Q, err, nexcl, mask = run_reference_units(ref, wl["x"], wl["y"], wl["lik"], P, 15.0, 2)
    np.savez_compressed(os.path.join(GOLDEN, "tri_cfg1_demo.npz"), P=P, x=wl["x"], y=wl["y"], w=wl["lik"],
                        params=np.array([15.0, 2.0]), Q=Q, err=err, nexcl=nexcl, mask=mask)
    print(f"  cfg1 demo: {np.isfinite(err).mean() * 100:.1f}% triangulated, mean err {np.nanmean(err):.3f} px, "
          f"mean nexcl {nexcl.mean():.3f}")

    # 4. association ----------------------------------------------------------------------------
    out = {}
    association_cases(ref, out)
    np.savez_compressed(os.path.join(GOLDEN, "assoc_random_frames.npz"), **out)
    print("golden vectors written to", GOLDEN)


if __name__ == "__main__":
    main()
