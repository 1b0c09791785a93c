#!/usr/bin/env python
"""Calibration handling side by side with the UNMODIFIED reference (no GPU needed): projection matrices with and
without `undistort` (`common.computeP` / cv2.getOptimalNewCameraMatrix vs `calib.compute_P`) on calibration files with
a metadata section and other non-camera tables, cameras in any key order, 4 / 5 / 8 distortion coefficients, strong
distortion, non-square pixels.  Build-container tool (needs /root/reference).

    python oracle/diff_calib_live.py | grep -E " SAME | DIFFERENT "
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, os.path.join(ROOT, "tests"), ROOT):
    sys.path.insert(0, p)

import ref_shim  # noqa: E402
from pose2sim_b200 import calib  # noqa: E402

CAM = """[{name}]
name = "{name}"
size = [ {w}, {h}]
matrix = [ [ {fx}, 0.0, {cx}], [ 0.0, {fy}, {cy}], [ 0.0, 0.0, 1.0]]
distortions = {dist}
rotation = [ {r0}, {r1}, {r2}]
translation = [ {t0}, {t1}, {t2}]
fisheye = false

"""
META = """[metadata]
adjusted = false
error = 0.0

[capture_volume]
size = [ 3.0, 3.0, 2.0]

"""


def cam(name, g, dist, w=1920.0, h=1080.0):
    return CAM.format(name=name, w=w, h=h, fx=g.uniform(900, 1800), fy=g.uniform(900, 1800), cx=w / 2 + g.uniform(-40, 40),
                      cy=h / 2 + g.uniform(-40, 40), dist=list(dist), r0=g.uniform(-2, 2), r1=g.uniform(-2, 2), r2=g.uniform(-2, 2),
                      t0=g.uniform(-3, 3), t1=g.uniform(-3, 3), t2=g.uniform(1, 5))


def files():
    g = np.random.default_rng(8)
    yield "metadata_first_4_coeffs", META + "".join(cam(f"cam_{i:02d}", g, g.normal(0, 0.02, 4)) for i in range(4))
    yield "metadata_last_5_coeffs", "".join(cam(f"int_cam{i}_img", g, g.normal(0, 0.03, 5)) for i in range(3)) + META
    yield "keys_not_sorted", "".join(cam(n, g, g.normal(0, 0.02, 4)) for n in ("camB", "cam10", "cam2", "camA"))
    yield "strong_distortion_8_coeffs", "".join(cam(f"c{i}", g, np.r_[-0.35, 0.15, 1e-3, -1e-3, -0.03, 0.01, 0.0, 0.0]) for i in range(2))
    yield "zero_distortion_portrait", "".join(cam(f"c{i}", g, np.zeros(4), w=1080.0, h=1920.0) for i in range(2))


def main():
    ref = ref_shim.load_reference()
    with tempfile.TemporaryDirectory() as td:
        for name, text in files():
            path = os.path.join(td, name + ".toml")
            open(path, "w").write(text)
            for und in (False, True):
                Pr = np.array(ref.common.computeP(path, undistort=und))
                Po = np.asarray(calib.compute_P(path, undistort=und))
                same = Pr.shape == Po.shape and np.allclose(Pr, Po, rtol=0, atol=1e-9 * max(1.0, np.abs(Pr).max()))
                print(f"computeP {name} undistort={und}", "SAME" if same else "DIFFERENT", Pr.shape, Po.shape,
                      float(np.abs(Pr - Po).max()) if Pr.shape == Po.shape else None)
            cp = ref.common.retrieve_calib_params(path)
            models = calib.camera_models(path)
            same = len(models) == len(cp["K"]) and all(
                np.allclose(m["K"], k) and np.allclose(np.asarray(m["dist"])[:len(d)], d) and np.allclose(m["newK"], ok, rtol=0, atol=1e-7)
                and np.allclose(m["R"], rm) and np.allclose(m["T"], t)
                for m, k, d, ok, rm, t in zip(models, cp["K"], cp["dist"], cp["optim_K"], cp["R_mat"], cp["T"]))
            print(f"retrieve_calib_params {name}", "SAME" if same else "DIFFERENT")


if __name__ == "__main__":
    main()
