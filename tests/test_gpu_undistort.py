"""`[triangulation] undistort_points = true` on the device: the stage kernel's lens inversion against
cv2.undistortPoints (values stored by the golden generator, and cv2 itself when importable), and the
search with distorted re-projection against the reference's per-unit outputs."""
import io
import tomllib
import warnings

import numpy as np
import pytest

import p2s_oracle as orc
from pose2sim_b200 import calib

pytestmark = pytest.mark.gpu


def _lens(g):
    toml = tomllib.load(io.BytesIO(str(g["calib"]).encode()))
    out = []
    for name in [str(c) for c in g["cams"]]:
        cam = toml[name]
        K = np.array(cam["matrix"], float)
        out.append({"K": K, "dist": np.array(cam["distortions"], float), "R": calib.rodrigues(cam["rotation"]),
                    "T": np.array(cam["translation"], float), "size": cam["size"],
                    "newK": calib.optimal_new_camera_matrix(K, cam["distortions"], cam["size"])})
    return out


def test_stage_kernel_undistorts_like_cv2(engine, golden):
    import torch
    g = golden("e2e_tri_undistort.npz")
    lens = _lens(g)
    x, y, lik = (torch.from_numpy(g[k]).cuda() for k in ("unit_x", "unit_y", "unit_lik"))
    obs = engine.stage_observations(x, y, lik, None, lens=lens).cpu().numpy()          # [C, U, 4]
    assert np.array_equal(obs[:, :, 0].T, g["unit_ux"])                                  # bit-exact float32
    assert np.array_equal(obs[:, :, 1].T, g["unit_uy"])
    assert np.array_equal(obs[:, :, 2].T, g["unit_lik"])
    # a larger random set, against cv2 directly when it is importable on this box
    rng = np.random.default_rng(3)
    U = 200_000
    rx = rng.uniform(-50, 1130, (U, 4)).astype(np.float32)
    ry = rng.uniform(-50, 1970, (U, 4)).astype(np.float32)
    rl = np.ones((U, 4), np.float32)
    obs = engine.stage_observations(*(torch.from_numpy(a).cuda() for a in (rx, ry, rl)), None, lens=lens).cpu().numpy()
    for c, L in enumerate(lens):
        ex, ey = orc.undistort_points(rx[:, c], ry[:, c], L["K"], L["dist"], L["newK"])
        assert np.array_equal(obs[c, :, 0], ex) and np.array_equal(obs[c, :, 1], ey)
        try:
            import cv2
        except ImportError:
            continue
        pts = np.stack([rx[:, c], ry[:, c]], 1).reshape(-1, 1, 2)
        u = cv2.undistortPoints(pts, L["K"], L["dist"], None, L["newK"]).reshape(-1, 2)
        assert np.array_equal(obs[c, :, 0], u[:, 0]) and np.array_equal(obs[c, :, 1], u[:, 1])


def test_undistort_mode_units_match_reference(engine, golden):
    g = golden("e2e_tri_undistort.npz")
    lens = _lens(g)
    out = engine.triangulate_host(g["unit_x"], g["unit_y"], g["unit_lik"], g["unit_P"], 0.3, 15.0, 2, lens=lens)
    assert np.array_equal(out["nexcl"], g["unit_nexcl"].astype(np.uint8))
    assert np.array_equal(out["mask"], g["unit_mask"])
    assert np.array_equal(np.isnan(out["err"]), np.isnan(g["unit_err"]))
    assert np.allclose(out["Q"], g["unit_Q"], atol=1e-6, rtol=0, equal_nan=True)
    assert np.allclose(out["err"], g["unit_err"], atol=1e-6, rtol=0, equal_nan=True)
    assert float(np.nanmax(np.abs(out["Q"] - g["unit_Q"]))) < 1e-9
