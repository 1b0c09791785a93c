"""Calibration TOML -> projection matrices (host, once per run).

Mirrors Pose2Sim/common.py:291-324 `computeP` and :254-288 `retrieve_calib_params`: cameras are the
TOML tables other than metadata / capture_volume / charuco / checkerboard, in file order;
P = [K | 0] . [[R, T], [0, 1]] with R = Rodrigues(rotation).
"""
import glob
import os

import numpy as np

try:  # py >= 3.11
    import tomllib as _toml

    def _load(path):
        with open(path, "rb") as f:
            return _toml.load(f)
except ImportError:  # pragma: no cover
    import toml as _toml

    def _load(path):
        return _toml.load(path)

_NOT_CAMERAS = ("metadata", "capture_volume", "charuco", "checkerboard")


def rodrigues(rvec):
    """Rotation vector -> matrix (what cv2.Rodrigues computes: R = cos t I + (1-cos t) r r^T + sin t [r]x)."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    theta = float(np.linalg.norm(r))
    if theta < 2.220446049250313e-16:
        return np.eye(3)
    k = r / theta
    Kx = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    c, s = np.cos(theta), np.sin(theta)
    return c * np.eye(3) + (1.0 - c) * np.outer(k, k) + s * Kx


def load_calibration(calib_file):
    calib = _load(calib_file)
    keys = [k for k, v in calib.items() if k not in _NOT_CAMERAS and isinstance(v, dict)]
    return calib, keys


def compute_P(calib_file, undistort=False):
    """List of 3x4 float64 projection matrices, camera order = TOML order (common.py:291-324)."""
    if undistort:
        raise NotImplementedError("undistort_points is a §8(f) 'next' row: not available in the B200 path")
    calib, keys = load_calibration(calib_file)
    P = []
    for cam in keys:
        K = np.array(calib[cam]["matrix"], dtype=np.float64)
        Kh = np.hstack([K, np.zeros((3, 1))])
        R = rodrigues(calib[cam]["rotation"])
        T = np.array(calib[cam]["translation"], dtype=np.float64).reshape(3, 1)
        H = np.vstack([np.hstack([R, T]), [0.0, 0.0, 0.0, 1.0]])
        P.append(Kh @ H)
    return P


def camera_names(calib_file):
    calib, keys = load_calibration(calib_file)
    return [calib[k].get("name") or k for k in keys]


def first_camera_scale(calib_file):
    """(fm, Dm) of the recap message: focal of camera 1 and its distance to the origin
    (triangulation.py:306-308)."""
    calib, keys = load_calibration(calib_file)
    cam = calib[keys[0]]
    return float(cam["matrix"][0][0]), float(np.sqrt(np.sum(np.square(np.asarray(cam["translation"], float)))))


def find_calibration_file(session_dir):
    """triangulation.py:698-706: newest .toml (ctime) in the first directory whose name contains 'calib'."""
    try:
        calib_dir = [os.path.join(session_dir, c) for c in os.listdir(session_dir)
                     if os.path.isdir(os.path.join(session_dir, c)) and "calib" in c.lower()][0]
    except Exception:
        raise Exception("No .toml calibration direcctory found.")
    try:
        files = glob.glob(os.path.join(calib_dir, "*.toml"))
        return max(files, key=os.path.getctime)
    except Exception:
        raise Exception(f"No .toml calibration file found in the {calib_dir}.")


def session_dir_of(project_dir):
    """triangulation.py:680-682: parent directory if it holds Config.toml (batch), else the cwd."""
    parent = os.path.realpath(os.path.join(project_dir, ".."))
    return parent if "Config.toml" in os.listdir(parent) else os.getcwd()


def write_calibration_toml(path, names, sizes, Ks, dists, rvecs, tvecs):
    """Writer in the layout of calibration.py:1521-1533 (used for synthetic projects / fixtures)."""
    lines = []
    for n, s, K, d, r, t in zip(names, sizes, Ks, dists, rvecs, tvecs):
        lines += [f"[{n}]", f'name = "{n}"', f"size = [ {float(s[0])!r}, {float(s[1])!r}]",
                  "matrix = [ " + ", ".join("[ " + ", ".join(repr(float(v)) for v in row) + "]" for row in K) + "]",
                  "distortions = [ " + ", ".join(repr(float(v)) for v in d) + "]",
                  "rotation = [ " + ", ".join(repr(float(v)) for v in r) + "]",
                  "translation = [ " + ", ".join(repr(float(v)) for v in t) + "]",
                  "fisheye = false", ""]
    lines += ["[metadata]", "adjusted = false", "error = 0.0", ""]
    with open(path, "w") as f:
        f.write("\n".join(lines))


def rotation_to_rodrigues(R):
    """Inverse of `rodrigues` for fixtures (axis-angle from a rotation matrix)."""
    R = np.asarray(R, float)
    c = (np.trace(R) - 1.0) / 2.0
    theta = np.arccos(np.clip(c, -1.0, 1.0))
    if theta < 1e-12:
        return np.zeros(3)
    ax = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if np.linalg.norm(ax) < 1e-9:                    # theta ~ pi
        w, v = np.linalg.eigh((R + R.T) / 2.0)
        ax = v[:, np.argmax(w)]
        return ax / np.linalg.norm(ax) * theta
    return ax / np.linalg.norm(ax) * theta
