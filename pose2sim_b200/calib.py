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


_PARSED = {}          # (path, mtime_ns, size) -> parsed TOML: one run reads the same file for P, the lens models and the recap


def load_calibration(calib_file):
    """The parsed calibration TOML (read-only for the callers) and its camera tables in file order."""
    try:
        st = os.stat(calib_file)
        key = (os.path.abspath(calib_file), st.st_mtime_ns, st.st_size)
    except OSError:
        key = None
    calib = _PARSED.get(key) if key is not None else None
    if calib is None:
        calib = _load(calib_file)
        if key is not None:
            if len(_PARSED) >= 16:
                _PARSED.clear()
            _PARSED[key] = calib
    keys = [k for k, v in calib.items() if k not in _NOT_CAMERAS and isinstance(v, dict)]
    return calib, keys


def undistort_normalized(u, v, K, dist, iterations=5):
    """OpenCV's fixed-point inversion of the radial/tangential lens model (cv2.undistortPoints without a
    new camera matrix): pixel -> normalised undistorted coordinates, 5 iterations, double precision."""
    k = np.zeros(8)
    k[:len(dist)] = dist
    x0 = (np.asarray(u, float) - K[0, 2]) * (1.0 / K[0, 0])
    y0 = (np.asarray(v, float) - K[1, 2]) * (1.0 / K[1, 1])
    x, y = x0.copy(), y0.copy()
    for _ in range(iterations):
        r2 = x * x + y * y
        icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
        dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x)
        dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y
        x, y = (x0 - dx) * icdist, (y0 - dy) * icdist
    return x, y


def optimal_new_camera_matrix(K, dist, size):
    """cv2.getOptimalNewCameraMatrix(K, dist, size, alpha=1, newImgSize=size)[0] (common.py:311-312,
    :279): the 9 x 9 grid of the image is undistorted to normalised coordinates and the new intrinsics
    map its circumscribed rectangle onto the image.  Agrees with OpenCV 4.13 to ~5e-13."""
    K = np.asarray(K, float)
    w, h = int(size[0]), int(size[1])
    n = 9
    gx, gy = np.meshgrid(np.arange(n) * (w - 1) / (n - 1), np.arange(n) * (h - 1) / (n - 1))
    x, y = undistort_normalized(gx.ravel(), gy.ravel(), K, np.asarray(dist, float).reshape(-1))
    fx = (w - 1) / (x.max() - x.min())
    fy = (h - 1) / (y.max() - y.min())
    return np.array([[fx, 0.0, -fx * x.min()], [0.0, fy, -fy * y.min()], [0.0, 0.0, 1.0]])


def camera_models(calib_file):
    """Per camera {K, dist, R, T, newK, size} for `undistort_points = true` (common.py:254-288)."""
    calib, keys = load_calibration(calib_file)
    out = []
    for cam in keys:
        K = np.array(calib[cam]["matrix"], dtype=np.float64)
        dist = np.array(calib[cam]["distortions"], dtype=np.float64).reshape(-1)
        size = [int(v) for v in calib[cam]["size"]]
        out.append({"K": K, "dist": dist, "R": rodrigues(calib[cam]["rotation"]),
                    "T": np.array(calib[cam]["translation"], dtype=np.float64), "size": size,
                    "newK": optimal_new_camera_matrix(K, dist, size)})
    return out


def compute_P(calib_file, undistort=False):
    """List of 3x4 float64 projection matrices, camera order = TOML order (common.py:291-324); with
    `undistort` the optimal new camera matrix replaces K (:310-313)."""
    calib, keys = load_calibration(calib_file)
    models = camera_models(calib_file) if undistort else None
    P = []
    for i, cam in enumerate(keys):
        K = models[i]["newK"] if undistort else np.array(calib[cam]["matrix"], dtype=np.float64)
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
