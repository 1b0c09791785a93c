"""Synthetic on-disk trials in the reference's project layout (inputs only, no results):

    <project>/Config.toml                     (marker of a single-trial session)
    <project>/calibration/Calib.toml          (calibration.py:1521-1533 layout)
    <project>/pose/<cam>_json/<cam>_<frame>.json   OpenPose 1.3 (poseEstimation.py:260-273)

Used by the end-to-end parity fixtures (oracle/make_golden_e2e.py, tests/) and by the drop-in
timing in bench.py.  Keypoint values are float32-exact, like the files Pose2Sim's pose stage writes.
"""
import json
import os
import shutil

import numpy as np


def base_config(project_dir, pose_model="HALPE_26", multi_person=False, frame_range="all", frame_rate=60, **tri):
    """A config_dict with every key the two stages read (Demo_SinglePerson/Config.toml values)."""
    t = {"reproj_error_threshold_triangulation": 15, "likelihood_threshold_triangulation": 0.3,
         "min_cameras_for_triangulation": 2, "max_distance_m": 1.0, "interp_if_gap_smaller_than": 20,
         "interpolation": "linear", "remove_incomplete_frames": False, "sections_to_keep": "all",
         "min_chunk_size": 10, "fill_large_gaps_with": "last_value", "show_interp_indices": True,
         "make_c3d": False}
    t.update(tri)
    return {"project": {"project_dir": project_dir, "multi_person": multi_person, "frame_range": frame_range,
                        "frame_rate": frame_rate},
            "pose": {"pose_model": pose_model, "vid_img_extension": "mp4"},
            "personAssociation": {"likelihood_threshold_association": 0.3,
                                  "single_person": {"likelihood_threshold_association": 0.3,
                                                    "reproj_error_threshold_association": 20, "tracked_keypoint": "Neck"},
                                  "multi_person": {"reconstruction_error_threshold": 0.1, "min_affinity": 0.2}},
            "triangulation": t}


def write_pose_dir(pose_root, cam_names, keypoints, frames=None, present=None):
    """keypoints: float array [F, C, Npeople, 3*J] (x, y, likelihood interleaved, OpenPose order);
    present[F, C, Npeople] (optional bool): people that exist in the file (others are omitted);
    a person whose values are all NaN is written as a person with NaN values (json NaN literal)."""
    F, C, Np, _ = keypoints.shape
    frames = list(range(F)) if frames is None else list(frames)
    for c, cam in enumerate(cam_names):
        d = os.path.join(pose_root, f"{cam}_json")
        os.makedirs(d, exist_ok=True)
        for fi, f in enumerate(frames):
            people = []
            for p in range(Np):
                if present is not None and not present[fi, c, p]:
                    continue
                people.append({"person_id": [-1], "pose_keypoints_2d": [float(v) for v in keypoints[fi, c, p]],
                               "face_keypoints_2d": [], "hand_left_keypoints_2d": [], "hand_right_keypoints_2d": [],
                               "pose_keypoints_3d": [], "face_keypoints_3d": [], "hand_left_keypoints_3d": [],
                               "hand_right_keypoints_3d": []})
            with open(os.path.join(d, f"{cam}_{f:06d}.json"), "w") as js:
                json.dump({"version": 1.3, "people": people}, js)


def write_project(project_dir, calib_toml_text, cam_names, keypoints, frames=None, present=None, pose_subdir="pose"):
    """Create the trial directory (removing a previous one) and return its path."""
    if os.path.exists(project_dir):
        shutil.rmtree(project_dir)
    os.makedirs(os.path.join(project_dir, "calibration"))
    with open(os.path.join(project_dir, "Config.toml"), "w") as f:
        f.write("# synthetic trial\n")
    with open(os.path.join(project_dir, "calibration", "Calib.toml"), "w") as f:
        f.write(calib_toml_text)
    write_pose_dir(os.path.join(project_dir, pose_subdir), cam_names, keypoints, frames, present)
    if pose_subdir != "pose":                      # camera directories are always discovered under pose/
        for cam in cam_names:
            os.makedirs(os.path.join(project_dir, "pose", f"{cam}_json"), exist_ok=True)
            # the reference probes the first camera directory for at least one entry (triangulation.py:754)
        open(os.path.join(project_dir, "pose", f"{cam_names[0]}_json", "placeholder.txt"), "w").close()
    return project_dir


def pack_openpose(x, y, lik, keypoint_ids, n_json_keypoints):
    """x, y, lik [..., K] in skeleton order -> [..., 3*J] in JSON (id) order; unused ids are zeros."""
    out = np.zeros(x.shape[:-1] + (3 * n_json_keypoints,), np.float32)
    ids = np.asarray(keypoint_ids)
    out[..., 3 * ids] = x
    out[..., 3 * ids + 1] = y
    out[..., 3 * ids + 2] = lik
    return out


def ring_calibration_toml(C, **kw):
    """TOML text + P for `synth.ring_cameras(C)` (Rodrigues vectors from the rotation matrices)."""
    import io
    import tempfile
    from . import calib, synth
    P, Ks, Rs, ts = synth.ring_cameras(C, **kw)
    names = [f"cam{c + 1:02d}" for c in range(C)]
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c.toml")
        calib.write_calibration_toml(path, names, [(1080.0, 1920.0)] * C, Ks, [[0.0] * 4] * C,
                                     [calib.rotation_to_rodrigues(R) for R in Rs], ts)
        text = open(path).read()
    return text, names, P
