"""Helpers shared by the drop-in tests: rebuild a golden trial on disk, parse / compare TRC files."""
import contextlib
import glob
import io
import json
import os

import numpy as np

from pose2sim_b200 import synth_project


@contextlib.contextmanager
def in_dir(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def rebuild_trial(g, tmp_path, name):
    """Golden npz (oracle/make_golden_e2e.py) -> trial directory + config_dict."""
    proj = synth_project.write_project(os.path.join(str(tmp_path), name), str(g["calib"]), [str(c) for c in g["cams"]],
                                       g["kp"], present=g["present"])
    multi = bool(g["multi_person"]) if "multi_person" in g.files else False
    extra = json.loads(str(g["extra"])) if "extra" in g.files else {}
    return proj, synth_project.base_config(proj, multi_person=multi, **extra)


def parse_trc(text):
    lines = text.splitlines()
    header = lines[:5]
    data = np.genfromtxt(io.StringIO("\n".join(lines[5:])), delimiter="\t")
    return header, np.atleast_2d(data)


def assert_trc_equal(got_text, ref_text, tol=1e-6):
    """Header lines identical; frame numbers identical; time and coordinates within tol (metres);
    empty cells (NaN) at the same places."""
    gh, gd = parse_trc(got_text)
    rh, rd = parse_trc(ref_text)
    assert gh == rh
    assert gd.shape == rd.shape
    assert np.array_equal(gd[:, 0], rd[:, 0])
    assert np.array_equal(np.isnan(gd), np.isnan(rd))
    assert np.allclose(gd, rd, atol=tol, rtol=0, equal_nan=True), float(np.nanmax(np.abs(gd - rd)))
    return float(np.nanmax(np.abs(gd[:, 2:] - rd[:, 2:]), initial=0.0))


def written_trcs(proj):
    return {os.path.basename(p): open(p).read() for p in sorted(glob.glob(os.path.join(proj, "pose-3d", "*.trc")))}


def golden_trcs(g):
    return {str(n): str(g[f"trc{i}"]) for i, n in enumerate(g["trc_names"])}


def associated_people(proj, cams, n_frames, n_values):
    """chosen[F, C, n_values] (NaN when {} / missing) and exists[F, C] of pose-associated/."""
    chosen = np.full((n_frames, len(cams), n_values), np.nan)
    exists = np.zeros((n_frames, len(cams)), bool)
    for c, cam in enumerate(cams):
        for f in range(n_frames):
            path = os.path.join(proj, "pose-associated", f"{cam}_json", f"{cam}_{f:06d}.json")
            if not os.path.exists(path):
                continue
            exists[f, c] = True
            people = json.load(open(path))["people"]
            assert len(people) == 1
            if people[0]:
                chosen[f, c] = people[0]["pose_keypoints_2d"]
    return chosen, exists


def assert_multi_person_json_equal(proj, g):
    """pose-associated/ of a multi-person trial against tests/golden/e2e_assoc_multi.npz."""
    cams = [str(c) for c in g["cams"]]
    F, C, S, V = g["chosen"].shape
    for c, cam in enumerate(cams):
        for f in range(F):
            path = os.path.join(proj, "pose-associated", f"{cam}_json", f"{cam}_{f:06d}.json")
            assert os.path.exists(path) == bool(g["exists"][f, c])
            if not g["exists"][f, c]:
                continue
            people = json.load(open(path))["people"]
            assert len(people) == int(g["n_people"][f, c]), (f, c)
            for p, person in enumerate(people):
                ref = g["chosen"][f, c, p]
                if person:
                    assert np.array_equal(np.asarray(person["pose_keypoints_2d"], np.float32), ref), (f, c, p)
                else:
                    assert np.isnan(ref).all(), (f, c, p)


def rebuild_variant(g_single, g_var, i, tmp_path):
    """Variant i of tests/golden/e2e_tri_variants.npz on the inputs of e2e_tri_single.npz."""
    proj, _ = rebuild_trial(g_single, tmp_path, "trial_demo")
    cams = [str(c) for c in g_single["cams"]]
    for c, f in g_var[f"v{i}_missing"]:
        os.remove(os.path.join(proj, "pose", f"{cams[int(c)]}_json", f"{cams[int(c)]}_{int(f):06d}.json"))
    over = json.loads(str(g_var[f"v{i}_over"]))
    prj = {k: over[k] for k in over if k in ("frame_range", "frame_rate")}
    cfg = synth_project.base_config(proj, **{k: over[k] for k in over if k not in prj and k != "pose_model"})
    cfg["project"].update(prj)
    if "pose_model" in over:
        cfg["pose"]["pose_model"] = over["pose_model"]
    return proj, cfg


def check_variant_outcome(g_var, i, proj, run):
    """Run `run()` in the trial directory and compare with what the reference did for variant i: the one TRC it wrote
    (header identical, coordinates within 1e-6 m), or nothing, and the exception it raised if any (third batch,
    tests/golden/e2e_tri_variants3.npz stores `v{i}_exc`)."""
    want_exc = str(g_var[f"v{i}_exc"]) if f"v{i}_exc" in g_var.files else ""
    got_exc = ""
    with in_dir(proj):
        try:
            run()
        except Exception as e:
            got_exc = type(e).__name__ + ": " + str(e)[:60]
    assert got_exc == want_exc, (str(g_var["names"][i]), got_exc, want_exc)
    got = written_trcs(proj)
    name = str(g_var[f"v{i}_trc_name"])
    assert list(got) == ([name] if name else []), (str(g_var["names"][i]), list(got), name)
    if name:
        assert_trc_equal(got[name], str(g_var[f"v{i}_trc"]), tol=1e-6)
