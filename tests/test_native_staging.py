"""The native JSON reader (csrc/p2s_json.cpp, host threads, no GPU) against Python's json.load on the
golden trials and on hostile files: the value for (person, keypoint) must be exactly what
`js['people'][n]['pose_keypoints_2d'][3 id : 3 id + 3]` gives, NaN wherever that lookup would raise."""
import json
import os

import numpy as np
import pytest

from dropin_util import rebuild_trial
from pose2sim_b200 import skeletons, staging


def _both(input_dir, cam_dirs, files, f_range, ids, n_persons):
    x, y, lik, inexact = staging.stage_triangulation(input_dir, cam_dirs, files, f_range, ids, n_persons)
    px, py, pl = staging.stage_triangulation_python(input_dir, cam_dirs, files, f_range, ids, n_persons)
    return (x, y, lik, inexact), (px, py, pl)


@pytest.mark.parametrize("tag,n_persons", [("e2e_tri_single", 1), ("e2e_tri_multi", 3), ("e2e_assoc_single", 3)])
def test_native_reader_equals_python_on_golden_trials(golden, tmp_path, tag, n_persons):
    g = golden(tag + ".npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial")
    dirs = staging.PoseDirs(proj)
    cam_dirs = dirs.camera_dirs()
    input_dir, files = dirs.files_for_triangulation(cam_dirs)
    ids, _ = skeletons.keypoints("HALPE_26")
    F = min(len(f) for f in files)
    (x, y, lik, inexact), (px, py, pl) = _both(input_dir, cam_dirs, files, [0, F], ids, n_persons)
    assert inexact == 0 == staging.float32_inexact(px, py, pl)
    for a, b in ((x, px), (y, py), (lik, pl)):
        assert a.dtype == np.float32 and a.shape == b.shape
        assert np.array_equal(a, b.astype(np.float32), equal_nan=True)
    assert staging.count_persons(input_dir, cam_dirs, files) == max(
        len(json.load(open(os.path.join(input_dir, d, n)))["people"]) for d, fl in zip(cam_dirs, files) for n in fl)


HOSTILE = {
    "ok.json": '{"version": 1.3, "people": [{"person_id": [-1], "pose_keypoints_2d": [1.5, 2.25, 0.5, 3, 4, 1e0, -7.125, 8E1, 0.25]}]}',
    "nan_literals.json": '{"people": [{"pose_keypoints_2d": [NaN, 2.0, 0.5, Infinity, -Infinity, 0.9, 1, 2, 3]}]}',
    "null_bool.json": '{"people": [{"pose_keypoints_2d": [null, true, false, 1, 2, 3, 4, 5, 6]}]}',
    "short_list.json": '{"people": [{"pose_keypoints_2d": [1, 2, 3, 4]}]}',
    "empty_person.json": '{"people": [{}, {"pose_keypoints_2d": [9, 8, 7, 6, 5, 4, 3, 2, 1]}]}',
    "no_people.json": '{"version": 1.3}',
    "people_not_list.json": '{"people": 3}',
    "person_not_object.json": '{"people": [5, {"pose_keypoints_2d": [1, 2, 3, 4, 5, 6, 7, 8, 9]}]}',
    "kp_not_list.json": '{"people": [{"pose_keypoints_2d": "abc"}]}',
    "string_entry.json": '{"people": [{"pose_keypoints_2d": [1, "x", 3, 4, 5, 6, 7, 8, 9]}]}',
    "nested_entry.json": '{"people": [{"pose_keypoints_2d": [1, [2], 3, 4, 5, 6, 7, 8, 9]}]}',
    "duplicate_keys.json": '{"people": [], "people": [{"pose_keypoints_2d": [0], "pose_keypoints_2d": [1, 2, 3, 4, 5, 6, 7, 8, 9]}]}',
    "escapes.json": '{"a\\"b\\u00e9": "q\\\\\\n", "people": [{"x": {"y": [1, {"z": null}]}, "pose_keypoints_2d": [1, 2, 3, 4, 5, 6, 7, 8, 9]}]}',
    "whitespace.json": ' \n\t{ "people" :\r\n [ { "pose_keypoints_2d" : [ 1 , 2 , 3 , 4 , 5 , 6 , 7 , 8 , 9 ] } ] } \n',
    "inexact.json": '{"people": [{"pose_keypoints_2d": [0.1, 0.2, 0.3, 16777217, 5, 6, 7, 8, 9]}]}',
    "trailing_garbage.json": '{"people": [{"pose_keypoints_2d": [1, 2, 3, 4, 5, 6, 7, 8, 9]}]} x',
    "truncated.json": '{"people": [{"pose_keypoints_2d": [1, 2, 3, 4, 5',
    "bad_number.json": '{"people": [{"pose_keypoints_2d": [01, 2, 3, 4, 5, 6, 7, 8, 9]}]}',
    "bad_number2.json": '{"people": [{"pose_keypoints_2d": [1., 2, 3, 4, 5, 6, 7, 8, 9]}]}',
    "trailing_comma.json": '{"people": [{"pose_keypoints_2d": [1, 2, 3, 4, 5, 6, 7, 8, 9,]}]}',
    "single_quotes.json": "{'people': []}",
    "empty.json": '',
    "top_scalar.json": '42',
    "top_array.json": '[{"people": [{"pose_keypoints_2d": [1, 2, 3, 4, 5, 6, 7, 8, 9]}]}]',
    "control_char.json": '{"a": "x\ty", "people": [{"pose_keypoints_2d": [1, 2, 3, 4, 5, 6, 7, 8, 9]}]}',
    "big_numbers.json": '{"people": [{"pose_keypoints_2d": [1e400, -1e400, 123456789012345678901234567890, 4e-400, 5, 6, 7, 8, 9]}]}',
}


def test_native_reader_on_hostile_files(tmp_path):
    cam = tmp_path / "cam1_json"
    cam.mkdir()
    names = sorted(HOSTILE)
    for i, name in enumerate(names):
        (cam / f"f_{i:03d}.json").write_text(HOSTILE[name])
    files = [[f"f_{i:03d}.json" for i in range(len(names))]]
    ids = [2, 0, 1]                                          # order differs from the file order on purpose
    F = len(names) + 1                                       # + one frame without a file
    with np.errstate(all="ignore"):
        (x, y, lik, inexact), (px, py, pl) = _both(str(tmp_path), ["cam1_json"], files, [0, F], ids, 2)
        for a, b, what in ((x, px, "x"), (y, py, "y"), (lik, pl, "lik")):
            for i, name in enumerate(names + ["<missing>"]):
                assert np.array_equal(a[i], b[i].astype(np.float32), equal_nan=True), (name, what, a[i], b[i])
        assert inexact == staging.float32_inexact(px, py, pl) > 0
    ok = names.index("ok.json")
    assert x[ok, 0, :, 0].tolist() == [-7.125, 1.5, 3.0] and lik[ok, 0, :, 0].tolist() == [0.25, 0.5, 1.0]
    assert np.isnan(x[ok, 1]).all()                          # second person absent


def test_count_persons_raises_on_unparsable(tmp_path):
    cam = tmp_path / "cam1_json"
    cam.mkdir()
    (cam / "a_0.json").write_text('{"people": [{}, {}]}')
    assert staging.count_persons(str(tmp_path), ["cam1_json"], [["a_0.json"]]) == 2
    (cam / "a_1.json").write_text('{"people": [')
    with pytest.raises(ValueError):
        staging.count_persons(str(tmp_path), ["cam1_json"], [["a_0.json", "a_1.json"]])


def test_custom_model_string_ids_follow_the_reference(golden, tmp_path):
    """`[pose.CUSTOM]` from Config.toml: the reference turns the string id 'None' into None for the ROOT only
    (triangulation.py:727-729); a deeper node with a string id stays a marker whose lookups always fail, i.e. a marker
    that is NaN in every frame (side by side with the live reference: oracle/diff_models_live.py, model CUSTOM)."""
    custom = {"name": "Hip", "id": "None", "children": [
        {"name": "RKnee", "id": 0, "children": [{"name": "RFoot", "id": 5}]},
        {"name": "Spine", "id": "None", "children": [{"name": "Neck", "id": 2}]}]}
    cfg = {"pose": {"pose_model": "CUSTOM", "CUSTOM": custom}}
    ids, names = skeletons.keypoints("CUSTOM", cfg)
    assert names == ["RKnee", "RFoot", "Spine", "Neck"]                      # the root is dropped, 'Spine' is kept
    assert ids == [0, 5, skeletons.UNREADABLE_ID, 2]
    g = golden("e2e_tri_single.npz")
    proj, _ = rebuild_trial(g, tmp_path, "trial")
    dirs = staging.PoseDirs(proj)
    cam_dirs = dirs.camera_dirs()
    input_dir, files = dirs.files_for_triangulation(cam_dirs)
    (x, y, lik, _), (px, py, pl) = _both(input_dir, cam_dirs, files, [0, 5], ids, 1)
    assert np.isnan(x[:, :, 2]).all() and np.isnan(lik[:, :, 2]).all() and np.isfinite(x[:, :, [0, 1, 3]]).all()
    assert np.array_equal(x, px.astype(np.float32), equal_nan=True) and np.array_equal(lik, pl.astype(np.float32), equal_nan=True)
