"""The native JSON reader (csrc/p2s_json.cpp, host threads, no GPU) against Python's json.load on the
golden trials and on hostile files: the value for (person, keypoint) must be exactly what
`js['people'][n]['pose_keypoints_2d'][3 id : 3 id + 3]` gives, NaN wherever that lookup would raise."""
import json
import os

import numpy as np
import pytest

from dropin_util import rebuild_trial
from pose2sim_b200 import skeletons, staging
from pose2sim_b200 import staging as stg


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


# ---- association stage: native people reader / rewriter against the Python statements --------------------------------
ASSOC_DOCS = [
    # what a pose estimator writes
    '{"version": 1.3, "people": [{"person_id": [-1], "pose_keypoints_2d": [1.5, 2.25, 0.9, 10.0, 20.0, 0.5], "face_keypoints_2d": []},'
    ' {"person_id": [-1], "pose_keypoints_2d": [3.0, 4.0, 0.8, 30.5, 40.5, 0.25]}]}',
    # other number spellings, compact separators, nested values, extra keys before and after `people`
    '{"a":{"b":[1,2.0,-0,-0.0,1e3,1E-7,123456789012345678901234567890,0.1e1,5e-324,1e400]},"people":[{"pose_keypoints_2d":[1,2,3,4,5,6]},'
    '{"pose_keypoints_2d":[7.125,8,9,1.0e1,11,12],"x":null,"y":true,"z":false}],"tail":[[],{},[{}]],"n":NaN,"i":-Infinity}',
    # a person with an empty list (not counted, not listed) and one with two values (counted, not listed)
    '{"people": [{"pose_keypoints_2d": []}, {"pose_keypoints_2d": [5.0, 6.0]}, {"pose_keypoints_2d": [1.0, 2.0, 0.5, 3.0, 4.0, 0.25]}]}',
    '{"people": []}',
    '{ "people" : [ { "pose_keypoints_2d" : [ 0.1 , 0.2 , 0.3 , 0.4 , 0.5 , 0.6 ] } ] , "version" : 1.3 }\n',
]
ASSOC_IRREGULAR = [
    '{"people": [{"pose_keypoints_2d": [1.0, null, 0.5, 2.0, 3.0, 0.5]}]}',             # not a plain number
    '{"people": [{"pose_keypoints_2d": [1.0, 2.0, 0.5]}, 7]}',                           # a person that is not an object
    '{"people": [{"pose": [1.0, 2.0, 0.5]}]}',                                           # no keypoint list
    '{"people": [{"pose_keypoints_2d": [1.0, 2.0, 0.5], "name": "J\\u00e9r\\u00f4me"}]}',  # a string with escapes (writer only)
    '{"people": [{"pose_keypoints_2d": [1.0, 2.0, 0.5]}], "people": [{"pose_keypoints_2d": [4.0, 5.0, 0.5]}]}',   # duplicate key
    '{"nobody": 1}',
    '[1, 2, 3]',
]


def _python_stage(path, t3, NP):
    js = stg.load_json(path)
    obs = np.full((NP, 3), np.nan)
    if js is None:
        return 0, obs
    n = stg.persons_per_camera(js)
    people = stg.read_people(js)
    for p in range(min(n, len(people), NP)):
        v = people[p][t3:t3 + 3]
        if len(v) == 3:
            obs[p] = v
    return n, obs


def test_native_people_reader_matches_the_python_statements(tmp_path):
    paths = []
    for i, doc in enumerate(ASSOC_DOCS + ["{not json", ""]):
        p = str(tmp_path / f"doc{i}.json")
        open(p, "w").write(doc)
        paths.append(p)
    paths.append(str(tmp_path / "missing.json"))
    for t3 in (0, 3):
        obs, named, listed, llen, status, _ = stg.read_people_files([paths], t3, 3, 4)
        for c, p in enumerate(paths):
            n, want = _python_stage(p, t3, 4)
            assert status[0, c] == (1 if c < len(ASSOC_DOCS) else 0), (c, status[0, c])
            assert named[0, c] == n, (c, named[0, c], n)
            have = np.arange(4) < min(named[0, c], listed[0, c])
            got = np.where(have[:, None], obs[0, c], np.nan)
            assert np.array_equal(got, want.astype(np.float32), equal_nan=True), (c, got, want)
    assert llen[0, 0] == 6 and llen[0, 2] == 6 and llen[0, 3] == 0
    # irregular content is handed to the Python path, never guessed at
    irr = []
    for i, doc in enumerate(ASSOC_IRREGULAR[:3] + ASSOC_IRREGULAR[5:]):
        p = str(tmp_path / f"irr{i}.json")
        open(p, "w").write(doc)
        irr.append(p)
    _, _, _, _, status, _ = stg.read_people_files([irr], 0, 3, 4)
    assert (status[0] == 2).all(), status
    _, _, _, _, status, _ = stg.read_people_files([[paths[0]]], 0, 3, 1)
    assert status[0, 0] == 3                                           # more people than max_persons


def test_native_people_writer_is_json_dumps_byte_for_byte(tmp_path):
    """`p2s_rewrite_people_files` against the reference's statements (personAssociation.py:552-580: json.load, replace
    `people`, json.dumps) on regular documents with assorted number spellings; irregular ones get status 2."""
    import json
    from pose2sim_b200 import personAssociation as pa
    src, dst, ref = [], [], []
    docs = ASSOC_DOCS + ASSOC_IRREGULAR
    for i, doc in enumerate(docs):
        s = str(tmp_path / f"src{i}.json")
        open(s, "w").write(doc)
        src.append(s)
        dst.append(str(tmp_path / f"dst{i}.json"))
        ref.append(str(tmp_path / f"ref{i}.json"))
    proposals = [np.array([[0.0] * len(docs), [np.nan] * len(docs), [1.0] * len(docs)])]
    status = stg.rewrite_people_files([src], [dst], proposals)
    pa.rewrite_frame(ref, [stg.load_json(s) for s in src], proposals[0])
    left_to_python = set(range(len(ASSOC_DOCS) + 3, len(docs)))        # escapes, duplicate key, no `people`, top-level list
    for i in range(len(docs)):
        if i in left_to_python:
            assert status[0, i] == 2 and not os.path.exists(dst[i]), (i, status[0, i])
            continue
        assert status[0, i] in (0, 1), (i, status[0, i])
        assert os.path.exists(dst[i]) == os.path.exists(ref[i]) == (status[0, i] == 1), i
        if status[0, i] == 1:
            assert open(dst[i], "rb").read() == open(ref[i], "rb").read(), (i, open(dst[i]).read(), open(ref[i]).read())
            json.loads(open(dst[i]).read())
    assert (status[0, :len(ASSOC_DOCS)] == 1).sum() >= 3
    # a missing source: no file, and a stale one is removed
    open(dst[0], "w").write("stale")
    st2 = stg.rewrite_people_files([[str(tmp_path / "nope.json")]], [[dst[0]]], proposals)
    assert st2[0, 0] == 0 and not os.path.exists(dst[0])


def test_native_index_equals_the_python_statements(tmp_path):
    """Listing, `sort_by_last_number`, `frame_file_table` and `frame_paths` against their native twin on folders with
    the naming oddities the reference's rules react to: no zero padding, extra numbers, duplicate frame numbers (which
    shift the later cameras), a numbering offset, other files in the folder."""
    base = tmp_path / "pose"
    names = {"cam1_json": [f"cam1_{i:06d}.json" for i in range(12)] + ["notes.txt", "cam1_000003.json.bak"],
             "cam2_json": [f"take2_cam2_{i}.json" for i in range(12)],                       # no padding, an extra number
             "cam3_json": [f"cam3_{i:04d}.json" for i in range(3, 15)] + ["cam3_5.json", "cam3_05.json"],   # offset + duplicates
             "cam10_json": [f"c10_{i:03d}_keypoints.json" for i in range(0, 12, 2)]}
    for d, files in names.items():
        os.makedirs(base / d)
        for f in files:
            (base / d / f).write_text("{}")
    cam_dirs = staging.sort_by_last_number(list(names))
    assert cam_dirs == ["cam1_json", "cam2_json", "cam3_json", "cam10_json"]
    want_files = [staging.sort_by_last_number(n) for n in staging.PoseDirs._list(str(base), cam_dirs)]
    ix = staging.NativeIndex(str(base), cam_dirs)
    got_files = ix.names()
    assert [sorted(g) for g in got_files] == [sorted(w) for w in want_files]
    for g, w in zip(got_files, want_files):                                    # same order up to ties of equal numbers,
        assert [staging.frame_number(n) for n in g] == [staging.frame_number(n) for n in w]     # which keep the listing order
    assert got_files == want_files
    for f_range in ([0, 12], [2, 9], [5, 40], [9, 2], [0, 0]):
        assert ix.build_table(f_range)
        want = staging.frame_paths(str(base), cam_dirs, staging.frame_file_table(want_files, f_range))
        assert ix.table_paths() == want, f_range
    sig = ix.signature()
    (base / "cam2_json" / "take2_cam2_4.json").write_text('{"people": []}')
    ix.build_table([0, 12])
    assert ix.signature() != sig                                               # a rewritten file changes the cache key
    # a name without a number: the native table declines, the Python statements raise the reference's IndexError
    (base / "cam1_json" / "readme.json").write_text("{}")
    ix2 = staging.NativeIndex(str(base), cam_dirs)
    assert not ix2.build_table([0, 12])
    with pytest.raises(IndexError):
        staging.frame_file_table([staging.sort_by_last_number(n) for n in staging.PoseDirs._list(str(base), cam_dirs)], [0, 12])
    with pytest.raises(OSError):
        staging.NativeIndex(str(tmp_path / "nowhere"), cam_dirs)


def test_staging_cache_hits_and_misses(golden, tmp_path, monkeypatch):
    """A parsed trial comes back memory-mapped while no file changed; touching, rewriting or removing a file misses."""
    g = golden("e2e_tri_single.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial")
    monkeypatch.setenv("P2S_CACHE_DIR", str(tmp_path / "cache"))
    monkeypatch.setattr(staging, "CACHE_MIN_FILES", 10)
    dirs = staging.PoseDirs(proj)
    cam_dirs = dirs.camera_dirs()
    ids, _ = skeletons.keypoints("HALPE_26")

    def stage():
        input_dir, ix = dirs.index_for_triangulation(cam_dirs)
        out = staging.stage_triangulation_indexed(ix, [0, 100], ids, 1)
        ix.close()
        return out
    a = stage()
    b = stage()
    assert isinstance(b[0], np.memmap) and not isinstance(a[0], np.memmap)
    for u, v in zip(a[:3], b[:3]):
        assert np.array_equal(u, v, equal_nan=True)
    victim = os.path.join(proj, "pose", cam_dirs[1], sorted(os.listdir(os.path.join(proj, "pose", cam_dirs[1])))[7])
    os.utime(victim, ns=(1, 1))
    c = stage()
    assert not isinstance(c[0], np.memmap)
    os.remove(victim)
    d = stage()
    assert not isinstance(d[0], np.memmap) and np.isnan(d[0][7, 0, :, 1]).all()
    assert isinstance(stage()[0], np.memmap)
