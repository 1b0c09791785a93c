"""Host logic of the drop-in stages on a CPU box: staging (JSON -> arrays), multi-person re-ID,
interpolation / trimming / fill, TRC writer, association JSON rewrite — checked against what the
UNMODIFIED reference wrote for the same trials (tests/golden/e2e_*.npz).

The device call in the middle is replaced HERE (test code only) by the NumPy oracle; the product
functions `triangulate_all` / `associate_all` have no such seam and need the CUDA library."""
import logging
import os
import warnings

import numpy as np
import pytest

import p2s_oracle as orc
from dropin_util import (assert_multi_person_json_equal, assert_trc_equal, associated_people, golden_trcs, in_dir,
                         rebuild_trial, written_trcs)
from pose2sim_b200 import personAssociation as pa
from pose2sim_b200 import staging, triangulation as tri


def oracle_units(st):
    F, N, K, C = st.x.shape
    U = F * N * K
    s = st.settings
    x, y, w = (a.reshape(U, C).astype(np.float64) for a in (st.x, st.y, st.lik))
    if st.lens is not None:                                 # undistort_points (triangulation.py:808-813)
        for c, L in enumerate(st.lens):
            ux, uy = orc.undistort_points(x[:, c], y[:, c], L["K"], L["dist"], L["newK"])
            x[:, c], y[:, c] = ux, uy
    with np.errstate(invalid="ignore"):
        low = w < s["lik_thr"]                              # triangulation.py:817-821
    x[low] = np.nan; y[low] = np.nan; w[low] = np.nan
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        partner = tri.swapped_keypoint_indices(st.keypoints_names) if s["handle_LR_swap"] else None
        Q, err, nexcl, mask = orc.triangulate_units(x, y, w, st.P, s["reproj_thr"], s["min_cams"], lens=st.lens, partner=partner)
    return {"Q": Q.reshape(F, N, K, 3), "err": err.reshape(F, N, K), "nexcl": nexcl.reshape(F, N, K).astype(np.int64),
            "mask": mask.reshape(F, N, K)}


@pytest.mark.parametrize("tag", ["e2e_tri_single", "e2e_tri_multi", "e2e_tri_undistort", "e2e_tri_undistort_lrswap"])
def test_triangulation_host_pipeline_matches_reference_trc(golden, tmp_path, tag, caplog):
    g = golden(tag + ".npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_demo")
    with in_dir(proj), caplog.at_level(logging.INFO):
        st = tri.stage_project(cfg)
        assert st.inexact == 0
        res = oracle_units(st)
        if st.settings["multi_person"]:
            res = tri.reidentify(res, st.f_range, st.n_cams, st.settings["max_distance_m"])
        tri.write_outputs(st, res)
    got, ref = written_trcs(proj), golden_trcs(g)
    assert sorted(got) == sorted(ref)
    for name in ref:
        assert_trc_equal(got[name], ref[name], tol=1e-6)
    # the recap lines the reference logged appear in ours too (same numbers)
    ref_lines = [l.strip() for l in str(g["log"]).splitlines()
                 if l.startswith("--> Mean reprojection error") or l.startswith("Camera ") or l.startswith("In average")]
    ours = caplog.text
    assert ref_lines
    for line in ref_lines:
        assert line in ours, line


def test_staging_matches_reference_order(golden, tmp_path):
    """Keypoints come out in skeleton pre-order and units are (frame, person, keypoint)-major."""
    g = golden("e2e_tri_single.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_demo")
    with in_dir(proj):
        st = tri.stage_project(cfg)
    assert st.keypoints_ids == [19, 12, 14, 16, 21, 23, 25, 11, 13, 15, 20, 22, 24, 18, 17, 0, 1, 2, 3, 4, 6, 8, 10, 5, 7, 9]
    kp = g["kp"]                                            # [F, C, 1, 78]
    ids = np.asarray(st.keypoints_ids)
    assert st.x.shape == (100, 1, 26, 4)
    assert np.array_equal(st.x[:, 0], kp[:, :, 0, :][:, :, 3 * ids].transpose(0, 2, 1))
    assert np.array_equal(st.lik[:, 0], kp[:, :, 0, :][:, :, 3 * ids + 2].transpose(0, 2, 1))


def test_file_selection_rules():
    assert staging.sort_by_last_number(["json1", "zero", "js4on2.b", "aaaa", "eypoints_0000003.json", "ajson0", "json10"]) == \
        ["ajson0", "json1", "js4on2.b", "eypoints_0000003.json", "json10", "aaaa", "zero"]
    assert staging.frame_number("cam01_000012.json") == 12
    table = staging.frame_file_table([["a_0.json", "a_2.json"], ["b_0.json", "b_1.json", "b_2.json"]], [0, 3])
    assert table == [["a_0.json", "b_0.json"], ["none", "b_1.json"], ["a_2.json", "b_2.json"]]


def test_valid_chunk_and_gap_filling():
    s = np.array([np.nan, 1, 1, 1, np.nan, 1, 1, 1, 1, np.nan, 1])
    assert tri.valid_chunk(s, 3, "all") == (1, 9)
    assert tri.valid_chunk(s, 3, "largest") == (5, 9)
    assert tri.valid_chunk(s, 3, "first") == (1, 4)
    assert tri.valid_chunk(s, 3, "last") == (5, 9)
    assert tri.valid_chunk(s, 5, "all") == (0, 0)
    col = np.array([1.0, 2.0, np.nan, 4.0, 5.0, 0.0, 0.0, 0.0, 9.0, 10.0])
    out = tri.fill_small_gaps(col, np.arange(10), 2, "linear")
    assert out[2] == 3.0 and np.isnan(out[5:8]).all() and out[8] == 9.0


def test_swapped_keypoint_indices_follow_the_reference_rule():
    """triangulation.py:741-749: initial R <-> L, then leading right <-> left; one missing partner disables all."""
    names = ["Hip", "RHip", "RKnee", "LHip", "LKnee", "Neck", "right_eye", "left_eye"]
    assert tri.swapped_keypoint_indices(names) == [0, 3, 4, 1, 2, 5, 7, 6]
    assert tri.swapped_keypoint_indices(["Hip", "RHip", "LHip", "RKnee"]) == [0, 1, 2, 3]      # no LKnee: identity
    halpe = ["Hip", "RHip", "RKnee", "RAnkle", "RBigToe", "RSmallToe", "RHeel", "LHip", "LKnee", "LAnkle", "LBigToe", "LSmallToe",
             "LHeel", "Neck", "Head", "Nose", "RShoulder", "RElbow", "RWrist", "LShoulder", "LElbow", "LWrist"]
    idx = tri.swapped_keypoint_indices(halpe)
    assert all(idx[idx[k]] == k for k in range(len(halpe))) and idx[1] == 7 and idx[13] == 13


@pytest.mark.parametrize("workers", [None, "3"])
def test_association_host_pipeline_matches_reference_json(golden, tmp_path, monkeypatch, workers):
    """workers = "3": the file I/O of the stage runs on a process pool (frame blocks parsed and rewritten by workers)."""
    if workers:
        monkeypatch.setenv("P2S_HOST_WORKERS", workers)
    g = golden("e2e_assoc_single.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_assoc")
    with in_dir(proj):
        st = pa.stage_project(cfg)
        assert st.tracked_keypoint_id == 18                  # 'Neck' in HALPE_26
        F, C = st.count.shape
        err, comb, Q = np.empty(F), np.empty((F, C)), np.empty((F, 3))
        s = st.settings
        for f in range(F):
            ob = [[st.obs[f, c, p, :3].astype(float) for p in range(st.count[f, c])] for c in range(C)]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                err[f], comb[f], Q[f] = orc.associate_frame(ob, list(st.count[f]), st.P, s["reproj_thr"], s["lik_thr"], s["min_cams"])
        pa.write_outputs(st, {"err": err, "comb": comb, "Q": Q})
    chosen, exists = associated_people(proj, [str(c) for c in g["cams"]], F, g["chosen"].shape[2])
    assert np.array_equal(exists, g["exists"])
    assert np.array_equal(np.isnan(chosen), np.isnan(g["chosen"]))
    assert np.array_equal(np.nan_to_num(chosen).astype(np.float32), np.nan_to_num(g["chosen"]))


@pytest.mark.parametrize("workers", [None, "3"])
def test_multi_person_association_matches_reference_json(golden, tmp_path, monkeypatch, workers):
    """Host half of `associate_all` with multi_person = true (staging, proposal bookkeeping, JSON rewrite)
    with the NumPy oracle standing in for the device call: same people, in the same order, as the
    reference wrote (tests/golden/e2e_assoc_multi.npz).  This also pins oracle/p2s_oracle_mp.py."""
    import p2s_oracle_mp as omp
    from pose2sim_b200 import multi_person as mp
    if workers:
        monkeypatch.setenv("P2S_HOST_WORKERS", workers)
    g = golden("e2e_assoc_multi.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_massoc")
    with in_dir(proj):
        st = pa.stage_project(cfg)
        obs, count, models = pa.stage_multi_person(st)
        s = st.settings
        cams = omp.camera_ray_params(models)
        proposals = []
        for f in range(len(count)):
            det = [[obs[f, c, p].astype(float) for p in range(count[f, c])] for c in range(st.n_cams)]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                aff, cum = omp.frame_affinity(det, cams, s["reconstruction_error_threshold"], s["min_affinity"])
                ref = omp.proposals_from_affinity(aff, cum, s["min_cams"])
            got = mp.proposals_from_rows(omp.argmax_rows(aff, cum), s["min_cams"])
            assert np.array_equal(got, ref, equal_nan=True)
            proposals.append(got)
        pa.write_outputs_multi_person(st, proposals)
    assert_multi_person_json_equal(proj, g)


@pytest.mark.parametrize("i", range(18))
def test_edge_config_variants_match_reference(golden, tmp_path, i):
    """Configuration values at their edges (third batch): the reference's TRC, its writing nothing, or its exception."""
    from dropin_util import check_variant_outcome, rebuild_variant
    gs, gv = golden("e2e_tri_single.npz"), golden("e2e_tri_variants3.npz")
    proj, cfg = rebuild_variant(gs, gv, i, tmp_path)

    def run():
        st = tri.stage_project(cfg)
        tri.write_outputs(st, oracle_units(st))
    check_variant_outcome(gv, i, proj, run)


@pytest.mark.parametrize("batch,i", [("e2e_tri_variants.npz", i) for i in range(7)] + [("e2e_tri_variants2.npz", i) for i in range(6)])
def test_config_variants_match_reference_trc(golden, tmp_path, batch, i):
    """Frame ranges, trimming / fill / interpolation modes, missing files, other thresholds; second batch: the other
    interpolation kinds, the empty-list frame range, mostly-failing units, files missing at the trial's edges, every
    option non-default at once."""
    from dropin_util import rebuild_variant
    gs, gv = golden("e2e_tri_single.npz"), golden(batch)
    proj, cfg = rebuild_variant(gs, gv, i, tmp_path)
    with in_dir(proj):
        st = tri.stage_project(cfg)
        tri.write_outputs(st, oracle_units(st))
    got = written_trcs(proj)
    assert list(got) == [str(gv[f"v{i}_trc_name"])], (str(gv["names"][i]), list(got))
    assert_trc_equal(got[str(gv[f"v{i}_trc_name"])], str(gv[f"v{i}_trc"]), tol=1e-6)


def test_nothing_triangulated_raises_like_the_reference(golden, tmp_path):
    """A threshold nothing can meet: the reference raises `Exception('No persons have been triangulated. ...')`
    (triangulation.py:955-956) and leaves pose-3d empty; checked side by side with the live reference when this test was
    written."""
    import glob
    g = golden("e2e_tri_single.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_demo")
    cfg["triangulation"].update(reproj_error_threshold_triangulation=1, min_cameras_for_triangulation=4)
    with in_dir(proj):
        st = tri.stage_project(cfg)
        with pytest.raises(Exception, match="No persons have been triangulated"):
            tri.write_outputs(st, oracle_units(st))
    assert glob.glob(os.path.join(proj, "pose-3d", "*")) == []


def test_association_directory_probe_follows_the_reference(golden, tmp_path):
    """personAssociation.py:713-720 differs from the triangulation stage's probe (:752-758): the walk is outside the
    `try` (no pose folder -> StopIteration, not ValueError) and the probe looks into the first folder AFTER sorting, so an
    empty folder of another camera is accepted and that camera is absent in every frame.  Both checked side by side with
    the live reference (oracle/diff_errors_live.py)."""
    import glob
    import shutil
    g = golden("e2e_assoc_single.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_assoc")
    cams = [str(c) for c in g["cams"]]
    for f in glob.glob(os.path.join(proj, "pose", f"{cams[1]}_json", "*.json")):
        os.remove(f)
    with in_dir(proj):
        st = pa.stage_project(cfg)
    assert st.n_cams == len(cams) and (st.count[:, 1] == 0).all() and (st.count[:, 0] > 0).any()
    assert all(names[1] == "none" for names in st.table)
    shutil.rmtree(os.path.join(proj, "pose"))
    with in_dir(proj), pytest.raises(StopIteration):
        pa.stage_project(cfg)


def test_multi_person_crashes_of_the_reference_are_kept(golden, tmp_path):
    """Two accidents of the reference's multi-person triangulation a caller may be catching (side by side with the live
    reference in oracle/diff_errors_live.py): a truncated JSON surfaces `json.JSONDecodeError` from the person count
    (triangulation.py:88-89), and a trial with nobody in any file dies in `sort_people_sports2d` with the unpacking
    ValueError (:852)."""
    import glob
    import json
    g = golden("e2e_tri_multi.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_demo")
    files = sorted(glob.glob(os.path.join(proj, "pose", "*", "*.json")))
    text = open(files[3]).read()
    open(files[3], "w").write(text[:len(text) // 2])
    with in_dir(proj), pytest.raises(json.JSONDecodeError):
        tri.stage_project(cfg)
    for f in files:
        open(f, "w").write('{"version": 1.3, "people": []}')
    with in_dir(proj):
        st = tri.stage_project(cfg)
        with pytest.raises(ValueError, match="not enough values to unpack"):
            tri.reidentify(oracle_units(st), st.f_range, st.n_cams, st.settings["max_distance_m"])


def test_install_into_reference_runs_under_the_unchanged_orchestrator(golden, tmp_path, monkeypatch):
    """`install_into_reference()` rebinds `Pose2Sim.triangulation.triangulate_all` /
    `Pose2Sim.personAssociation.associate_all`, and the UNMODIFIED orchestrator (`Pose2Sim.Pose2Sim.triangulation(config)`,
    Pose2Sim.py:386-388 -> :241-248, which imports the stage function at call time) then runs this package.  Build
    container only (needs /root/reference through oracle/ref_shim.py); the device call is replaced by the oracle here
    because this box has no GPU — what is under test is the hook, not the kernel."""
    import importlib
    import logging
    import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("the reference is only present in the build container")
    ref = ref_shim.load_reference()
    import pose2sim_b200
    orchestrator = importlib.import_module("Pose2Sim.Pose2Sim")
    original = ref.triangulation.triangulate_all, ref.personAssociation.associate_all
    calls = []
    monkeypatch.setattr(tri, "solve_units", lambda st, engine=None, device=0: (calls.append("solve_units"), oracle_units(st))[1])
    try:
        pose2sim_b200.install_into_reference()
        assert ref.triangulation.triangulate_all is pose2sim_b200.triangulate_all
        assert ref.personAssociation.associate_all is pose2sim_b200.associate_all
        g = golden("e2e_tri_single.npz")
        proj, cfg = rebuild_trial(g, tmp_path, "trial_demo")
        cfg.setdefault("logging", {})["use_custom_logging"] = True          # keep the orchestrator from adding log handlers
        with in_dir(proj):
            orchestrator.triangulation(cfg)                                  # the reference's own entry point
        assert calls == ["solve_units"]                                      # ... reached THIS package's device seam
        got, want = written_trcs(proj), golden_trcs(g)
        assert sorted(got) == sorted(want)
        for name in want:
            assert_trc_equal(got[name], want[name], tol=1e-6)
    finally:
        ref.triangulation.triangulate_all, ref.personAssociation.associate_all = original
        logging.getLogger().handlers = [h for h in logging.getLogger().handlers if not isinstance(h, logging.FileHandler)]


def test_linear_gap_fill_and_recap_means_keep_the_reference_bits():
    """`fill_small_gaps` evaluates scipy's linear interp1d statements itself (no interpolator object per column) and
    `log_recap` takes all column means in one call: both must give the bits of the statements they replace
    (common.py:669-712 `interpolate_zeros_nans`; triangulation.py:315-328 `np.nanmean` per keypoint)."""
    from scipy import interpolate
    rng = np.random.default_rng(11)
    for trial in range(120):
        n = int(rng.integers(6, 300))
        index = np.arange(50, 50 + n)
        col = rng.normal(size=n) * rng.choice([1.0, 1e3, 1e-3])
        col[rng.random(n) < rng.choice([0.05, 0.3, 0.7])] = np.nan
        col[rng.random(n) < 0.02] = 0.0
        good = ~(np.isnan(col) | (col == 0))
        got = tri.fill_small_gaps(col, index, 10, "linear")
        if np.count_nonzero(good) <= 4:
            assert got is col
            continue
        f = interpolate.interp1d(index[good], col[good], kind="linear", fill_value="extrapolate", bounds_error=False)
        ref = np.where(good, col, f(index))
        bad = np.flatnonzero(~good)
        for seq in np.split(bad, np.flatnonzero(np.diff(bad) > 1) + 1):
            if len(seq) > 10:
                ref[seq] = np.nan
        assert np.array_equal(got, ref, equal_nan=True), trial
    for trial in range(20):
        e = rng.normal(size=(int(rng.integers(5, 3000)), 27)) * 10
        e[rng.random(e.shape) < 0.1] = np.nan
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            per_col = np.array([np.nanmean(e[:, k]) for k in range(27)])
            at_once = np.nanmean(np.ascontiguousarray(e.T), axis=1)
        assert np.array_equal(per_col, at_once, equal_nan=True)
    a = rng.normal(size=(40, 9))
    a[3:7, 2] = np.nan; a[0:2, 5] = np.nan; a[-3:, 7] = np.nan
    import pandas as pd
    assert np.array_equal(tri._ffill_bfill(a), pd.DataFrame(a).ffill().bfill().to_numpy())
    b = rng.normal(size=(10, 4))
    assert tri._ffill_bfill(b) is b


def test_batched_proposal_bookkeeping_equals_the_per_frame_statements():
    """`multi_person.proposals_from_rows_batch` (all frames on padded arrays, `np.argsort(counts)[::-1]` per frame) against
    `proposals_from_rows` (personAssociation.py:532-547 frame by frame) on random arg-max tables: clustered rows with
    noise, empty frames, rigs from 2 to 16 cameras, a min_cameras nothing passes."""
    from pose2sim_b200 import multi_person as mp
    rng = np.random.default_rng(17)
    for C, NM, maxp, mc in ((8, 48, 6, 2), (4, 20, 3, 2), (3, 10, 2, 3), (8, 64, 16, 2), (16, 40, 4, 3), (2, 6, 2, 1), (8, 48, 6, 9)):
        F = 120
        rows = np.full((F, NM, C), -1, np.int8)
        n = np.zeros(F, np.int64)
        for f in range(F):
            n[f] = rng.integers(0, NM + 1)
            pats = rng.integers(-1, maxp, size=(int(rng.integers(1, 6)), C))
            for i in range(n[f]):
                r = pats[rng.integers(len(pats))].copy()
                if rng.random() < 0.3:
                    r[rng.integers(C)] = rng.integers(-1, maxp)
                rows[f, i] = r
        ref = [mp.proposals_from_rows(rows[f, :n[f]], mc) for f in range(F)]
        got = mp.proposals_from_rows_batch(rows, n, mc)
        assert len(got) == F
        for a, b in zip(ref, got):
            assert a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)
    assert mp.proposals_from_rows_batch(np.zeros((0, 4, 3), np.int8), np.zeros(0, np.int64), 2) == []
