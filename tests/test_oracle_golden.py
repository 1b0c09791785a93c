"""The NumPy oracle restatement (oracle/p2s_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/*.npz, made by oracle/make_golden.py in the build container).  CPU only."""
import warnings

import numpy as np
import pytest

import p2s_oracle as orc
from conftest import tri_cases

Q_TOL = 1e-9      # metres; SVD backends differ by ~1e-14
E_TOL = 1e-8      # pixels


def _check_units(P, x, y, w, thr, mc, Q, err, nexcl, mask):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        q, e, n, m = orc.triangulate_units(x.astype(float), y.astype(float), w.astype(float), P, thr, mc)
    assert np.array_equal(n, nexcl)
    assert np.array_equal(m, mask)
    assert np.array_equal(np.isnan(e), np.isnan(err))
    assert np.allclose(q, Q, atol=Q_TOL, rtol=0, equal_nan=True)
    assert np.allclose(e, err, atol=E_TOL, rtol=0, equal_nan=True)


def test_edge_case_table(golden):
    g = golden("tri_edge_cases.npz")
    n = len(g["names"])
    assert n >= 16
    for case in tri_cases(g, "e{}_", n):
        _check_units(*case[1:])


def test_edge_case_known_answers(golden):
    """nb_cams_excluded / id_excluded_cams pinned by name."""
    g = golden("tri_edge_cases.npz")
    names = list(g["names"])
    expect = {"clean": (0, 0b0), "cam3_nan": (1, 0b1000), "cam3_nan_cam0_outlier": (2, 0b1001),
              "two_valid": (2, 0b1100), "two_valid_outlier": (2, 0b1100), "one_valid": (4, 0b1111),
              "all_nan": (4, 0b1111), "cam1_zero": (1, 0b0), "cam1_zero_cam0_outlier": (2, 0b1),
              "two_outliers": (2, 0b11), "min3_one_outlier": (1, 0b100), "min4_one_outlier": (0, 0),
              "min1_three_nan": (3, 0b1110), "c8_two_nan_one_outlier": (3, 0b11000100)}
    for name, (nx, mk) in expect.items():
        i = names.index(name)
        assert int(g[f"e{i}_nexcl"][0]) == nx, name
        assert int(g[f"e{i}_mask"][0]) == mk, name
    for name in ("two_valid_outlier", "one_valid", "all_nan", "three_outliers_same_dir", "min3_two_outliers",
                 "min4_one_outlier", "min1_three_nan"):
        i = names.index(name)
        assert np.isnan(g[f"e{i}_err"][0]) and np.isnan(g[f"e{i}_Q"]).all(), name


def test_random_units(golden):
    g = golden("tri_random_units.npz")
    for case in tri_cases(g, "r{}_", int(g["n"])):
        _check_units(*case[1:])


def test_deep_units(golden):
    """Wide rigs (12 / 13 / 16 / 20 cameras) walked by the unmodified reference through levels of thousands of camera
    subsets (oracle/make_golden_deep.py).  The NumPy restatement takes a few units per case (it costs what the reference
    costs per subset); the plain-C oracle below takes all of them."""
    g = golden("tri_deep_units.npz")
    for name, P, x, y, w, thr, mc, Q, err, nexcl, mask in tri_cases(g, "r{}_", int(g["n"])):
        order = np.argsort(-nexcl.astype(int), kind="stable")[:2]          # the two deepest units of the case
        sel = np.unique(np.concatenate([order, np.arange(0, len(nexcl), 9)]))
        _check_units(P, x[sel], y[sel], w[sel], thr, mc, Q[sel], err[sel], nexcl[sel], mask[sel])


def test_widest_units(golden):
    """24 / 28 / 32-camera rigs (BASELINE configs[4]'s widest, search capped at three or four exclusions by min_cameras),
    reference-generated (oracle/make_golden_deep.py): both oracles, every unit."""
    import c_oracle as co
    g = golden("tri_widest_units.npz")
    for name, P, x, y, w, thr, mc, Q, err, nexcl, mask in tri_cases(g, "r{}_", int(g["n"])):
        _check_units(P, x, y, w, thr, mc, Q, err, nexcl, mask)
        q, e, nx, m, lv, nc = co.triangulate_units(x, y, w, P, thr, mc)
        assert np.array_equal(nx, nexcl.astype(np.uint8)) and np.array_equal(m, mask), name
        assert np.allclose(q, Q, atol=Q_TOL, rtol=0, equal_nan=True)
        assert np.allclose(e, err, atol=E_TOL, rtol=0, equal_nan=True)


def test_cfg1_demo_cameras(golden):
    g = golden("tri_cfg1_demo.npz")
    thr, mc = g["params"]
    sel = slice(0, 2600, 4)          # a quarter of the 2600 units keeps the CPU suite short
    _check_units(g["P"], g["x"][sel], g["y"][sel], g["w"][sel], float(thr), int(mc),
                 g["Q"][sel], g["err"][sel], g["nexcl"][sel], g["mask"][sel])


def test_association_frames(golden):
    g = golden("assoc_random_frames.npz")
    for i in range(int(g["assoc_n"])):
        p = f"assoc{i}_"
        thr, lt, mc = g[p + "params"]
        obs = g[p + "obs"].astype(float)
        cnt = g[p + "count"]
        for f in range(0, obs.shape[0], 2):
            ob = [[obs[f, c, pp] for pp in range(cnt[f, c])] for c in range(obs.shape[1])]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                e, comb, Q = orc.associate_frame(ob, list(cnt[f]), g[p + "P"], float(thr), float(lt), int(mc))
            assert np.array_equal(np.nan_to_num(comb, nan=-1), np.nan_to_num(g[p + "comb"][f], nan=-1)), (i, f)
            assert np.isclose(e, g[p + "err"][f], atol=E_TOL, rtol=0, equal_nan=True) or (np.isinf(e) and np.isinf(g[p + "err"][f]))
            assert np.allclose(Q, g[p + "Q"][f], atol=Q_TOL, rtol=0, equal_nan=True)


def test_wide_likelihood_units(golden):
    """Valid likelihoods spanning up to 1e6 within a unit (likelihood threshold 0): both restatements factorise A itself
    like the reference's SVD, so they stay at rounding level where the normal matrix would be off by 1e-4 m."""
    import c_oracle as co
    g = golden("tri_wide_likelihood.npz")
    for name, P, x, y, w, thr, mc, Q, err, nexcl, mask in tri_cases(g, "r{}_", int(g["n"])):
        q, e, nx, m, lv, nc = co.triangulate_units(x, y, w, P, thr, mc)
        assert np.array_equal(nx, nexcl.astype(np.uint8)) and np.array_equal(m, mask), name
        assert np.allclose(q, Q, atol=Q_TOL, rtol=0, equal_nan=True)
        assert np.allclose(e, err, atol=E_TOL, rtol=0, equal_nan=True)
        sel = slice(0, len(x), 5)
        _check_units(P, x[sel], y[sel], w[sel], thr, mc, Q[sel], err[sel], nexcl[sel], mask[sel])


def test_six_person_association_frames(golden):
    """Six persons per camera (BASELINE configs[3]'s person count) on 4 and 5 cameras, reference-generated."""
    import c_oracle as co
    g = golden("assoc_six_persons.npz")
    for i in range(int(g["assoc_n"])):
        p = f"assoc{i}_"
        thr, lt, mc = g[p + "params"]
        e, comb, Q = co.associate_frames(g[p + "obs"], g[p + "count"], g[p + "P"], float(thr), float(lt), int(mc))
        assert np.array_equal(comb.astype(int), np.nan_to_num(g[p + "comb"], nan=-1).astype(int)), i
        assert np.allclose(e, g[p + "err"], atol=E_TOL, rtol=0)
        assert np.allclose(Q, g[p + "Q"], atol=Q_TOL, rtol=0, equal_nan=True)
    # the NumPy restatement on the cheapest frames of the first configuration
    p = "assoc0_"
    thr, lt, mc = g[p + "params"]
    obs, cnt = g[p + "obs"].astype(float), g[p + "count"]
    for f in range(0, 6):
        ob = [[obs[f, c, pp] for pp in range(cnt[f, c])] for c in range(obs.shape[1])]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            e, comb, Q = orc.associate_frame(ob, list(cnt[f]), g[p + "P"], float(thr), float(lt), int(mc))
        assert np.array_equal(np.nan_to_num(comb, nan=-1), np.nan_to_num(g[p + "comb"][f], nan=-1)), f
        assert np.allclose(Q, g[p + "Q"][f], atol=Q_TOL, rtol=0, equal_nan=True)


def test_wide_rig_association_frames(golden):
    """6 / 7 / 8-camera rigs (BASELINE configs[3]'s camera count) with 2-4 persons per camera, searches that reach
    levels 1-2 of thousands of rows, reference-generated (oracle/make_golden_wide_assoc.py)."""
    import c_oracle as co
    g = golden("assoc_wide_rigs.npz")
    off = 0
    for i in range(int(g["assoc_n"])):
        p = f"assoc{i}_"
        thr, lt, mc = g[p + "params"]
        e, comb, Q = co.associate_frames(g[p + "obs"], g[p + "count"], g[p + "P"], float(thr), float(lt), int(mc))
        assert np.array_equal(comb.astype(int), np.nan_to_num(g[p + "comb"], nan=-1).astype(int)), i
        fin = np.isfinite(g[p + "err"])
        assert np.array_equal(np.isinf(e), ~fin)
        assert np.allclose(e[fin], g[p + "err"][fin], atol=E_TOL, rtol=0)
        assert np.allclose(Q, g[p + "Q"], atol=Q_TOL, rtol=0, equal_nan=True)
        off = max(off, int(np.isnan(g[p + "comb"]).sum(1).max()))
    assert off >= 2                                   # some frame did drop two cameras
    # the NumPy restatement on the two cheapest configurations' first frames
    for p in ("assoc1_", "assoc3_"):
        thr, lt, mc = g[p + "params"]
        obs, cnt = g[p + "obs"].astype(float), g[p + "count"]
        for f in range(0, 3):
            ob = [[obs[f, c, pp] for pp in range(cnt[f, c])] for c in range(obs.shape[1])]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                e, comb, Q = orc.associate_frame(ob, list(cnt[f]), g[p + "P"], float(thr), float(lt), int(mc))
            assert np.array_equal(np.nan_to_num(comb, nan=-1), np.nan_to_num(g[p + "comb"][f], nan=-1)), (p, f)
            assert np.allclose(Q, g[p + "Q"][f], atol=Q_TOL, rtol=0, equal_nan=True)


def test_wide_likelihood_association_frames(golden):
    import c_oracle as co
    g = golden("assoc_wide_likelihood.npz")
    for i in range(int(g["assoc_n"])):
        p = f"assoc{i}_"
        thr, lt, mc = g[p + "params"]
        e, comb, Q = co.associate_frames(g[p + "obs"], g[p + "count"], g[p + "P"], float(thr), float(lt), int(mc))
        assert np.array_equal(comb.astype(int), np.nan_to_num(g[p + "comb"], nan=-1).astype(int)), i
        assert np.allclose(Q, g[p + "Q"], atol=Q_TOL, rtol=0, equal_nan=True)


def test_subset_order_is_itertools():
    """The kernel's candidate tables must follow itertools.combinations order (triangulation.py:411)."""
    import itertools
    from pose2sim_b200 import combos
    for n in (2, 3, 4, 5, 8, 11):
        for k in range(0, n + 1):
            ref = [sum(1 << c for c in cc) for cc in itertools.combinations(range(n), k)]
            assert combos.subset_masks(n, k).tolist() == ref
            for r in (0, len(ref) // 2, len(ref) - 1):
                assert combos.unrank_subset(n, k, r) == ref[r]


# ---- the plain-C oracle (oracle/p2s_oracle.c) against the same reference outputs -------------------
def test_c_oracle_triangulation(golden):
    import c_oracle as co
    for fn, fmt in (("tri_edge_cases.npz", "e{}_"), ("tri_random_units.npz", "r{}_")):
        g = golden(fn)
        n = len(g["names"]) if "names" in g else int(g["n"])
        for name, P, x, y, w, thr, mc, Q, err, nexcl, mask in tri_cases(g, fmt, n):
            q, e, nx, m, lv, nc = co.triangulate_units(x, y, w, P, thr, mc)
            assert np.array_equal(nx, nexcl.astype(np.uint8)) and np.array_equal(m, mask), name
            assert np.allclose(q, Q, atol=Q_TOL, rtol=0, equal_nan=True)
            assert np.allclose(e, err, atol=E_TOL, rtol=0, equal_nan=True)
    g = golden("tri_deep_units.npz")                                     # 12-20 cameras, levels of thousands of subsets
    deepest = 0
    for name, P, x, y, w, thr, mc, Q, err, nexcl, mask in tri_cases(g, "r{}_", int(g["n"])):
        q, e, nx, m, lv, nc = co.triangulate_units(x, y, w, P, thr, mc)
        assert np.array_equal(nx, nexcl.astype(np.uint8)) and np.array_equal(m, mask), name
        assert np.allclose(q, Q, atol=Q_TOL, rtol=0, equal_nan=True)
        assert np.allclose(e, err, atol=E_TOL, rtol=0, equal_nan=True)
        deepest = max(deepest, int(lv.max()))
    assert deepest >= 6                                                  # some unit did enumerate C(16, 5) and C(16, 6)
    g = golden("tri_cfg1_demo.npz")
    q, e, nx, m, lv, nc = co.triangulate_units(g["x"], g["y"], g["w"], g["P"], 15.0, 2)
    assert np.array_equal(nx, g["nexcl"].astype(np.uint8)) and np.array_equal(m, g["mask"])
    assert np.allclose(q, g["Q"], atol=Q_TOL, rtol=0, equal_nan=True)
    assert np.allclose(e, g["err"], atol=E_TOL, rtol=0, equal_nan=True)


def test_c_oracle_association(golden):
    import c_oracle as co
    g = golden("assoc_random_frames.npz")
    for i in range(int(g["assoc_n"])):
        p = f"assoc{i}_"
        thr, lt, mc = g[p + "params"]
        e, comb, Q = co.associate_frames(g[p + "obs"], g[p + "count"], g[p + "P"], float(thr), float(lt), int(mc))
        assert np.array_equal(comb.astype(int), np.nan_to_num(g[p + "comb"], nan=-1).astype(int))
        fin = np.isfinite(g[p + "err"])
        assert np.array_equal(np.isinf(e), ~fin)
        assert np.allclose(e[fin], g[p + "err"][fin], atol=E_TOL, rtol=0)
        assert np.allclose(Q, g[p + "Q"], atol=Q_TOL, rtol=0, equal_nan=True)


def test_c_and_numpy_oracles_agree_on_synthetic():
    import c_oracle as co
    from pose2sim_b200 import synth
    for C, mc in ((8, 2), (16, 12)):
        wl = synth.make_triangulation_workload(C, 6, 1, 26, seed=300 + C)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Q, err, nexcl, mask = orc.triangulate_units(wl["x"].astype(float), wl["y"].astype(float),
                                                        wl["lik"].astype(float), wl["P"], 15.0, mc)
        q, e, nx, m, lv, nc = co.triangulate_units(wl["x"], wl["y"], wl["lik"], wl["P"], 15.0, mc)
        assert np.array_equal(nx, nexcl.astype(np.uint8)) and np.array_equal(m, mask)
        assert np.allclose(q, Q, atol=Q_TOL, rtol=0, equal_nan=True)


def test_undistort_mode_units(golden):
    """`undistort_points = true`: the restated lens inversion (vs cv2.undistortPoints, stored by the
    generator) and the search with distorted re-projection (vs triangulation_from_best_cameras)."""
    import io
    import tomllib
    from pose2sim_b200 import calib
    g = golden("e2e_tri_undistort.npz")
    lens = []
    toml = tomllib.load(io.BytesIO(str(g["calib"]).encode()))
    for name in [str(c) for c in g["cams"]]:
        cam = toml[name]
        K = np.array(cam["matrix"], float)
        lens.append({"K": K, "dist": np.array(cam["distortions"], float), "R": calib.rodrigues(cam["rotation"]),
                     "T": np.array(cam["translation"], float),
                     "newK": calib.optimal_new_camera_matrix(K, cam["distortions"], cam["size"])})
    x, y, lik = g["unit_x"], g["unit_y"], g["unit_lik"]
    ux, uy = np.empty_like(x), np.empty_like(y)
    for c, L in enumerate(lens):
        ux[:, c], uy[:, c] = orc.undistort_points(x[:, c], y[:, c], L["K"], L["dist"], L["newK"])
    assert np.array_equal(ux, g["unit_ux"]) and np.array_equal(uy, g["unit_uy"])       # bit-exact vs cv2
    P = np.stack([np.hstack([L["newK"], np.zeros((3, 1))]) @ np.vstack([np.hstack([L["R"], L["T"][:, None]]), [0, 0, 0, 1]])
                  for L in lens])
    assert np.allclose(P, g["unit_P"], atol=1e-9, rtol=0)
    low = lik.astype(np.float64) < 0.3
    gx, gy, gl = (np.where(low, np.nan, a.astype(np.float64)) for a in (ux, uy, lik))
    sub = np.arange(0, len(gx), 3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        q, e, n, m = orc.triangulate_units(gx[sub], gy[sub], gl[sub], P, 15.0, 2, lens=lens)
    assert np.array_equal(n, g["unit_nexcl"][sub]) and np.array_equal(m, g["unit_mask"][sub])
    assert np.allclose(q, g["unit_Q"][sub], atol=Q_TOL, rtol=0, equal_nan=True)
    assert np.allclose(e, g["unit_err"][sub], atol=E_TOL, rtol=0, equal_nan=True)


def test_undistort_with_lr_swap_units(golden):
    """Both off-by-default flags at once (`undistort_points` + `handle_LR_swap`, limbs swapped in 20 % of the views):
    the restated swapped pass with the DISTORTED re-projection against the reference's per-unit outputs."""
    import io
    import tomllib
    from pose2sim_b200 import calib, skeletons
    g = golden("e2e_tri_undistort_lrswap.npz")
    toml = tomllib.load(io.BytesIO(str(g["calib"]).encode()))
    lens = []
    for name in [str(c) for c in g["cams"]]:
        cam = toml[name]
        K = np.array(cam["matrix"], float)
        lens.append({"K": K, "dist": np.array(cam["distortions"], float), "R": calib.rodrigues(cam["rotation"]),
                     "T": np.array(cam["translation"], float),
                     "newK": calib.optimal_new_camera_matrix(K, cam["distortions"], cam["size"])})
    partner = skeletons.swapped_indices(skeletons.keypoints("HALPE_26")[1])
    lik = g["unit_lik"]
    low = lik.astype(np.float64) < 0.3
    gx, gy, gl = (np.where(low, np.nan, a.astype(np.float64)) for a in (g["unit_ux"], g["unit_uy"], lik))
    n = 26 * 20                                                  # whole frames: partners live in the same frame
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        q, e, nx, m = orc.triangulate_units(gx[:n], gy[:n], gl[:n], g["unit_P"], 6.0, 2, lens=lens, partner=partner)
        q0 = orc.triangulate_units(gx[:n], gy[:n], gl[:n], g["unit_P"], 6.0, 2, lens=lens)[0]
    assert np.array_equal(nx, g["unit_nexcl"][:n]) and np.array_equal(m, g["unit_mask"][:n])
    assert np.allclose(q, g["unit_Q"][:n], atol=Q_TOL, rtol=0, equal_nan=True)
    assert np.allclose(e, g["unit_err"][:n], atol=E_TOL, rtol=0, equal_nan=True)
    assert (~np.isclose(q, q0, atol=1e-9, rtol=0, equal_nan=True).all(axis=1)).sum() >= 10     # the swapped pass mattered


def test_multi_person_oracle_reproduces_reference_frames(golden):
    """oracle/p2s_oracle_mp.py (and the product's integer bookkeeping `multi_person.proposals_from_rows`) against
    146 random frames matched by the live reference (oracle/make_golden_mp.py): thresholded affinity within 1e-9,
    proposals identical — including empty frames, single-view frames and all-NaN detections."""
    import p2s_oracle_mp as omp
    from pose2sim_b200 import multi_person as mp
    g = golden("mp_random_frames.npz")
    worst = 0.0
    for i in range(int(g["n"])):
        p = f"m{i}_"
        obs, count = g[p + "obs"], g[p + "count"]
        d_max, min_aff, min_cams = g[p + "params"]
        models = [{"K": K, "R": R, "T": T} for K, R, T in zip(g[p + "K"], g[p + "R"], g[p + "T"])]
        det = [[obs[c, q].astype(float) for q in range(count[c])] for c in range(len(count))]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            aff, cum = omp.frame_affinity(det, omp.camera_ray_params(models), float(d_max), float(min_aff))
            prop = omp.proposals_from_affinity(aff, cum, int(min_cams))
        ref_aff, ref_prop = g[p + "affinity"], g[p + "proposals"]
        assert aff.shape == ref_aff.shape, i
        worst = max(worst, float(np.abs(aff - ref_aff).max(initial=0.0)))
        assert np.array_equal(np.asarray(prop, float).reshape(-1, len(count)), ref_prop, equal_nan=True), i
        mine = mp.proposals_from_rows(omp.argmax_rows(aff, cum), int(min_cams))
        assert np.array_equal(np.asarray(mine, float).reshape(-1, len(count)), ref_prop, equal_nan=True), i
    assert worst < 1e-9, worst


def test_lr_swap_units(golden):
    """`handle_LR_swap = true`: the restatement of what the reference executes (oracle swapped_pass) against the
    reference's own outputs on 3000 units, ~15 % of which the swapped evaluation changes."""
    g = golden("lr_swap_units.npz")
    partner = g["partner"]
    for gi, (C, mc, thr, n_pairs, _) in enumerate(g["groups"]):
        x, y, w = (g[f"g{gi}_{k}"].astype(np.float64) for k in "xyw")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Q, err, nexcl, mask = orc.triangulate_units(x, y, w, g[f"g{gi}_P"], float(thr), int(mc), partner=partner)
        assert np.array_equal(nexcl, g[f"g{gi}_nexcl"]) and np.array_equal(mask, g[f"g{gi}_mask"])
        assert np.array_equal(np.isnan(err), np.isnan(g[f"g{gi}_err"]))
        assert np.allclose(Q, g[f"g{gi}_Q"], atol=1e-8, rtol=0, equal_nan=True)
        assert np.allclose(err, g[f"g{gi}_err"], atol=1e-7, rtol=0, equal_nan=True)


def test_lr_swap_units_c_oracle(golden):
    """The plain-C restatement of the swapped pass (oracle/p2s_oracle.c) against the same reference outputs."""
    import c_oracle as co
    g = golden("lr_swap_units.npz")
    for gi, (C, mc, thr, n_pairs, _) in enumerate(g["groups"]):
        Q, err, nexcl, mask = co.triangulate_units_lr_swap(g[f"g{gi}_x"], g[f"g{gi}_y"], g[f"g{gi}_w"], g["partner"],
                                                           g[f"g{gi}_P"], float(thr), int(mc))
        assert np.array_equal(nexcl, g[f"g{gi}_nexcl"]) and np.array_equal(mask, g[f"g{gi}_mask"])
        assert np.array_equal(np.isnan(err), np.isnan(g[f"g{gi}_err"]))
        assert np.allclose(Q, g[f"g{gi}_Q"], atol=1e-10, rtol=0, equal_nan=True)
        assert np.allclose(err, g[f"g{gi}_err"], atol=1e-9, rtol=0, equal_nan=True)
