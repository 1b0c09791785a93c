"""CUDA triangulation path (through the C ABI) against (a) the reference's own outputs stored as
golden vectors and (b) the NumPy oracle on seeded synthetic streams.

Tolerances (BASELINE.json north_star): |dQ| <= 1e-6 m; exclusion counts / masks bit-exact except
units whose error lies within eps = 1e-6 px of the threshold or of an arg-min tie (counted)."""
import warnings

import numpy as np
import pytest

import p2s_oracle as orc
from conftest import tri_cases
from pose2sim_b200 import synth

pytestmark = pytest.mark.gpu

Q_TOL = 1e-6     # metres (north_star)
E_TOL = 1e-6     # pixels
EPS_BAND = 1e-6  # pixels


def run_gpu(engine, P, x, y, w, thr, mc):
    out = engine.triangulate_host(np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32),
                                  np.ascontiguousarray(w, np.float32), P, None, thr, mc)
    return out


def compare(out, Q, err, nexcl, mask, thr, allow_band=True):
    """Returns the number of units that differ in a decision but sit inside the eps band."""
    dn = out["nexcl"].astype(np.int64) != np.asarray(nexcl, np.int64)
    dm = out["mask"].astype(np.uint32) != np.asarray(mask, np.uint32)
    dnan = np.isnan(out["err"]) != np.isnan(err)
    bad = dn | dm | dnan
    n_band = 0
    if bad.any():
        assert allow_band, f"{bad.sum()} decision mismatches"
        # a decision may only differ when the oracle's error is within eps of the threshold
        near = np.abs(np.nan_to_num(err, nan=thr) - thr) < EPS_BAND
        assert not (bad & ~near).any(), f"{(bad & ~near).sum()} decision mismatches outside the eps band"
        n_band = int(bad.sum())
    ok = ~bad
    assert np.allclose(out["Q"][ok], Q[ok], atol=Q_TOL, rtol=0, equal_nan=True)
    assert np.allclose(out["err"][ok], err[ok], atol=E_TOL, rtol=0, equal_nan=True)
    return n_band


def test_edge_case_table(engine, golden):
    g = golden("tri_edge_cases.npz")
    for name, P, x, y, w, thr, mc, Q, err, nexcl, mask in tri_cases(g, "e{}_", len(g["names"])):
        out = run_gpu(engine, P, x, y, w, thr, mc)
        assert compare(out, Q, err, nexcl, mask, thr, allow_band=False) == 0, name


def test_reference_random_units(engine, golden):
    g = golden("tri_random_units.npz")
    worst = 0.0
    for name, P, x, y, w, thr, mc, Q, err, nexcl, mask in tri_cases(g, "r{}_", int(g["n"])):
        out = run_gpu(engine, P, x, y, w, thr, mc)
        assert compare(out, Q, err, nexcl, mask, thr, allow_band=False) == 0, name
        worst = max(worst, float(np.nanmax(np.abs(out["Q"] - Q), initial=0.0)))
    assert worst < 1e-9          # in practice ~1e-13 m: far inside the 1e-6 m bar


@pytest.mark.parametrize("deep_min", [2048, 0, 64])
def test_reference_deep_units(engine, golden, deep_min):
    """Units of 12 / 13 / 16 / 20-camera rigs that the unmodified reference walked through levels of thousands of camera
    subsets (tests/golden/tri_deep_units.npz, oracle/make_golden_deep.py; triangulation.py:408-505): the wide-rig kernels
    with the deep levels parked for deep_search_kernel (default threshold), searched in place (0) and parked early (64)."""
    g = golden("tri_deep_units.npz")
    worst = 0.0
    try:
        engine.set_deep_search(deep_min)
        for name, P, x, y, w, thr, mc, Q, err, nexcl, mask in tri_cases(g, "r{}_", int(g["n"])):
            out = run_gpu(engine, P, x, y, w, thr, mc)
            assert compare(out, Q, err, nexcl, mask, thr, allow_band=False) == 0, name
            worst = max(worst, float(np.nanmax(np.abs(out["Q"] - Q), initial=0.0)))
    finally:
        engine.set_deep_search(2048)
    assert worst < 1e-9, worst


def test_wide_likelihood_spread_matches_reference(engine, golden):
    """Valid likelihoods of one unit spanning 1e-4 ... 1 and 1e-6 ... 1 (`likelihood_threshold_triangulation = 0` is a
    legal configuration): the reference takes the SVD of A (common.py:347-350); the normal matrix would be off by up to
    1.7e-4 m on these reference-generated units, the kernel's factorisation-of-A path (p2s_math.cuh) must keep the
    north_star bar of 1e-6 m — and is held to 1e-9 m here."""
    g = golden("tri_wide_likelihood.npz")
    worst = 0.0
    for name, P, x, y, w, thr, mc, Q, err, nexcl, mask in tri_cases(g, "r{}_", int(g["n"])):
        for lik_thr in (None, 0.0):                              # staged likelihoods as they are / through the gate at 0
            out = engine.triangulate_host(x, y, w, P, lik_thr, thr, mc)
            assert compare(out, Q, err, nexcl, mask, thr, allow_band=False) == 0, name
            worst = max(worst, float(np.nanmax(np.abs(out["Q"] - Q), initial=0.0)))
    assert worst < 1e-9, worst


def test_reference_cfg1_demo_cameras(engine, golden):
    g = golden("tri_cfg1_demo.npz")
    thr, mc = g["params"]
    out = run_gpu(engine, g["P"], g["x"], g["y"], g["w"], float(thr), int(mc))
    assert compare(out, g["Q"], g["err"], g["nexcl"], g["mask"], float(thr), allow_band=False) == 0
    st = out["stats"]
    assert sum(st["level_hist"]) + st["not_evaluated"] == 2600


@pytest.mark.parametrize("C,mc,thr,seed", [(8, 2, 15.0, 202), (16, 3, 15.0, 303), (4, 2, 15.0, 101),
                                          (5, 3, 10.0, 7), (12, 8, 15.0, 11), (32, 28, 15.0, 532),
                                          (6, 2, 15.0, 506), (24, 20, 15.0, 524), (7, 3, 15.0, 9)])
def test_oracle_synthetic(engine, C, mc, thr, seed):
    F = 12 if C <= 8 else (6 if C <= 16 else 3)
    wl = synth.make_triangulation_workload(C, F, 1, 26, seed=seed)
    out = run_gpu(engine, wl["P"], wl["x"], wl["y"], wl["lik"], thr, mc)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Q, err, nexcl, mask = orc.triangulate_units(wl["x"].astype(float), wl["y"].astype(float),
                                                    wl["lik"].astype(float), wl["P"], thr, mc)
    compare(out, Q, err, nexcl, mask, thr)


def test_likelihood_gate_matches_reference_rule(engine):
    """triangulation.py:817-821: lik < thr -> NaN triple (compared in float64 like the reference)."""
    import torch
    lik = np.array([[0.3, 0.29999998, 0.30000001, np.nan, 0.0, 1.0, 0.1, 0.5]], np.float32)
    x = np.arange(8, dtype=np.float32)[None] + 1
    y = x + 100
    xs, ys, ls = (torch.from_numpy(a).cuda() for a in (x, y, lik))
    obs = engine.stage_observations(xs, ys, ls, 0.3).cpu().numpy()      # [C, U, 4]
    exp_nan = (lik.astype(np.float64) < 0.3)[0] | np.isnan(lik[0])
    assert np.array_equal(np.isnan(obs[:, 0, 2]), exp_nan)
    assert np.array_equal(np.isnan(obs[:, 0, 0]), (lik.astype(np.float64) < 0.3)[0])
    keep = ~np.isnan(obs[:, 0, 0])
    assert np.array_equal(obs[keep, 0, 0], x[0][keep]) and np.array_equal(obs[keep, 0, 1], y[0][keep])


@pytest.mark.parametrize("thr", [0.3, 0.30000001192092896, 0.29999998211860657, 0.5, 1e-50, 0.0])
def test_fused_gate_is_the_float64_comparison(engine, thr):
    """The fused path gates in the float domain against the smallest float >= threshold; that must select exactly
    the likelihoods the reference's float64 comparison selects, also right at the rounding boundaries."""
    import torch
    wl = synth.make_triangulation_workload(8, 40, 1, 26, seed=5, lik_thr=None)
    lik = wl["lik"].copy()
    edge = np.array([0.3, 0.29999998, 0.30000004, np.nextafter(np.float32(0.3), np.float32(0)), 0.0, 1e-45, 0.5, 0.49999997], np.float32)
    lik[:64] = edge[np.arange(64 * 8).reshape(64, 8) % 8]
    x, y, l = (torch.from_numpy(a).cuda() for a in (wl["x"], wl["y"], lik))
    a = engine.triangulate_planes(x, y, l, wl["P"], thr, 15.0, 2)
    gx, gy, gl = synth.gate_likelihood(wl["x"], wl["y"], lik, thr)            # float64 comparison on the host
    b = engine.triangulate_planes(*(torch.from_numpy(t).cuda() for t in (gx, gy, gl)), wl["P"], None, 15.0, 2)
    torch.cuda.synchronize()
    for k in ("Q", "err", "nexcl", "mask"):
        assert torch.equal(torch.nan_to_num(a[k].double(), nan=-7.0), torch.nan_to_num(b[k].double(), nan=-7.0)), k


@pytest.mark.parametrize("C,U", [(8, 26 * 300), (8, 333), (5, 1001), (16, 26 * 40)])
def test_host_entry_zero_copy_equals_copy_pipeline(engine, C, U):
    """`p2s_triangulate_host` with pinned buffers (the kernel reads / writes host memory directly) against the chunked
    copy pipeline and against pageable buffers: the same bytes."""
    import torch
    wl = synth.make_triangulation_workload(C, -(-U // 26), 1, 26, seed=31 + C, lik_thr=None)
    xs, ys, ls = (np.ascontiguousarray(wl[k][:U]) for k in ("x", "y", "lik"))
    hx, hy, hl = (torch.from_numpy(a.copy()).pin_memory() for a in (xs, ys, ls))

    def pinned_out():
        return {"Q": torch.empty((U, 3), dtype=torch.float64).pin_memory().numpy(),
                "err": torch.empty(U, dtype=torch.float64).pin_memory().numpy(),
                "nexcl": torch.empty(U, dtype=torch.uint8).pin_memory().numpy(),
                "mask": torch.empty(U, dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
    mc = 2 if C < 16 else 3
    try:
        engine.set_host_mode("zero_copy")
        a = engine.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], 0.3, 15.0, mc, out=pinned_out())
        with pytest.raises(Exception):                       # pageable buffers cannot be mapped: refused, not copied silently
            engine.triangulate_host(xs, ys, ls, wl["P"], 0.3, 15.0, mc)
        engine.set_host_mode("pipeline")
        b = engine.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], 0.3, 15.0, mc, out=pinned_out())
        engine.set_host_mode("auto")
        c = engine.triangulate_host(xs, ys, ls, wl["P"], 0.3, 15.0, mc)                     # pageable -> pipeline
        d = engine.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], 0.3, 15.0, mc, out=pinned_out())
    finally:
        engine.set_host_mode("auto")
    for k in ("Q", "err", "nexcl", "mask"):
        for other in (b, c, d):
            assert np.array_equal(a[k], other[k], equal_nan=True) if a[k].dtype.kind == "f" else np.array_equal(a[k], other[k]), k
    assert a["stats"]["level_hist"] == b["stats"]["level_hist"] == c["stats"]["level_hist"]


def test_nan_coordinate_with_valid_likelihood(engine):
    """Only the likelihood decides whether a camera is valid (triangulation.py:435-436): a NaN x or y with a valid
    likelihood keeps the camera, poisons every candidate that keeps it (Q = NaN, error +inf) and is removed — and
    LISTED — by the exclusion search, exactly like an outlier (checked against the live reference when written)."""
    wl = synth.make_triangulation_workload(8, 6, 1, 26, seed=9, lik_thr=0.3)
    x, y, w = wl["x"].copy(), wl["y"].copy(), wl["lik"].copy()
    g = np.random.default_rng(3)
    m = g.random(x.shape)
    x[m < 0.06] = np.nan
    y[(m > 0.06) & (m < 0.10)] = np.nan
    for mc, thr in ((2, 15.0), (6, 15.0), (2, 1e-3)):
        out = run_gpu(engine, wl["P"], x, y, w, thr, mc)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Q, err, nexcl, mask = orc.triangulate_units(x.astype(float), y.astype(float), w.astype(float), wl["P"], thr, mc)
        compare(out, Q, err, nexcl, mask, thr)
    # staged-buffer path and fused path agree on it too
    import torch
    xs, ys, ls = (torch.from_numpy(a).cuda() for a in (x, y, w))
    a = engine.triangulate(engine.stage_observations(xs, ys, ls, None), wl["P"], 15.0, 2)
    b = engine.triangulate_planes(xs, ys, ls, wl["P"], None, 15.0, 2)
    for k in ("Q", "err", "nexcl", "mask"):
        assert torch.equal(torch.nan_to_num(a[k].double(), nan=-7.0), torch.nan_to_num(b[k].double(), nan=-7.0)), k


def test_solution_at_infinity_candidate(engine):
    """Regression (found by tests/perf/fuzz_parity.py): excluding camera 1 leaves two nearly parallel rays, the smallest
    eigenvector's last component is ~1e-3, so the secular root sits 5e-4 (relative) left of its pole and the safeguarded
    Newton needs ~40 steps.  That candidate has the smallest error of its level; the unit fails (all errors above the
    threshold) and must report cameras {1, 3} like the reference, not the runner-up {0, 3}."""
    P = synth.ring_cameras(4)[0]
    x = np.array([[443.91367, -123.24936, 654.57043, 1196.4896]], np.float32)
    y = np.array([[192.70348, 483.57507, 624.25165, 710.6486]], np.float32)
    w = np.array([[0.44252497, 0.8741765, 0.5474235, np.nan]], np.float32)
    out = run_gpu(engine, P, x, y, w, 5.0, 2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Q, err, nexcl, mask = orc.triangulate_units(x.astype(float), y.astype(float), w.astype(float), P, 5.0, 2)
    assert int(mask[0]) == 0b1010 and int(nexcl[0]) == 2 and np.isnan(err[0])
    assert int(out["mask"][0]) == 0b1010 and int(out["nexcl"][0]) == 2 and np.isnan(out["err"][0]) and np.isnan(out["Q"][0]).all()
    # the same candidate as a passing unit: a huge threshold keeps its far-away point
    out = run_gpu(engine, P, x, y, w, 100.0, 2)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Q, err, nexcl, mask = orc.triangulate_units(x.astype(float), y.astype(float), w.astype(float), P, 100.0, 2)
    assert int(out["mask"][0]) == int(mask[0]) == 0b1010
    assert np.allclose(out["Q"][0], Q[0], rtol=1e-7, atol=0) and abs(out["err"][0] - err[0]) < 1e-6


def test_ragged_and_empty(engine):
    P = synth.ring_cameras(8)[0]
    out = run_gpu(engine, P, np.zeros((0, 8), np.float32), np.zeros((0, 8), np.float32), np.zeros((0, 8), np.float32), 15.0, 2)
    assert out["Q"].shape == (0, 3)
    for U in (1, 31, 33, 257):                       # tiles that do not fill a warp
        wl = synth.make_triangulation_workload(8, 10, 1, 26, seed=U)
        x, y, w = wl["x"][:U], wl["y"][:U], wl["lik"][:U]
        out = run_gpu(engine, P, x, y, w, 15.0, 2)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Q, err, nexcl, mask = orc.triangulate_units(x.astype(float), y.astype(float), w.astype(float), P, 15.0, 2)
        compare(out, Q, err, nexcl, mask, 15.0)


def test_argument_errors(engine):
    from pose2sim_b200 import _lib
    P = synth.ring_cameras(8)[0]
    z = np.zeros((4, 8), np.float32)
    with pytest.raises(_lib.P2SError):
        engine.triangulate_host(z, z, z, P, None, 15.0, 0)            # min_cams < 1
    z1 = np.zeros((4, 1), np.float32)
    with pytest.raises(_lib.P2SError):
        engine.triangulate_host(z1, z1, z1, P[:1], None, 15.0, 1)     # n_cams < 2


def test_full_size_properties_cfg2(engine):
    """BASELINE config 2 at full size (8 cams x 26 kpts x 100k frames): size-independent properties.
    - determinism (two runs bit-identical), - invariance to a permutation of the units,
    - every finite unit has err <= thr and a finite Q; failed units are NaN in both,
    - histogram of levels sums to U,  - triangulated points lie near the generating truth."""
    C, F, thr, mc = 8, 100_000, 15.0, 2
    wl = synth.make_triangulation_workload(C, F, 1, 26, seed=202)
    out = run_gpu(engine, wl["P"], wl["x"], wl["y"], wl["lik"], thr, mc)
    U = F * 26
    st = out["stats"]
    assert sum(st["level_hist"]) + st["not_evaluated"] == U
    ok = np.isfinite(out["err"])
    assert ok.mean() > 0.99
    assert (out["err"][ok] <= thr).all() and np.isfinite(out["Q"][ok]).all()
    assert np.isnan(out["Q"][~ok]).all()
    assert np.median(np.linalg.norm(out["Q"][ok] - wl["truth"][ok], axis=1)) < 0.02
    out2 = run_gpu(engine, wl["P"], wl["x"], wl["y"], wl["lik"], thr, mc)
    for k in ("Q", "err", "nexcl", "mask"):
        assert np.array_equal(out[k], out2[k], equal_nan=True)
    perm = np.random.default_rng(0).permutation(U)
    outp = run_gpu(engine, wl["P"], wl["x"][perm], wl["y"][perm], wl["lik"][perm], thr, mc)
    for k in ("Q", "err", "nexcl", "mask"):
        assert np.array_equal(out[k][perm], outp[k], equal_nan=True)
    # oracle on a deterministic subsample of the full-size run
    sub = np.arange(0, U, U // 1500)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Q, err, nexcl, mask = orc.triangulate_units(wl["x"][sub].astype(float), wl["y"][sub].astype(float),
                                                    wl["lik"][sub].astype(float), wl["P"], thr, mc)
    compare({k: out[k][sub] for k in ("Q", "err", "nexcl", "mask")}, Q, err, nexcl, mask, thr)


def test_jacobi_solver_agrees(engine):
    """A/B: the north-star's nominal Jacobi eigen-solver gives the same decisions and Q."""
    wl = synth.make_triangulation_workload(8, 40, 1, 26, seed=9)
    a = run_gpu(engine, wl["P"], wl["x"], wl["y"], wl["lik"], 15.0, 2)
    engine.set_solver("jacobi")
    try:
        b = run_gpu(engine, wl["P"], wl["x"], wl["y"], wl["lik"], 15.0, 2)
    finally:
        engine.set_solver("secular")
    assert np.array_equal(a["nexcl"], b["nexcl"]) and np.array_equal(a["mask"], b["mask"])
    assert np.allclose(a["Q"], b["Q"], atol=1e-9, rtol=0, equal_nan=True)
    assert np.allclose(a["err"], b["err"], atol=1e-8, rtol=0, equal_nan=True)


@pytest.mark.parametrize("C,U", [(8, 2600), (8, 37), (5, 333), (4, 1), (16, 257), (3, 1000), (32, 65), (6, 999), (12, 300),
                                 (24, 130)])
def test_fused_planes_path_equals_staged_path(engine, C, U):
    """`p2s_triangulate_planes_device` (gate + float4 staging fused into the tile load) is bit-identical to
    `p2s_stage_observations_device` + `p2s_triangulate_device`, for ragged sizes and odd camera counts."""
    import torch
    wl = synth.make_triangulation_workload(C, -(-U // 26), 1, 26, seed=77 + C, lik_thr=None)
    x, y, lik = (torch.from_numpy(np.ascontiguousarray(wl[k][:U])).cuda() for k in ("x", "y", "lik"))
    mc = 2 if C < 24 else C - 4
    a = engine.triangulate(engine.stage_observations(x, y, lik, 0.3), wl["P"], 15.0, mc)
    b = engine.triangulate_planes(x, y, lik, wl["P"], 0.3, 15.0, mc)
    torch.cuda.synchronize()
    for k in ("Q", "err", "nexcl", "mask"):
        assert torch.equal(torch.nan_to_num(a[k].double(), nan=-7.0), torch.nan_to_num(b[k].double(), nan=-7.0)), k


def test_full_size_properties_cfg3_shard(engine):
    """BASELINE config 3 (16 cams x Body_with_feet, min_cameras = 3) at the per-GPU shard size bench.py uses
    (125 k of the 1 M frames = 3.25 M units): determinism, permutation invariance, decision consistency and
    the oracle on a subsample — through the fused raw-plane entry point and the host entry point."""
    import torch
    C, F, thr, mc = 16, 125_000, 15.0, 3
    wl = synth.make_triangulation_workload(C, F, 1, 26, seed=303, lik_thr=None)
    U = F * 26
    x, y, lik = (torch.from_numpy(wl[k]).cuda() for k in ("x", "y", "lik"))
    stats = engine.new_stats()
    a = engine.triangulate_planes(x, y, lik, wl["P"], 0.3, thr, mc, stats=stats)
    torch.cuda.synchronize()
    from pose2sim_b200 import ops
    st = ops.stats_dict(stats.cpu().numpy())
    assert sum(st["level_hist"]) + st["not_evaluated"] == U
    err, Q = a["err"].cpu().numpy(), a["Q"].cpu().numpy()
    ok = np.isfinite(err)
    assert ok.mean() > 0.99 and (err[ok] <= thr).all() and np.isfinite(Q[ok]).all() and np.isnan(Q[~ok]).all()
    nexcl, mask = a["nexcl"].cpu().numpy(), a["mask"].cpu().numpy().view(np.uint32)
    popc = np.array([bin(int(m)).count("1") for m in mask[:20000]])
    assert (popc == nexcl[:20000]).all()                     # no zero likelihoods here: listed == counted
    assert (nexcl[ok] <= C - mc).all()
    assert np.median(np.linalg.norm(Q[ok] - wl["truth"][ok], axis=1)) < 0.02
    perm = torch.randperm(U, generator=torch.Generator().manual_seed(1)).cuda()
    b = engine.triangulate_planes(x[perm].contiguous(), y[perm].contiguous(), lik[perm].contiguous(), wl["P"], 0.3, thr, mc)
    for k in ("Q", "err", "nexcl", "mask"):
        assert torch.equal(torch.nan_to_num(a[k][perm].double(), nan=-7.0), torch.nan_to_num(b[k].double(), nan=-7.0)), k
    sub = np.arange(0, U, U // 400)
    gx, gy, gl = synth.gate_likelihood(wl["x"][sub], wl["y"][sub], wl["lik"][sub], 0.3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oQ, oerr, onexcl, omask = orc.triangulate_units(gx.astype(float), gy.astype(float), gl.astype(float), wl["P"], thr, mc)
    compare({"Q": Q[sub], "err": err[sub], "nexcl": nexcl[sub], "mask": mask[sub]}, oQ, oerr, onexcl, omask, thr)


def test_lr_swap_units_match_reference(engine, golden):
    """`handle_LR_swap = true` (triangulation.py:509-579) in its own kernel (p2s_lrswap.cu): 3000 units with
    genuinely swapped limbs, outputs of the unmodified reference (oracle/make_golden_swap.py)."""
    import torch
    g = golden("lr_swap_units.npz")
    partner = g["partner"]
    for gi, (C, mc, thr, n_pairs, _) in enumerate(g["groups"]):
        xs, ys, ls = (torch.from_numpy(g[f"g{gi}_{k}"]).cuda() for k in "xyw")
        obs = engine.stage_observations(xs, ys, ls, None)
        res = engine.triangulate_lr_swap(obs, partner, g[f"g{gi}_P"], float(thr), int(mc))
        torch.cuda.synchronize()
        out = {"Q": res["Q"].cpu().numpy(), "err": res["err"].cpu().numpy(), "nexcl": res["nexcl"].cpu().numpy(),
               "mask": res["mask"].cpu().numpy().view(np.uint32)}
        assert compare(out, g[f"g{gi}_Q"], g[f"g{gi}_err"], g[f"g{gi}_nexcl"], g[f"g{gi}_mask"], float(thr), allow_band=False) == 0, gi
    # without partners (identity map) the swapped evaluation repeats the level's own candidates on fewer cameras;
    # with the flag's kernel and an identity map on data that never exceeds the threshold the main kernel's answer comes back
    wl = synth.make_triangulation_workload(8, 40, 1, 26, seed=21, lik_thr=0.3)
    xs, ys, ls = (torch.from_numpy(wl[k]).cuda() for k in ("x", "y", "lik"))
    obs = engine.stage_observations(xs, ys, ls, 0.3)
    a = engine.triangulate(obs, wl["P"], 1e9, 2)
    b = engine.triangulate_lr_swap(obs, np.arange(26), wl["P"], 1e9, 2)
    torch.cuda.synchronize()
    assert torch.equal(a["nexcl"], b["nexcl"]) and torch.equal(a["mask"], b["mask"])
    assert torch.allclose(a["Q"], b["Q"], atol=1e-9, rtol=0, equal_nan=True)
    with pytest.raises(ValueError):
        engine.triangulate_lr_swap(obs, [0, 5], wl["P"], 15.0, 2)


@pytest.mark.parametrize("C,mc,thr,frames", [(8, 3, 8.0, 12), (5, 2, 4.0, 20), (12, 9, 6.0, 4)])
def test_lr_swap_matches_oracle_on_swapped_limbs(engine, C, mc, thr, frames):
    """HALPE_26 units (ragged last tile, 26 does not divide 32) with the limbs swapped in 20 % of the (frame, camera)
    views: `lrswap_kernel`'s work-item dealing against the per-unit NumPy restatement."""
    import torch
    from pose2sim_b200 import skeletons
    names = skeletons.keypoints("HALPE_26")[1]
    partner = np.asarray(skeletons.swapped_indices(names))
    K = len(names)
    wl = synth.make_triangulation_workload(C, frames, 1, K, seed=31 + C, lik_thr=None, p_out=0.08)
    sw = np.random.default_rng(5 + C).random((frames, 1, C)) < 0.2
    planes = {k: np.ascontiguousarray(np.where(sw, wl[k].reshape(frames, K, C)[:, partner, :], wl[k].reshape(frames, K, C))
                                      .reshape(frames * K, C)) for k in ("x", "y", "lik")}
    obs = engine.stage_observations(*(torch.from_numpy(planes[k]).cuda() for k in ("x", "y", "lik")), 0.3)
    res = engine.triangulate_lr_swap(obs, partner, wl["P"], thr, mc)
    torch.cuda.synchronize()
    x, y, w = (planes[k].astype(np.float64) for k in ("x", "y", "lik"))
    low = w < 0.3
    x[low] = np.nan; y[low] = np.nan; w[low] = np.nan
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oQ, oerr, onexcl, omask = orc.triangulate_units(x, y, w, wl["P"], thr, mc, partner=partner)
        pQ = orc.triangulate_units(x, y, w, wl["P"], thr, mc)[0]
    out = {"Q": res["Q"].cpu().numpy(), "err": res["err"].cpu().numpy(), "nexcl": res["nexcl"].cpu().numpy(),
           "mask": res["mask"].cpu().numpy().view(np.uint32)}
    compare(out, oQ, oerr, onexcl, omask, thr)
    assert (~np.isclose(oQ, pQ, atol=1e-9, rtol=0, equal_nan=True).all(axis=1)).sum() > 0     # the swapped pass mattered


def test_lr_swap_c_abi_checks_the_partner_map(engine):
    """The C entry point reads the device-resident partner map back and refuses indices outside [0, K) and unit
    counts that K does not divide (P2S_EINVAL = 1) instead of indexing the observation buffer with them."""
    import torch
    wl = synth.make_triangulation_workload(4, 3, 1, 2, seed=3, lik_thr=0.3)
    obs = engine.stage_observations(*(torch.from_numpy(wl[k]).cuda() for k in ("x", "y", "lik")), None)
    U = obs.shape[1]
    P = np.ascontiguousarray(wl["P"], np.float64).reshape(4, 12)
    Q = torch.empty((U, 3), dtype=torch.float64, device="cuda")
    err = torch.empty(U, dtype=torch.float64, device="cuda")
    nexcl = torch.empty(U, dtype=torch.uint8, device="cuda")
    mask = torch.empty(U, dtype=torch.int32, device="cuda")

    def call(partner, K):
        part = torch.tensor(partner, dtype=torch.int32, device="cuda")
        rc = engine.lib.p2s_triangulate_lrswap_device(engine.h, obs.data_ptr(), part.data_ptr(), K, P.ctypes.data, None, U, 4,
                                                      15.0, 2, Q.data_ptr(), err.data_ptr(), nexcl.data_ptr(), mask.data_ptr(), None)
        torch.cuda.synchronize()
        return rc

    assert call([1, 0], 2) == 0
    assert call([1, 2], 2) == 1 and call([-1, 0], 2) == 1
    assert call([1, 0, 2, 3], 4) == 1                       # 6 units are not whole blocks of 4 keypoints


@pytest.mark.parametrize("C,U,mc,thr,lik_thr", [(8, 26 * 400, 2, 15.0, 0.3), (8, 26 * 400 - 7, 2, 1e-3, 0.3), (8, 4099, 6, 15.0, 0.3),
                                               (8, 3001, 2, 2.0, None), (8, 31, 2, 1e-3, 0.3), (8, 26 * 300, 7, 5.0, 0.3),
                                               (8, 26 * 300, 8, 5.0, 0.3), (4, 26 * 400 + 5, 2, 15.0, 0.3), (4, 2049, 2, 1e-3, 0.3),
                                               (4, 1000, 3, 4.0, None), (8, 26 * 200, 2, 1e9, 0.3)])
def test_pooled_kernel_equals_tile_kernel(engine, C, U, mc, thr, lik_thr):
    """`triangulate_pool_kernel` (`p2s_set_output_mode(h, 2)`: a device-resident call without statistics at 4 / 8 cameras runs
    its level-1 passes over a pool of unit slots fed by several tiles) against `triangulate_kernel` (the same call WITH the statistics block): bit
    for bit, incl. thresholds that send every unit through all levels (pool overflow: 32 pending units meet waiting slots),
    thresholds nobody fails, `min_cameras` that forbid level 1 / level 2, ragged last tiles, NaN coordinates under a valid
    likelihood, zero and NaN likelihoods."""
    import torch
    F = (U + 25) // 26
    wl = synth.make_triangulation_workload(C, F, 1, 26, seed=1000 + C + U % 97, lik_thr=None, p_out=0.12, p_low=0.1)
    x, y, w = (np.ascontiguousarray(wl[k][:U]).copy() for k in ("x", "y", "lik"))
    g = np.random.default_rng(U)
    m = g.random(x.shape)
    x[m < 0.01] = np.nan
    y[(m > 0.01) & (m < 0.02)] = np.nan
    w[(m > 0.02) & (m < 0.04)] = 0.0
    w[(m > 0.04) & (m < 0.06)] = np.nan
    xs, ys, ls = (torch.from_numpy(a).cuda() for a in (x, y, w))
    engine.set_output_mode("pooled")
    try:
        pooled = engine.triangulate_planes(xs, ys, ls, wl["P"], lik_thr, thr, mc)
        torch.cuda.synchronize()
    finally:
        engine.set_output_mode("vector")
    st = engine.new_stats()
    tile = engine.triangulate_planes(xs, ys, ls, wl["P"], lik_thr, thr, mc, stats=st)
    torch.cuda.synchronize()
    for k in ("Q", "err", "nexcl", "mask"):
        a, b = pooled[k].cpu().numpy(), tile[k].cpu().numpy()
        assert np.array_equal(a, b, equal_nan=True), (k, int((~((a == b) | ((a != a) & (b != b)))).sum()))
    from pose2sim_b200 import ops
    hist = ops.stats_dict(st.cpu().numpy())["level_hist"]
    if thr == 1e-3 and mc == 2:
        assert sum(hist[1:]) > 0.8 * U                              # the overflow case really sends (almost) everyone on


def _deep_case(C, U, seed):
    F = (U + 25) // 26
    wl = synth.make_triangulation_workload(C, F, 1, 26, seed=seed, lik_thr=None, p_out=0.12, p_low=0.1)
    x, y, w = (np.ascontiguousarray(wl[k][:U]).copy() for k in ("x", "y", "lik"))
    g = np.random.default_rng(seed + U)
    m = g.random(x.shape)
    x[m < 0.01] = np.nan
    y[(m > 0.01) & (m < 0.02)] = np.nan
    w[(m > 0.02) & (m < 0.04)] = 0.0
    w[(m > 0.04) & (m < 0.06)] = np.nan
    return wl["P"], x, y, w


@pytest.mark.parametrize("C,U,mc,thr,lik_thr,deep_min", [
    (16, 26 * 12, 3, 1e-3, 0.3, 2048),       # default threshold, every unit walks every level: C(16, 5..8) parked
    (16, 26 * 60 + 5, 3, 15.0, 0.3, 100),    # cfg3's settings, parked from level 2 (120 subsets)
    (16, 26 * 40, 9, 4.0, None, 16),         # parked at level 1, min_cameras ends the search early
    (12, 26 * 30, 2, 1e-3, 0.3, 200),        # exact-count kernel (downdate rule), parked from level 3
    (13, 26 * 20, 2, 1e-3, 0.3, 500),        # 13 of 16: the sum-of-kept-blocks form beyond level 6
    (24, 26 * 8, 20, 1e-3, 0.3, 2048),       # C(24, 3) = 2024 stays, C(24, 4) = 10 626 is parked
    (32, 26 * 3, 28, 1e-3, 0.3, 2048),       # C(32, 3) = 4 960, C(32, 4) = 35 960
    (7, 26 * 100 + 3, 2, 1e-3, 0.3, 7),      # small rig, everything parked at level 1; ragged last tile
    (5, 26 * 100, 3, 2.0, None, 10),         # parked at level 2 only
    (6, 26 * 800, 2, 1e-3, 0.3, 2),          # 20 800 parked units: more than the list holds (16 384), the rest stay in the kernel
])
def test_deep_levels_equal_single_kernel_search(engine, C, U, mc, thr, lik_thr, deep_min):
    """`deep_search_kernel` (units pending at a level of >= deep_min camera subsets are parked by the lean search kernel and
    searched by a cluster of 512-thread CTAs each) against the search that never parks — the same call with `p2s_set_deep_search(h, 0)`
    and the statistics launch, which never parks either: bit for bit, incl. NaN coordinates under a valid likelihood, zero
    and NaN likelihoods, thresholds that send every unit through every level, list overflow, ragged tiles."""
    import torch
    P, x, y, w = _deep_case(C, U, 4000 + C)
    xs, ys, ls = (torch.from_numpy(a).cuda() for a in (x, y, w))
    try:
        engine.set_deep_search(deep_min)
        deep = engine.triangulate_planes(xs, ys, ls, P, lik_thr, thr, mc)
        torch.cuda.synchronize()
        engine.set_deep_search(0)
        flat = engine.triangulate_planes(xs, ys, ls, P, lik_thr, thr, mc)
        torch.cuda.synchronize()
    finally:
        engine.set_deep_search(2048)
    st = engine.new_stats()
    counted = engine.triangulate_planes(xs, ys, ls, P, lik_thr, thr, mc, stats=st)
    torch.cuda.synchronize()
    for other in (flat, counted):
        for k in ("Q", "err", "nexcl", "mask"):
            a, b = deep[k].cpu().numpy(), other[k].cpu().numpy()
            assert np.array_equal(a, b, equal_nan=True), (k, int((~((a == b) | ((a != a) & (b != b)))).sum()))
    from pose2sim_b200 import ops
    hist = ops.stats_dict(st.cpu().numpy())["level_hist"]
    assert sum(hist[1:]) > 0                                        # the case does reach the levels it parks


@pytest.mark.parametrize("C,mc,thr,deep_min", [(8, 2, 1e-3, 8), (8, 3, 15.0, 28), (16, 3, 15.0, 16), (4, 2, 1e-3, 4)])
def test_deep_levels_staged_path_and_oracle(engine, C, mc, thr, deep_min):
    """The staged-buffer entry point (`p2s_triangulate_device`, also at 4 / 8 cameras where the raw-plane path runs the RAW
    instantiation, which never parks) with parked units against the NumPy oracle."""
    import torch
    wl = synth.make_triangulation_workload(C, 10, 1, 26, seed=900 + C, lik_thr=None, p_out=0.15, p_low=0.1)
    x, y, lik = (torch.from_numpy(wl[k]).cuda() for k in ("x", "y", "lik"))
    try:
        engine.set_deep_search(deep_min)
        out = engine.triangulate(engine.stage_observations(x, y, lik, 0.3), wl["P"], thr, mc)
        torch.cuda.synchronize()
    finally:
        engine.set_deep_search(2048)
    gx, gy, gl = synth.gate_likelihood(wl["x"], wl["y"], wl["lik"], 0.3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Q, err, nexcl, mask = orc.triangulate_units(gx.astype(float), gy.astype(float), gl.astype(float), wl["P"], thr, mc)
    got = {k: v.cpu().numpy() for k, v in out.items()}
    got["mask"] = got["mask"].view(np.uint32)
    compare(got, Q, err, nexcl, mask, thr)
