"""Both off-by-default flags at once on the device — `undistort_points = true` AND `handle_LR_swap = true`: the
stage kernel's lens inversion feeding `lrswap_kernel<.., DISTORT = true>` (distorted re-projection inside the swapped
pass), against what the UNMODIFIED reference produced for a trial with lens distortion and the limbs swapped in 20 % of
the views (tests/golden/e2e_tri_undistort_lrswap.npz, oracle/make_golden_e2e.py undistort_lrswap).

The oracle side of this combination is pinned on the CPU (tests/test_oracle_golden.py::test_undistort_with_lr_swap_units,
tests/test_dropin_host.py)."""
import numpy as np
import pytest

from dropin_util import assert_trc_equal, golden_trcs, in_dir, rebuild_trial, written_trcs
from test_gpu_undistort import _lens

pytestmark = pytest.mark.gpu


def test_units_match_reference(engine, golden):
    import torch
    from pose2sim_b200 import skeletons
    g = golden("e2e_tri_undistort_lrswap.npz")
    lens = _lens(g)
    partner = skeletons.swapped_indices(skeletons.keypoints("HALPE_26")[1])
    x, y, lik = (torch.from_numpy(g[k]).cuda() for k in ("unit_x", "unit_y", "unit_lik"))
    obs = engine.stage_observations(x, y, lik, 0.3, lens=lens)
    res = engine.triangulate_lr_swap(obs, partner, g["unit_P"], 6.0, 2, lens=lens)
    torch.cuda.synchronize()
    Q, err = res["Q"].cpu().numpy(), res["err"].cpu().numpy()
    assert np.array_equal(res["nexcl"].cpu().numpy(), g["unit_nexcl"].astype(np.uint8))
    assert np.array_equal(res["mask"].cpu().numpy().view(np.uint32), g["unit_mask"])
    assert np.array_equal(np.isnan(err), np.isnan(g["unit_err"]))
    assert np.allclose(Q, g["unit_Q"], atol=1e-6, rtol=0, equal_nan=True)
    assert np.allclose(err, g["unit_err"], atol=1e-6, rtol=0, equal_nan=True)


def test_triangulate_all_writes_the_reference_trc(golden, tmp_path):
    import pose2sim_b200
    g = golden("e2e_tri_undistort_lrswap.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_demo")
    with in_dir(proj):
        assert pose2sim_b200.triangulate_all(cfg) is None
    got, ref = written_trcs(proj), golden_trcs(g)
    assert sorted(got) == sorted(ref)
    for n in ref:
        assert_trc_equal(got[n], ref[n], tol=1e-6)
