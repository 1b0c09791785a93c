"""The drop-in entry points end to end on a B200: `triangulate_all(config)` / `associate_all(config)`
(CUDA library in the middle) against the files the UNMODIFIED reference wrote for the same trials
(tests/golden/e2e_*.npz, made by oracle/make_golden_e2e.py)."""
import numpy as np
import pytest

from dropin_util import assert_trc_equal, associated_people, golden_trcs, in_dir, rebuild_trial, written_trcs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["e2e_tri_single", "e2e_tri_multi", "e2e_tri_undistort"])
def test_triangulate_all_writes_the_reference_trc(golden, tmp_path, tag):
    import pose2sim_b200
    g = golden(tag + ".npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_demo")
    with in_dir(proj):
        assert pose2sim_b200.triangulate_all(cfg) is None
    got, ref = written_trcs(proj), golden_trcs(g)
    assert sorted(got) == sorted(ref)
    worst = max(assert_trc_equal(got[n], ref[n], tol=1e-6) for n in ref)     # north_star: <= 1e-6 m
    assert worst < 1e-9


def test_associate_all_writes_the_reference_json(golden, tmp_path):
    import pose2sim_b200
    g = golden("e2e_assoc_single.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_assoc")
    with in_dir(proj):
        assert pose2sim_b200.associate_all(cfg) is None
    chosen, exists = associated_people(proj, [str(c) for c in g["cams"]], g["kp"].shape[0], g["chosen"].shape[2])
    assert np.array_equal(exists, g["exists"])
    assert np.array_equal(np.isnan(chosen), np.isnan(g["chosen"]))
    assert np.array_equal(np.nan_to_num(chosen).astype(np.float32), np.nan_to_num(g["chosen"]))


def test_association_then_triangulation_chain(golden, tmp_path):
    """pose/ -> associate_all -> pose-associated/ -> triangulate_all picks pose-associated (triangulation.py:760-771)."""
    import glob
    import os
    import pose2sim_b200
    g = golden("e2e_assoc_single.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_assoc")
    with in_dir(proj):
        pose2sim_b200.associate_all(cfg)
        pose2sim_b200.triangulate_all(cfg)
    trcs = glob.glob(os.path.join(proj, "pose-3d", "*.trc"))
    assert len(trcs) == 1


VARIANTS = ([("e2e_tri_variants.npz", i) for i in range(7)] + [("e2e_tri_variants2.npz", i) for i in range(6)] +
            [("e2e_tri_variants3.npz", i) for i in range(18)])


@pytest.mark.parametrize("batch,i", VARIANTS)
def test_triangulate_all_config_variants(golden, tmp_path, batch, i):
    """All 31 reference-generated configuration variants of the single-person trial THROUGH THE DEVICE: frame ranges,
    trimming / fill / interpolation modes, missing files, thresholds (batches 1-2) and the edge values — likelihood
    threshold 0 and 1, more cameras required than exist, `min_cameras` 1, unknown option values, reversed and
    overshooting frame ranges (batch 3, where the reference's exception or its writing nothing is the golden)."""
    import pose2sim_b200
    from dropin_util import check_variant_outcome, rebuild_variant
    gs, gv = golden("e2e_tri_single.npz"), golden(batch)
    proj, cfg = rebuild_variant(gs, gv, i, tmp_path)
    check_variant_outcome(gv, i, proj, lambda: pose2sim_b200.triangulate_all(cfg))
