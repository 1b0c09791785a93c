"""Multi-person cross-view association on the B200 (`mp_associate_kernel` through p2s_associate_multi_host)
against the NumPy restatement of personAssociation.py:277-549 (oracle/p2s_oracle_mp.py, pinned by the
reference-written JSON of tests/golden/e2e_assoc_multi.npz), and `associate_all(multi_person=true)` end to
end against those files.

Tolerances: the matched affinity is the result of <= 20 ADMM steps with an SVD each; the device SVD is a
one-sided Jacobi, NumPy's is LAPACK gesdd, so values agree to rounding (1e-9 here), and the INTEGER outputs
(per-view arg-max rows, proposals) must be identical."""
import warnings

import numpy as np
import pytest

import p2s_oracle_mp as omp
from dropin_util import in_dir, rebuild_trial
from pose2sim_b200 import multi_person as mp
from pose2sim_b200 import synth

pytestmark = pytest.mark.gpu


def oracle_frames(w, d_max, min_aff):
    cams = omp.camera_ray_params(w["models"])
    out = []
    for f in range(w["F"]):
        det = [[w["obs"][f, c, p].astype(float) for p in range(w["count"][f, c])] for c in range(w["C"])]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out.append(omp.frame_affinity(det, cams, d_max, min_aff))
    return out


@pytest.mark.parametrize("C,n_persons,F,seed", [(4, 3, 40, 11), (8, 6, 24, 404), (5, 1, 10, 7), (3, 8, 12, 21),
                                                (16, 4, 6, 5)])
def test_rows_and_affinity_match_the_oracle(engine, C, n_persons, F, seed):
    w = synth.make_multi_person_workload(C, F, n_persons, seed=seed, p_missing=0.25)
    n_max = max(1, int(w["count"].sum(axis=1).max()))
    out = engine.associate_multi_host(w["obs"], w["count"], w["models"], 0.1, 0.2, n_max=n_max, want_affinity=True)
    ref = oracle_frames(w, 0.1, 0.2)
    worst = 0.0
    for f, (aff, cum) in enumerate(ref):
        n = cum[-1]
        assert n == w["count"][f].sum()
        got = out["affinity"][f, :n, :n]
        worst = max(worst, float(np.abs(got - aff).max(initial=0.0)))
        assert np.array_equal(out["rows"][f, :n], omp.argmax_rows(aff, cum)), f
        assert (out["rows"][f, n:] == -1).all()
        p_ref = omp.proposals_from_affinity(aff, cum, 2)
        p_got = mp.proposals_from_rows(out["rows"][f, :n], 2)
        assert np.array_equal(p_got, p_ref, equal_nan=True), f
    assert worst < 1e-9, worst


def test_reference_random_frames(engine, golden):
    """The kernel against 146 random frames matched by the LIVE reference (tests/golden/mp_random_frames.npz)."""
    g = golden("mp_random_frames.npz")
    worst, ties = 0.0, 0
    for i in range(int(g["n"])):
        p = f"m{i}_"
        obs, count = g[p + "obs"], g[p + "count"]
        d_max, min_aff, min_cams = g[p + "params"]
        models = [{"K": K, "R": R, "T": T} for K, R, T in zip(g[p + "K"], g[p + "R"], g[p + "T"])]
        n = int(count.sum())
        out = engine.associate_multi_host(obs[None], count[None], models, float(d_max), float(min_aff), n_max=max(n, 1),
                                          want_affinity=True)
        ref_aff, ref_prop = g[p + "affinity"], g[p + "proposals"]
        worst = max(worst, float(np.abs(out["affinity"][0, :n, :n] - ref_aff).max(initial=0.0)))
        # The per-view arg-max is only defined up to the affinity's rounding: when the two best detections of a view
        # are closer than 1e-7 (e.g. a frame whose detections are all undetected joints: every cross-view affinity
        # is exactly 1 and the SVD's rounding noise picks the winner) the frame is counted as a tie, not compared.
        cum = np.concatenate([[0], np.cumsum(count)])
        tie = False
        for v in range(len(count)):
            seg = ref_aff[:, cum[v]:cum[v + 1]]
            if seg.shape[1] >= 2:
                top = np.sort(seg, axis=1)[:, -2:]
                tie |= bool(((top[:, 1] - top[:, 0] < 1e-7) & (top[:, 1] > 0)).any())
        ties += int(tie)
        if tie:
            continue
        prop = mp.proposals_from_rows(out["rows"][0, :n], int(min_cams))
        assert np.array_equal(np.asarray(prop, float).reshape(-1, len(count)), ref_prop, equal_nan=True), i
    assert worst < 1e-9, worst
    assert ties <= 3, ties


def test_people_are_recovered(engine):
    """Size-independent property: with clean observations every proposal groups detections of ONE true person."""
    w = synth.make_multi_person_workload(8, 200, 6, seed=404, p_out=0.0, p_low=0.0, p_missing=0.0, p_nan=0.0)
    props = mp.associate_frames(engine, w["obs"], w["count"], w["models"], 0.1, 0.2, 2)
    for f, prop in enumerate(props):
        assert prop.shape == (6, 8), (f, prop.shape)
        true = np.take_along_axis(w["perm"][f], prop.astype(int).T, axis=1).T      # [proposal, camera] -> true person
        assert (true == true[:, :1]).all(), f
        assert sorted(true[:, 0]) == list(range(6))


def test_empty_and_single_view_frames(engine):
    w = synth.make_multi_person_workload(4, 6, 3, seed=3)
    w["count"][0] = 0                                   # nobody anywhere
    w["count"][1] = [2, 0, 0, 0]                        # one view only: no cross-view pair
    w["obs"][2] = np.nan                                # everything undetected
    out = engine.associate_multi_host(w["obs"], w["count"], w["models"], 0.1, 0.2, want_affinity=True)
    ref = oracle_frames(w, 0.1, 0.2)
    for f, (aff, cum) in enumerate(ref):
        n = cum[-1]
        if n == 0:
            continue
        assert np.allclose(out["affinity"][f, :n, :n], aff, atol=1e-9, rtol=0)
        assert np.array_equal(out["rows"][f, :n], omp.argmax_rows(aff, cum))
    assert mp.proposals_from_rows(out["rows"][0, :0], 2).size == 0


def test_argument_errors(engine):
    from pose2sim_b200 import _lib
    w = synth.make_multi_person_workload(4, 2, 3, seed=3)
    with pytest.raises(_lib.P2SError):
        engine.associate_multi_host(w["obs"], w["count"], w["models"], 0.1, 0.2, n_max=65)
    with pytest.raises(_lib.P2SError):
        engine.associate_multi_host(w["obs"], w["count"], w["models"], 0.0, 0.2)


def test_associate_all_multi_person_writes_the_reference_json(golden, tmp_path):
    import pose2sim_b200
    from dropin_util import assert_multi_person_json_equal
    g = golden("e2e_assoc_multi.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_massoc")
    with in_dir(proj):
        assert pose2sim_b200.associate_all(cfg) is None
    assert_multi_person_json_equal(proj, g)
