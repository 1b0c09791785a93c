"""The counter-based workload generator's NumPy twin (pose2sim_b200/synth_philox.py): Philox4x32-10 known-answer
vectors (Random123's kat_vectors), purity in (seed, unit, camera), and the scene statistics of SURVEY.md 8(d)."""
import numpy as np

from pose2sim_b200 import synth, synth_philox as sp


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = sp.philox4x32(*ctr, key[0], key[1])
        assert tuple(int(v) for v in got) == want


def test_pure_function_of_unit_and_camera():
    """Any sub-range (or permutation) of units regenerates the same values: what lets a rank generate its shard and the
    oracle check a subsample."""
    P = synth.ring_cameras(8)[0]
    full = sp.make_workload(8, 1000, 520, seed=508)
    part = sp.make_workload(8, 1260, 130, seed=508)
    for k in ("x", "y", "lik"):
        assert np.array_equal(full[k][260:390], part[k])
    units = np.array([1003, 5, 1003, 2 ** 33 + 7], np.int64)
    x, y, lik, Q = sp.observations(units, P, 26, 508)
    assert np.array_equal(x[0], x[2]) and not np.array_equal(x[0], x[1])
    x5, _, _, _ = sp.observations(units, P[:5], 26, 508)
    assert np.array_equal(x5, x[:, :5])                     # a camera's value does not depend on the other cameras
    other = sp.make_workload(8, 1000, 520, seed=509)
    assert not np.array_equal(full["x"], other["x"])


def test_scene_statistics():
    wl = sp.make_workload(8, 0, 26 * 4000, seed=508)
    P = wl["P"].reshape(8, 12)
    Qh = np.concatenate([wl["truth"], np.ones((len(wl["truth"]), 1))], axis=1)
    d = Qh @ P[:, 8:12].T
    dx = wl["x"].astype(np.float64) - Qh @ P[:, 0:4].T / d
    dy = wl["y"].astype(np.float64) - Qh @ P[:, 4:8].T / d
    inl = np.hypot(dx, dy) < 12                             # 6 sigma; outliers sit 50-300 px away
    assert 0.94 < inl.mean() < 0.96                         # ~5 % outliers
    assert abs(dx[inl].std() - 2.0) < 0.03 and abs(dy[inl].std() - 2.0) < 0.03 and abs(dx[inl].mean()) < 0.02
    low = wl["lik"] < 0.3
    assert 0.045 < low.mean() < 0.055
    assert wl["lik"].min() >= 0.0 and wl["lik"].max() < 1.0
    assert wl["x"].dtype == np.float32 and np.isfinite(wl["x"]).all()
