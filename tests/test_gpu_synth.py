"""Device workload generator (csrc/p2s_synth.cu) against its NumPy twin: bit-identical float32 observations for any
unit range, and the search on device-generated inputs against the C oracle on the twin's."""
import numpy as np
import pytest

from pose2sim_b200 import synth, synth_philox as sp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,unit0,n,seed", [(8, 0, 26 * 300, 508), (4, 123457, 5000, 504), (32, 2 ** 33 + 11, 4096, 532), (16, 77, 33, 516)])
def test_device_generator_equals_numpy_twin(engine, C, unit0, n, seed):
    P = synth.ring_cameras(C)[0]
    dev = engine.synth_observations(P, unit0, n, 26, seed, want_truth=True)
    tw = sp.make_workload(C, unit0, n, 26, seed)
    for k in ("x", "y", "lik"):
        assert np.array_equal(dev[k].cpu().numpy(), tw[k]), k
    assert np.array_equal(dev["truth"].cpu().numpy(), tw["truth"])


def test_search_on_device_generated_shard_matches_oracle(engine):
    import c_oracle as co
    C, unit0, n, seed, mc = 12, 26 * 10 ** 6, 26 * 500, 512, 8
    P = synth.ring_cameras(C)[0]
    dev = engine.synth_observations(P, unit0, n, 26, seed)
    out = engine.triangulate_planes(dev["x"], dev["y"], dev["lik"], P, 0.3, 15.0, mc)
    tw = sp.make_workload(C, unit0, n, 26, seed)
    x, y, lik = synth.gate_likelihood(tw["x"], tw["y"], tw["lik"], 0.3)
    q, e, nx, m, lv, nc = co.triangulate_units(x, y, lik, P, 15.0, mc)
    assert np.array_equal(out["nexcl"].cpu().numpy(), nx)
    assert np.array_equal(out["mask"].cpu().numpy().view(np.uint32), m)
    assert np.allclose(out["Q"].cpu().numpy(), q, atol=1e-6, rtol=0, equal_nan=True)
    assert np.allclose(out["err"].cpu().numpy(), e, atol=1e-6, rtol=0, equal_nan=True)
