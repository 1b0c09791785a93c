"""The gather fused into the search kernel (include/pose2sim_b200.h, `p2s_triangulate_planes_push_device`,
`p2s_peer_*`; pose2sim_b200/sharding.py `PeerGather`) on ONE B200: the producer and the consumer are the same
GPU here, so the flag protocol, the vectorised tile stores and the buffer rotation are exercised without NVLink;
tools/push_multi_gpu.py runs the same protocol across GPUs under torchrun."""
import numpy as np
import pytest

from pose2sim_b200 import sharding, synth

pytestmark = pytest.mark.gpu


def _planes(F, C=8, seed=202):
    import torch
    wl = synth.make_triangulation_workload(C, F, 1, 26, seed=seed, lik_thr=None)
    return wl, tuple(torch.from_numpy(wl[k]).cuda() for k in ("x", "y", "lik"))


@pytest.mark.parametrize("F", [64, 37, 1])          # U = 1664 (aligned planes), 962 and 26 (ragged tile, unaligned planes)
def test_push_equals_plain_path(engine, F):
    import torch
    wl, (x, y, lik) = _planes(F)
    U = x.shape[0]
    ref = engine.triangulate_planes(x, y, lik, wl["P"], 0.3, 15.0, 2)
    pg = sharding.PeerGather(engine, [U], n_buffers=2)
    try:
        for step in range(5):                        # rotates both buffers, exercises wait / done / ack
            engine.triangulate_planes_push(x, y, lik, wl["P"], 0.3, 15.0, 2, **pg.push_args(step))
            pg.collect(step)
            torch.cuda.synchronize()
            got = pg.views(step % 2)[0]
            for k in ("Q", "err", "mask", "nexcl"):
                a, b = got[k].cpu().numpy(), ref[k].cpu().numpy()
                assert np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b), (step, k)
        assert engine.peer_error() == 0
    finally:
        pg.close()


@pytest.mark.parametrize("F", [64, 37])
def test_bulk_store_mode_equals_vector_stores(engine, F):
    """`p2s_set_output_mode(1)`: full tiles leave by cp.async.bulk (TMA) stores — same bytes as the vector stores."""
    import torch
    wl, (x, y, lik) = _planes(F)
    ref = engine.triangulate_planes(x, y, lik, wl["P"], 0.3, 15.0, 2)
    torch.cuda.synchronize()
    engine.set_output_mode("bulk")
    try:
        got = engine.triangulate_planes(x, y, lik, wl["P"], 0.3, 15.0, 2)
        pg = sharding.PeerGather(engine, [x.shape[0]], n_buffers=2)
        try:
            for step in range(3):
                engine.triangulate_planes_push(x, y, lik, wl["P"], 0.3, 15.0, 2, **pg.push_args(step))
                pg.collect(step)
            torch.cuda.synchronize()
            pushed = pg.views(0)[0]
            for k in ("Q", "err", "mask", "nexcl"):
                for t in (got[k], pushed[k]):
                    a, b = t.cpu().numpy(), ref[k].cpu().numpy()
                    assert np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a, b), k
            assert engine.peer_error() == 0
        finally:
            pg.close()
    finally:
        engine.set_output_mode("vector")


def test_producer_times_out_instead_of_hanging(engine):
    """A buffer that is never released: the kernel gives up after its bounded wait and reports it."""
    import torch
    wl, (x, y, lik) = _planes(8)
    pg = sharding.PeerGather(engine, [x.shape[0]], n_buffers=1)
    try:
        args = pg.push_args(0)
        engine.triangulate_planes_push(x, y, lik, wl["P"], 0.3, 15.0, 2, **args)
        args = pg.push_args(1)                       # step 0 was never collected: ack stays 0 < 1
        engine.triangulate_planes_push(x, y, lik, wl["P"], 0.3, 15.0, 2, **args)
        torch.cuda.synchronize()
        assert engine.peer_error() & 1
        assert engine.peer_error() == 0              # cleared by the read
    finally:
        pg.close()


def test_peer_layout_is_aligned_and_disjoint():
    offs, total = sharding.peer_layout([2_600_000, 2_600_000, 13, 0, 7])
    assert all(o % sharding.SLOT_ALIGN == 0 for o in offs) and total % sharding.SLOT_ALIGN == 0
    ends = [o + sharding.PACK_BYTES * u for o, u in zip(offs, [2_600_000, 2_600_000, 13, 0, 7])]
    assert all(e <= n for e, n in zip(ends, offs[1:] + [total]))
