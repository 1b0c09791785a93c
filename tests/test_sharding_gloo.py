"""N > 1 host plumbing on CPU: world_size-2 (and 3, ragged) gloo processes shard frames, fill their
packed output block and gather it on rank 0 — the same code path bench.py / the GPU box run over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pose2sim_b200 import sharding


def test_frame_blocks_partition():
    for F in (0, 1, 7, 100, 100_001):
        for G in (1, 2, 3, 4, 8):
            blocks = sharding.frame_blocks(F, G)
            assert blocks[0][0] == 0 and blocks[-1][1] == F
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _fake_results(f0, f1, K):
    """Deterministic stand-in for the device outputs of frames [f0, f1)."""
    u = np.arange(f0 * K, f1 * K)
    Q = np.stack([u * 0.5, u * 0.25 + 1, -u.astype(float)], axis=1)
    err = np.where(u % 7 == 0, np.nan, u * 1e-3)
    return Q, err, (u % 251).astype(np.uint32), (u % 5).astype(np.uint8)


def _worker(rank, world, port, F, K, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blocks = sharding.frame_blocks(F, world)
        units = [(b - a) * K for a, b in blocks]
        f0, f1 = blocks[rank]
        buf = torch.empty(sharding.PACK_BYTES * units[rank], dtype=torch.uint8)
        v = sharding.packed_views(buf, units[rank])
        Q, err, mask, nexcl = _fake_results(f0, f1, K)
        v["Q"].copy_(torch.from_numpy(Q)); v["err"].copy_(torch.from_numpy(err))
        v["mask"].copy_(torch.from_numpy(mask.view(np.int32))); v["nexcl"].copy_(torch.from_numpy(nexcl))
        bufs, _ = sharding.gather_packed(buf, units, dst=0)
        if rank == 0:
            out = sharding.unpack_concat(bufs, units)
            np.savez(out_path, **out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,F", [(2, 10), (2, 11), (3, 10)])
def test_gather_packed_outputs_gloo(tmp_path, world, F):
    K = 26
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, port, F, K, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    Q, err, mask, nexcl = _fake_results(0, F, K)
    assert np.array_equal(got["Q"], Q) and np.array_equal(got["err"], err, equal_nan=True)
    assert np.array_equal(got["mask"], mask) and np.array_equal(got["nexcl"], nexcl)
