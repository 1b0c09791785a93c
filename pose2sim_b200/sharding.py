"""Multi-GPU plumbing: contiguous frame blocks per rank, ONE gather of the packed per-unit outputs.

Units (frame, person, keypoint) are independent (Pose2Sim/triangulation.py:831-845 keeps no
cross-unit state; the only cross-frame state, multi-person re-ID :847-865, is a host post-pass on
rank 0), so the path shards without any data-path collective.  Rank r of G owns frames
[r*F/G, (r+1)*F/G); every rank holds all projection matrices (<= 3 KB).  The only exchange is the
final gather of 37 bytes per unit (Q 24 | err 8 | mask 4 | nexcl 1) to rank 0 — NCCL over
NVLink on the GPU box, gloo in the CPU tests.
"""
import numpy as np

PACK_BYTES = 37


def frame_block(n_frames, rank, world):
    """Half-open frame range of `rank`: sizes differ by at most one, concatenation = [0, n_frames)."""
    base, extra = divmod(int(n_frames), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def frame_blocks(n_frames, world):
    return [frame_block(n_frames, r, world) for r in range(world)]


def packed_views(buf, n_units):
    """Views into a uint8 torch tensor of 37*n_units bytes: planes Q | err | mask | nexcl, each plane
    contiguous (so that the kernel's coalesced stores and ONE collective both work on the same memory)."""
    import torch
    U = int(n_units)
    assert buf.dtype == torch.uint8 and buf.numel() == PACK_BYTES * U
    return {"Q": buf[:24 * U].view(torch.float64).view(U, 3), "err": buf[24 * U:32 * U].view(torch.float64),
            "mask": buf[32 * U:36 * U].view(torch.int32), "nexcl": buf[36 * U:]}


def gather_packed(buf, units_per_rank, dst=0, group=None, async_op=False):
    """Gather every rank's packed output block on `dst` (ragged blocks allowed).
    Returns (list of per-rank uint8 tensors on dst | None, work handle | None)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    bufs = None
    if rank == dst:
        bufs = [torch.empty(PACK_BYTES * int(u), dtype=torch.uint8, device=buf.device) for u in units_per_rank]
    if len(set(int(u) for u in units_per_rank)) == 1:
        work = dist.gather(buf, bufs, dst=dst, group=group, async_op=async_op)
        return bufs, work
    # ragged: point-to-point (gather needs equal sizes)
    reqs = []
    if rank == dst:
        bufs[dst].copy_(buf)
        for r in range(len(units_per_rank)):
            if r != dst and units_per_rank[r] > 0:
                reqs.append(dist.irecv(bufs[r], src=r, group=group))
    elif buf.numel() > 0:
        reqs.append(dist.isend(buf, dst=dst, group=group))
    if not async_op:
        for q in reqs:
            q.wait()
        return bufs, None
    return bufs, reqs


def unpack_concat(bufs, units_per_rank):
    """Per-rank packed blocks -> one dict of numpy arrays in global unit order."""
    out = {"Q": [], "err": [], "mask": [], "nexcl": []}
    for b, u in zip(bufs, units_per_rank):
        v = packed_views(b, u)
        out["Q"].append(v["Q"].cpu().numpy())
        out["err"].append(v["err"].cpu().numpy())
        out["mask"].append(v["mask"].cpu().numpy().view(np.uint32))
        out["nexcl"].append(v["nexcl"].cpu().numpy())
    return {k: np.concatenate(v) for k, v in out.items()}
