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


# ---------------------------------------------------------------------------------------------------------------
# The gather fused into the search kernel: producers store their outputs straight into the consumer's memory.
# ---------------------------------------------------------------------------------------------------------------
SLOT_ALIGN = 256
FLAG_WORDS = 1024                       # per rank: arrive[b][r] at b * 16 + r (used on dst), ack[b] at 512 + b


def peer_layout(units_per_rank):
    """Byte offset of every rank's packed block inside ONE gather buffer and the buffer's size; blocks start on
    256-byte boundaries so that the kernel's 16-byte vector stores stay aligned whenever U_r % 4 == 0."""
    offs, total = [], 0
    for u in units_per_rank:
        offs.append(total)
        total += -(-PACK_BYTES * int(u) // SLOT_ALIGN) * SLOT_ALIGN
    return offs, max(total, SLOT_ALIGN)


def plane_offsets(n_units):
    U = int(n_units)
    return {"Q": 0, "err": 24 * U, "mask": 32 * U, "nexcl": 36 * U}


class _DeviceBytes:
    """Zero-copy view of a raw device allocation for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerGather:
    """One-node gather of the packed outputs WITHOUT a collective: rank r's search kernel writes its 37 bytes per
    unit directly into rank `dst`'s buffer over NVLink (mapped with CUDA IPC through the C ABI), the last CTA raises
    an arrival flag in dst's memory, dst releases a buffer for reuse through an acknowledgement flag in the
    producer's memory.  `n_buffers` gather buffers rotate, so step i + n_buffers waits for the collection of step i.

    torch.distributed is only used to exchange the 64-byte IPC handles (any backend)."""

    def __init__(self, engine, units_per_rank, group=None, dst=0, n_buffers=2):
        import torch.distributed as dist
        self.eng, self.dst, self.nb, self.group = engine, int(dst), int(n_buffers), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        assert len(units_per_rank) == self.world and self.world <= 16 and self.nb * 16 <= 512
        self.units = [int(u) for u in units_per_rank]
        self.offs, self.buf_bytes = peer_layout(self.units)
        self.flags_ptr, fh = engine.peer_alloc(FLAG_WORDS * 4)
        self.out_ptr, oh = (engine.peer_alloc(self.buf_bytes * self.nb) if self.rank == self.dst else (0, None))
        mine = (fh, oh)
        if self.world > 1:
            every = [None] * self.world
            dist.all_gather_object(every, mine, group=group)
        else:
            every = [mine]
        self.opened = []
        if self.rank == self.dst:
            self.peer_flags = [self.flags_ptr if r == self.rank else self._open(every[r][0]) for r in range(self.world)]
            self.dst_flags, self.dst_out = self.flags_ptr, self.out_ptr
        else:
            self.peer_flags = None
            self.dst_flags, self.dst_out = self._open(every[self.dst][0]), self._open(every[self.dst][1])
        self.step = 0

    def _open(self, handle):
        p = self.eng.peer_open(handle)
        self.opened.append(p)
        return p

    # ---- producer side -------------------------------------------------------------------------------------------
    def out_ptrs(self, b, rank=None):
        r = self.rank if rank is None else rank
        base = self.dst_out + b * self.buf_bytes + self.offs[r]
        return {k: base + o for k, o in plane_offsets(self.units[r]).items()}

    def push_args(self, step):
        """Flag arguments of `Engine.triangulate_planes_push` for global step number `step` (0, 1, 2 ...)."""
        b = step % self.nb
        return {"out_ptrs": self.out_ptrs(b),
                "wait_flag": self.flags_ptr + 4 * (512 + b), "wait_value": max(0, step - self.nb + 1),
                "done_flag": self.dst_flags + 4 * (b * 16 + self.rank), "done_value": step + 1}

    # ---- consumer side -------------------------------------------------------------------------------------------
    def collect(self, step, stream=None):
        """dst only: enqueue the wait for every producer's step `step` and the release of that buffer."""
        assert self.rank == self.dst
        b = step % self.nb
        acks = [self.peer_flags[r] + 4 * (512 + b) for r in range(self.world)]
        self.eng.peer_collect(self.flags_ptr + 4 * (b * 16), self.world, step + 1, acks, stream=stream)

    def views(self, b):
        """dst only: per-rank dicts of torch views (Q, err, mask, nexcl) into gather buffer b."""
        import torch
        assert self.rank == self.dst
        whole = torch.as_tensor(_DeviceBytes(self.out_ptr + b * self.buf_bytes, self.buf_bytes), device=f"cuda:{self.eng.device}")
        return [packed_views(whole[self.offs[r]:self.offs[r] + PACK_BYTES * self.units[r]], self.units[r]) for r in range(self.world)]

    def close(self):
        """Collective: every rank unmaps what it opened, then (after a barrier) the owners free their buffers."""
        for p in self.opened:
            self.eng.peer_close(p)
        self.opened = []
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)
        if self.out_ptr:
            self.eng.peer_free(self.out_ptr)
            self.out_ptr = 0
        if self.flags_ptr:
            self.eng.peer_free(self.flags_ptr)
            self.flags_ptr = 0
