"""Guard bands around every caller-owned device buffer of the `*_device` entry points.

`compute-sanitizer` is closed on this GPU pool (profiles/r2i_compute_sanitizer_refused.txt), so out-of-bounds
WRITES to global memory are looked for the plain way: while an entry point runs, every tensor the host layer
allocates (`torch.empty / zeros / full` inside pose2sim_b200.ops) and every input of the test sits in the middle
of a larger allocation whose 4 KB on either side hold a byte pattern, and the pattern must be intact after the
kernels have finished.  Sizes are ragged on purpose (partial tiles, odd unit counts, missing detections,
every camera-count template) and the results are compared with the unguarded call, so a kernel that stays in
bounds only by luck of the allocator's rounding shows up.  What this cannot see: out-of-bounds reads and
shared-memory overruns (those are bounded by construction)."""
import contextlib

import numpy as np
import pytest

from pose2sim_b200 import calib, synth

pytestmark = pytest.mark.gpu

GUARD = 4096
PATTERN = 0xA5


class Guarded:
    """Allocates tensors between two guard bands and checks the bands afterwards."""

    def __init__(self, torch, device):
        self.torch, self.device, self.arenas = torch, device, []
        self._empty = torch.empty

    def alloc(self, shape, dtype):
        torch = self.torch
        if isinstance(shape, int):
            shape = (shape,)
        n = int(np.prod(shape)) * self._empty((), dtype=dtype).element_size()
        body = (n + 255) // 256 * 256                    # the tail padding up to 256 bytes is pattern as well
        arena = self._empty((GUARD + body + GUARD,), dtype=torch.uint8, device=self.device)
        arena.fill_(PATTERN)
        self.arenas.append((arena, n))
        return arena[GUARD:GUARD + n].view(dtype).reshape(shape)

    def copy_in(self, t):
        g = self.alloc(tuple(t.shape), t.dtype)
        g.copy_(t)
        return g

    def check(self):
        self.torch.cuda.synchronize()
        for arena, n in self.arenas:
            head = arena[:GUARD]
            tail = arena[GUARD + n:]
            assert bool((head == PATTERN).all()), "bytes written BEFORE a buffer"
            assert bool((tail == PATTERN).all()), "bytes written PAST a buffer"
        k = len(self.arenas)
        self.arenas = []
        return k


@contextlib.contextmanager
def guarded_allocations(engine):
    """torch.empty / zeros / full on the engine's device hand out guarded buffers inside the block."""
    import torch
    g = Guarded(torch, torch.device("cuda", engine.device))
    saved = torch.empty, torch.zeros, torch.full

    def on_device(kw):
        d = kw.get("device")
        return d is not None and torch.device(d).type == "cuda"

    def empty(*shape, **kw):
        if not on_device(kw):
            return saved[0](*shape, **kw)
        shp = shape[0] if len(shape) == 1 and not isinstance(shape[0], int) else shape
        return g.alloc(tuple(shp), kw.get("dtype", torch.float32))

    def zeros(*shape, **kw):
        if not on_device(kw):
            return saved[1](*shape, **kw)
        return empty(*shape, **kw).zero_()

    def full(shape, value, **kw):
        if not on_device(kw):
            return saved[2](shape, value, **kw)
        return empty(shape, **kw).fill_(value)

    torch.empty, torch.zeros, torch.full = empty, zeros, full
    try:
        yield g
    finally:
        torch.empty, torch.zeros, torch.full = saved


def same(a, b):
    import torch
    for k in a:
        if isinstance(a[k], torch.Tensor):
            x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
            assert np.array_equal(x, y, equal_nan=True), k


@pytest.mark.parametrize("C,min_cams", [(2, 2), (3, 2), (4, 2), (5, 3), (6, 2), (8, 2), (12, 8), (16, 3), (24, 21), (32, 28)])
def test_triangulation_entries_stay_inside_their_buffers(engine, C, min_cams):
    import torch
    wl = synth.make_triangulation_workload(C, 9, 1, 26, seed=30 + C, lik_thr=None, p_out=0.15, p_low=0.1)
    for U in (1, 31, 26 * 9 - 5):                         # a lone unit, one short of a tile, full tiles plus a ragged one
        host = [np.ascontiguousarray(wl[k][:U]) for k in ("x", "y", "lik")]
        plain_in = [torch.from_numpy(h).cuda() for h in host]
        plain = engine.triangulate_planes(*plain_in, wl["P"], 0.3, 15.0, min_cams)
        torch.cuda.synchronize()
        with guarded_allocations(engine) as g:
            x, y, lik = (g.copy_in(t) for t in plain_in)
            st = engine.new_stats()
            fused = engine.triangulate_planes(x, y, lik, wl["P"], 0.3, 15.0, min_cams, stats=st)
            staged_buf = engine.stage_observations(x, y, lik, 0.3)
            staged = engine.triangulate(staged_buf, wl["P"], 15.0, min_cams)
            assert g.check() >= 13
        same(plain, fused)
        same(plain, staged)


def test_lr_swap_and_undistort_entries_stay_inside_their_buffers(engine):
    import torch
    C = 4
    P, Ks, Rs, ts = synth.ring_cameras(C)
    dist = [-0.05, 0.02, 1e-3, -5e-4, 0.01]
    lens = [{"K": Ks[c], "dist": dist, "R": Rs[c], "T": ts[c],
             "newK": calib.optimal_new_camera_matrix(Ks[c], dist, (1080, 1920))} for c in range(C)]
    wl = synth.make_triangulation_workload(C, 7, 1, 26, seed=9, lik_thr=None, p_out=0.1)
    partner = np.arange(26, dtype=np.int32)
    partner[[1, 2]] = [2, 1]
    partner[[5, 6]] = [6, 5]
    plain_in = [torch.from_numpy(np.ascontiguousarray(wl[k])).cuda() for k in ("x", "y", "lik")]
    ref_u = engine.triangulate(engine.stage_observations(*plain_in, 0.3, lens=lens), wl["P"], 15.0, 2, lens=lens)
    ref_s = engine.triangulate_lr_swap(engine.stage_observations(*plain_in, 0.3), partner, wl["P"], 15.0, 2)
    torch.cuda.synchronize()
    with guarded_allocations(engine) as g:
        x, y, lik = (g.copy_in(t) for t in plain_in)
        und = engine.stage_observations(x, y, lik, 0.3, lens=lens)
        got_u = engine.triangulate(und, wl["P"], 15.0, 2, lens=lens)
        got_s = engine.triangulate_lr_swap(engine.stage_observations(x, y, lik, 0.3), partner, wl["P"], 15.0, 2)
        assert g.check() >= 13
    same(ref_u, got_u)
    same(ref_s, got_s)


@pytest.mark.parametrize("C,n_persons,F", [(4, 3, 9), (8, 2, 7), (8, 6, 3), (16, 2, 5)])
def test_association_entry_stays_inside_its_buffers(engine, C, n_persons, F):
    import torch
    aw = synth.make_association_workload(C, F, n_persons, seed=C + n_persons, p_out=0.2, p_low=0.1, p_missing=0.2)
    obs = np.zeros(aw["obs"].shape[:3] + (4,), np.float32)
    obs[..., :aw["obs"].shape[3]] = aw["obs"]
    d_obs, d_cnt = torch.from_numpy(obs).cuda(), torch.from_numpy(np.ascontiguousarray(aw["count"], dtype=np.int32)).cuda()
    plain = engine.associate(d_obs, d_cnt, aw["P"], 20.0, 0.3, 2, want_stats=True)
    torch.cuda.synchronize()
    with guarded_allocations(engine) as g:
        got = engine.associate(g.copy_in(d_obs), g.copy_in(d_cnt), aw["P"], 20.0, 0.3, 2, want_stats=True)
        assert g.check() >= 6
    same({k: plain[k] for k in ("err", "comb", "Q")}, got)


@pytest.mark.parametrize("C,n_persons,F", [(4, 3, 11), (8, 6, 5), (3, 8, 4), (16, 4, 3)])
def test_multi_person_entry_stays_inside_its_buffers(engine, C, n_persons, F):
    import torch
    w = synth.make_multi_person_workload(C, F, n_persons, seed=77 + C, p_missing=0.25)
    n_max = max(1, int(w["count"].sum(axis=1).max()))
    d_obs, d_cnt = torch.from_numpy(w["obs"]).cuda(), torch.from_numpy(np.ascontiguousarray(w["count"], dtype=np.int32)).cuda()
    plain = engine.associate_multi(d_obs, d_cnt, w["models"], 0.1, 0.2, n_max, want_affinity=True)
    torch.cuda.synchronize()
    with guarded_allocations(engine) as g:
        got = engine.associate_multi(g.copy_in(d_obs), g.copy_in(d_cnt), w["models"], 0.1, 0.2, n_max, want_affinity=True)
        assert g.check() >= 5
    same(plain, got)


def test_device_generator_stays_inside_its_buffers(engine):
    import torch
    P, *_ = synth.ring_cameras(8)
    plain = engine.synth_observations(P, 1000, 26 * 11 - 3, want_truth=True)
    torch.cuda.synchronize()
    with guarded_allocations(engine) as g:
        got = engine.synth_observations(P, 1000, 26 * 11 - 3, want_truth=True)
        assert g.check() >= 4
    same(plain, got)
