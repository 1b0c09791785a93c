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


# ---- the drop-in itself under a 2-rank job: sharded staging, ONE gather, rank 0 writes the TRC -------------
def _dropin_worker(rank, world, port, proj, cfg, tag):
    import sys
    import warnings
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import p2s_oracle as orc
    from pose2sim_b200 import triangulation as tri
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def oracle_solve(st, engine=None, device=0):          # TEST stand-in for the device call (no GPU here)
        F, N, K, C = st.x.shape
        U = F * N * K
        s = st.settings
        x, y, w = (a.reshape(U, C).astype(np.float64) for a in (st.x, st.y, st.lik))
        with np.errstate(invalid="ignore"):
            low = w < s["lik_thr"]
        x[low] = np.nan; y[low] = np.nan; w[low] = np.nan
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Q, err, nexcl, mask = orc.triangulate_units(x, y, w, st.P, s["reproj_thr"], s["min_cams"])
        return {"Q": Q.reshape(F, N, K, 3), "err": err.reshape(F, N, K), "nexcl": nexcl.reshape(F, N, K).astype(np.int64),
                "mask": mask.reshape(F, N, K)}

    tri.solve_units = oracle_solve
    try:
        os.chdir(proj)
        tri.triangulate_all(cfg)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tag", ["e2e_tri_single", "e2e_tri_multi"])
def test_dropin_two_ranks_writes_the_reference_trc(golden, tmp_path, tag):
    from dropin_util import assert_trc_equal, golden_trcs, rebuild_trial, written_trcs
    g = golden(tag + ".npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_demo")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_dropin_worker, args=(2, port, proj, cfg, tag), nprocs=2, join=True)
    got, ref = written_trcs(proj), golden_trcs(g)
    assert sorted(got) == sorted(ref)
    for name in ref:
        assert_trc_equal(got[name], ref[name], tol=1e-6)


def _broken_file_worker(rank, world, port, proj, cfg, out_dir):
    from pose2sim_b200 import triangulation as tri
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        os.chdir(proj)
        try:
            tri.triangulate_all(cfg)
            outcome = "returned"
        except Exception as e:                                   # noqa: BLE001 — the outcome is what the test reads
            outcome = type(e).__name__
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write(outcome)
    finally:
        dist.destroy_process_group()


def test_unparsable_file_on_one_rank_raises_on_every_rank(golden, tmp_path):
    """Multi-person person count (triangulation.py:784, :77-90) under a 2-rank job: each rank parses its stride of the
    files and the counts are all-reduced.  A file only ONE rank reads is broken: both ranks must raise (the reference's
    JSONDecodeError on the rank that read it), none may be left waiting in the collective."""
    import glob
    from dropin_util import rebuild_trial
    g = golden("e2e_tri_multi.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_broken")
    cam_dirs = sorted(d for d in glob.glob(os.path.join(proj, "pose*", "*")) if os.path.isdir(d))
    files = sorted(glob.glob(os.path.join(cam_dirs[0], "*.json")))
    with open(files[1], "w") as f:                               # index 1 of the camera's list: rank 1's stride only
        f.write('{"version": 1.3, "people": [')
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_dir = str(tmp_path)
    ctx = mp.spawn(_broken_file_worker, args=(2, port, proj, cfg, out_dir), nprocs=2, join=False)
    deadline = 120
    import time
    t0 = time.time()
    while not ctx.join(timeout=5):
        if time.time() - t0 > deadline:
            for p in ctx.processes:
                p.kill()
            pytest.fail("a rank was left waiting in the person-count all-reduce")
    outcomes = [open(os.path.join(out_dir, f"rank{r}.txt")).read() for r in range(2)]
    assert outcomes[1] == "JSONDecodeError", outcomes
    assert outcomes[0] == "RuntimeError", outcomes


# ---- rank-local post-processing: the sharded writer must produce the single-process file byte for byte ----------
def _variant_worker(rank, world, port, proj, cfg, log_path):
    import logging
    import sys
    import warnings
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from pose2sim_b200 import triangulation as tri
    import test_dropin_host as tdh
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if rank == 0:
        logging.basicConfig(filename=log_path, level=logging.INFO, format="%(message)s", force=True)
    tri.solve_units = lambda st, engine=None, device=0: tdh.oracle_units(st)      # TEST stand-in for the device call
    try:
        os.chdir(proj)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                tri.triangulate_all(cfg)
                outcome = ""
            except Exception as e:
                outcome = type(e).__name__ + ": " + str(e)[:60]
        with open(os.path.join(proj, f"outcome_{rank}.txt"), "w") as f:
            f.write(outcome)
        dist.barrier()
    finally:
        dist.destroy_process_group()


SHARDED_VARIANTS = [("e2e_tri_variants.npz", 1, 3), ("e2e_tri_variants.npz", 2, 2), ("e2e_tri_variants.npz", 4, 3),
                    ("e2e_tri_variants2.npz", 3, 3), ("e2e_tri_variants2.npz", 4, 4), ("e2e_tri_variants3.npz", 0, 3),
                    ("e2e_tri_variants3.npz", 10, 2), ("e2e_tri_variants3.npz", 9, 5)]


@pytest.mark.parametrize("batch,i,world", SHARDED_VARIANTS)
def test_sharded_writer_equals_single_process(golden, tmp_path, batch, i, world):
    """`write_outputs_sharded` (results stay rank-local, every rank writes its byte range of the TRC) against
    `write_outputs` in one process on reference-generated configuration variants: frame ranges, gaps that cross rank
    boundaries (interpolated and too long), every fill / trimming mode, missing files, failing trials.  Same bytes, same
    exception, same log lines."""
    import logging
    import warnings
    import test_dropin_host as tdh
    from dropin_util import in_dir, rebuild_variant, written_trcs
    from pose2sim_b200 import triangulation as tri
    gs, gv = golden("e2e_tri_single.npz"), golden(batch)
    # one process
    proj1, cfg1 = rebuild_variant(gs, gv, i, tmp_path / "one")
    log1 = str(tmp_path / "one.log")
    h = logging.FileHandler(log1)
    h.setFormatter(logging.Formatter("%(message)s"))
    logging.getLogger().addHandler(h)
    logging.getLogger().setLevel(logging.INFO)
    want_exc = ""
    try:
        with in_dir(proj1), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            st = tri.stage_project(cfg1)
            assert tri.rank_local_supported(st), "pick variants the sharded writer covers"
            try:
                tri.write_outputs(st, tdh.oracle_units(st))
            except Exception as e:
                want_exc = type(e).__name__ + ": " + str(e)[:60]
    finally:
        logging.getLogger().removeHandler(h)
        h.close()
    want = written_trcs(proj1)
    # `world` ranks
    proj, cfg = rebuild_variant(gs, gv, i, tmp_path / "many")
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    logn = str(tmp_path / "many.log")
    mp.spawn(_variant_worker, args=(world, port, proj, cfg, logn), nprocs=world, join=True)
    got = written_trcs(proj)
    assert sorted(got) == sorted(want)
    for name in want:
        assert got[name] == want[name], name                    # byte-identical
    for r in range(world):
        assert open(os.path.join(proj, f"outcome_{r}.txt")).read() == want_exc
    mask = lambda text: text.replace(str(tmp_path / "one"), "<tmp>").replace(str(tmp_path / "many"), "<tmp>")
    assert mask(open(logn).read()) == mask(open(log1).read())


# ---- associate_all under a 2-rank job: frames sharded, every rank writes its own files, no gather -------------
def _assoc_worker(rank, world, port, proj, cfg, multi):
    import sys
    import warnings
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import p2s_oracle as orc
    import p2s_oracle_mp as omp
    from pose2sim_b200 import multi_person as mpx
    from pose2sim_b200 import personAssociation as pa
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def oracle_single(st, engine=None, device=0):           # TEST stand-ins for the device calls (no GPU here)
        F, C = st.count.shape
        err, comb, Q = np.empty(F), np.empty((F, C)), np.empty((F, 3))
        s = st.settings
        for f in range(F):
            ob = [[st.obs[f, c, p, :3].astype(float) for p in range(st.count[f, c])] for c in range(C)]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                err[f], comb[f], Q[f] = orc.associate_frame(ob, list(st.count[f]), st.P, s["reproj_thr"], s["lik_thr"], s["min_cams"])
        return {"err": err, "comb": comb, "Q": Q}

    def oracle_multi(st, engine=None, device=0):
        obs, count, models = pa.stage_multi_person(st)
        s = st.settings
        cams = omp.camera_ray_params(models)
        out = []
        for f in range(len(count)):
            det = [[obs[f, c, p].astype(float) for p in range(count[f, c])] for c in range(st.n_cams)]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                aff, cum = omp.frame_affinity(det, cams, s["reconstruction_error_threshold"], s["min_affinity"])
            out.append(mpx.proposals_from_rows(omp.argmax_rows(aff, cum), s["min_cams"]))
        return out

    pa.solve_frames, pa.solve_frames_multi_person = oracle_single, oracle_multi
    try:
        os.chdir(proj)
        pa.associate_all(cfg)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("tag", ["e2e_assoc_single", "e2e_assoc_multi"])
def test_associate_all_two_ranks_writes_the_reference_json(golden, tmp_path, tag):
    from dropin_util import associated_people, rebuild_trial
    g = golden(tag + ".npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_assoc")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_assoc_worker, args=(2, port, proj, cfg, tag.endswith("multi")), nprocs=2, join=True)
    if tag.endswith("multi"):
        from dropin_util import assert_multi_person_json_equal
        assert_multi_person_json_equal(proj, g)
        return
    chosen, exists = associated_people(proj, [str(c) for c in g["cams"]], g["kp"].shape[0], g["chosen"].shape[2])
    assert np.array_equal(exists, g["exists"])
    assert np.array_equal(np.isnan(chosen), np.isnan(g["chosen"]))
    assert np.array_equal(np.nan_to_num(chosen).astype(np.float32), np.nan_to_num(g["chosen"]))


def _assoc_failing_worker(rank, world, port, proj, cfg, out_dir):
    from pose2sim_b200 import personAssociation as pa
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def device_call(st, engine=None, device=0):              # TEST stand-in for the device call: rank 1's fails
        if rank == 1:
            raise RuntimeError("device call failed on this rank")
        F, C = st.count.shape
        return {"err": np.full(F, np.inf), "comb": np.full((F, C), np.nan), "Q": np.full((F, 3), np.nan)}

    pa.solve_frames = pa.solve_frames_multi_person = device_call
    try:
        os.chdir(proj)
        try:
            pa.associate_all(cfg)
            outcome = "returned"
        except Exception as e:                               # noqa: BLE001 — the outcome is what the test reads
            outcome = f"{type(e).__name__}: {e}"
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write(outcome)
    finally:
        dist.destroy_process_group()


def test_associate_all_failure_on_one_rank_raises_on_every_rank(golden, tmp_path):
    """associate_all under a 2-rank job does its rank-local work first and meets in ONE object collective: a rank whose
    device call fails reports that through the collective — it raises its own error, the other rank raises too, and
    nobody is left waiting."""
    import time
    from dropin_util import rebuild_trial
    g = golden("e2e_assoc_single.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_assoc_fail")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.spawn(_assoc_failing_worker, args=(2, port, proj, cfg, str(tmp_path)), nprocs=2, join=False)
    t0 = time.time()
    while not ctx.join(timeout=5):
        if time.time() - t0 > 120:
            for p in ctx.processes:
                p.kill()
            pytest.fail("a rank was left waiting in associate_all's collective")
    outcomes = [open(os.path.join(str(tmp_path), f"rank{r}.txt")).read() for r in range(2)]
    assert outcomes[1] == "RuntimeError: device call failed on this rank", outcomes
    assert outcomes[0].startswith("RuntimeError: rank 1 failed in associate_all: RuntimeError: device call failed"), outcomes


def _tri_failing_worker(rank, world, port, proj, cfg, out_dir):
    from pose2sim_b200 import triangulation as tri
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def device_call(st, engine=None, device=0):              # TEST stand-in for the device call: rank 1's fails
        if rank == 1:
            raise RuntimeError("device call failed on this rank")
        F, N, K, C = st.x.shape
        return {"Q": np.full((F, N, K, 3), np.nan), "err": np.full((F, N, K), np.nan),
                "nexcl": np.zeros((F, N, K), np.int64), "mask": np.zeros((F, N, K), np.uint32)}

    tri.solve_units = device_call
    try:
        os.chdir(proj)
        try:
            tri.triangulate_all(cfg)
            outcome = "returned"
        except Exception as e:                               # noqa: BLE001 — the outcome is what the test reads
            outcome = f"{type(e).__name__}: {e}"
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write(outcome)
    finally:
        dist.destroy_process_group()


def test_triangulate_all_failure_on_one_rank_raises_on_every_rank(golden, tmp_path):
    """triangulate_all under a 2-rank job: a rank whose device call fails raises its own error, the other rank is told
    through `raise_together` and raises too instead of waiting in the sharded writer's first collective."""
    import time
    from dropin_util import rebuild_trial
    g = golden("e2e_tri_single.npz")
    proj, cfg = rebuild_trial(g, tmp_path, "trial_tri_fail")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.spawn(_tri_failing_worker, args=(2, port, proj, cfg, str(tmp_path)), nprocs=2, join=False)
    t0 = time.time()
    while not ctx.join(timeout=5):
        if time.time() - t0 > 120:
            for p in ctx.processes:
                p.kill()
            pytest.fail("a rank was left waiting in triangulate_all's collectives")
    outcomes = [open(os.path.join(str(tmp_path), f"rank{r}.txt")).read() for r in range(2)]
    assert outcomes[1] == "RuntimeError: device call failed on this rank", outcomes
    assert outcomes[0].startswith("RuntimeError: rank 1 failed in triangulate_all: RuntimeError: device call failed"), outcomes
