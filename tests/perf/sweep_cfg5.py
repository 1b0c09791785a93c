#!/usr/bin/env python
"""BASELINE.json configs[4] ("stress sweep 4-32 cams x 10M frames (combination explosion) sharded over 8xB200").

    python tests/perf/sweep_cfg5.py [--frames F] [--cams 4,8,16,32] [--check N]                       one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tests/perf/sweep_cfg5.py --frames 10000000   the config itself

Per camera count C: F frames x 26 keypoints, min_cameras = max(2, C - 4) (the search is capped at exclusion level 4: at
most sum_k<=4 C(32,k) = 41 449 candidates per unit — SURVEY.md 8(d)), seed 500 + C.  Frames are sharded in contiguous
blocks over the ranks; every rank GENERATES its shard on the device (csrc/p2s_synth.cu, a pure function of (seed, unit,
camera)), runs the fused search kernel on it and keeps its results; nothing crosses ranks but the timing and the counters.
Every rank regenerates `--check` random units of its own shard with the NumPy twin (synth_philox.py) and compares the
device results with the plain-C oracle on them.

One JSON line per C on stdout (rank 0) and appended to gpurun_out/sweep_cfg5.jsonl: units/s and candidates/s of the whole
job (units of all ranks / max-over-ranks device time), the level histogram, the generator's rate, the parity sample.
(tests/perf-style harness: it imports oracle/ as the checker, never as the thing measured.)"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200_000)
    ap.add_argument("--cams", default="4,6,8,12,16,24,32")
    ap.add_argument("--check", type=int, default=1500, help="units per rank compared with the C oracle")
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from pose2sim_b200 import ops, sharding, synth, synth_philox
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = ops.get_engine(local)
    K = 26
    b0, b1 = sharding.frame_block(args.frames, rank, world)
    unit0, U = b0 * K, (b1 - b0) * K

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for C in [int(c) for c in args.cams.split(",")]:
        mc, seed = max(2, C - 4), 500 + C
        P = synth.ring_cameras(C)[0]
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        barrier()
        e0.record()
        wl = eng.synth_observations(P, unit0, U, K, seed)
        e1.record()
        stats = eng.new_stats()
        out = eng.triangulate_planes(wl["x"], wl["y"], wl["lik"], P, 0.3, 15.0, mc, stats=stats)
        torch.cuda.synchronize()
        st = ops.stats_dict(stats.cpu().numpy())
        barrier()
        e1.record()
        for _ in range(args.steps):
            eng.triangulate_planes(wl["x"], wl["y"], wl["lik"], P, 0.3, 15.0, mc, out=out)
        e2.record()
        barrier()
        gen_ms = 0.0
        ms = e1.elapsed_time(e2) / args.steps
        # generator rate: timed separately (one more pass into the same buffers)
        e0.record()
        eng.synth_observations(P, unit0, U, K, seed, out={k: wl[k] for k in ("x", "y", "lik")})
        e1.record()
        torch.cuda.synchronize()
        gen_ms = e0.elapsed_time(e1)
        # parity sample of this rank's shard against the C oracle on the NumPy twin's inputs
        import c_oracle as co
        g = np.random.default_rng(1000 * C + rank)
        n_chk = min(args.check, U)
        sel = np.sort(g.choice(U, n_chk, replace=False)) if n_chk else np.zeros(0, np.int64)
        x, y, lik, _ = synth_philox.observations(unit0 + sel, P, K, seed)
        xg, yg, lg = synth.gate_likelihood(x, y, lik, 0.3)
        q, e, nx, m, lv, nc = co.triangulate_units(xg, yg, lg, P, 15.0, mc)
        sel_t = torch.from_numpy(sel).to(dev)
        dQ = out["Q"][sel_t].cpu().numpy()
        dn, dm = out["nexcl"][sel_t].cpu().numpy(), out["mask"][sel_t].cpu().numpy().view(np.uint32)
        same_in = bool(np.array_equal(wl["x"][sel_t].cpu().numpy(), x) and np.array_equal(wl["lik"][sel_t].cpu().numpy(), lik))
        differing = int(((dn != nx) | (dm != m) | (np.isnan(dQ).any(1) != np.isnan(q).any(1))).sum())
        ok = ~np.isnan(q).any(1) & ~np.isnan(dQ).any(1)
        max_dq = float(np.abs(dQ[ok] - q[ok]).max(initial=0.0))
        hist = np.zeros(33, np.int64)
        hist[:len(st["level_hist"])] = st["level_hist"]
        vec = torch.tensor([ms, gen_ms, max_dq], dtype=torch.float64, device=dev)
        cnt = torch.tensor([U, st["candidates"], st["failed"], n_chk, differing, int(same_in)] + hist.tolist(), dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(vec, op=dist.ReduceOp.MAX)
            mn = torch.tensor([int(same_in)], dtype=torch.int64, device=dev)
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            same_in = bool(mn.item())
        ms, gen_ms, max_dq = (float(v) for v in vec.cpu())
        cnt = cnt.cpu().numpy()
        if rank == 0:
            lh = cnt[6:].tolist()
            while len(lh) > 1 and lh[-1] == 0:
                lh.pop()
            line = {"bench": "cfg5_sweep", "n_gpus": world, "cams": C, "min_cams": mc, "frames": args.frames, "units": int(cnt[0]),
                    "kernel_ms": ms, "units_per_s": float(cnt[0]) / ms * 1e3, "candidates_per_unit": float(cnt[1]) / float(cnt[0]),
                    "candidates_per_s": float(cnt[1]) / ms * 1e3, "level_hist": lh, "failed_units": int(cnt[2]),
                    "search_cap": "min_cameras = max(2, C - 4): exclusion level <= 4",
                    "inputs": "generated on the device per shard (p2s_synth_observations_device), not timed in kernel_ms",
                    "generator_ms": gen_ms, "generator_units_per_s_per_gpu": U / gen_ms * 1e3 if gen_ms else None,
                    "oracle_checked_units": int(cnt[3]), "oracle_differing_decisions": int(cnt[4]), "oracle_max_abs_dQ_m": max_dq,
                    "device_inputs_equal_numpy_twin": same_in, "grid": eng.last_grid()}
            print(json.dumps(line), flush=True)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "sweep_cfg5.jsonl"), "a") as f:
                f.write(json.dumps(line) + "\n")
        del wl, out
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
