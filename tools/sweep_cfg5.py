#!/usr/bin/env python
"""BASELINE.json configs[4] ("stress sweep 4-32 cams, combination explosion"): the fused triangulation
kernel over camera counts C with min_cameras = max(2, C - 4) (search capped at level 4: at most
sum_k<=4 C(32,k) = 41 449 candidates per unit), seeds 500 + C.  Frames per C are scaled
down from the 10 M of the config to what one B200 call needs to show the rate (the kernel is persistent
and tile-scheduled, so units/s does not depend on F beyond a few 100 k units).

    python tools/sweep_cfg5.py [frames]        -> one JSON line per C, appended to gpurun_out/sweep_cfg5.jsonl
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from pose2sim_b200 import ops, synth
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
    eng = ops.get_engine(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for C in (4, 6, 8, 12, 16, 24, 32):
        mc = max(2, C - 4)
        wl = synth.make_triangulation_workload(C, F, 1, 26, seed=500 + C, lik_thr=None)
        x, y, lik = (torch.from_numpy(wl[k]).cuda() for k in ("x", "y", "lik"))
        stats = eng.new_stats()
        out = eng.triangulate_planes(x, y, lik, wl["P"], 0.3, 15.0, mc, stats=stats)
        torch.cuda.synchronize()
        st = ops.stats_dict(stats.cpu().numpy())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 5
        e0.record()
        for _ in range(steps):
            eng.triangulate_planes(x, y, lik, wl["P"], 0.3, 15.0, mc, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        U = x.shape[0]
        ok = torch.isfinite(out["err"])
        truth = torch.from_numpy(wl["truth"]).cuda()
        med = float(torch.median(torch.linalg.norm(out["Q"][ok] - truth[ok], dim=1)).item())
        line = {"bench": "cfg5_sweep", "cams": C, "min_cams": mc, "frames": F, "units": U, "kernel_ms": ms,
                "units_per_s": U / ms * 1e3, "candidates_per_unit": st["candidates"] / U,
                "candidates_per_s": st["candidates"] / ms * 1e3, "level_hist": st["level_hist"],
                "failed_units": st["failed"], "triangulated_fraction": float(ok.float().mean().item()),
                "median_error_vs_truth_m": med}
        print(json.dumps(line), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "sweep_cfg5.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
