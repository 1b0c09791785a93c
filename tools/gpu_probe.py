"""Development probe (not part of the product): device-resident timing of the triangulation kernel
at a BASELINE config size, FP64 peak, level histogram.  Writes JSON under gpurun_out/."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pose2sim_b200 import ops, synth  # noqa: E402


def main():
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
    mc = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    solver = sys.argv[4] if len(sys.argv) > 4 else "secular"
    eng = ops.get_engine(0)
    eng.set_solver(solver)
    res = {"device": eng.info, "C": C, "F": F, "min_cams": mc, "solver": solver}
    res["fp64_peak_tflops"] = eng.fp64_peak()
    t0 = time.time()
    wl = synth.make_triangulation_workload(C, F, 1, 26, seed=202 if C == 8 else 300 + C)
    res["gen_s"] = time.time() - t0
    U = wl["x"].shape[0]
    x, y, lik = (torch.from_numpy(wl[k]).cuda() for k in ("x", "y", "lik"))
    obs = eng.stage_observations(x, y, lik, None)
    stats = eng.new_stats()
    out = eng.triangulate(obs, wl["P"], 15.0, mc, stats=stats)
    torch.cuda.synchronize()
    res["stats"] = ops.stats_dict(stats.cpu().numpy())
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    times, stimes = [], []
    for it in range(8):
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        eng.stage_observations(x, y, lik, None, out=obs)
        e1.record()
        eng.triangulate(obs, wl["P"], 15.0, mc, out=out)
        e2.record()
        torch.cuda.synchronize()
        stimes.append(e0.elapsed_time(e1))
        times.append(e1.elapsed_time(e2))
    res["stage_ms"] = stimes
    res["tri_ms"] = times
    best = min(times[2:])
    res["units"] = U
    res["units_per_s_kernel"] = U / (best * 1e-3)
    res["cands_per_s"] = res["stats"]["candidates"] / (best * 1e-3)
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"probe_C{C}_F{F}_{solver}.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
