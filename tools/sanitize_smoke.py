#!/usr/bin/env python
"""Small invocations of every kernel for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
all camera-count templates, ragged tiles, deep levels, the unranked (beyond-table) path is not reachable
at these sizes; association with missing detections; undistort mode."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from pose2sim_b200 import calib, ops, synth
    eng = ops.get_engine(0)
    for C, mc in ((2, 2), (3, 2), (4, 2), (5, 3), (8, 2), (12, 8), (16, 3), (32, 28)):
        wl = synth.make_triangulation_workload(C, 7, 1, 26, seed=C, lik_thr=None, p_out=0.15, p_low=0.1)
        U = 26 * 7 - 5
        x, y, lik = (torch.from_numpy(np.ascontiguousarray(wl[k][:U])).cuda() for k in ("x", "y", "lik"))
        st = eng.new_stats()
        a = eng.triangulate_planes(x, y, lik, wl["P"], 0.3, 15.0, mc, stats=st)
        b = eng.triangulate(eng.stage_observations(x, y, lik, 0.3), wl["P"], 15.0, mc)
        torch.cuda.synchronize()
        assert torch.equal(a["mask"], b["mask"])
        h = eng.triangulate_host(wl["x"][:U], wl["y"][:U], wl["lik"][:U], wl["P"], 0.3, 15.0, mc)
        print("tri C", C, ops.stats_dict(st.cpu().numpy())["level_hist"], int(np.isfinite(h["err"]).sum()))
    for C, Np in ((4, 3), (8, 2), (16, 2)):
        aw = synth.make_association_workload(C, 9, Np, seed=C, p_out=0.2, p_low=0.1, p_missing=0.2)
        o = eng.associate_host(aw["obs"], aw["count"], aw["P"], 20.0, 0.3, 2, want_stats=True)
        print("assoc C", C, int((o["err"] < 20).sum()), int(o["stats"][:, 1].sum()))
    P, Ks, Rs, ts = synth.ring_cameras(4)
    lens = [{"K": Ks[c], "dist": [-0.05, 0.02, 1e-3, -5e-4, 0.01], "R": Rs[c], "T": ts[c],
             "newK": calib.optimal_new_camera_matrix(Ks[c], [-0.05, 0.02, 1e-3, -5e-4, 0.01], (1080, 1920))} for c in range(4)]
    wl = synth.make_triangulation_workload(4, 5, 1, 26, seed=9, lik_thr=None)
    h = eng.triangulate_host(wl["x"], wl["y"], wl["lik"], wl["P"], 0.3, 15.0, 2, lens=lens)
    print("undistort", int(np.isfinite(h["err"]).sum()))
    eng.set_solver("jacobi")
    h = eng.triangulate_host(wl["x"], wl["y"], wl["lik"], wl["P"], 0.3, 15.0, 2)
    eng.set_solver("secular")
    print("jacobi", int(np.isfinite(h["err"]).sum()), "fp64 peak", round(eng.fp64_peak(), 1))
    eng.close()
    print("sanitize smoke done")


if __name__ == "__main__":
    main()
