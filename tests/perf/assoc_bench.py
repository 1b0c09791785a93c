#!/usr/bin/env python
"""Association-search benchmark (BASELINE.json configs[3]: Demo_MultiPerson-shaped synthetic, 8 cameras x
6 persons per camera => 6^8 = 1 679 616 person combinations per frame, single-person search mode).

    python tests/perf/assoc_bench.py [frames] [persons] [cams]

Prints one JSON line (also appended to gpurun_out/assoc_bench.jsonl): frames/s, combination rows/s and
candidate solves/s of `associate_kernel` with the inputs resident in HBM, the same through
`p2s_associate_host`, and a NumPy-oracle parity check on the first frames that are cheap enough."""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    from pose2sim_b200 import ops, synth
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    Np = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    C = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    thr, lik_thr, mc = 20.0, 0.3, 2
    wl = synth.make_association_workload(C, F, Np, seed=404)
    eng = ops.get_engine(0)
    obs4 = np.zeros((F, C, Np, 4), np.float32)
    obs4[..., :3] = wl["obs"]
    d_obs = torch.from_numpy(obs4).cuda()
    d_cnt = torch.from_numpy(wl["count"]).cuda()
    out = eng.associate(d_obs, d_cnt, wl["P"], thr, lik_thr, mc, want_stats=True)
    torch.cuda.synchronize()
    st = out["stats"].cpu().numpy().astype(np.int64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    e0.record()
    for _ in range(steps):
        eng.associate(d_obs, d_cnt, wl["P"], thr, lik_thr, mc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t0 = time.perf_counter()
    ho = eng.associate_host(obs4, wl["count"], wl["P"], thr, lik_thr, mc)
    host_ms = (time.perf_counter() - t0) * 1e3
    # parity on frames whose search is short enough for the per-candidate NumPy oracle
    import p2s_oracle as orc
    cheap = np.flatnonzero(st[:, 0] <= 300)[:20]
    bad = 0
    t_np = time.perf_counter()
    for f in cheap:
        ob = [[wl["obs"][f, c, p].astype(float) for p in range(wl["count"][f, c])] for c in range(C)]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            e, comb, q = orc.associate_frame(ob, list(wl["count"][f]), wl["P"], thr, lik_thr, mc)
        bad += int(not np.array_equal(ho["comb"][f].astype(int), np.nan_to_num(comb, nan=-1).astype(int)))
    t_np = time.perf_counter() - t_np
    # full-size parity against the plain-C oracle (OpenMP) when the search space is small enough for it
    c_checked = c_bad = 0
    c_maxdq, t_c = 0.0, None
    if Np ** C <= 10_000:
        import c_oracle as co
        t_c = time.perf_counter()
        ce, cc, cq = co.associate_frames(wl["obs"], wl["count"], wl["P"], thr, lik_thr, mc)
        t_c = time.perf_counter() - t_c
        c_checked = F
        c_bad = int((cc != ho["comb"]).any(axis=1).sum())
        both = np.isfinite(cq).all(axis=1) & np.isfinite(ho["Q"]).all(axis=1)
        c_maxdq = float(np.abs(cq[both] - ho["Q"][both]).max(initial=0.0))
    line = {"bench": "associate", "frames": F, "c_oracle_checked_frames": c_checked, "c_oracle_mismatching_frames": c_bad,
            "c_oracle_max_abs_dQ_m": c_maxdq,
            "cpu_c_oracle_frames_per_s": (F / t_c) if t_c else None, "cpu_c_oracle_threads": (co.max_threads() if t_c else None),
            "cpu_numpy_oracle_frames_per_s_1core_cheapest_frames": (len(cheap) / t_np) if len(cheap) else None, "cams": C, "persons_per_cam": Np, "rows_per_frame_full": Np ** C,
            "kernel_ms": ms, "frames_per_s": F / ms * 1e3, "rows_visited_per_frame": float(st[:, 0].mean()),
            "rows_per_s": float(st[:, 0].sum()) / ms * 1e3, "candidate_solves_per_s": float(st[:, 1].sum()) / ms * 1e3,
            "host_api_ms": host_ms, "host_api_frames_per_s": F / host_ms * 1e3,
            "under_threshold": float((ho["err"] < thr).mean()), "oracle_checked_frames": int(len(cheap)), "oracle_mismatches": bad,
            "grid": eng.last_grid()}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "assoc_bench.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
