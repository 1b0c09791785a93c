#!/usr/bin/env python
"""Multi-person matching benchmark (BASELINE.json configs[3] in `multi_person = true` mode: 8 cameras x 6
persons per camera => 48 detections per frame, ray affinity + matchSVT + arg-max rows per frame).

    python tests/perf/mp_bench.py [frames] [persons] [cams]

Prints one JSON line (also appended to gpurun_out/mp_bench.jsonl): frames/s of `mp_associate_kernel` with
the inputs resident in HBM, the same through `p2s_associate_multi_host`, the NumPy restatement's frames/s on
one host core over a sample, and the integer-output parity on that sample."""
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
    import p2s_oracle_mp as omp
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    Np = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    C = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    d_max, min_aff = 0.1, 0.2
    w = synth.make_multi_person_workload(C, F, Np, seed=404)
    n_max = max(1, int(w["count"].sum(axis=1).max()))
    eng = ops.get_engine(0)
    d_obs = torch.from_numpy(w["obs"]).cuda()
    d_cnt = torch.from_numpy(w["count"]).cuda()
    out = eng.associate_multi(d_obs, d_cnt, w["models"], d_max, min_aff, n_max)
    torch.cuda.synchronize()
    iters = out["iters"].cpu().numpy()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 3
    e0.record()
    for _ in range(steps):
        eng.associate_multi(d_obs, d_cnt, w["models"], d_max, min_aff, n_max)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t0 = time.perf_counter()
    ho = eng.associate_multi_host(w["obs"], w["count"], w["models"], d_max, min_aff, n_max=n_max)
    host_ms = (time.perf_counter() - t0) * 1e3
    cams = omp.camera_ray_params(w["models"])
    sample = min(F, 100)
    bad = 0
    t0 = time.perf_counter()
    for f in range(sample):
        det = [[w["obs"][f, c, p].astype(float) for p in range(w["count"][f, c])] for c in range(C)]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            aff, cum = omp.frame_affinity(det, cams, d_max, min_aff)
        bad += int(not np.array_equal(ho["rows"][f, :cum[-1]], omp.argmax_rows(aff, cum)))
    cpu_s = time.perf_counter() - t0
    line = {"bench": "associate_multi_person", "frames": F, "cams": C, "persons_per_cam": Np, "n_max": n_max,
            "kernel_ms": ms, "frames_per_s": F / ms * 1e3, "svt_iterations_mean": float(iters.mean()),
            "host_api_ms": host_ms, "host_api_frames_per_s": F / host_ms * 1e3,
            "numpy_oracle_frames_per_s_1core": sample / cpu_s, "oracle_checked_frames": sample, "oracle_mismatches": bad,
            "grid": eng.last_grid()}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "mp_bench.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
