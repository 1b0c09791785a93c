#!/usr/bin/env python
"""Full-size parity report of the triangulation path (SURVEY.md §8(d) "Parity report per run"): the CUDA path
through `p2s_triangulate_host` against the plain-C oracle (oracle/p2s_oracle.c, pinned by the reference's golden
vectors) on EVERY unit of a BASELINE workload.

    python tests/perf/parity_report.py [cfg2|cfg3|cfg5-<cams>] [frames]      (cfg5-C: min_cameras = max(2, C - 4), seed 500 + C)

One JSON line (also gpurun_out/parity_report.jsonl): max / percentiles of |dQ| (m) and |d err| (px), units whose
nb_cams_excluded / id_excluded_cams / NaN-ness differ, and how many of those sit inside the eps-band
(|error - threshold| < eps at an evaluated level, or best and runner-up candidate errors closer than eps)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import bench
    import c_oracle as co
    from pose2sim_b200 import ops, synth
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    if name.startswith("cfg5-"):
        Cn = int(name[5:])
        cfg = dict(C=Cn, F=2000, N=1, K=26, seed=500 + Cn, thr=15.0, min_cams=max(2, Cn - 4), lik_thr=0.3,
                   name=f"cfg5 sweep point: synthetic {Cn} cams x HALPE_26, min_cameras = {max(2, Cn - 4)}")
    else:
        cfg = bench.WORKLOADS[name]
    F = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["F"]
    wl = synth.make_triangulation_workload(cfg["C"], F, cfg["N"], cfg["K"], seed=cfg["seed"], lik_thr=None)
    eng = ops.get_engine(0)
    eps = 1e-6
    eng.set_band_eps(eps)
    out = eng.triangulate_host(wl["x"], wl["y"], wl["lik"], wl["P"], cfg["lik_thr"], cfg["thr"], cfg["min_cams"])
    x, y, w = synth.gate_likelihood(wl["x"], wl["y"], wl["lik"], cfg["lik_thr"])
    t0 = time.perf_counter()
    Q, err, nexcl, mask, level, ncand = co.triangulate_units(x, y, w, wl["P"], cfg["thr"], cfg["min_cams"])
    t_cpu = time.perf_counter() - t0
    U = len(err)
    nan_diff = np.isnan(Q).any(axis=1) != np.isnan(out["Q"]).any(axis=1)
    dec_diff = (nexcl != out["nexcl"]) | (mask != out["mask"]) | nan_diff
    ok = ~np.isnan(Q).any(axis=1) & ~np.isnan(out["Q"]).any(axis=1) & ~dec_diff
    dq = np.abs(Q[ok] - out["Q"][ok]).max(axis=1)
    de = np.abs(err[ok] - out["err"][ok])
    # a differing decision is acceptable only inside the eps-band of the ORACLE's error around the threshold
    in_band = dec_diff & (np.abs(np.nan_to_num(err, nan=np.inf) - cfg["thr"]) < eps)
    st = out["stats"]
    line = {"tool": "parity_report", "workload": cfg["name"], "units": int(U), "oracle": "oracle/p2s_oracle.c (OpenMP)",
            "oracle_s": t_cpu, "oracle_threads": co.max_threads(), "candidates_gpu": st["candidates"], "candidates_oracle": int(ncand),
            "compared_units": int(ok.sum()),
            "dQ_m": {"max": float(dq.max(initial=0.0)), "p999": float(np.percentile(dq, 99.9)), "p99": float(np.percentile(dq, 99)),
                     "p50": float(np.percentile(dq, 50))},
            "derr_px": {"max": float(de.max(initial=0.0)), "p99": float(np.percentile(de, 99)), "p50": float(np.percentile(de, 50))},
            "units_with_differing_decision": int(dec_diff.sum()), "of_which_inside_eps_band": int(in_band.sum()),
            "eps_px": eps, "gpu_band_threshold_units": st["band_threshold"], "gpu_band_argmin_units": st["band_argmin"],
            "level_hist": st["level_hist"], "failed_units": st["failed"], "tolerance_m": 1e-6,
            "within_tolerance": bool(dq.max(initial=0.0) <= 1e-6 and int(dec_diff.sum()) == int(in_band.sum()))}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
