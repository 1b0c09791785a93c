#!/usr/bin/env python
"""`handle_LR_swap = true` benchmark: cfg2-shaped synthetic (8 cameras x HALPE_26) where in a fraction of the
(frame, camera) views the pose estimator swapped the left and right limbs — the situation the flag exists for.

    python tests/perf/lrswap_bench.py [frames] [swap_fraction] [threshold_px] [numpy_sample_units]

Prints one JSON line (also appended to gpurun_out/lrswap_bench.jsonl): units/s of `lrswap_kernel`
(p2s_lrswap.cu, staged input resident in HBM, CUDA events on the launching stream), the main kernel on the same
staged buffer beside it (what the flag costs), EVERY unit against the plain-C oracle (OpenMP), and the NumPy
oracle's rate and parity on a bounded sample on one core (decisions bit-exact, Q / error within 1e-6)."""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def swap_limbs(wl, partner, frac, seed):
    """Exchange the observations of every keypoint with its partner's in a random `frac` of the (frame, camera) views."""
    F, K, C = wl["F"], wl["K"], wl["C"]
    g = np.random.default_rng(seed)
    sw = g.random((F, 1, C)) < frac                                           # [F, 1, C]
    part = np.asarray(partner)
    out = {}
    for k in ("x", "y", "lik"):
        a = wl[k].reshape(F, K, C)
        out[k] = np.ascontiguousarray(np.where(sw, a[:, part, :], a).reshape(F * K, C))
    return out, float(sw.mean())


def main():
    import torch
    from pose2sim_b200 import ops, skeletons, synth
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.15
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 15.0
    n_check = int(sys.argv[4]) if len(sys.argv) > 4 else 2600
    C, mc, lik_thr = 8, 2, 0.3
    names = skeletons.keypoints("HALPE_26")[1]
    partner = skeletons.swapped_indices(names)
    wl = synth.make_triangulation_workload(C, F, 1, len(names), seed=606, lik_thr=None)
    planes, frac_seen = swap_limbs(wl, partner, frac, 607)
    U = F * len(names)
    eng = ops.get_engine(0)
    d = [torch.from_numpy(planes[k]).cuda() for k in ("x", "y", "lik")]
    obs = eng.stage_observations(*d, lik_thr)
    res = eng.triangulate_lr_swap(obs, partner, wl["P"], thr, mc)
    base = eng.triangulate(obs, wl["P"], thr, mc)
    torch.cuda.synchronize()

    def timed(fn, steps=5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ms_swap = timed(lambda: eng.triangulate_lr_swap(obs, partner, wl["P"], thr, mc))
    ms_main = timed(lambda: eng.triangulate(obs, wl["P"], thr, mc))
    changed = int((~torch.isclose(res["Q"], base["Q"], atol=1e-9, rtol=0, equal_nan=True).all(dim=1)).sum())

    gQ, gerr = res["Q"].cpu().numpy(), res["err"].cpu().numpy()
    gn, gm = res["nexcl"].cpu().numpy(), res["mask"].cpu().numpy().view(np.uint32)
    # every unit against the plain-C oracle (gate applied like the stage kernel: float32 values, compared in float64)
    import c_oracle as co
    x, y, w = (planes[k].copy() for k in ("x", "y", "lik"))
    low = w.astype(np.float64) < lik_thr
    x[low] = np.nan; y[low] = np.nan; w[low] = np.nan
    t0 = time.perf_counter()
    cQ, cerr, cnexcl, cmask = co.triangulate_units_lr_swap(x, y, w, partner, wl["P"], thr, mc)
    t_c = time.perf_counter() - t0
    c_differing = int(((gn != cnexcl) | (gm != cmask) | (np.isnan(gerr) != np.isnan(cerr))).sum())
    c_both = np.isfinite(cQ).all(axis=1) & np.isfinite(gQ).all(axis=1)
    c_max_dq = float(np.abs(cQ[c_both] - gQ[c_both]).max(initial=0.0))
    # bounded sample: the first n_check units (whole frames), NumPy oracle on one core
    import p2s_oracle as orc
    n_check = min(n_check - n_check % len(names), U)
    xs, ys, ws = (a[:n_check].astype(np.float64) for a in (x, y, w))
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        oQ, oerr, onexcl, omask = orc.triangulate_units(xs, ys, ws, wl["P"], thr, mc, partner=partner)
    t_cpu = time.perf_counter() - t0
    differing = int(((gn[:n_check] != onexcl) | (gm[:n_check] != omask) | (np.isnan(gerr[:n_check]) != np.isnan(oerr))).sum())
    both = np.isfinite(oQ).all(axis=1) & np.isfinite(gQ[:n_check]).all(axis=1)
    gQ, gerr = gQ[:n_check], gerr[:n_check]
    line = {"bench": "lr_swap", "cams": C, "keypoints": len(names), "frames": F, "units": U, "threshold_px": thr,
            "min_cameras": mc, "swapped_view_fraction": frac_seen,
            "units_changed_by_the_swapped_pass": changed, "units_triangulated": float(torch.isfinite(res["err"]).float().mean()),
            "lrswap_kernel_ms": ms_swap, "lrswap_units_per_s": U / ms_swap * 1e3,
            "main_kernel_same_buffer_ms": ms_main, "main_units_per_s": U / ms_main * 1e3,
            "c_oracle_checked_units": U, "c_oracle_units_with_differing_decision": c_differing, "c_oracle_max_abs_dQ_m": c_max_dq,
            "cpu_c_oracle_units_per_s": U / t_c, "cpu_c_oracle_threads": co.max_threads(),
            "cpu_numpy_oracle_units_per_s_1core": (n_check / t_cpu) if n_check else None, "cpu_sample_units": n_check,
            "parity_units_with_differing_decision": differing,
            "parity_max_abs_dQ_m": float(np.abs(oQ[both] - gQ[both]).max(initial=0.0)),
            "parity_max_abs_derr_px": float(np.nanmax(np.abs(np.where(np.isfinite(oerr) & np.isfinite(gerr), oerr - gerr, 0.0)), initial=0.0))}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "lrswap_bench.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
