#!/usr/bin/env python
"""Randomised differential test of the single-person association search: CUDA (p2s_associate_host, both team widths)
against the plain-C oracle over random camera counts, persons per camera (incl. cameras without anybody), thresholds,
min_cameras, outlier rates and likelihood gates.

    python tests/perf/fuzz_assoc.py [cases] [seed]

One JSON line (also gpurun_out/fuzz_assoc.jsonl)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import c_oracle as co
    from pose2sim_b200 import ops, synth
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4242
    g = np.random.default_rng(seed)
    eng = ops.get_engine(0)
    frames = bad = 0
    worst = 0.0
    offenders = []
    for case in range(n_cases):
        C = int(g.integers(2, 9))
        Np = int(g.integers(1, 5 if C <= 6 else 4))
        F = int(g.integers(1, 60))
        thr = float(g.choice([2.0, 5.0, 20.0, 1e-3, 1e5]))
        lik_thr = float(g.choice([0.3, 0.0, 0.55]))
        mc = int(g.integers(1, min(C, 4) + 1))
        wl = synth.make_association_workload(C, F, Np, seed=int(g.integers(1, 1 << 30)), sigma=float(g.choice([0.5, 2.0, 6.0])),
                                             p_out=float(g.choice([0.0, 0.1, 0.4])), p_low=float(g.choice([0.0, 0.1, 0.5])),
                                             p_missing=float(g.choice([0.0, 0.3])))
        obs = wl["obs"].copy()
        m = g.random(obs.shape[:3])
        obs[..., 2][m < 0.03] = 0.0
        obs[..., 2][(m > 0.03) & (m < 0.05)] = np.nan
        team = int(g.choice([0, 1, 8]))
        eng.set_assoc_team(team)
        out = eng.associate_host(obs, wl["count"], wl["P"], thr, lik_thr, mc)
        ce, cc, cq = co.associate_frames(obs, wl["count"], wl["P"], thr, lik_thr, mc)
        diff = (cc != out["comb"]).any(axis=1) | (np.isinf(ce) != np.isinf(out["err"]))
        both = np.isfinite(cq).all(axis=1) & np.isfinite(out["Q"]).all(axis=1) & ~diff
        dq = float((np.abs(cq[both] - out["Q"][both]).max(axis=1) / np.maximum(1.0, np.abs(cq[both]).max(axis=1))).max(initial=0.0))
        # a differing choice is acceptable only when the oracle's best error sits within 1e-6 px of the threshold
        band = diff & (np.abs(np.nan_to_num(ce, nan=np.inf, posinf=np.inf) - thr) < 1e-6)
        frames += F
        bad += int(diff.sum()) - int(band.sum())
        worst = max(worst, dq)
        if (int(diff.sum()) > int(band.sum()) or dq > 1e-6) and len(offenders) < 5:
            f = int(np.flatnonzero(diff)[0]) if diff.any() else -1
            offenders.append({"case": case, "C": C, "persons": Np, "F": F, "thr": thr, "lik_thr": lik_thr, "min_cams": mc, "team": team,
                              "differing": int(diff.sum()), "max_rel_dQ": dq,
                              "first": None if f < 0 else {"f": f, "count": wl["count"][f].tolist(), "oracle": [float(ce[f]), cc[f].tolist()],
                                                           "gpu": [float(out["err"][f]), out["comb"][f].tolist()]}})
    eng.set_assoc_team(0)
    line = {"tool": "fuzz_assoc", "cases": n_cases, "seed": seed, "frames": frames, "frames_with_unexplained_difference": bad,
            "max_rel_abs_dQ": worst, "offenders": offenders, "ok": bad == 0 and worst <= 1e-6}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fuzz_assoc.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
