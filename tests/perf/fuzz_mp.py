#!/usr/bin/env python
"""Randomised differential test of the multi-person matching: CUDA (p2s_associate_multi_host — every team width of
`mp_associate_kernel`: 128 / 256 / 512 threads for <= 16 / 32 / 64 detections per frame) against the NumPy restatement
of personAssociation.py:277-549 (oracle/p2s_oracle_mp.py) over random camera counts, persons per camera, missing
detections, noise, outliers, undetected joints, `max_distance`, `min_affinity` and `min_cameras`.

    python tests/perf/fuzz_mp.py [cases] [seed]

Per frame: the matched affinity within 1e-9 and, unless the frame is a tie (two best detections of a view closer than
1e-7 in the oracle's affinity — the arg-max is then the SVD's rounding), identical arg-max rows and proposals.
One JSON line (also gpurun_out/fuzz_mp.jsonl)."""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def is_tie(aff, cum):
    for v in range(len(cum) - 1):
        seg = aff[:, cum[v]:cum[v + 1]]
        if seg.shape[1] >= 2:
            top = np.sort(seg, axis=1)[:, -2:]
            if bool(((top[:, 1] - top[:, 0] < 1e-7) & (top[:, 1] > 0)).any()):
                return True
    return False


def main():
    import p2s_oracle_mp as omp
    from pose2sim_b200 import multi_person as mp
    from pose2sim_b200 import ops, synth
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4242
    g = np.random.default_rng(seed)
    eng = ops.get_engine(0)
    frames = ties = bad_rows = bad_props = 0
    worst = 0.0
    widths = {128: 0, 256: 0, 512: 0}
    offenders = []
    for case in range(n_cases):
        C = int(g.integers(2, 13))
        Np = int(g.integers(1, max(2, min(8, 64 // C) + 1)))
        F = int(g.integers(1, 13))
        K = int(g.choice([17, 26, 5]))
        d_max = float(g.choice([0.05, 0.1, 0.3]))
        min_aff = float(g.choice([0.0, 0.2, 0.6]))
        mc = int(g.integers(2, min(C, 4) + 1))
        w = synth.make_multi_person_workload(C, F, Np, seed=int(g.integers(1, 1 << 30)), K=K, sigma=float(g.choice([0.5, 2.0, 8.0])),
                                             p_out=float(g.choice([0.0, 0.05, 0.3])), p_low=float(g.choice([0.0, 0.05, 0.4])),
                                             p_missing=float(g.choice([0.0, 0.25, 0.6])), p_nan=float(g.choice([0.0, 0.02, 0.3])))
        n_max = max(1, int(w["count"].sum(axis=1).max()))
        widths[128 if n_max <= 16 else 256 if n_max <= 32 else 512] += 1
        out = eng.associate_multi_host(w["obs"], w["count"], w["models"], d_max, min_aff, n_max=n_max, want_affinity=True)
        cams = omp.camera_ray_params(w["models"])
        for f in range(F):
            det = [[w["obs"][f, c, p].astype(float) for p in range(w["count"][f, c])] for c in range(C)]
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                aff, cum = omp.frame_affinity(det, cams, d_max, min_aff)
            n = int(cum[-1])
            frames += 1
            d = float(np.abs(out["affinity"][f, :n, :n] - aff).max(initial=0.0))
            worst = max(worst, d)
            if is_tie(aff, cum):
                ties += 1
                continue
            r_ok = np.array_equal(out["rows"][f, :n], omp.argmax_rows(aff, cum)) and bool((out["rows"][f, n:] == -1).all())
            p_ok = np.array_equal(mp.proposals_from_rows(out["rows"][f, :n], mc), omp.proposals_from_affinity(aff, cum, mc),
                                  equal_nan=True)
            bad_rows += int(not r_ok)
            bad_props += int(not p_ok)
            if (not r_ok or not p_ok or d > 1e-9) and len(offenders) < 5:
                offenders.append({"case": case, "frame": f, "C": C, "persons": Np, "K": K, "d_max": d_max, "min_affinity": min_aff,
                                  "min_cams": mc, "n": n, "max_abs_d_affinity": d, "rows_equal": r_ok, "proposals_equal": p_ok})
    line = {"tool": "fuzz_mp", "cases": n_cases, "seed": seed, "frames": frames, "tie_frames_not_compared": ties,
            "cases_per_team_width": widths, "frames_with_differing_rows": bad_rows, "frames_with_differing_proposals": bad_props,
            "max_abs_d_affinity": worst, "offenders": offenders, "ok": bad_rows == 0 and bad_props == 0 and worst <= 1e-9}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fuzz_mp.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
