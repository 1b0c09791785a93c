#!/usr/bin/env python
"""Randomised differential test of the triangulation path: CUDA (through p2s_triangulate_host, pinned and pageable
buffers alternating) against the plain-C oracle over random camera counts, thresholds, min_cameras, outlier / NaN /
zero-likelihood rates and ragged sizes.

    python tests/perf/fuzz_parity.py [cases] [seed]

One JSON line (also gpurun_out/fuzz_parity.jsonl): cases run, units compared, units with a differing decision (and how
many of those inside the eps-band), max |dQ|; the first few offending cases are printed for reproduction."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    import c_oracle as co
    from pose2sim_b200 import ops, synth
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    # > 0: duplicate camera 0 as camera 1 in that share of the cases.  Two identical cameras alone are a rank-deficient
    # problem (every point of the ray has zero error), so Q and the arg-min among the zero-error candidates are decided
    # by rounding noise on BOTH sides — useful to see that nothing crashes, not a parity criterion.  Default off.
    dup_rate = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    g = np.random.default_rng(seed)
    eng = ops.get_engine(0)
    eps = 1e-6
    tot_units = bad = in_band = 0
    worst = 0.0
    offenders = []
    for case in range(n_cases):
        C = int(g.choice([2, 3, 4, 5, 6, 7, 8, 9, 12, 13, 16, 17, 24, 32], p=[.04, .08, .12, .08, .1, .06, .16, .05, .08, .04, .08, .03, .04, .04]))
        deep = C <= 12
        mc = int(g.integers(1, C + 1)) if deep and g.random() < 0.7 else max(1, C - int(g.integers(0, 5)))
        thr = float(g.choice([2.0, 5.0, 15.0, 30.0, 1e-3, 1e6]))
        lik_thr = float(g.choice([0.3, 0.0, 0.55, 0.30000001192092896]))
        F = int(g.integers(1, 120 if C <= 16 else 30))
        wl = synth.make_triangulation_workload(C, F, 1, 26, seed=int(g.integers(1, 1 << 30)), lik_thr=None,
                                               sigma=float(g.choice([0.0, 0.5, 2.0, 6.0])), p_out=float(g.choice([0.0, 0.05, 0.2, 0.5])),
                                               p_low=float(g.choice([0.0, 0.05, 0.3, 0.8])))
        U = int(g.integers(1, F * 26 + 1))
        x, y, lik = (np.ascontiguousarray(wl[k][:U]).copy() for k in ("x", "y", "lik"))
        m = g.random((U, C))
        lik[m < 0.03] = 0.0                                       # zero likelihoods: counted, not listed
        lik[(m > 0.03) & (m < 0.05)] = np.nan
        x[(m > 0.05) & (m < 0.06)] = np.nan                       # NaN coordinate with a valid likelihood: camera stays valid
        if C >= 3 and g.random() < dup_rate:                      # duplicated cameras: exactly tied candidates
            P = wl["P"].copy(); P[1] = P[0]; x[:, 1] = x[:, 0]; y[:, 1] = y[:, 0]; lik[:, 1] = lik[:, 0]
        else:
            P = wl["P"]
        if g.random() < 0.5:
            hx, hy, hl = (torch.from_numpy(a).pin_memory().numpy() for a in (x, y, lik))
        else:
            hx, hy, hl = x, y, lik
        out = eng.triangulate_host(hx, hy, hl, P, lik_thr, thr, mc)
        gx, gy, gl = synth.gate_likelihood(x, y, lik, lik_thr)
        Q, err, nexcl, mask, level, ncand = co.triangulate_units(gx, gy, gl, P, thr, mc)
        nan_diff = np.isnan(Q).any(axis=1) != np.isnan(out["Q"]).any(axis=1)
        dec = (nexcl != out["nexcl"]) | (mask != out["mask"]) | nan_diff
        band = dec & (np.abs(np.nan_to_num(err, nan=np.inf) - thr) < eps)
        ok = ~np.isnan(Q).any(axis=1) & ~np.isnan(out["Q"]).any(axis=1) & ~dec
        # relative to the scale of the solution: ill-conditioned two-view units far from the cameras are legitimate
        # the normal-matrix formulation loses (w_max / w_min)^2 relative to an SVD of A: units whose valid likelihoods span
        # more than 100x (only possible with a likelihood threshold near 0) are checked for decisions, not for |dQ|
        wv = np.where(np.isnan(gl) | (gl == 0), np.nan, gl).astype(np.float64)
        with np.errstate(all="ignore"):
            spread = np.nanmax(wv, axis=1) / np.nanmin(wv, axis=1)
        ok &= ~(spread > 100.0)
        rel = np.abs(Q[ok] - out["Q"][ok]).max(axis=1) / np.maximum(1.0, np.abs(Q[ok]).max(axis=1))
        dq = float(rel.max(initial=0.0))
        wu = int(np.flatnonzero(ok)[int(np.argmax(rel))]) if ok.any() else -1
        worst = max(worst, dq)
        tot_units += U
        bad += int(dec.sum()); in_band += int(band.sum())
        if (int(dec.sum()) > int(band.sum()) or dq > 1e-6 or out["stats"]["candidates"] != ncand) and len(offenders) < 5:
            offenders.append({"case": case, "C": C, "min_cams": mc, "thr": thr, "lik_thr": lik_thr, "U": U, "differing": int(dec.sum()),
                              "in_band": int(band.sum()), "max_rel_dQ": dq, "cands_gpu": out["stats"]["candidates"], "cands_oracle": int(ncand),
                              "first_units": np.flatnonzero(dec)[:5].tolist(),
                              "worst_unit": {"u": wu, "Q_oracle": Q[wu].tolist(), "Q_gpu": out["Q"][wu].tolist(), "err": [float(err[wu]), float(out["err"][wu])],
                                             "nexcl": int(nexcl[wu]), "mask": int(mask[wu]), "lik": [None if v != v else float(v) for v in gl[wu]]} if wu >= 0 else None,
                              "detail": [{"u": int(u), "lik": [None if v != v else float(v) for v in gl[u]], "oracle": [float(err[u]), int(nexcl[u]), int(mask[u]), int(level[u])],
                                          "gpu": [float(out["err"][u]), int(out["nexcl"][u]), int(out["mask"][u])]} for u in np.flatnonzero(dec)[:3]]})
    line = {"tool": "fuzz_parity", "cases": n_cases, "seed": seed, "units": tot_units, "units_with_differing_decision": bad,
            "of_which_inside_eps_band": in_band, "max_rel_abs_dQ": worst, "offenders": offenders,
            "ok": bad == in_band and worst <= 1e-6 and not offenders}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fuzz_parity.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
