#!/usr/bin/env python
"""Randomised differential test of `handle_LR_swap = true`: `lrswap_kernel` (through the staged-buffer entry point)
against the plain-C oracle's swapped pass over random camera counts, keypoint counts and partner maps (incl. the
identity and keypoints without a partner), thresholds, min_cameras, swapped-view / outlier / NaN / zero-likelihood
rates.

    python tests/perf/fuzz_lrswap.py [cases] [seed]

One JSON line (also gpurun_out/fuzz_lrswap.jsonl): cases run, units compared, units with a differing decision (and how
many of those inside the eps-band), max relative |dQ|; the first offending cases are printed for reproduction.
GPU runs: profiles/r2a_fuzz_lrswap_gpu_seed*.json, r3f_fuzz_lrswap.jsonl; `--cpu-selfcheck` exercises the CPU half alone (workload
generation and the C oracle against the NumPy oracle)."""
import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def random_partner_map(g, K):
    """An involution on 0..K-1: random disjoint pairs, the other keypoints map to themselves (like Hip / Neck / Nose)."""
    partner = np.arange(K, dtype=np.int32)
    if g.random() < 0.15:
        return partner                                             # a missing partner name disables the swap (identity)
    idx = g.permutation(K)
    n_pairs = int(g.integers(1, K // 2 + 1))
    for a, b in idx[:2 * n_pairs].reshape(-1, 2):
        partner[a], partner[b] = b, a
    return partner


def make_case(g, synth):
    C = int(g.choice([3, 4, 5, 6, 7, 8, 9, 12, 16], p=[.08, .14, .1, .12, .08, .22, .08, .1, .08]))
    K = int(g.choice([2, 5, 17, 26]))
    mc = int(g.integers(2, C + 1)) if C <= 9 else max(2, C - int(g.integers(0, 4)))
    thr = float(g.choice([2.0, 5.0, 15.0, 30.0, 1e-3, 1e6]))
    lik_thr = float(g.choice([0.3, 0.0, 0.55]))
    F = int(g.integers(1, 60 if C <= 9 else 12))
    partner = random_partner_map(g, K)
    wl = synth.make_triangulation_workload(C, F, 1, K, seed=int(g.integers(1, 1 << 30)), lik_thr=None,
                                           sigma=float(g.choice([0.0, 0.5, 2.0, 6.0])), p_out=float(g.choice([0.0, 0.05, 0.2])),
                                           p_low=float(g.choice([0.0, 0.05, 0.3])))
    sw = g.random((F, 1, C)) < float(g.choice([0.0, 0.15, 0.5, 1.0]))
    planes = [np.ascontiguousarray(np.where(sw, wl[k].reshape(F, K, C)[:, partner, :], wl[k].reshape(F, K, C)).reshape(F * K, C))
              for k in ("x", "y", "lik")]
    x, y, lik = planes
    m = g.random(x.shape)
    lik[m < 0.03] = 0.0                                            # zero likelihoods: counted, not listed
    lik[(m > 0.03) & (m < 0.05)] = np.nan
    x[(m > 0.05) & (m < 0.06)] = np.nan                            # NaN coordinate with a valid likelihood
    return {"C": C, "K": K, "min_cams": mc, "thr": thr, "lik_thr": lik_thr, "F": F, "partner": partner, "P": wl["P"],
            "x": x, "y": y, "lik": lik}


def oracle_side(case, synth, co):
    gx, gy, gl = synth.gate_likelihood(case["x"], case["y"], case["lik"], case["lik_thr"])
    return (gx, gy, gl) + tuple(co.triangulate_units_lr_swap(gx, gy, gl, case["partner"], case["P"], case["thr"], case["min_cams"]))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n_cases = int(args[0]) if len(args) > 0 else 120
    seed = int(args[1]) if len(args) > 1 else 2468
    import c_oracle as co
    from pose2sim_b200 import synth
    g = np.random.default_rng(seed)
    if "--cpu-selfcheck" in sys.argv:                              # no GPU: the C oracle against the NumPy oracle on the same cases
        import p2s_oracle as orc
        tot = bad = 0
        for case_i in range(n_cases):
            case = make_case(g, synth)
            if case["x"].shape[0] * (2 ** min(case["C"], 12)) > 3e6:
                continue                                           # keep the per-candidate NumPy oracle to seconds
            gx, gy, gl, Q, err, nexcl, mask = oracle_side(case, synth, co)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                oQ, oerr, on, om = orc.triangulate_units(gx.astype(float), gy.astype(float), gl.astype(float), case["P"], case["thr"],
                                                         case["min_cams"], partner=case["partner"])
            d = (on != nexcl) | (om != mask) | (np.isnan(oerr) != np.isnan(err))
            tot += len(d); bad += int(d.sum())
        print(json.dumps({"tool": "fuzz_lrswap", "mode": "cpu-selfcheck (C oracle vs NumPy oracle)", "cases": n_cases, "seed": seed,
                          "units": tot, "units_with_differing_decision": bad, "ok": bad == 0}))
        return
    import torch
    from pose2sim_b200 import ops
    eng = ops.get_engine(0)
    eps = 1e-6
    tot_units = bad = in_band = 0
    worst = 0.0
    offenders = []
    for case_i in range(n_cases):
        case = make_case(g, synth)
        C, thr = case["C"], case["thr"]
        obs = eng.stage_observations(*(torch.from_numpy(case[k]).cuda() for k in ("x", "y", "lik")), case["lik_thr"])
        res = eng.triangulate_lr_swap(obs, case["partner"], case["P"], thr, case["min_cams"])
        torch.cuda.synchronize()
        out = {"Q": res["Q"].cpu().numpy(), "err": res["err"].cpu().numpy(), "nexcl": res["nexcl"].cpu().numpy(),
               "mask": res["mask"].cpu().numpy().view(np.uint32)}
        gx, gy, gl, Q, err, nexcl, mask = oracle_side(case, synth, co)
        dec = (nexcl != out["nexcl"]) | (mask != out["mask"]) | (np.isnan(err) != np.isnan(out["err"]))
        # a unit may only differ when an error the decision hangs on sits within eps of the threshold; the final error is the
        # one available here (a level's un-swapped error within eps of the threshold is the other possibility, reported as offender)
        band = dec & (np.abs(np.nan_to_num(err, nan=np.inf) - thr) < eps)
        ok = ~np.isnan(Q).any(axis=1) & ~np.isnan(out["Q"]).any(axis=1) & ~dec
        wv = np.where(np.isnan(gl) | (gl == 0), np.nan, gl).astype(np.float64)
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spread = np.nanmax(wv, axis=1) / np.nanmin(wv, axis=1)
        ok &= ~(spread > 100.0)                                    # normal-matrix formulation: see fuzz_parity.py
        rel = np.abs(Q[ok] - out["Q"][ok]).max(axis=1) / np.maximum(1.0, np.abs(Q[ok]).max(axis=1))
        dq = float(rel.max(initial=0.0))
        worst = max(worst, dq)
        tot_units += len(dec)
        bad += int(dec.sum()); in_band += int(band.sum())
        if (int(dec.sum()) > int(band.sum()) or dq > 1e-6) and len(offenders) < 5:
            offenders.append({"case": case_i, "C": C, "K": case["K"], "min_cams": case["min_cams"], "thr": thr, "lik_thr": case["lik_thr"],
                              "F": case["F"], "partner": case["partner"].tolist(), "differing": int(dec.sum()), "in_band": int(band.sum()),
                              "max_rel_dQ": dq, "first_units": np.flatnonzero(dec)[:5].tolist(),
                              "detail": [{"u": int(u), "oracle": [float(err[u]), int(nexcl[u]), int(mask[u])],
                                          "gpu": [float(out["err"][u]), int(out["nexcl"][u]), int(out["mask"][u])]} for u in np.flatnonzero(dec)[:3]]})
    line = {"tool": "fuzz_lrswap", "cases": n_cases, "seed": seed, "units": tot_units, "units_with_differing_decision": bad,
            "of_which_inside_eps_band": in_band, "max_rel_abs_dQ": worst, "offenders": offenders,
            "ok": bad == in_band and worst <= 1e-6 and not offenders}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "fuzz_lrswap.jsonl"), "a") as f:
        f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
