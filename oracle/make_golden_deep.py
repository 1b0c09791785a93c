"""TEST INFRASTRUCTURE — tests/golden/tri_deep_units.npz: units of WIDE rigs that the UNMODIFIED reference walks through
DEEP levels of its exclusion search (Pose2Sim/triangulation.py:408-505: level k enumerates all C(n_cams, k) camera
subsets — thousands at 12-20 cameras), same key layout as tri_random_units.npz.

Run in the build container only (needs /root/reference; ~10 minutes, the reference spends ~0.15 ms per subset):

    python oracle/make_golden_deep.py [--widest-only]

Why a separate set: tri_random_units.npz stops at 8 cameras.  Beyond that the CUDA path runs other instantiations
(12 / 16 exact-count kernels, 13 / 20 cameras on the next wider one), forms a candidate's normal matrix by the
downdate-or-sum rule, unranks subsets instead of reading a table, and parks units for deep_search_kernel — all pinned
to the oracle by the GPU tests; this set pins the oracle (and the CUDA path) to the reference itself at those depths.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from make_golden import GOLDEN, random_units, run_reference_units  # noqa: E402

# (cameras, units, min_cameras, threshold px, outlier rate, NaN rate, zero-likelihood rate)
CASES = [
    (12, 24, 2, 1e-3, 0.10, 0.10, 0.04),    # threshold nobody meets: every unit walks every level it may (2^12 subsets)
    (12, 40, 4, 10.0, 0.30, 0.10, 0.04),    # many outliers, search ends on the threshold at levels 2-6
    (13, 30, 3, 8.0, 0.30, 0.08, 0.04),     # not an exact-count rig
    (16, 40, 3, 15.0, 0.28, 0.05, 0.03),    # cfg3's settings with enough outliers for levels 4-7 (C(16, 5) = 4 368 ...)
    (16, 12, 10, 1e-3, 0.10, 0.05, 0.03),   # min_cameras ends the search at level 6, every level evaluated
    (20, 16, 16, 15.0, 0.15, 0.04, 0.02),   # C(20, 4) = 4 845
]

# tests/golden/tri_widest_units.npz: BASELINE configs[4]'s widest rigs (search capped at four exclusions by min_cameras,
# like tests/perf/sweep_cfg5.py); pins the ORACLES to the reference there (CPU tests) — the CUDA path is pinned to the
# oracles at these widths by the fuzzers and the cfg5 sweep
CASES_WIDEST = [
    (24, 14, 20, 15.0, 0.10, 0.03, 0.02),   # C(24, 3) = 2 024, C(24, 4) = 10 626
    (32, 10, 29, 15.0, 0.05, 0.02, 0.01),   # C(32, 3) = 4 960
    (28, 10, 25, 10.0, 0.06, 0.03, 0.02),   # not an exact-count rig
]


def generate(ref, cases, seed0, fname):
    out = {}
    for i, (C, U, mc, thr, p_out, p_nan, p_zero) in enumerate(cases):
        t0 = time.time()
        P, x, y, w = random_units(C, U, seed=seed0 + i, p_out=p_out, p_nan=p_nan, p_zero=p_zero)
        Q, err, nexcl, mask = run_reference_units(ref, x, y, w, P, thr, mc)
        pre = f"r{i}_"
        out[pre + "P"], out[pre + "x"], out[pre + "y"], out[pre + "w"] = P, x, y, w
        out[pre + "params"] = np.array([thr, mc], float)
        out[pre + "Q"], out[pre + "err"], out[pre + "nexcl"], out[pre + "mask"] = Q, err, nexcl, mask
        print(f"  deep units C={C} U={U} min_cams={mc} thr={thr}: nexcl histogram "
              f"{np.bincount(nexcl, minlength=C + 1).tolist()}, {int(np.isnan(err).sum())} failed, {time.time() - t0:.0f} s",
              flush=True)
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(GOLDEN, fname), **out)


def main():
    ref = ref_shim.load_reference()
    if "--widest-only" not in sys.argv:
        generate(ref, CASES, 7000, "tri_deep_units.npz")
    generate(ref, CASES_WIDEST, 7100, "tri_widest_units.npz")


if __name__ == "__main__":
    main()
