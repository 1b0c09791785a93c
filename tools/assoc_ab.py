#!/usr/bin/env python
"""A/B of the association kernel over several builds of the library (pose2sim_b200/ab/libp2s_*.so, built by
`tools/kernel_ab.py build name:"-DFLAG"`), one process per build, three shapes of BASELINE configs[3]:

    python tools/assoc_ab.py            # on the GPU box; lines to stdout and gpurun_out/assoc_ab.jsonl"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = ((20000, 3, 8), (2000, 6, 8), (100, 6, 8))


def one():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from pose2sim_b200 import ops, synth
    eng = ops.get_engine(0)
    if os.environ.get("P2S_ASSOC_MODE") and hasattr(eng, "set_search_mode"):
        eng.set_search_mode(os.environ["P2S_ASSOC_MODE"])
    for F, Np, C in SHAPES:
        wl = synth.make_association_workload(C, F, Np, seed=404)
        obs4 = np.zeros((F, C, Np, 4), np.float32)
        obs4[..., :3] = wl["obs"]
        d_obs, d_cnt = torch.from_numpy(obs4).cuda(), torch.from_numpy(wl["count"]).cuda()
        out = eng.associate(d_obs, d_cnt, wl["P"], 20.0, 0.3, 2, want_stats=True)
        torch.cuda.synchronize()
        st = out["stats"].cpu().numpy().astype(np.int64)
        for _ in range(5):                                   # warm-up: module load, clocks
            eng.associate(d_obs, d_cnt, wl["P"], 20.0, 0.3, 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng.associate(d_obs, d_cnt, wl["P"], 20.0, 0.3, 2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        chk = float(torch.nansum(out["Q"]).item()) + float(out["comb"].long().sum().item())
        print(json.dumps({"lib": os.path.basename(os.environ.get("P2S_LIB", "default")), "mode": os.environ.get("P2S_ASSOC_MODE", "default"), "frames": F, "persons": Np, "cams": C,
                          "kernel_ms": ms, "frames_per_s": F / ms * 1e3, "rows_per_s": float(st[:, 0].sum()) / ms * 1e3,
                          "cands_per_s": float(st[:, 1].sum()) / ms * 1e3, "grid": eng.last_grid(), "checksum": chk}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        libs = [None] + sorted(glob.glob(os.path.join(ROOT, "pose2sim_b200", "ab", "libp2s_*.so")))
        with open(os.path.join(ROOT, "gpurun_out", "assoc_ab.jsonl"), "a") as log:
            runs = [(lib, None) for lib in libs] + [(None, "exhaustive")]
            for lib, mode in runs:
                env = dict(os.environ)
                if lib:
                    env["P2S_LIB"] = lib
                if mode:
                    env["P2S_ASSOC_MODE"] = mode
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, capture_output=True, text=True)
                out = r.stdout.strip() or ("FAILED " + r.stderr[-400:])
                print(out, flush=True)
                log.write(out + "\n")
