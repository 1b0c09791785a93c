#!/usr/bin/env python
"""Experiment: the search kernel reading its input planes from, and writing its outputs to, PINNED HOST memory
directly (UVA: pinned allocations are device-accessible at the same address) instead of the chunked
H2D -> kernel -> D2H pipeline of p2s_triangulate_host.  One JSON line (also gpurun_out/zero_copy_probe.json).

    python tools/zero_copy_probe.py
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from pose2sim_b200 import _lib, ops, synth
    cfg = bench.WORKLOADS["cfg2"]
    wl = synth.make_triangulation_workload(cfg["C"], cfg["F"], 1, 26, seed=cfg["seed"], lik_thr=None)
    U, Cn = wl["x"].shape
    eng = ops.get_engine(0)
    hx, hy, hl = (torch.from_numpy(wl[k].copy()).pin_memory() for k in ("x", "y", "lik"))
    Q = torch.empty((U, 3), dtype=torch.float64).pin_memory()
    err = torch.empty(U, dtype=torch.float64).pin_memory()
    nexcl = torch.empty(U, dtype=torch.uint8).pin_memory()
    mask = torch.empty(U, dtype=torch.int32).pin_memory()
    P = np.ascontiguousarray(wl["P"].reshape(Cn, 12))
    st = torch.cuda.current_stream().cuda_stream
    res = {}
    eng.set_output_mode(os.environ.get("P2S_OUTPUT_MODE", "vector"))
    res["output_mode"] = os.environ.get("P2S_OUTPUT_MODE", "vector")
    for mode in ("inputs_and_outputs_in_host_memory", "inputs_in_host_memory"):
        if mode == "inputs_in_host_memory":
            dQ, derr, dn, dm = (t.cuda() for t in (Q, err, nexcl, mask))
            outs = (dQ, derr, dn, dm)
        else:
            outs = (Q, err, nexcl, mask)

        def run():
            _lib.check(eng.h, eng.lib.p2s_triangulate_planes_device(
                eng.h, hx.data_ptr(), hy.data_ptr(), hl.data_ptr(), P.ctypes.data, U, Cn, float(cfg["lik_thr"]),
                float(cfg["thr"]), int(cfg["min_cams"]), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                outs[3].data_ptr(), None, st))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 8
        for _ in range(n):
            run()
        torch.cuda.synchronize()
        res[mode + "_ms"] = (time.perf_counter() - t0) / n * 1e3
    eng.set_host_mode("pipeline")
    ref = eng.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], cfg["lik_thr"], cfg["thr"], cfg["min_cams"], want_stats=False)
    same = bool(np.array_equal(ref["Q"], Q.numpy(), equal_nan=True) and np.array_equal(ref["mask"], mask.numpy().view(np.uint32)))
    t0 = time.perf_counter()
    for _ in range(8):
        eng.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], cfg["lik_thr"], cfg["thr"], cfg["min_cams"], out=ref, want_stats=False)
    res["pipeline_ms"] = (time.perf_counter() - t0) / 8 * 1e3
    res.update(tool="zero_copy_probe", units=U, outputs_equal_pipeline=same)
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "zero_copy_probe.json"), "w"))


if __name__ == "__main__":
    main()
