#!/usr/bin/env python
"""e2e (host buffers through p2s_triangulate_host) against the pipeline chunk size and host-thread placement.

    python tools/e2e_sweep.py [cfg2|cfg3]

One JSON line per setting (also gpurun_out/e2e_sweep.jsonl): ms per pass over the whole workload, units/s and
the PCIe payload rate (H2D bytes / time)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from pose2sim_b200 import ops, synth
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    cfg = bench.WORKLOADS[name]
    wl = synth.make_triangulation_workload(cfg["C"], cfg["F"], cfg["N"], cfg["K"], seed=cfg["seed"], lik_thr=None)
    U, C = wl["x"].shape
    eng = ops.get_engine(0)
    out_f = open(os.path.join(ROOT, "gpurun_out", "e2e_sweep.jsonl"), "a")
    for bind in (False, True):
        prev = ops.bind_host_threads_to_gpu(0) if bind else None
        hx, hy, hl = (torch.from_numpy(wl[k].copy()).pin_memory() for k in ("x", "y", "lik"))
        ho = {"Q": torch.empty((U, 3), dtype=torch.float64).pin_memory().numpy(),
              "err": torch.empty(U, dtype=torch.float64).pin_memory().numpy(),
              "nexcl": torch.empty(U, dtype=torch.uint8).pin_memory().numpy(),
              "mask": torch.empty(U, dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
        for chunk in (0, 1 << 16, 1 << 18, 1 << 20):                  # 0 = automatic (quarter of the call, tapered tail)
            eng.set_chunk_units(chunk)
            for _ in range(3):
                eng.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], cfg["lik_thr"], cfg["thr"], cfg["min_cams"], out=ho, want_stats=False)
            torch.cuda.synchronize()
            n = 15
            t0 = time.perf_counter()
            for _ in range(n):
                eng.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], cfg["lik_thr"], cfg["thr"], cfg["min_cams"], out=ho, want_stats=False)
            dt = (time.perf_counter() - t0) / n
            line = {"tool": "e2e_sweep", "workload": name, "bound_to_gpu_numa": bool(bind and prev is not None), "chunk_units": chunk,
                    "ms": dt * 1e3, "units_per_s": U / dt, "h2d_GBps": 12 * C * U / dt / 1e9, "d2h_GBps": 37 * U / dt / 1e9,
                    "cpus": len(os.sched_getaffinity(0))}
            print(json.dumps(line), flush=True)
            out_f.write(json.dumps(line) + "\n")
        if prev is not None:
            os.sched_setaffinity(0, prev)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
