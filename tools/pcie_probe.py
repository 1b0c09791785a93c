#!/usr/bin/env python
"""PCIe reference rates of the box for the e2e roofline: pinned H2D alone, D2H alone, and both at once on two streams
(sizes of the cfg2 step: 249.6 MB in, 96.2 MB out).  One JSON line (also gpurun_out/pcie_probe.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    n_in, n_out = 249_600_000, 96_200_000
    hin = torch.empty(n_in, dtype=torch.uint8).pin_memory()
    hout = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    din = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    dout = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(h2d, d2h, reps=10):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    run(True, True, 2)
    a, b, c = run(True, False), run(False, True), run(True, True)
    line = {"tool": "pcie_probe", "h2d_alone_GBps": n_in / a / 1e6, "d2h_alone_GBps": n_out / b / 1e6,
            "both_ms": c, "both_h2d_GBps": n_in / c / 1e6, "both_d2h_GBps": n_out / c / 1e6,
            "cfg2_step_bound_ms": c, "cfg2_e2e_bound_units_per_s": 2_600_000 / c * 1e3}
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(line, open(os.path.join(ROOT, "gpurun_out", "pcie_probe.json"), "w"))


if __name__ == "__main__":
    main()
