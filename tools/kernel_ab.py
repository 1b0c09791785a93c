#!/usr/bin/env python
"""Kernel A/B harness: times the triangulation kernel of several builds of the library on the same
workload in ONE gpurun call.

    python tools/kernel_ab.py build name1:"-DFOO=1" name2:"-DBAR"   # here (CPU box): builds pose2sim_b200/ab/libp2s_<name>.so
    python tools/kernel_ab.py run [cfg2|cfg3] [steps]               # on the GPU box: one line per build

Each build is timed in its own process (P2S_LIB selects the library), CUDA events around `steps`
back-to-back launches after 5 warm-ups; results go to stdout and gpurun_out/kernel_ab.jsonl."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AB = os.path.join(ROOT, "pose2sim_b200", "ab")
CSRC = os.environ.get("P2S_AB_SRC", os.path.join(ROOT, "pose2sim_b200", "csrc"))   # P2S_AB_SRC: build another checkout's sources
NVCC = ["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
        "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def build(specs):
    os.makedirs(AB, exist_ok=True)
    for spec in specs:
        name, _, flags = spec.partition(":")
        flags = flags.split()
        objs = []
        for src in ("p2s_capi.cu", "p2s_triangulate.cu", "p2s_associate.cu", "p2s_multiperson.cu", "p2s_lrswap.cu", "p2s_synth.cu"):
            if not os.path.exists(os.path.join(CSRC, src)):
                continue
            obj = os.path.join(AB, f"{name}_{src[:-3]}.o")
            r = subprocess.run(NVCC + flags + ["-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
            if r.returncode:
                sys.exit(r.stderr)
            if src == "p2s_triangulate.cu":
                regs = [l for l in r.stderr.splitlines() if "Used" in l or "spill" in l]
                open(os.path.join(AB, f"{name}.ptxas.log"), "w").write(r.stderr)
            objs.append(obj)
        jobj = os.path.join(AB, f"{name}_p2s_json.o")
        subprocess.run(["g++", "-O3", "-std=c++17", "-fPIC", "-pthread", "-c", os.path.join(CSRC, "p2s_json.cpp"), "-o", jobj], check=True)
        objs.append(jobj)
        out = os.path.join(AB, f"libp2s_{name}.so")
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs, check=True)
        for o in objs:
            os.remove(o)
        print("built", out)


def time_one(workload, steps):
    sys.path.insert(0, ROOT)
    import torch
    import bench
    from pose2sim_b200 import ops, synth
    cfg = bench.WORKLOADS[workload]
    F = min(cfg["F"], 100_000)
    wl = synth.make_triangulation_workload(cfg["C"], F, cfg["N"], cfg["K"], seed=cfg["seed"], lik_thr=None)
    eng = ops.get_engine(0)
    if "P2S_DEEP_MIN" in os.environ:                     # A/B of the deep-level kernel: 0 = the single-kernel search
        eng.set_deep_search(int(os.environ["P2S_DEEP_MIN"]))
    x, y, lik = (torch.from_numpy(wl[k]).cuda() for k in ("x", "y", "lik"))
    stats = eng.new_stats()
    out = eng.triangulate_planes(x, y, lik, wl["P"], cfg["lik_thr"], cfg["thr"], cfg["min_cams"], stats=stats)
    torch.cuda.synchronize()
    st = ops.stats_dict(stats.cpu().numpy())
    for _ in range(5):
        eng.triangulate_planes(x, y, lik, wl["P"], cfg["lik_thr"], cfg["thr"], cfg["min_cams"], out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        eng.triangulate_planes(x, y, lik, wl["P"], cfg["lik_thr"], cfg["thr"], cfg["min_cams"], out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    chk = float(torch.nansum(out["Q"]).item()) + float(torch.nansum(out["err"]).item()) + float(out["mask"].sum().item())
    U = x.shape[0]
    print(json.dumps({"lib": os.path.basename(os.environ.get("P2S_LIB", "default")), "deep_min": os.environ.get("P2S_DEEP_MIN", "default"),
                      "workload": workload, "kernel_ms": ms,
                      "units_per_s": U / ms * 1e3, "grid": eng.last_grid(), "ctas_per_sm": eng.last_grid() / eng.info["sm_count"],
                      "cands": st["candidates"], "solver_steps_per_cand": st["solver_steps"] / max(st["candidates"], 1),
                      "checksum": chk}))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "one":
        time_one(sys.argv[2], int(sys.argv[3]))
    else:
        workload = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
        steps = sys.argv[3] if len(sys.argv) > 3 else "30"
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        libs = [None] + sorted(glob.glob(os.path.join(AB, "libp2s_*.so")))
        with open(os.path.join(ROOT, "gpurun_out", "kernel_ab.jsonl"), "a") as log:
            for lib in libs:
                env = dict(os.environ)
                if lib:
                    env["P2S_LIB"] = lib
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "one", workload, steps], env=env, capture_output=True, text=True)
                line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-400:]
                print(line, flush=True)
                log.write(line + "\n")
