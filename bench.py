#!/usr/bin/env python
"""Benchmark of the triangulation hot path (BASELINE.json metric) — one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input resident in HBM: ONE search kernel (likelihood
gate + SoA staging fused into its tile load, weighted DLT + camera-exclusion search) followed by the wide-likelihood /
arrival-flag kernel that returns at once on this workload.  At N > 1 every rank owns its own frame block (weak scaling,
no data-path collective) and — default `--gather local` — the results STAY in the rank's HBM, because that is what the
product does: `triangulate_all` under torchrun post-processes every frame block on its own rank and writes its byte
range of the TRC (`triangulation.write_outputs_sharded`).  `--gather push` (every finished tile is stored into rank 0's
memory over NVLink by the search kernel itself) and `--gather nccl` (`dist.gather` after the kernel) measure the
gather-to-one alternatives; both are bounded by rank 0's NVLink ingress.
The same line carries a `cfg3` sub-object: BASELINE.json configs[2] (16 cameras, min_cameras 3, 125 k frames per GPU,
inputs generated on the device) timed the same way, so that a 1/2/4/8-GPU scaling run shows that configuration too.

Keys beyond the base contract: `roofline` (FP64 CUDA-core bound, plus the HBM view), `cpu_baseline`
(the oracle port on this box's host cores), `e2e` (host buffers through the C ABI, copies timed; `link_bound_ms` =
the same bytes as plain pinned copies), `parity` (this run's CUDA results on the CPU leg's sample against the
plain-C oracle: max |dQ|, max |d err|, differing decisions and how many of them sit in the eps-band).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

JSON_OUT = sys.stdout
METRIC = "keypoint triangulations/sec w/ exclusion search"
UNIT = "triangulations/s"

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "cfg2": dict(C=8, F=100_000, N=1, K=26, seed=202, thr=15.0, min_cams=2, lik_thr=0.3,
                 name="cfg2: synthetic 8 cams x HALPE_26 x 100k frames, likelihood-weighted DLT + reproj exclusion"),
    # BASELINE.json configs[2] at 1/8 of its frames per GPU (1M frames over 8 GPUs)
    "cfg3": dict(C=16, F=125_000, N=1, K=26, seed=303, thr=15.0, min_cams=3, lik_thr=0.3,
                 name="cfg3 shard: synthetic 16 cams x Body_with_feet x 125k frames (1M/8), min_cameras=3"),
}

# FP64 operations per unit of ALGORITHMIC work of the triangulation kernel (FMA = 2 flops,
# MUFU seeds not counted).  The work items are counted by the kernel itself (statistics block), so the
# figure follows the measured exclusion-level histogram instead of an assumed one.
FLOPS = {
    "direct_cams": 64,    # level 0: two weighted DLT rows (24) + two rank-1 updates of the 4x4 (40) per valid camera
    "blocks": 54,         # levels >= 1: rows (24) + 10-entry block a a^T + b b^T (30) per valid camera, unit and level
    "entry_adds": 1,      # levels >= 1: M_all and M_all -/+ excluded/kept blocks
    "solver_steps": 70,   # one secular-Newton step: 3x3 LDL^T, two triangular solves, f, g, step
    "solved": 21 + 11,    # first-order final update of q + the mean (reciprocal of m with one correction)
    "cam_solves": 44,     # reprojection distance per valid camera: 3 dot-4, N, one refined rsqrt
}


def algorithmic_flops(st):
    return sum(w * st[k] for k, w in FLOPS.items())


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` capture of this same command (profiles/): 249.7 MB read + 73.5 MB written on cfg2
# (algorithmic: 249.6 MB of planes in, 96.2 MB of results out — the outputs are partly still in L2).
NCU_TRAFFIC = {("cfg2", 1): 321.4e6, ("cfg3", 1): 739.8e6}      # cfg3: profiles/r1s_triangulate_cfg3_ncu_full.csv
NCU_TRAFFIC_SOURCE = "profiles/r3p_triangulate_ncu_full.csv"


def algorithmic_bytes(U, C):
    return (12 * C + 37) * U     # x, y, likelihood planes in (12 B per camera), Q + err + mask + count out


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while `active` (the timed region)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.active, self.stop_flag = index, [], False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self.stop_flag:
            if self.active:
                try:
                    self.samples.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                         nv.nvmlDeviceGetCurrentClocksEventReasons(self.h),
                                         nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
                except Exception:
                    pass
            time.sleep(0.002)

    def summary(self):
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        nv = self.nv
        names = {getattr(nv, "nvmlClocksEventReasonGpuIdle", 0x1): "gpu_idle",
                 getattr(nv, "nvmlClocksEventReasonApplicationsClocksSetting", 0x2): "applications_clocks_setting",
                 getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                 getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                 getattr(nv, "nvmlClocksEventReasonSyncBoost", 0x10): "sync_boost",
                 getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                 getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown"}
        bits = 0
        for _, r, _ in self.samples:
            bits |= r
        reasons = sorted(n for b, n in names.items() if bits & b and n != "gpu_idle")
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median([s[0] for s in self.samples])), "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.samples), "power_w_max": max(s[2] for s in self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------
def _port_worker(args):
    import warnings
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import p2s_oracle as orc
    x, y, w, P, thr, mc = args
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        orc.triangulate_units(x.astype(float), y.astype(float), w.astype(float), P, thr, mc)
    return len(x)


_REF = None


def _ref_init():
    """Pool initializer of the live-reference arm: import the UNMODIFIED reference through oracle/ref_shim.py."""
    global _REF
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    _REF = ref_shim.load_reference()


def _ref_worker(args):
    """The reference's own per-unit function (Pose2Sim/triangulation.py:363 `triangulation_from_best_cameras`), called the
    way `triangulate_all` calls it (:837-840) on the gated (3, C) slices."""
    import warnings
    x, y, w, P, thr, mc = args
    cfg = {"triangulation": {"reproj_error_threshold_triangulation": thr, "min_cameras_for_triangulation": mc,
                             "handle_LR_swap": False, "undistort_points": False}}
    Plist = [P[c] for c in range(P.shape[0])]
    fn = _REF.triangulation.triangulation_from_best_cameras
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for u in range(len(x)):
            coords = np.array([x[u].astype(np.float64), y[u].astype(np.float64), w[u].astype(np.float64)])
            fn(cfg, coords, coords, Plist, None)
    return len(x)


def live_reference_available():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_shim
        return ref_shim.reference_available()
    except Exception:
        return False


class PythonPort:
    """The reference's per-unit path on all host cores (multiprocessing): the UNMODIFIED reference itself when it is
    present (`live=True`: build container, or a vendored copy named by P2S_REFERENCE_ROOT), else its NumPy restatement."""

    def __init__(self, wl, cfg, live=False):
        import multiprocessing as mp
        self.cores = len(os.sched_getaffinity(0))
        self.live = live
        self.worker = _ref_worker if live else _port_worker
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_ref_init if live else None)
        self.wl, self.cfg = wl, cfg

    def run(self, n_units):
        """Time the first n_units units of the workload; returns seconds."""
        wl, cfg = self.wl, self.cfg
        n_units = min(n_units, wl["x"].shape[0])
        per = max(1, -(-n_units // (self.cores * 4)))
        jobs = [(wl["x"][i:min(i + per, n_units)], wl["y"][i:min(i + per, n_units)], wl["lik"][i:min(i + per, n_units)],
                 wl["P"], cfg["thr"], cfg["min_cams"]) for i in range(0, n_units, per)]
        t0 = time.perf_counter()
        done = sum(self.pool.map(self.worker, jobs, chunksize=1))
        dt = time.perf_counter() - t0
        assert done == n_units
        return dt, n_units

    def close(self):
        self.pool.terminate()


def c_port_rate(wl, cfg, n_units, want_results=False):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as co
    n_units = min(n_units, wl["x"].shape[0])
    co.triangulate_units(wl["x"][:2000], wl["y"][:2000], wl["lik"][:2000], wl["P"], cfg["thr"], cfg["min_cams"])
    t0 = time.perf_counter()
    res = co.triangulate_units(wl["x"][:n_units], wl["y"][:n_units], wl["lik"][:n_units], wl["P"], cfg["thr"], cfg["min_cams"])
    dt = time.perf_counter() - t0
    if want_results:
        return n_units / dt, co.max_threads(), n_units, res
    return n_units / dt, co.max_threads(), n_units


def parity_block(eng, cwl, cfg, n_units, oracle_res, eps=1e-6):
    """Parity of THIS run: the CUDA path through the host entry point on the CPU leg's sample against
    the plain-C oracle's results for the same units (the oracle is the checker here, never the thing measured)."""
    Q, err, nexcl, mask = oracle_res[:4]
    out = eng.triangulate_host(cwl["x"][:n_units], cwl["y"][:n_units], cwl["lik"][:n_units], cwl["P"], None,
                               cfg["thr"], cfg["min_cams"], want_stats=False)
    nan_diff = np.isnan(Q).any(axis=1) != np.isnan(out["Q"]).any(axis=1)
    dec = (nexcl != out["nexcl"]) | (mask != out["mask"]) | nan_diff
    ok = ~np.isnan(Q).any(axis=1) & ~np.isnan(out["Q"]).any(axis=1) & ~dec
    in_band = dec & (np.abs(np.nan_to_num(err, nan=np.inf) - cfg["thr"]) < eps)
    dq = np.abs(Q[ok] - out["Q"][ok]).max(axis=1) if ok.any() else np.zeros(1)
    return {"units": int(n_units), "oracle": "oracle/p2s_oracle.c", "tolerance_m": 1e-6, "eps_px": eps,
            "max_abs_dQ_m": float(dq.max(initial=0.0)), "p99_abs_dQ_m": float(np.percentile(dq, 99)),
            "max_abs_derr_px": float(np.abs(err[ok] - out["err"][ok]).max(initial=0.0)),
            "units_with_differing_decision": int(dec.sum()), "of_which_inside_eps_band": int(in_band.sum())}


def run_reference_arm(args, cfg, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores, on a bounded sample of
    the same workload.  Where the reference is importable (`/root/reference` in the build container, or a copy named by
    P2S_REFERENCE_ROOT) that is the UNMODIFIED `triangulation_from_best_cameras` through oracle/ref_shim.py
    (`kind: "reference"`); on the GPU box, where it is absent (pure Python: it cannot travel as a compiled oracle/_ref),
    the NumPy port with the same per-unit / per-candidate structure (`kind: "port"`)."""
    if rank != 0:
        return
    from pose2sim_b200 import synth
    sample_frames = 4000
    wl = synth.make_triangulation_workload(cfg["C"], sample_frames, cfg["N"], cfg["K"], seed=cfg["seed"], lik_thr=cfg["lik_thr"])
    live = live_reference_available()
    port = PythonPort(wl, cfg, live=live)
    dt, n = port.run(port.cores * 64)                          # calibration (also warms the pool)
    rate = n / dt
    budget = min(20.0, 150.0 / max(1, args.steps + args.warmup))
    per_step = int(max(port.cores * 16, min(wl["x"].shape[0], rate * budget)))
    for _ in range(args.warmup):
        port.run(per_step)
    t = 0.0
    for _ in range(args.steps):
        dt, n = port.run(per_step)
        t += dt
    port.close()
    value = per_step * args.steps / t
    c_rate, c_threads, c_n = c_port_rate(wl, cfg, 100_000)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "n_cams": cfg["C"], "keypoints": cfg["K"],
                       "reproj_error_threshold_triangulation": cfg["thr"], "min_cameras_for_triangulation": cfg["min_cams"],
                       "likelihood_threshold_triangulation": cfg["lik_thr"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": port.cores, "kind": "reference" if live else "port",
                             "sample": f"first {per_step} units of the workload per step, "
                                       + ("the UNMODIFIED reference's triangulation_from_best_cameras (Pose2Sim/triangulation.py:363) "
                                          "through oracle/ref_shim.py" if live else
                                          "NumPy per-unit port of triangulation_from_best_cameras (the reference is absent on this box)")
                                       + ", multiprocessing over all host cores",
                             "c_port": {"value": c_rate, "threads": c_threads, "sample_units": c_n,
                                        "what": "plain-C restatement (one-sided Jacobi SVD), OpenMP"}},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=JSON_OUT, flush=True)


def cfg3_leg(eng, torch, barrier, dev, rank, steps=10):
    """BASELINE.json configs[2] beside the headline workload: 16 cameras, min_cameras 3, 125 k frames per GPU (1 M / 8),
    inputs generated ON THE DEVICE (csrc/p2s_synth.cu, a pure function of (seed, unit, camera) — every rank draws its own
    frame block), timed like the main step: barrier, `steps` launches, CUDA events, max over ranks by the caller."""
    from pose2sim_b200 import ops, synth
    cfg = WORKLOADS["cfg3"]
    C, F, K = cfg["C"], cfg["F"], cfg["K"]
    U = F * K
    P = synth.ring_cameras(C)[0]
    # The 1 M-frame stream is dealt to the configuration's 8 GPUs BLOCK-CYCLICALLY (blocks of 600 frames): the cost of a
    # frame drifts along the stream (the walk carries the subject through regions with more or fewer deep exclusion
    # levels: contiguous eighths differ by up to 25 % in kernel time, profiles/r2p_bench_n4_contiguous.json), so every
    # shard is an interleaved sample of the whole stream and N GPUs process N of the 8 shards.
    # A block is one period of the generator's walk (600 frames), so all shards see the same mix of geometry.
    BLK, SHARDS = 600, 8
    wl = {k: torch.empty((U, C), dtype=torch.float32, device=dev) for k in ("x", "y", "lik")}
    for j in range((F + BLK - 1) // BLK):
        a, b = j * BLK * K, min((j + 1) * BLK, F) * K
        eng.synth_observations(P, (j * SHARDS + rank % SHARDS) * BLK * K, b - a, K, cfg["seed"],
                               out={k: v[a:b] for k, v in wl.items()})
    stats = eng.new_stats()
    out = eng.triangulate_planes(wl["x"], wl["y"], wl["lik"], P, cfg["lik_thr"], cfg["thr"], cfg["min_cams"], stats=stats)
    torch.cuda.synchronize()
    st = ops.stats_dict(stats.cpu().numpy())
    for _ in range(3):
        eng.triangulate_planes(wl["x"], wl["y"], wl["lik"], P, cfg["lik_thr"], cfg["thr"], cfg["min_cams"], out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        eng.triangulate_planes(wl["x"], wl["y"], wl["lik"], P, cfg["lik_thr"], cfg["thr"], cfg["min_cams"], out=out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    del wl, out
    torch.cuda.empty_cache()
    return {"workload": cfg["name"] + "; inputs from the device generator (p2s_synth_observations_device)", "n_cams": C,
            "min_cameras_for_triangulation": cfg["min_cams"], "units_per_gpu": U, "steps": steps, "ms_per_step": ms,
            "kernel_ms": ms, "flops": algorithmic_flops(st), "level_hist": st["level_hist"],
            "candidates_per_unit": st["candidates"] / U, "gather": "none: results stay in the rank's HBM",
            "sharding": f"block-cyclic: 1 M frames in blocks of {BLK}, block b belongs to shard b mod {SHARDS}; rank r processes shard r"}


# ---------------------------------------------------------------------------------------------------
def _claim_stdout():
    """Route fd 1 to stderr for the duration of the run (NCCL and torch print banners on stdout) and
    return a file object on the ORIGINAL stdout for the one JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--output-mode", default="vector", choices=["vector", "bulk"],
                    help="how a finished tile leaves the kernel: 16-byte vector stores or cp.async.bulk (TMA) stores")
    ap.add_argument("--gather", default="local", choices=["local", "push", "nccl"],
                    help="N > 1: 'local' = results stay in the rank's HBM (the product's rank-local post-processing), "
                         "'push' = the kernel stores its outputs straight into rank 0's memory over NVLink "
                         "(sharding.PeerGather), 'nccl' = dist.gather after the kernel")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the cfg3 sub-object")
    args = ap.parse_args()
    global JSON_OUT
    JSON_OUT = _claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = WORKLOADS[args.workload]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pose2sim_b200 import ops, sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = ops.get_engine(local)
    eng.set_output_mode(args.output_mode)

    # ---- workload: every rank owns its own block of frames (weak scaling) ---------------------------
    C, F, thr, mc = cfg["C"], cfg["F"], cfg["thr"], cfg["min_cams"]
    # raw likelihoods: the gate (triangulation.py:817-821) is applied on the device by the stage kernel
    wl = synth.make_triangulation_workload(C, F, cfg["N"], cfg["K"], seed=cfg["seed"] + 1000 * rank,
                                           lik_thr=None, frame0=rank * F)
    U = wl["x"].shape[0]
    # pinned staging buffers are first-touched on the cores NVML calls local to this GPU (matters at N > 1, where
    # every rank streams 250 MB per step through its own PCIe root); restored before the CPU legs
    prev_affinity = ops.bind_host_threads_to_gpu(local)
    hx, hy, hl = (torch.from_numpy(wl[k].copy()).pin_memory() for k in ("x", "y", "lik"))
    x, y, lik = (t.to(dev, non_blocking=True) for t in (hx, hy, hl))
    # per-unit outputs packed contiguously (Q | err | mask | nexcl = 37 B/unit) so that the final
    # gather is ONE NCCL call; two buffers alternate so the gather of step i overlaps step i+1
    packs = [torch.empty(sharding.PACK_BYTES * U, dtype=torch.uint8, device=dev) for _ in range(2)]
    outs = [sharding.packed_views(p, U) for p in packs]
    out = outs[0]
    stats = eng.new_stats()
    pending = [None, None]
    step_no = [0]
    # ---- N > 1: how the packed outputs reach rank 0 ------------------------------------------------------
    pg, side, gather_mode, gather_note = None, None, "none", ""
    if world > 1 and args.gather == "push":
        ok = torch.ones(1, device=dev)
        try:
            pg = sharding.PeerGather(eng, [U] * world, dst=0, n_buffers=2)
        except Exception as e:                                  # e.g. CUDA IPC not permitted in this container
            ok.zero_()
            gather_note = f"push unavailable ({type(e).__name__}: {e}); "
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() < 1:
            if pg is not None:
                pg.close()                                      # collective (barrier inside) ...
            else:
                dist.barrier()                                  # ... so the ranks whose set-up failed join it here
            pg = None
        else:
            gather_mode = "push"
            side = torch.cuda.Stream(device=dev)
    if world > 1 and pg is None:
        gather_mode = "nccl" if args.gather != "local" else "local"
    gather_lists = [[torch.empty_like(packs[0]) for _ in range(world)] for _ in range(2)] \
        if (gather_mode == "nccl" and rank == 0) else [None, None]

    def step(record=None):
        s_no = step_no[0]
        b = s_no & 1
        step_no[0] += 1
        if pending[b] is not None:
            pending[b].wait()              # stream-side wait: the buffer's previous gather must be done
        if record is not None:
            record[0].record()
        # ONE kernel: likelihood gate + float4 SoA staging (in shared memory) + exclusion search
        if gather_mode == "push":
            # ... whose stores land in rank 0's gather buffer (NVLink), followed by the arrival flag
            eng.triangulate_planes_push(x, y, lik, wl["P"], cfg["lik_thr"], thr, mc, **pg.push_args(s_no))
            if rank == 0:
                pg.collect(s_no, stream=side.cuda_stream)       # device-side wait for all ranks + buffer release
        else:
            eng.triangulate_planes(x, y, lik, wl["P"], cfg["lik_thr"], thr, mc, out=outs[b])
        if record is not None:
            record[1].record()
        if gather_mode == "nccl":
            pending[b] = dist.gather(packs[b], gather_lists[b], dst=0, async_op=True)
        return pending[b]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one counted pass for the level histogram / algorithmic work of this workload
    eng.triangulate_planes(x, y, lik, wl["P"], cfg["lik_thr"], thr, mc, out=out, stats=stats)
    torch.cuda.synchronize()
    st = ops.stats_dict(stats.cpu().numpy())
    fp64_peak = eng.fp64_peak()

    for _ in range(args.warmup):
        w = step()
        if w is not None:
            w.wait()
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = eng.launch_count()
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    e_begin.record()
    works = []
    for i in range(args.steps):
        works.append(step(evs[i]))
    for w in works:
        if w is not None:
            w.wait()
    if side is not None and rank == 0:
        torch.cuda.current_stream().wait_stream(side)   # every rank's last step has arrived in rank 0's buffer
    e_end.record()
    barrier()
    wall = time.perf_counter() - t0
    sampler.active = False
    launches = eng.launch_count() - launches0
    dev_ms = e_begin.elapsed_time(e_end)
    tri_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    peer_err = 0
    if pg is not None:
        # the gathered bytes must be the results: compare rank 0's own slot and one peer slot with a plain launch
        chk = eng.triangulate_planes(x, y, lik, wl["P"], cfg["lik_thr"], thr, mc)
        torch.cuda.synchronize()
        if rank == 0:
            mine = pg.views((step_no[0] - 1) & 1)[0]
            for k in ("Q", "err", "mask", "nexcl"):
                same = torch.equal(torch.nan_to_num(mine[k].double()), torch.nan_to_num(chk[k].double()))
                if not same:
                    raise SystemExit(f"push gather: rank 0's slot differs from a plain launch in {k}")
        peer_err = eng.peer_error()
        pe = torch.tensor([peer_err], device=dev)
        dist.all_reduce(pe, op=dist.ReduceOp.MAX)
        peer_err = int(pe.item())
        if peer_err:
            raise SystemExit(f"push gather: flag wait timed out (bits {peer_err}) — the measurement is void")

    # ---- e2e: host buffers through the C ABI (H2D + stage + search + D2H inside the timed region) -----
    ho = {"Q": torch.empty((U, 3), dtype=torch.float64).pin_memory().numpy(),
          "err": torch.empty(U, dtype=torch.float64).pin_memory().numpy(),
          "nexcl": torch.empty(U, dtype=torch.uint8).pin_memory().numpy(),
          "mask": torch.empty(U, dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        eng.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], cfg["lik_thr"], thr, mc, out=ho, want_stats=False)
    barrier()
    sampler.active = True
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], cfg["lik_thr"], thr, mc, out=ho, want_stats=False)
    barrier()
    e2e_s = (time.perf_counter() - t1) / e2e_steps
    sampler.active = False
    # the same call forced onto the chunked H2D -> kernel -> D2H copy pipeline (what pageable buffers get)
    eng.set_host_mode("pipeline")
    for _ in range(2):
        eng.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], cfg["lik_thr"], thr, mc, out=ho, want_stats=False)
    barrier()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.triangulate_host(hx.numpy(), hy.numpy(), hl.numpy(), wl["P"], cfg["lik_thr"], thr, mc, out=ho, want_stats=False)
    barrier()
    e2e_pipe_s = (time.perf_counter() - t1) / e2e_steps
    eng.set_host_mode("auto")
    sampler.stop_flag = True
    # the link's own bound for exactly these volumes: the three input planes H2D and the packed outputs D2H as plain
    # copies on two streams, no kernel (what `e2e` can reach at best on this box)
    link_ms = float("nan")
    try:
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        hpack = torch.empty(sharding.PACK_BYTES * U, dtype=torch.uint8).pin_memory()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = []
        for rep in range(4):
            barrier()                                               # all ranks load the host links at the same time
            e0.record()
            s_in.wait_event(e0); s_out.wait_event(e0)
            with torch.cuda.stream(s_in):
                for h, d in ((hx, x), (hy, y), (hl, lik)):
                    d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s_out):
                hpack.copy_(packs[0], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s_in); torch.cuda.current_stream().wait_stream(s_out)
            e1.record()
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1))
        link_ms = float(np.mean(reps[1:]))                          # mean of 3 after one warm-up, max over ranks below
    except Exception:
        link_ms = float("nan")
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)

    # ---- max over ranks --------------------------------------------------------------------------------
    cfg3 = None if args.no_cfg3 or args.workload == "cfg3" else cfg3_leg(eng, torch, barrier, dev, rank)
    tm = torch.tensor([dev_ms, wall * 1e3, e2e_s * 1e3, tri_ms, e2e_pipe_s * 1e3, link_ms if link_ms == link_ms else 0.0,
                       cfg3["ms_per_step"] if cfg3 else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    per_rank = None
    if world > 1:                                               # per-rank device times (which rank is the slowest, and by how much)
        mine = torch.tensor([dev_ms / args.steps, cfg3["ms_per_step"] if cfg3 else 0.0], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"step_ms": [round(float(t[0]), 4) for t in allr], "cfg3_ms": [round(float(t[1]), 4) for t in allr]}
    dev_ms, wall_ms, e2e_ms, tri_ms_max, e2e_pipe_ms, link_max, cfg3_ms = (float(v) for v in tm.cpu())
    link_ms = link_max if link_max > 0 else None
    step_ms = max(dev_ms, 0.0) / args.steps
    total_units = U * world
    value = total_units / (step_ms * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        flops = algorithmic_flops(st)
        tf = flops / (tri_ms * 1e-3) / 1e12
        gbs = algorithmic_bytes(U, C) / (tri_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "n_cams": C, "keypoints": cfg["K"], "frames_per_gpu": F,
                       "units_per_gpu": U, "reproj_error_threshold_triangulation": thr,
                       "min_cameras_for_triangulation": mc, "likelihood_threshold_triangulation": cfg["lik_thr"],
                       "seed": cfg["seed"], "l2": f"inputs {(12 * C * U) >> 20} MiB + outputs {(37 * U) >> 20} MiB per step > 126 MB L2, no flush",
                       "step": "one search kernel (TMA-staged raw planes; at 4 / 8 cameras every lane reads its own unit's row, likelihood "
                               "gate and level 0 run from registers and only the units that go on write their slab column, wider rigs "
                               "transpose the tile in shared memory; exclusion search) + the wide-likelihood / arrival-flag kernel "
                               "behind it (returns at once here)"
                               + ("" if world == 1 else
                                  "; results stay in the rank's HBM (rank-local post-processing, triangulation.write_outputs_sharded)"
                                  if gather_mode == "local" else
                                  " whose stores land in rank 0's memory over NVLink (peer-mapped gather buffer, arrival / "
                                  "release flags, no collective)" if gather_mode == "push" else
                                  " + NCCL gather to rank 0"),
                       "gather": gather_note + gather_mode, "output_mode": args.output_mode,
                       "level_hist": st["level_hist"], "candidates_per_unit": st["candidates"] / U,
                       "failed_units": st["failed"], "eps_band_px": 1e-6,
                       "band_threshold_units": st["band_threshold"], "band_argmin_units": st["band_argmin"]},
            "roofline": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                         "traffic": NCU_TRAFFIC.get((args.workload, 1)), "traffic_source": NCU_TRAFFIC_SOURCE,
                         "traffic_measured_in_run": False,
                         "kernel": f"triangulate_kernel<{next(m for m in (4, 6, 8, 12, 16, 24, 32) if C <= m)},secular,exact,lean{',raw' if C in (4, 8) else ''}>",
                         "grid_ctas": eng.last_grid(),
                         "kernel_ms": tri_ms, "algorithmic_flops_per_launch": flops,
                         "peak_source": "dependent-chain DFMA microbenchmark in this run (p2s_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                 "peak_source": hbm_src, "algorithmic_bytes_per_launch": algorithmic_bytes(U, C)}},
            "e2e": {"value": total_units / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 12 * C * U,
                    "d2h_bytes_per_step": 37 * U, "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "api": "p2s_triangulate_host with pinned host buffers: ONE kernel that TMA-reads its input tiles from host "
                           "memory and writes its results to host memory over PCIe (zero-copy; every input and output byte "
                           "crosses the link inside the timed call)",
                    "copy_pipeline_value": total_units / (e2e_pipe_ms * 1e-3), "copy_pipeline_ms_per_step": e2e_pipe_ms,
                    "copy_pipeline_note": "the same call forced onto chunked cudaMemcpyAsync H2D -> kernel -> D2H on 4 streams "
                                          "(what pageable buffers get)",
                    "host_threads_bound_to_gpu": prev_affinity is not None,
                    "link_bound_ms": link_ms,
                    "link_bound_note": "the same H2D + D2H volumes as plain pinned copies on two streams, no kernel, on ALL ranks at "
                                       "the same time (barrier before every repetition, mean of 3, max over ranks)"},
            "gpu_launches": launches,
            "per_rank_ms": per_rank,
            "clocks": sampler.summary(),
            "wall_ms_per_step": wall_ms / args.steps,
        }
        if cfg3 is not None:
            cfg3["ms_per_step"] = cfg3_ms
            cfg3["value"] = cfg3["units_per_gpu"] * world / (cfg3_ms * 1e-3)
            cfg3["roofline_frac"] = cfg3.pop("flops") / (cfg3.pop("kernel_ms") * 1e-3) / 1e12 / fp64_peak
            line["cfg3"] = cfg3
        if world == 1 and not args.no_cpu_baseline:
            # bounded CPU sample: the oracle port on this box's host cores
            sample_frames = 2000
            cwl = synth.make_triangulation_workload(C, sample_frames, cfg["N"], cfg["K"], seed=cfg["seed"], lik_thr=cfg["lik_thr"])
            live = live_reference_available()
            port = PythonPort(cwl, cfg, live=live)
            port.run(port.cores * 32)                                  # warms the pool
            cdt, cn = port.run(port.cores * 64)                        # calibrates the rate
            n_units = int(min(cwl["x"].shape[0], max(port.cores * 64, 12.0 * cn / cdt)))    # ~12 s of CPU work
            dt, n = port.run(n_units)
            port.close()
            c_rate, c_threads, c_n, c_res = c_port_rate(cwl, cfg, 52_000, want_results=True)
            line["parity"] = parity_block(eng, cwl, cfg, c_n, c_res)
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": port.cores, "kind": "reference" if live else "port",
                                    "value_per_core": n / dt / port.cores,
                                    "full_config_wall_s_extrapolated": U / (n / dt),
                                    "sample": f"first {n} units (of {U}) of the same workload, "
                                              + ("the UNMODIFIED reference's triangulation_from_best_cameras through oracle/ref_shim.py"
                                                 if live else "NumPy per-unit port of triangulation_from_best_cameras "
                                                 "(oracle/p2s_oracle.py; the reference is absent on this box)") + ", one process per core",
                                    "c_port": {"value": c_rate, "threads": c_threads, "sample_units": c_n,
                                               "what": "plain-C restatement (oracle/p2s_oracle.c), OpenMP"}}
        print(json.dumps(line), file=JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        if pg is not None:
            pg.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
