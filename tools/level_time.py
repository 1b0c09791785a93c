#!/usr/bin/env python
"""Kernel time of cfg2 by exclusion level: all levels, level 0 only (threshold 1e9), levels 0-1 and 0-2 (min_cameras 7 / 6).
python tools/level_time.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from pose2sim_b200 import ops, synth
eng=ops.get_engine(0)
wl=synth.make_triangulation_workload(8,100000,1,26,seed=202,lik_thr=None)
x,y,lik=(torch.from_numpy(wl[k]).cuda() for k in ("x","y","lik"))
def t(thr, mc, n=20):
    out=eng.triangulate_planes(x,y,lik,wl["P"],0.3,thr,mc)
    for _ in range(3): eng.triangulate_planes(x,y,lik,wl["P"],0.3,thr,mc,out=out)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): eng.triangulate_planes(x,y,lik,wl["P"],0.3,thr,mc,out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
print(json.dumps({"all_levels_ms": t(15.0,2), "level0_only_ms": t(1e9,2), "levels_0_1_ms": t(15.0,7), "levels_0_2_ms": t(15.0,6)}))
