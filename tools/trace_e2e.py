#!/usr/bin/env python
"""Device-side timeline of the p2s_triangulate_host pipeline on cfg2 (P2S_TRACE=1): per chunk, when its H2D copies,
its kernel and its D2H copies finished.  python tools/trace_e2e.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pose2sim_b200 import ops, synth
wl=synth.make_triangulation_workload(8,100000,1,26,seed=202,lik_thr=None)
U=wl["x"].shape[0]; eng=ops.get_engine(0)
hx,hy,hl=(torch.from_numpy(wl[k].copy()).pin_memory() for k in ("x","y","lik"))
ho={"Q":torch.empty((U,3),dtype=torch.float64).pin_memory().numpy(),"err":torch.empty(U,dtype=torch.float64).pin_memory().numpy(),"nexcl":torch.empty(U,dtype=torch.uint8).pin_memory().numpy(),"mask":torch.empty(U,dtype=torch.int32).pin_memory().numpy().view(np.uint32)}
for _ in range(3): eng.triangulate_host(hx.numpy(),hy.numpy(),hl.numpy(),wl["P"],0.3,15.0,2,out=ho,want_stats=False)
os.environ["P2S_TRACE"]="1"
t=time.perf_counter(); eng.triangulate_host(hx.numpy(),hy.numpy(),hl.numpy(),wl["P"],0.3,15.0,2,out=ho,want_stats=False); print("wall ms", (time.perf_counter()-t)*1e3)
