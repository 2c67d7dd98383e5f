#!/usr/bin/env python3
"""Developer probe (GPU box): time of pt_scene_create (scene conversion + H2D) per scene; with PT_B200_TIMING=1 the library prints its phases."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pt = ge.load_package()
ctx = pt.Context(0)
for sid, w in ((6, 1920), (70, 600), (5, 600)):
    scene = pt.Scene.build(sid, width=w, spp=1, seed=1)
    for k in range(4):
        t0 = time.perf_counter(); dev = ctx.upload(scene); t1 = time.perf_counter(); dev.close(); t2 = time.perf_counter()
        print(f"scene {sid}: upload {1e3*(t1-t0):.2f} ms close {1e3*(t2-t1):.2f} ms", flush=True)
