#!/usr/bin/env python3
"""Developer probe: trace/shade time by maximum path depth (primary-only vs full) on one scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pt = ge.load_package()
sid, width, spp = (int(x) for x in sys.argv[1].split(":"))
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
scene = pt.Scene.build(sid, width=width, spp=spp, seed=1)
ctx = pt.Context(0); dev = ctx.upload(scene); ctx.set_profiling(1)
for depth in (1, 2, 3, 50):
    cam = scene.camera_copy(max_depth=depth)
    dev.render(camera=cam, spp=2, seed=1, flags=flags)
    _, st = dev.render(camera=cam, spp=spp, seed=2, nan_policy=1, flags=flags)
    print(f"scene {sid} max_depth {depth:2d}: segments {st.segments/1e6:.1f} M | device {st.device_ms:.1f} ms | gen {st.raygen_ms:.1f} trace {st.trace_ms:.1f} shade {st.shade_ms:.1f} | trace {st.segments/st.trace_ms/1e3:.0f} Mrays/s | iters {st.iterations}", flush=True)
