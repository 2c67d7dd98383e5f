#!/usr/bin/env python3
"""Developer probe (GPU box): does one B200 render faster when TWO (or more) half-size wavefronts run concurrently?
pt_render_multi with the same device listed k times = k contexts, k host threads, spp split k ways on ONE GPU.
usage: concurrency_probe.py [scene width spp]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
pt = ge.load_package()

def main():
    sid, width, spp = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (6, 1920, 128)))
    scene = pt.Scene.build(sid, width=width, spp=spp, seed=1)
    for k, pool in ((1, 0), (2, 16 << 20), (2, 32 << 20), (3, 11 << 20), (4, 8 << 20), (1, 0)):
        devs = [0] * k
        pt.render_multi(scene, devs, spp=min(spp, 2 * k), seed=1, nan_policy=pt.PT_NAN_DROP, pool_paths=pool)  # warm-up: contexts, pools, scene copies
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            img, st = pt.render_multi(scene, devs, spp=spp, seed=2, nan_policy=pt.PT_NAN_DROP, pool_paths=pool)
            best = min(best, time.perf_counter() - t0)
        print(f"scene {sid} {width}px {spp} spp: {k} context(s) on device 0, pool {pool or 'default'}: wall {best * 1e3:.1f} ms, device (slowest share) {st.device_ms:.1f} ms, "
              f"{st.segments / best / 1e6:.1f} Mrays/s wall, mean {img.mean():.5f}", flush=True)
    pt.device_lib().pt_render_multi_release()

if __name__ == "__main__":
    main()
