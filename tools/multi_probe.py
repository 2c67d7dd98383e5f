#!/usr/bin/env python3
"""Developer probe (GPU box with N GPUs): pt_render_multi — Camera::render on all GPUs from ONE process (a host thread per device,
spp split, the reduce on the device over NVLink peer memory) — scene 6 at FHD.  usage: multi_probe.py [spp]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402
pt = ge.load_package()
import torch  # noqa: E402

def main():
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n = torch.cuda.device_count()
    scene = pt.Scene.build(6, width=1920, spp=spp, seed=1)
    for devs in ([0], list(range(n)), list(range(n)), list(range(n))):
        t0 = time.perf_counter()
        img, st = pt.render_multi(scene, devs, spp=spp, seed=3, nan_policy=pt.PT_NAN_DROP)
        wall = time.perf_counter() - t0
        print(f"pt_render_multi {len(devs)} devices: wall {wall:.3f} s, device (slowest share) {st.device_ms * 1e-3:.3f} s, {st.segments / wall / 1e6:.0f} Mrays/s wall, "
              f"p2p shares {st.p2p_shares}, mean {img.mean():.5f}", flush=True)
    pt.device_lib().pt_render_multi_release()

if __name__ == "__main__":
    main()
