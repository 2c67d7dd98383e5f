#!/usr/bin/env python3
"""Developer probe (GPU box): how much would sorting rays for coherence buy the traversal stage?  The same bounce rays of
scene 6 go through pt_trace_closest_wavefront in path order, shuffled, and sorted on the host by (direction octant, Morton code of
the origin); per-kernel-family times from pt_debug_stage_ms."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

pt, orc = ge.load_package(), ge.load_oracle()


def morton3(q):
    def spread(x):
        x = x.astype(np.uint64) & 0x3FF
        x = (x | (x << 16)) & 0x30000FF
        x = (x | (x << 8)) & 0x300F00F
        x = (x | (x << 4)) & 0x30C30C3
        x = (x | (x << 2)) & 0x9249249
        return x
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


def main():
    scene = pt.Scene.build(6, width=1920, spp=4, seed=1)
    ora = orc.OracleScene(scene.desc, pt)
    rays = ora.dump_path_rays(scene.camera, 11, 1, 4, 1, 8388608)
    print(len(rays), "bounce rays")
    ctx = pt.Context(0); dev = ctx.upload(scene); ctx.set_profiling(1)
    o, d = rays["origin"], rays["direction"]
    octant = (d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << 1) | ((d[:, 2] < 0).astype(np.uint64) << 2)
    lo, hi = np.percentile(o, 1, axis=0), np.percentile(o, 99, axis=0)
    q = np.clip((o - lo) / (hi - lo) * 1023, 0, 1023).astype(np.uint32)
    orders = {"path order": np.arange(len(rays)), "shuffled": np.random.default_rng(1).permutation(len(rays)),
              "octant only": np.argsort(octant, kind="stable"), "octant + morton(origin)": np.argsort((octant << np.uint64(30)) | morton3(q), kind="stable"),
              "morton(origin) only": np.argsort(morton3(q), kind="stable")}
    ref = None
    for name, idx in orders.items():
        r = rays[idx]
        dev.trace_closest_wavefront(r)  # warm-up
        ctx.stage_ms(reset=True)
        for _ in range(3):
            h, st = dev.trace_closest_wavefront(r)
        sm = ctx.stage_ms()
        if ref is None:
            ref = h
        print(f"{name:26s} " + " ".join(f"{k} {v / 3:.3f}" for k, v in sm.items() if v) + f"  (two-pass {st.two_pass_iterations}, queue errors {st.queue_errors})")
    dev.close(); ctx.close()


if __name__ == "__main__":
    main()
