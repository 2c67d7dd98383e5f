#!/usr/bin/env python3
"""Small renders and ray batches through every kernel flavour, for compute-sanitizer (tools/sanitize.sh).
Sizes are just above the 65 536-ray floor of the multi-pass traversals so that those run too."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import __graft_entry__ as ge  # noqa: E402

pt = ge.load_package()


def main():
    from test_oracle import fog_world, mixed_lights_world, small_light_world, sun_world
    ctx = pt.Context(0)
    done = []

    def render(name, scene, spp=1, **kw):
        dev = ctx.upload(scene)
        if kw.get("flags", 0) & pt.PT_RENDER_ENV_IMPORTANCE:
            dev.build_env_sampler()
        img, st = dev.render(spp=spp, seed=3, nan_policy=pt.PT_NAN_DROP, **kw)
        done.append(f"{name}: {st.paths} paths, {st.segments} segments, {st.kernel_launches} launches, two-pass iterations {st.two_pass_iterations}")
        dev.close()

    s6 = pt.Scene.build(6, width=384, spp=1, seed=1)          # 384 x 216 = 82 944 paths per sample: above the floor
    render("scene 6: k_top + k_mesh_enter + k_mesh_walk + k_mesh_multi, forked shade, tail megakernel", s6)
    render("scene 6, 800 px (360 000 paths): k_mesh_multi on its side stream", pt.Scene.build(6, width=800, spp=1, seed=1))
    render("scene 6: BVH kernels, k_trace<DEFER> + k_trace_blas_refill", s6, flags=0x400000)
    render("scene 6: BVH kernels, grid-stride mesh rounds", s6, flags=0x400000 | 0x200000)
    render("scene 6: fused BVH kernel, unforked shade, wavefront iterations to the last path (no tail megakernel)", s6, flags=0x100000 | 0x2000 | 0x4000)
    render("scene 70: k_top + mesh rounds", pt.Scene.build(70, width=288, spp=1, seed=1))
    render("scene 3: k_top, quad light", pt.Scene.build(3, width=288, spp=1, seed=1))
    render("scene 3: NEE shadow paths", pt.Scene.build(3, width=96, spp=2, seed=1), flags=pt.PT_RENDER_NEE)
    render("scene 1: 4-wide BVH kernel (484 spheres)", pt.Scene.build(1, width=384, spp=1, seed=1))
    render("scene 5: principled grid + env map", pt.Scene.build(5, width=384, spp=1, seed=1))
    render("media: k_trace<VOL>", fog_world(pt, 96), spp=2)
    render("every light kind: k_shade<*, 2>", mixed_lights_world(pt, 96), spp=2)
    render("environment importance sampling: k_shade<*, 3>", sun_world(pt, 96, True), spp=2, flags=pt.PT_RENDER_ENV_IMPORTANCE)
    render("small pool (128 paths): many tiny iterations", pt.Scene.build(3, width=32, spp=2, seed=1), pool_paths=128)
    # parity entry points
    dev = ctx.upload(s6)
    rng = np.random.default_rng(1)
    rays = np.zeros(70000, dtype=pt.RAY_DTYPE)
    rays["origin"] = rng.uniform(-3, 3, size=(70000, 3)) + [0, 2, 5]
    d = rng.normal(size=(70000, 3)); rays["direction"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    for flags in (0, 0x400000, 0x100000):
        h, st = dev.trace_closest_wavefront(rays, flags=flags)
        done.append(f"pt_trace_closest_wavefront flags {flags:#x}: {int(h['hit'].sum())} hits, queue errors {st.queue_errors}")
    r2, h2, st = dev.trace_camera_wavefront(seed=5, sample=1)
    done.append(f"pt_trace_camera_wavefront: {int(h2['hit'].sum())} hits of {len(h2)}")
    dev.trace_closest(rays[:5000]); dev.trace_any(rays[:5000], np.full(5000, 10.0))
    multi, stm = pt.render_multi(s6, [0, 0], spp=2, seed=4, nan_policy=1)
    done.append(f"pt_render_multi on [0, 0]: {stm.paths} paths")
    dev.close(); ctx.close()
    print("\n".join(done))


if __name__ == "__main__":
    main()
