#!/usr/bin/env python3
"""Developer probe (GPU box): per-scene throughput and per-stage split of the wavefront integrator.
usage: perf_probe.py scene:width:spp[:pool] ...   e.g. 6:1920:16 3:600:100"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

pt = ge.load_package()


def main():
    ctx = pt.Context(0)
    for spec in sys.argv[1:]:
        parts = [int(x) for x in spec.split(":")]
        sid, width, spp = parts[:3]
        pool = parts[3] if len(parts) > 3 else 0
        flags = parts[4] if len(parts) > 4 else 0
        t0 = time.time(); scene = pt.Scene.build(sid, width=width, spp=spp, seed=1); tb = time.time() - t0
        t0 = time.time(); dev = ctx.upload(scene); tu = time.time() - t0
        dev.render(spp=min(spp, 2), seed=1, pool_paths=pool)  # warm-up
        for prof in (0, 1, 2):  # production path / per-stage events / + traversal work counters
            ctx.set_profiling(prof)
            ctx.stage_ms(reset=True)
            t0 = time.time(); img, st = dev.render(spp=spp, seed=2, nan_policy=pt.PT_NAN_DROP, pool_paths=pool, flags=flags); tw = time.time() - t0
            sm = ctx.stage_ms()
            print(f"scene {sid} {width}x{st.height} spp {spp} pool {pool or 'default'} flags {flags} prof={prof}:build {tb:.2f}s upload {tu * 1e3:.1f} ms ({dev.device_bytes / 1e6:.1f} MB) | "
                  f"device {st.device_ms:.1f} ms wall {tw * 1e3:.1f} ms | {st.segments / st.device_ms / 1e3:.1f} Mrays/s {st.paths / st.device_ms * 1e3:.3e} samples/s | "
                  f"seg/path {st.segments / st.paths:.2f} iters {st.iterations} launches {st.kernel_launches} nonfinite {st.nonfinite} | "
                  f"gen {st.raygen_ms:.1f} trace {st.trace_ms:.1f} shade {st.shade_ms:.1f} ms"
                  + (" [" + " ".join(f"{k} {v:.1f}" for k, v in sm.items() if v) + "]" if prof == 1 else "")
                  + (f" | per ray: {st.node_pairs / st.segments:.2f} x64B nodes, {st.ref_boxes / st.segments:.2f} ref boxes, {st.prim_tests / st.segments:.2f} f64 tests"
                     if prof == 2 else ""), flush=True)
        dev.close()
    ctx.close()


if __name__ == "__main__":
    main()
