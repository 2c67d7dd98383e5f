#!/usr/bin/env python3
"""Progressive render with checkpoint / resume and a noise-floor read-out.
usage: render_progressive.py --scene 6 --width 1920 --spp 4000 --batch 250 --checkpoint ck.npz --out scene6.png [--env-importance]
Re-running the same command after an interruption continues from the checkpoint."""
import argparse
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", type=int, default=6); ap.add_argument("--width", type=int, default=600); ap.add_argument("--spp", type=int, default=100)
    ap.add_argument("--batch", type=int, default=50); ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--checkpoint", default=""); ap.add_argument("--out", default=""); ap.add_argument("--env-importance", action="store_true")
    a = ap.parse_args()
    pt = ge.load_package()
    P = importlib.import_module("pt_b200.progressive")
    scene = pt.Scene.build(a.scene, width=a.width, spp=a.spp, seed=a.seed)
    ctx = pt.Context(0)
    dev = ctx.upload(scene)
    flags = 0
    if a.env_importance and scene.camera.env_is_map:
        dev.build_env_sampler(); flags = pt.PT_RENDER_ENV_IMPORTANCE
    pr = P.ProgressiveRender(dev, seed=a.seed, nan_policy=pt.PT_NAN_DROP, flags=flags)
    if a.checkpoint and os.path.exists(a.checkpoint):
        pr.load(a.checkpoint)
        print(f"resumed at {pr.spp} spp from {a.checkpoint}")
    while pr.spp < a.spp:
        pr.advance(min(a.batch, a.spp - pr.spp))
        if a.checkpoint:
            pr.save(a.checkpoint)
        print(f"{pr.spp:6d} spp  {pr.device_ms / 1e3:8.2f} s device  {pr.segments / max(pr.device_ms, 1e-9) / 1e3:8.1f} Mrays/s  expected relRMSE from noise {pr.noise_floor():.4f}", flush=True)
    mean, _ = pr.result()
    if a.out:
        pt.write_png(a.out, pt.tonemap_rgb8(mean))
        print("wrote", a.out)
    dev.close(); ctx.close()


if __name__ == "__main__":
    main()
