#!/usr/bin/env python3
"""Benchmark of the integrator hot path (Camera::render -> Camera::trace, reference src/camera.rs:79-228).

  python bench.py --gpus N --steps K --warmup W            our CUDA path, N ranks (torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W     the reference algorithm on the host CPU (oracle port)

Workload (BASELINE.json): scene 6 "everything" at FHD 1920x1080 (reference src/main.rs:371-532, `-q` geometry).
A step = one pass of the wavefront integrator over one batch of samples: `--spp` samples per pixel PER RANK
(weak scaling: rank g renders sample indices g, g+N, ... so N ranks deliver N*spp samples per pixel per step), followed
by the single reduce(sum) of the fp32 accumulators to rank 0.  Throughput does not depend on spp (4000 spp = 4000/spp steps).
metric: Mrays/s, ray = one World::intersect_all call (camera.rs:179); samples/s is reported alongside.
After the timed region the north-star run itself is done once (`target_render`): scene 6 FHD x 4000 spp split over the N
ranks, one reduce, tonemap, relRMSE against the reference's own demo/scene6.png (tests/golden/demo_scene6.npz).
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

SCENE, WIDTH = 6, 1920   # the headline workload; --scene / --width select the other BASELINE.json configs
SCENE_NAMES = {3: "scene3_cornell_box", 1: "scene1_bouncing_balls", 5: "scene5_principled_grid_envmap", 7: "scene7_normal_mapped_cornell",
               70: "scene7m_cornell_bunny_teapot", 6: "scene6_everything", 2: "scene2_earth", 4: "scene4_lights"}
# SURVEY §8(d) byte model (kept as roofline.survey_model): bytes one segment would move if every box / primitive the
# REFERENCE's un-narrowed recursion touches came from HBM
B_RAY, B_HIT, B_STATE = 64, 16, 96   # ray record (o, d, time f64 + pixel, sample), HitRec, full path state (ray + throughput + ids)
B_NODE, B_REF, B_SPHERE, B_QUAD, B_TRI = 32, 32, 64, 128, 80

# Must-move DRAM bytes and issue efficiency per kernel family, from the ncu --set full capture of THIS workload
# (profiles/r2_kernels_ncu.md, written by tools/collect_evidence.sh; `dram_B_per_seg` = (dram__bytes_read.sum +
# dram__bytes_write.sum) of the kernel's launch / rays of that wavefront iteration; issue = smsp__issue_active %,
# lanes = smsp__thread_inst_executed_per_inst_executed).  The live part of the roofline (kernel times, segments) is measured
# by this script; these per-segment constants are re-captured every round.
NCU_SOURCE = "profiles/r2_kernels_ncu.md"
try:
    NCU = json.load(open(os.path.join(ROOT, "profiles", "r2_kernels_ncu.json")))
except (OSError, ValueError):
    NCU = {}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(args, H, world):
    """The `config` both arms print (identical by construction: the reference arm times a bounded sample OF this workload)."""
    return {"workload": f"{SCENE_NAMES.get(args.scene, args.scene)}_{args.width}x{H}", "spp_per_step_per_gpu": args.spp, "max_depth": 50,
            "parallelism": f"spp-split x{world}",
            "l2": f"per-step path-state working set (2 x min(32Mi, {args.width * H * args.spp}) paths x 96 B) exceeds the 126 MB L2 unless the step is tiny; "
                  "the scene (BVH, primitives, textures; a few MB) is read through L2 by design"}


def oracle_sample(pt, orc, scene, seconds, threads=0):
    """The reference algorithm (C++ restatement, oracle/) on the host CPU over a bounded sample of the workload: one
    sample per pixel to learn the speed, then as many as fit in about `seconds` (at most 64), timed in one call."""
    ora = orc.OracleScene(scene.desc, pt)
    threads = threads or host_threads()
    _, st = ora.render(scene.camera, 1, seed=1, nan_policy=pt.PT_NAN_DROP, threads=threads)
    spp = int(max(1, min(64, seconds / max(st.seconds, 1e-3))))
    if spp > 1:
        _, st = ora.render(scene.camera, spp, seed=2, nan_policy=pt.PT_NAN_DROP, threads=threads)
    ora.close()
    return st, spp


def host_threads():
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, so ask explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Rust binary cannot be built (no cargo/rustc
    in this image), so this arm times the oracle port (built -O3, no fast-math, no FMA contraction) with all host threads.
    Same workload and `config` as our arm; each step is a bounded sample of it: `--ref-spp` of the step's `--spp` samples per
    pixel (Mrays/s does not depend on the sample count).  The scene description comes from the host mirror alone: the CUDA
    library is never loaded in this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pt, orc = ge.load_package(), ge.load_oracle()
    scene = pt.Scene.build(args.scene, width=args.width, spp=args.spp, seed=1)
    H = scene.image_height()
    ora = orc.OracleScene(scene.desc, pt)
    times, segs, paths = [], 0, 0
    for i in range(args.warmup + args.steps):
        _, st = ora.render(scene.camera, args.ref_spp, seed=1 + i, nan_policy=pt.PT_NAN_DROP, threads=host_threads())
        if i >= args.warmup:
            times.append(st.seconds); segs += st.segments; paths += st.paths
    total = sum(times)
    v = segs / total / 1e6
    cores = host_threads()
    sample = (f"each step renders {args.width}x{H} x {args.ref_spp} of the step's {args.spp} samples per pixel on {cores} host threads "
              f"({paths // args.steps} paths, {total / args.steps:.2f} s per step); C++ restatement of the reference (oracle/, -O3), not the Rust binary")
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "samples_per_s": paths / total, "config": workload_config(args, H, args.gpus),
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "cuda_library_loaded": "libptb200.so" in open("/proc/self/maps").read()}
    print(json.dumps(line))


def compare_with_demo(mean_full, scene):
    """relRMSE of SURVEY §8(d) against the reference's own 4000-spp demo render, on the 8x8-pixel cells of
    tests/golden/demo_scene<N>.npz that hold no clipped reference pixel."""
    g = np.load(os.path.join(ROOT, "tests", "golden", f"demo_scene{scene}.npz"))
    ref, clipped, w, h, f = g["mean"].astype(np.float64), g["clipped"].astype(np.float64), int(g["width"]), int(g["height"]), int(g["factor"])
    clip = 0.999 ** 2
    ours = np.clip(np.nan_to_num(np.asarray(mean_full, np.float64), nan=0.0, posinf=1.0), 0.0, clip)
    cells = ours.reshape(h // f, f, w // f, f, 3).mean(axis=(1, 3))
    ok = clipped == 0
    d = (cells[ok] - ref[ok]) ** 2 / (ref[ok] ** 2 + 1e-2)
    return float(np.sqrt(d.mean())), float(ok.mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--spp", type=int, default=128, help="samples per pixel per rank per step")
    ap.add_argument("--ref-spp", type=int, default=4, help="samples per pixel the CPU reference arm renders per step (a bounded sample of --spp)")
    ap.add_argument("--pool", type=int, default=0, help="in-flight path pool (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-target-render", action="store_true", help="skip the 4000-spp north-star render after the timed region")
    ap.add_argument("--target-spp", type=int, default=4000)
    ap.add_argument("--scene", type=int, default=SCENE, help="reference scene id (main.rs -s N; 70 = scene 7 + bunny + teapot)")
    ap.add_argument("--width", type=int, default=0, help="image width (default: 1920 for scene 6 like `-q`, else the reference's 600)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work of the cpu_baseline sample")
    args = ap.parse_args()
    args.width = args.width or (1920 if args.scene == SCENE else 600)
    if args.impl == "reference":
        return run_reference(args)
    assert args.warmup >= 3 or os.environ.get("PT_BENCH_ALLOW_SHORT"), "timing rules: at least 3 warm-up steps"

    import torch
    import torch.distributed as dist
    pt = ge.load_package()
    D = __import__("importlib").import_module("pt_b200.distributed")
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    WIDTH = args.width
    scene = pt.Scene.build(args.scene, width=WIDTH, spp=args.spp, seed=1)
    cam = scene.camera
    H = scene.image_height()
    ctx = pt.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    dev = ctx.upload(scene)
    accum = torch.zeros((H, WIDTH, 3), dtype=torch.float32, device="cuda")
    host_out = torch.empty((H, WIDTH, 3), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, dscene):
        """rank's share of one step: spp samples per pixel (indices rank + k*world, offset by the step), then the reduce."""
        accum.zero_()
        st = dscene.render_accumulate(accum.data_ptr(), camera=cam, spp=args.spp, seed=1000 + i, sample_begin=rank, sample_stride=world,
                                      nan_policy=pt.PT_NAN_DROP, pool_paths=args.pool)
        D.reduce_accumulators(accum, 0)
        return st

    def timed(n_steps, first, profiling):
        ctx.set_profiling(profiling)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        stats = [step(first + i, dev) for i in range(n_steps)]
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        tot = torch.tensor([float(sum(s.segments for s in stats)), float(sum(s.paths for s in stats)), float(sum(s.kernel_launches for s in stats)),
                            float(sum(s.nonfinite for s in stats))], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)   # max over ranks
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)  # whole-job totals
        return float(ms.item()), tot.tolist(), stats

    for i in range(args.warmup):
        step(i, dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, (segs, paths, launches, nonfinite), _ = timed(args.steps, 100, False)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel CUDA-event times for the roofline (separate pass: events are recorded on the launching stream around
    #      every kernel family of every wavefront iteration; their overhead is kept out of `value`)
    ctx.stage_ms(reset=True)
    ms_p, (segs_p, paths_p, _, _), pstats = timed(args.steps, 100, 1)
    fam_ms = ctx.stage_ms()
    trace_ms, shade_ms = (sum(getattr(s, k) for s in pstats) for k in ("trace_ms", "shade_ms"))
    iters = sum(s.iterations for s in pstats)
    segs_rank = sum(s.segments for s in pstats); paths_rank = sum(s.paths for s in pstats)
    # traversal work counters: one more step with the counting kernel variants (slower, so kept out of the stage times)
    _, _, cstats = timed(1, 100, 2)
    csegs = sum(s.segments for s in cstats)

    ctx.set_profiling(False)   # the e2e leg below runs the production path (forked shade streams, no per-stage events)

    # ---- end to end through the C ABI with host buffers: scene upload (H2D) + render + reduce + image D2H, every step
    def e2e_step(i):
        d2 = ctx.upload(scene)                      # pt_scene_create: H2D of the flattened scene from host memory
        st = step(i, d2)
        if rank == 0:
            host_out.copy_(accum, non_blocking=True)  # D2H of the step's result
        torch.cuda.synchronize()
        nbytes = d2.device_bytes
        d2.close()
        return st, nbytes
    for i in range(2):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    e2e_stats = [e2e_step(200 + i) for i in range(args.steps)]
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    e2e_segs = torch.tensor([float(sum(s.segments for s, _ in e2e_stats))], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_segs, op=dist.ReduceOp.SUM)

    # ---- the north-star run: scene 6 FHD x 4000 spp over the N ranks, one reduce, tonemap, relRMSE vs the reference's demo
    target = None
    if not args.no_target_render and args.scene == SCENE and WIDTH == 1920:
        barrier()
        t0 = time.perf_counter()
        d3 = ctx.upload(scene)
        tacc, tst = D.render_distributed(d3, cam, args.target_spp, seed=77, nan_policy=pt.PT_NAN_DROP, pool_paths=args.pool)
        rgb8 = None
        if rank == 0:
            rgb8 = ctx.tonemap_rgb8(tacc.data_ptr(), 1.0 / args.target_spp, H, WIDTH)  # camera.rs:109-114 on the device, RGB8 to the host
            mean = (tacc / float(args.target_spp)).cpu().numpy()
        torch.cuda.synchronize()
        wall = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        tt = torch.tensor([float(tst.segments if tst else 0), float(tst.paths if tst else 0), float(tst.device_ms if tst else 0)], device="cuda", dtype=torch.float64)
        tmax = tt.clone()
        if world > 1:
            dist.all_reduce(wall, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        d3.close()
        if rank == 0:
            err, frac = compare_with_demo(mean, SCENE)
            target = {"workload": f"scene6_everything_1920x1080 x {args.target_spp} spp, split over {world} GPU(s) (strong scaling: the total is fixed)",
                      "wall_s": wall.item(), "device_s_max_rank": tmax[2].item() * 1e-3, "mrays_per_s": tt[0].item() / wall.item() / 1e6,
                      "samples_per_s": tt[1].item() / wall.item(), "rel_rmse_vs_reference_demo": err, "tolerance": 0.03, "within_tolerance": bool(err < 0.03),
                      "cells_compared": frac, "rgb8_mean": float(rgb8.mean()),
                      "includes": "pt_scene_create (scene H2D), render, reduce(sum) to rank 0, tonemap + RGB8 D2H, mean image D2H"}

    if rank == 0:
        hbm, peak_src = peaks()
        # ---- CPU baseline on the box's host cores (bounded sample) + the oracle's work counters for the SURVEY byte model
        cpu = None
        orc = ge.load_oracle()
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline leg runs on rank 0 at N=1 only
            ost, cpu_spp = oracle_sample(pt, orc, scene, args.cpu_seconds)
            cpu = {"value": ost.segments / ost.seconds / 1e6, "unit": "Mrays/s", "cores": ost.threads, "kind": "port",
                   "sample": f"{WIDTH}x{H} x {cpu_spp} spp of the same scene ({ost.paths} paths, {ost.seconds:.1f} s); oracle/ built -O3, no fast-math",
                   "samples_per_s": ost.paths / ost.seconds}
        else:  # no timed CPU leg: the oracle only counts the reference's work per segment, on a small image of the same scene
            small = pt.Scene.build(args.scene, width=min(WIDTH, 240), spp=1, seed=1)
            ost, _ = oracle_sample(pt, orc, small, 0.0)
        n_node, n_sph, n_quad, n_tri = (getattr(ost, k) / ost.segments for k in ("boxes", "spheres", "quads", "triangles"))
        seg_per_path = segs_rank / max(paths_rank, 1)
        surv = 1.0 - 1.0 / seg_per_path                     # fraction of segments whose path continues
        b_trace = B_RAY + B_HIT + n_node * B_NODE + n_sph * B_SPHERE + n_quad * B_QUAD + n_tri * B_TRI + (n_sph + n_quad + n_tri) * B_REF
        b_shade = B_STATE + B_HIT + surv * B_STATE + 12.0   # read state+hit, write the survivor's state, ~one fp32x3 accumulate per path
        d_pairs, d_refs, d_prims = (sum(getattr(s, k) for s in cstats) / max(csegs, 1) for k in ("node_pairs", "ref_boxes", "prim_tests"))
        # ---- roofline.  The traversal and shade kernels of this workload are NOT bound by HBM (the scene lives in L2 / L1;
        # ncu: DRAM throughput 5-35 % of peak) but by instruction issue x divergence.  `frac` is therefore the DRAM-STREAM
        # fraction of the dominant kernel family: bytes that must cross HBM per segment (ncu dram__bytes of that kernel /
        # segments of the iteration, NCU_SOURCE) x segments / its live CUDA-event time / measured HBM peak.  The §8(d) byte
        # model is kept beside it as `survey_model`, and the resource that does bind is reported as `issue`.
        fams = {k: v for k, v in fam_ms.items() if v > 0}
        dom = max(fams, key=fams.get) if fams else "mesh_walk"
        kernels = {}
        for k, t_ms in fams.items():
            c = NCU.get(k, {})
            e = {"ms_per_step": t_ms / args.steps, "share_of_step": t_ms / max(sum(fams.values()), 1e-9)}
            if c:
                gbs = c["dram_B_per_seg"] * segs_rank / (t_ms * 1e-3) / 1e9
                e.update({"kernel": c.get("kernel"), "dram_B_per_segment": c["dram_B_per_seg"], "dram_GBps": gbs, "hbm_frac": gbs / hbm,
                          "issue_slot_pct": c["issue_pct"], "active_lanes": c["lanes"], "useful_lane_issue": c["issue_pct"] / 100.0 * c["lanes"] / 32.0})
            kernels[k] = e
        dk = kernels.get(dom, {})
        ach = dk.get("dram_GBps", 0.0)
        survey_gbs = segs_rank * b_trace / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else 0.0
        line = {
            "metric": "Mrays/s", "value": segs / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "samples_per_s": paths / (ms * 1e-3),
            "config": workload_config(args, H, world),
            "pool_paths": args.pool or 32 << 20, "scene_device_bytes": int(dev.device_bytes),
            "gpu_launches": int(launches), "segments_per_path": seg_per_path, "nonfinite_samples": int(nonfinite),
            "e2e": {"value": e2e_segs.item() / e2e_s.item() / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(e2e_stats[0][1]) * world,
                    "d2h_bytes_per_step": H * WIDTH * 3 * 4, "ms_per_step": 1e3 * e2e_s.item() / args.steps,
                    "what": "pt_scene_create (scene H2D) + pt_render_accumulate + reduce + D2H of the fp32 image, every step"},
            "roofline": {"bound": "hbm", "limiter": "issue", "kernel": dk.get("kernel", dom), "kernel_family": dom,
                         "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm if hbm else None,
                         "traffic": dk.get("dram_B_per_segment", 0.0) * segs_rank / max(iters, 1) if dk.get("dram_B_per_segment") else None,
                         "traffic_note": f"ncu dram__bytes_read+write of that kernel per launch = per-segment figure of {NCU_SOURCE} x segments per wavefront iteration",
                         "peak_source": peak_src,
                         "what": "DRAM-stream fraction of the dominant kernel family: must-move bytes (ncu) x live segments / live CUDA-event time / measured HBM peak; "
                                 "the kernels are bound by instruction issue x divergence, not HBM: see `issue` and `kernels`",
                         "avg_launch_ms": fams.get(dom, 0.0) / max(iters, 1), "launches": iters,
                         "issue": {k: v["useful_lane_issue"] for k, v in kernels.items() if "useful_lane_issue" in v},
                         "issue_note": f"issue-slot utilisation x active lanes / 32 per kernel family ({NCU_SOURCE})",
                         "kernels": kernels,
                         "stage_ms_per_step": {"traversal": trace_ms / args.steps, "shade": shade_ms / args.steps, "tail_megakernel": fam_ms.get("tail", 0.0) / args.steps},
                         "survey_model": {"bytes_per_segment": b_trace, "achieved": survey_gbs, "frac": survey_gbs / hbm,
                                          "note": "SURVEY §8(d): the reference's per-segment box / primitive counts (oracle counters) x struct sizes / traversal-stage time; "
                                                  "over-states the device, which fetches far fewer nodes and serves them from L1/L2",
                                          "oracle_counters_per_segment": {"boxes": n_node, "spheres": n_sph, "quads": n_quad, "triangles": n_tri},
                                          "shade_bytes_per_segment": b_shade},
                         "device_counters_per_segment": {"node_fetch_64B_units": d_pairs, "ref_boxes": d_refs, "f64_prim_tests": d_prims}},
            "cpu_baseline": cpu, "clocks": clocks, "target_render": target,
        }
        print(json.dumps(line))
    dev.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
