#!/usr/bin/env python3
"""Benchmark of the integrator hot path (Camera::render -> Camera::trace, reference src/camera.rs:79-228).

  python bench.py --gpus N --steps K --warmup W            our CUDA path, N ranks (torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W     the reference algorithm on the host CPU (oracle port)

Workload (BASELINE.json): scene 6 "everything" at FHD 1920x1080 (reference src/main.rs:371-532, `-q` geometry).
A step = one pass of the wavefront integrator over one batch of samples: `--spp` samples per pixel PER RANK
(weak scaling: rank g renders sample indices g, g+N, ... so N ranks deliver N*spp samples per pixel per step), followed
by the single reduce(sum) of the fp32 accumulators to rank 0.  Throughput does not depend on spp (4000 spp = 4000/spp steps).
metric: Mrays/s, ray = one World::intersect_all call (camera.rs:179); samples/s is reported alongside.
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
# device structs (csrc/device_scene.cuh, csrc/kernels.cuh): bytes one segment moves through HBM per stage
B_RAY, B_HIT, B_STATE = 56, 16, 96   # ray (o,d,time f64), HitRec, full path state (ray + throughput f64x3 + ids uint4)
B_NODE, B_REF, B_SPHERE, B_QUAD, B_TRI = 32, 32, 64, 128, 80
# dram__bytes_read.sum + dram__bytes_write.sum per ray of the traversal stage from the ncu --set full capture of this workload
# (profiles/r1_l_k_trace_ncu.md, first bounce iteration of 79 364 x 32 rays: k_trace<DEFER> 152.2 + 66.6 MB, mesh round 0
# 120.2 + 8.1 MB; rounds 1-2 are two orders of magnitude smaller).  The fused kernel moved 95 B per ray
# (profiles/r1_h_k_trace_ncu.md); the two-pass traversal re-reads the ray and the hit record of every queued mesh visit.
# The ray stream, hit records, queues and stack spills come from HBM, the 5 MB scene from L2.
K_TRACE_DRAM_BYTES_PER_RAY = (152.23e6 + 66.61e6 + 120.23e6 + 8.12e6) / (79364 * 32)


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
    in this image), so this arm times the oracle port with all host threads; each step = FHD x `ref_spp` samples."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pt, orc = ge.load_package(), ge.load_oracle()
    scene = pt.Scene.build(args.scene, width=args.width, spp=args.ref_spp, seed=1)
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
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "samples_per_s": paths / total,
            "config": {"workload": f"{SCENE_NAMES.get(args.scene, args.scene)}_{args.width}x{H}", "spp_per_step": args.ref_spp, "max_depth": 50,
                       "note": "C++ restatement of the reference (oracle/), not the Rust binary"},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": f"{args.width}x{H} x {args.ref_spp} spp per step, {args.steps} steps"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--spp", type=int, default=128, help="samples per pixel per rank per step")
    ap.add_argument("--ref-spp", type=int, default=4, help="samples per pixel per step of the CPU reference arm")
    ap.add_argument("--pool", type=int, default=0, help="in-flight path pool (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    # ---- per-stage CUDA-event times for the roofline of the dominant kernel (separate pass: the per-stage events
    #      are recorded on the launching stream around every launch; their overhead is kept out of `value`)
    ms_p, (segs_p, paths_p, _, _), pstats = timed(args.steps, 100, 1)
    trace_ms, shade_ms, gen_ms = (sum(getattr(s, k) for s in pstats) for k in ("trace_ms", "shade_ms", "raygen_ms"))
    iters = sum(s.iterations for s in pstats)
    segs_rank = sum(s.segments for s in pstats); paths_rank = sum(s.paths for s in pstats)
    # traversal work counters: one more step with the counting kernel variants (slower, so kept out of the stage times)
    _, _, cstats = timed(1, 100, 2)
    csegs = sum(s.segments for s in cstats)

    ctx.set_profiling(False)   # the e2e leg below runs the production path (forked shade streams, no per-stage events)

    # ---- end to end through the C ABI with host buffers: scene upload (H2D) + render + reduce + image D2H, every step
    def e2e_step(i):
        t_a = time.perf_counter()
        d2 = ctx.upload(scene)                      # pt_scene_create: H2D of the flattened scene from host memory
        t_b = time.perf_counter()
        st = step(i, d2)
        t_c = time.perf_counter()
        if rank == 0:
            host_out.copy_(accum, non_blocking=True)  # D2H of the step's result
        torch.cuda.synchronize()
        nbytes = d2.device_bytes
        d2.close()
        if rank == 0 and os.environ.get("PT_BENCH_VERBOSE"):
            print(f"e2e step {i}: upload {1e3 * (t_b - t_a):.1f} ms, render+reduce {1e3 * (t_c - t_b):.1f} ms, d2h+free {1e3 * (time.perf_counter() - t_c):.1f} ms", file=sys.stderr)
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

    if rank == 0:
        hbm, peak_src = peaks()
        # ---- CPU baseline on the box's host cores (bounded sample) + the oracle's work counters for the algorithmic bytes
        cpu = None
        n_node = n_sph = n_quad = n_tri = None
        orc = ge.load_oracle()
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline leg runs on rank 0 at N=1 only
            ost, cpu_spp = oracle_sample(pt, orc, scene, args.cpu_seconds)
            cpu = {"value": ost.segments / ost.seconds / 1e6, "unit": "Mrays/s", "cores": ost.threads, "kind": "port",
                   "sample": f"{WIDTH}x{H} x {cpu_spp} spp of the same scene ({ost.paths} paths, {ost.seconds:.1f} s)", "samples_per_s": ost.paths / ost.seconds}
        else:  # no timed CPU leg: the oracle only counts the reference's work per segment, on a small image of the same scene
            small = pt.Scene.build(args.scene, width=min(WIDTH, 240), spp=1, seed=1)
            ost, _ = oracle_sample(pt, orc, small, 0.0)
        n_node, n_sph, n_quad, n_tri = (getattr(ost, k) / ost.segments for k in ("boxes", "spheres", "quads", "triangles"))
        seg_per_path = segs_rank / max(paths_rank, 1)
        surv = 1.0 - 1.0 / seg_per_path                     # fraction of segments whose path continues
        b_trace = B_RAY + B_HIT + n_node * B_NODE + n_sph * B_SPHERE + n_quad * B_QUAD + n_tri * B_TRI + (n_sph + n_quad + n_tri) * B_REF
        b_shade = B_STATE + B_HIT + surv * B_STATE + 12.0   # read state+hit, write the survivor's state, ~one fp32x3 accumulate per path
        b_gen = B_STATE / seg_per_path
        dom = "k_trace" if trace_ms >= shade_ms else "k_shade"
        dom_ms, dom_b = (trace_ms, b_trace) if dom == "k_trace" else (shade_ms, b_shade)
        ach = segs_rank * dom_b / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        b_seg = b_trace + b_shade + b_gen
        # the same kernel judged by what the DEVICE requests (live counters of the profiling pass): ordered traversal with
        # culling touches far fewer nodes than the reference's un-narrowed recursion that the oracle counters describe
        d_pairs, d_refs, d_prims = (sum(getattr(s, k) for s in cstats) / max(csegs, 1) for k in ("node_pairs", "ref_boxes", "prim_tests"))
        prim_bytes = (n_sph * B_SPHERE + n_quad * B_QUAD + n_tri * B_TRI) / max(n_sph + n_quad + n_tri, 1e-9)
        b_trace_dev = B_RAY + B_HIT + d_pairs * 2 * B_NODE + d_refs * B_REF + d_prims * prim_bytes
        ach_dev = segs_rank * b_trace_dev / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else 0.0
        step_gbs = (segs / world) * b_seg / (ms * 1e-3) / 1e9
        line = {
            "metric": "Mrays/s", "value": segs / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "samples_per_s": paths / (ms * 1e-3),
            "config": {"workload": f"{SCENE_NAMES.get(args.scene, args.scene)}_{WIDTH}x{H}", "spp_per_step_per_gpu": args.spp, "max_depth": cam.max_depth,
                       "parallelism": f"spp-split x{world}", "pool_paths": args.pool or 32 << 20, "scene_device_bytes": int(dev.device_bytes),
                       "l2": f"per-step path-state working set (2 x min(32Mi, {WIDTH * H * args.spp}) paths x 112 B) exceeds the 126 MB L2 "
                             f"unless the step is tiny; the {dev.device_bytes / 1e6:.1f} MB scene (BVH, primitives, textures) is read through L2 by design"},
            "gpu_launches": int(launches), "segments_per_path": seg_per_path, "nonfinite_samples": int(nonfinite),
            "e2e": {"value": e2e_segs.item() / e2e_s.item() / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(e2e_stats[0][1]) * world,
                    "d2h_bytes_per_step": H * WIDTH * 3 * 4, "ms_per_step": 1e3 * e2e_s.item() / args.steps,
                    "what": "pt_scene_create (scene H2D) + pt_render_accumulate + reduce + D2H of the fp32 image, every step"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                         "traffic": (K_TRACE_DRAM_BYTES_PER_RAY * segs_rank / max(iters, 1)) if dom == "k_trace" and args.scene == SCENE else None,
                         "traffic_note": "DRAM bytes per launch = ncu dram read+write per ray of k_trace<DEFER> + mesh round 0 (profiles/r1_l_k_trace_ncu.md, 137 B) x rays per launch; far below the algorithmic bytes because the scene is L2-resident",
                         "peak_source": peak_src,
                         "kernel_note": "k_trace = the traversal stage of one wavefront iteration: k_trace (top level) + the k_trace_blas_refill mesh rounds on "
                                        "scenes with meshes; a 'launch' below is one iteration's stage, timed with CUDA events around it",
                         "bytes_per_segment": dom_b, "bytes_per_launch": dom_b * segs_rank / max(iters, 1),
                         "avg_launch_ms": dom_ms / max(iters, 1), "launches": iters,
                         "stage_ms_per_step": {"k_generate": gen_ms / args.steps, "k_trace": trace_ms / args.steps, "k_shade": shade_ms / args.steps},
                         "whole_step": {"bytes_per_segment": b_seg, "achieved": step_gbs, "frac": step_gbs / hbm},
                         # node fetches in 64-byte units: a binary pair = 1 unit, a 4-wide node = 2 units
                         "k_trace_device_counters": {"node_fetch_64B_units_per_segment": d_pairs, "ref_boxes_per_segment": d_refs, "f64_prim_tests_per_segment": d_prims,
                                                     "bytes_per_segment": b_trace_dev, "achieved": ach_dev, "frac": ach_dev / hbm,
                                                     "note": "bytes the device actually requests (mostly served by L1/L2: see config.scene_device_bytes)"},
                         "oracle_counters_per_segment": {"boxes": n_node, "spheres": n_sph, "quads": n_quad, "triangles": n_tri}},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    dev.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
