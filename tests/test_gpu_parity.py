"""GPU parity tests (run on the B200 with -m gpu).  Every call goes through the C ABI of include/pt_b200.h; the
oracle (CPU restatement of the reference) is the checker on identical inputs.

Bars (BASELINE.json north_star): closest-hit primitive / instance IDs bit-exact and hit t within 4 ulp (we get 0);
BSDF eval/pdf within 1e-5 relative; renders with the same counter-based RNG streams agree per pixel, and renders at
the reference's demo resolution match the reference's own demo images within a stated relRMSE."""
import os

import numpy as np
import pytest

import helpers as H
from test_oracle import _material_world, _ray, fog_world, mixed_lights_world, small_light_world, sun_world, tie_world

pytestmark = pytest.mark.gpu

T_ULP = 4          # north_star: "hit t within 4 ulp"
BSDF_REL = 1e-5    # north_star: "BSDF eval/pdf within 1e-5 relative"


class Pair:
    """A scene uploaded to the device and rebuilt in the oracle from the same pt_scene_desc."""

    def __init__(self, pt, orc, ctx, scene):
        self.scene, self.dev, self.ora = scene, ctx.upload(scene), orc.OracleScene(scene.desc, pt)

    def close(self):
        self.dev.close(); self.ora.close()


@pytest.fixture(scope="module")
def pairs(pt, orc, ctx):
    cache = {}

    def get(scene_id, width=160):
        key = (scene_id, width)
        if key not in cache:
            cache[key] = Pair(pt, orc, ctx, pt.Scene.build(scene_id, width=width, spp=4, seed=1))
        return cache[key]
    yield get
    for p in cache.values():
        p.close()


def assert_hits_equal(pt, a, b, what):
    assert np.array_equal(a["hit"], b["hit"]), f"{what}: hit flags differ"
    both = a["hit"] == 1
    for f in ("prim_kind", "prim_index", "instance", "is_light", "material", "front_face"):
        assert np.array_equal(a[f][both], b[f][both]), f"{what}: {f} differs"
    assert H.ulp_diff(a["t"][both], b["t"][both]).max(initial=0) <= T_ULP, f"{what}: t differs by more than {T_ULP} ulp"
    for f, tol in (("point", 1e-9), ("geometric_normal", 1e-12), ("shading_normal", 1e-12), ("u", 1e-12), ("v", 1e-12)):
        assert np.abs(a[f][both] - b[f][both]).max(initial=0) <= tol, f"{what}: {f}"


# ---------------------------------------------------------------- closest hit
@pytest.mark.parametrize("scene_id", [3, 1, 7, 6, 70, 4, 2, 5])
def test_closest_hit_matches_oracle(pt, orc, ctx, pairs, scene_id):
    p = pairs(scene_id)
    cam = p.scene.camera
    h, w = p.scene.image_height(), cam.image_width
    rows, cols = np.divmod(np.arange(w * h, dtype=np.uint32), w)
    rays = orc.camera_rays(cam, 7, rows, cols, np.zeros_like(rows), pt)
    dev_rays = pt.camera_rays(ctx, cam, 7, rows, cols, np.zeros_like(rows))       # Camera::generate_ray on the device
    assert np.abs(rays["origin"] - dev_rays["origin"]).max() < 1e-12 and np.abs(rays["direction"] - dev_rays["direction"]).max() < 1e-12
    assert np.array_equal(rays["time"], dev_rays["time"])                         # same Philox stream
    assert_hits_equal(pt, p.dev.trace_closest(rays), p.ora.trace_closest(rays), f"scene {scene_id} camera rays")
    bounce = p.ora.dump_path_rays(cam, 11, 3, 2, 1, 150000)                       # incoherent rays from bounces >= 1
    assert len(bounce) > 1000
    assert_hits_equal(pt, p.dev.trace_closest(bounce), p.ora.trace_closest(bounce), f"scene {scene_id} bounce rays")
    # light-pdf style rays use t_min = 0 (quad.rs:90)
    assert_hits_equal(pt, p.dev.trace_closest(bounce[:20000], 0.0), p.ora.trace_closest(bounce[:20000], 0.0), f"scene {scene_id} t_min=0")


# ---------------------------------------------------------------- the render's own traversal stage, ID for ID
TWO_PASS_MIN = 1 << 16   # launch_trace (csrc/api.cu) takes the two-pass path from this many live rays up


@pytest.mark.parametrize("scene_id,width", [(6, 400), (70, 300)])
def test_render_traversal_stage_matches_oracle(pt, orc, ctx, pairs, scene_id, width):
    """pt_trace_closest_wavefront pushes a host batch through launch_trace — the kernels, grids and queues that
    pt_render_accumulate launches per wavefront iteration (k_top + k_mesh_enter / k_mesh_walk on these scenes) — and the result
    must equal the oracle's World::intersect_all ID for ID (t within 4 ulp; we get 0): camera rays, >= 150 k bounce rays,
    t_min = 1e-3 and 0, persistent-lane rounds and plain grid-stride rounds (flag 0x200000), and the fused kernel (0x100000)."""
    p = pairs(scene_id, width)
    cam = p.scene.camera
    h, w = p.scene.image_height(), cam.image_width
    rows, cols = np.divmod(np.arange(w * h, dtype=np.uint32), w)
    rays = orc.camera_rays(cam, 7, rows, cols, np.zeros_like(rows), pt)
    assert len(rays) >= TWO_PASS_MIN
    want = p.ora.trace_closest(rays)
    got, st = p.dev.trace_closest_wavefront(rays)
    assert st.two_pass_iterations == 1 and st.segments == len(rays)        # the production flavour ran, not the fused kernel
    assert st.queue_errors == 0                                              # every ray joined the shade queue of its hit's material, once
    assert_hits_equal(pt, got, want, f"scene {scene_id} camera rays, two-pass")
    bounce = p.ora.dump_path_rays(cam, 11, 1, 4, 1, 400000)                 # incoherent rays from bounces >= 1
    assert len(bounce) >= 150000
    want = p.ora.trace_closest(bounce)
    frac_mesh = (want["prim_kind"][want["hit"] == 1] == 2).mean()
    assert frac_mesh > 0.05                                                  # a good share of them ends on a mesh triangle
    # 0: flat top level + k_mesh_enter / k_mesh_walk rounds (what the benchmark runs); 0x400000: BVH kernels, k_trace<DEFER> +
    # k_trace_blas_refill; + 0x200000: grid-stride mesh rounds; 0x100000: one fused BVH kernel
    for flags, two_pass in ((0, 1), (0x400000, 1), (0x400000 | 0x200000, 1), (0x100000, 0)):
        got, st = p.dev.trace_closest_wavefront(bounce, flags=flags)
        assert st.two_pass_iterations == two_pass and st.queue_errors == 0, hex(flags)
        assert_hits_equal(pt, got, want, f"scene {scene_id} bounce rays, flags {flags:#x}")
    want0 = p.ora.trace_closest(bounce, 0.0)                                 # light-pdf style rays use t_min = 0 (quad.rs:90)
    got0, st = p.dev.trace_closest_wavefront(bounce, 0.0)
    assert st.two_pass_iterations == 1
    assert_hits_equal(pt, got0, want0, f"scene {scene_id} t_min=0, two-pass")
    assert scene_id != 6 or not np.array_equal(want0["t"], want["t"])      # on scene 6 a few of these rays do hit inside [0, 1e-3)
    small, st = p.dev.trace_closest_wavefront(bounce[:TWO_PASS_MIN - 1])     # one ray under the floor: the fused kernel
    assert st.two_pass_iterations == 0
    assert_hits_equal(pt, small, want[:TWO_PASS_MIN - 1], f"scene {scene_id} below the two-pass floor")


@pytest.mark.parametrize("scene_id,width", [(6, 400), (70, 300), (3, 128), (5, 160), (1, 160), (3, 150), (4, 1001)])  # 150 / 1001: no 8x4 pixel tiles, row = pixel / width
def test_start_of_path_stage_matches_oracle(pt, orc, ctx, pairs, scene_id, width):
    """pt_trace_camera_wavefront = the first iteration of a render: on flat scenes k_top<PRIMARY> generates the camera ray in
    registers and traces it in the same launch (there is no k_generate).  Rays equal the oracle's generate_ray for the same
    (seed, pixel, sample); hits equal the oracle's World::intersect_all of those very rays, ID for ID."""
    p = pairs(scene_id, width)
    cam = p.scene.camera
    h, w = p.scene.image_height(), cam.image_width
    rows, cols = np.divmod(np.arange(w * h, dtype=np.uint32), w)
    for sample in (0, 3):
        rays, hits, st = p.dev.trace_camera_wavefront(seed=7, sample=sample)
        want_rays = orc.camera_rays(cam, 7, rows, cols, np.full_like(rows, sample), pt)
        assert np.abs(rays["origin"] - want_rays["origin"]).max() < 1e-12 and np.abs(rays["direction"] - want_rays["direction"]).max() < 1e-12
        assert np.array_equal(rays["time"], want_rays["time"])
        assert (st.two_pass_iterations > 0) == (scene_id in (6, 70)) and st.queue_errors == 0
        assert_hits_equal(pt, hits, p.ora.trace_closest(rays), f"scene {scene_id} start of path, sample {sample}")
    rays2, hits2, st = p.dev.trace_camera_wavefront(seed=7, sample=3, flags=0x400000)      # k_generate + the BVH kernels
    assert np.array_equal(rays2, rays) and st.two_pass_iterations == (1 if scene_id in (6, 70) else 0)
    assert_hits_equal(pt, hits2, hits, f"scene {scene_id} start of path, BVH kernels")


def test_render_traversal_stage_other_flavours(pt, orc, ctx, pairs):
    """The same entry on scenes without meshes (binary-pair and 4-wide fused flavours) and with media."""
    for scene_id in (3, 1):
        p = pairs(scene_id)
        bounce = p.ora.dump_path_rays(p.scene.camera, 11, 2, 2, 1, 100000)
        got, st = p.dev.trace_closest_wavefront(bounce)
        assert st.two_pass_iterations == 0
        assert_hits_equal(pt, got, p.ora.trace_closest(bounce), f"scene {scene_id} wavefront stage")
    scene = fog_world(pt, 64)
    dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
    bounce = ora.dump_path_rays(scene.camera, 5, 1, 2, 1, 60000)
    got, st = dev.trace_closest_wavefront(bounce)
    assert_hits_equal(pt, got, ora.trace_closest(bounce), "fog world wavefront stage")
    dev.close(); ora.close()


@pytest.mark.parametrize("scene_id", [3, 6, 1])
def test_any_hit_matches_oracle(pt, pairs, scene_id):
    p = pairs(scene_id)
    rays = p.ora.dump_path_rays(p.scene.camera, 5, 5, 1, 1, 40000)
    rng = np.random.default_rng(scene_id)
    t_max = rng.uniform(0.1, 30.0 if scene_id != 3 else 600.0, size=len(rays))
    a, b = p.dev.trace_any(rays, t_max), p.ora.trace_any(rays, t_max)
    assert np.array_equal(a, b) and 0 < b.mean() < 1


@pytest.mark.parametrize("order,winner", [(["quad", "quad"], 1), (["quad", "quad", "quad"], 2), (["sphere", "sphere"], 0),
                                          (["quad", "cuboid"], 1), (["cuboid", "quad"], 1), (["quad", "inst_cuboid"], 1), (["inst_cuboid", "quad"], 1)])
@pytest.mark.parametrize("bvh,filler", [(True, 0), (False, 0), (True, 9)])
def test_exact_tie_rules(pt, orc, ctx, order, winner, bvh, filler):
    """Exact-t ties resolve as the reference's recursion does — same leaf, linear list, across leaves."""
    scene = tie_world(pt, order, bvh, filler)
    dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
    r = _ray(pt, (0, 0, 0), (0, 0, -1))
    a, b = dev.trace_closest(r)[0], ora.trace_closest(r)[0]
    assert a["t"] == 5.0 and b["t"] == 5.0
    assert a["material"] == b["material"] and a["prim_index"] == b["prim_index"] and a["instance"] == b["instance"]
    if filler == 0:
        assert a["material"] == winner
    dev.close(); ora.close()


def test_light_object_tie_and_empty_world(pt, orc, ctx):
    w = pt.World()
    w.add_light(pt.Quad((-1, -1, -5), (2, 0, 0), (0, 2, 0), pt.DiffuseLight((5, 5, 5))))
    w.add_object(pt.Quad((-1, -1, -5), (2, 0, 0), (0, 2, 0), pt.DiffuseBRDF((0.5, 0.5, 0.5))))
    w.build_bvh()
    scene = pt.Scene.from_world(w, pt.make_camera(8))
    dev = ctx.upload(scene)
    h = dev.trace_closest(_ray(pt, (0, 0, 0), (0, 0, -1)))[0]
    assert h["hit"] == 1 and h["is_light"] == 0 and h["t"] == 5.0          # world.rs:55-59: object beats light
    dev.close()
    empty = pt.Scene.from_world(pt.World(), pt.make_camera(8, env_color=(0.25, 0.5, 0.75), samples_per_pixel=2))
    dev = ctx.upload(empty)
    assert dev.trace_closest(_ray(pt, (0, 0, 0), (0, 0, -1)))[0]["hit"] == 0
    img, st = dev.render(spp=2)
    assert np.allclose(img, [0.25, 0.5, 0.75]) and st.segments == st.paths      # every path: one miss, env colour
    dev.close()


def test_ragged_meshes_and_motion_blur(pt, orc, ctx):
    """A mesh with vertex normals + texcoords (Q6, Q7), an un-built (linear) world and moving spheres."""
    rng = np.random.default_rng(3)
    n = 40
    pos = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    idx = rng.integers(0, n, size=(60, 3)).astype(np.uint32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    tex = rng.uniform(size=(n, 2)).astype(np.float32)
    mat = pt.DiffuseBRDF((0.5, 0.6, 0.7))
    w = pt.World()
    w.add_object(pt.Instance(pt.TriangleMesh.from_arrays(1.3, pos, idx, mat, tex, nrm), (0.2, 1.0, 0.1), 0.7, (0.1, 0.2, -4)))
    w.add_object(pt.TriangleMesh.from_arrays(0.9, pos, idx, mat))
    for k in range(12):
        w.add_object(pt.Sphere.new_moving(0.3, rng.uniform(-2, 2, 3), rng.uniform(-2, 2, 3), mat))
    rays = np.zeros(20000, dtype=pt.RAY_DTYPE)
    rays["origin"] = rng.uniform(-3, 3, size=(20000, 3))
    rays["direction"] = H.rand_dirs(rng, 20000)
    rays["time"] = rng.uniform(size=20000)
    for build in (False, True):
        if build:
            w.build_bvh()
        scene = pt.Scene.from_world(w, pt.make_camera(8))
        dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
        a, b = dev.trace_closest(rays), ora.trace_closest(rays)
        assert b["hit"].mean() > 0.05
        assert_hits_equal(pt, a, b, f"ragged world (bvh={build})")
        dev.close(); ora.close()


# ---------------------------------------------------------------- BSDF
def test_shared_reciprocal_division_is_the_ieee_division(ctx):
    """`DVec3 / f64` of the shade kernels shares one reciprocal between its three divisions (device_scene.cuh: div3_shared).  On
    6e8 quotients — every exponent incl. denormals, inf, NaN, zero, and a stream restricted to the ranges shading divides in —
    the bits must be those of the `/` operator."""
    assert ctx.div_check(200_000_000, seed=1) == 0
    assert ctx.div_check(1_000_003, seed=0xDEADBEEFCAFE) == 0


def test_bsdf_eval_pdf_sample_match_oracle(pt, orc, ctx):
    scene, n_mat = _material_world(pt)
    dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
    rng = np.random.default_rng(2)
    for m in range(n_mat):
        for tilt in (0.0, 0.3):
            q = H.random_bsdf_queries(pt, rng, 20000, tilt)
            a, b = dev.bsdf_eval_pdf(m, q), ora.bsdf_eval_pdf(m, q)
            for f in ("eval", "pdf", "emitted"):
                assert H.max_rel_err(a[f], b[f]) <= BSDF_REL, f"material {m} {f}"
            u8 = rng.uniform(size=(20000, 8))
            sa, sb = dev.bsdf_sample(m, q, u8), ora.bsdf_sample(m, q, u8)
            assert np.array_equal(sa["valid"], sb["valid"]) and np.array_equal(sa["n_uniforms"], sb["n_uniforms"])
            ok = (sb["valid"] == 1) & np.isfinite(sb["dir"]).all(axis=1)
            assert np.abs(sa["dir"][ok] - sb["dir"][ok]).max(initial=0) < 1e-9, f"material {m} sampled direction"
    dev.close(); ora.close()


@pytest.mark.parametrize("scene_id", [3, 7, 6, 5])
def test_scene_materials_match_oracle(pt, pairs, scene_id):
    """Every material of the shipped scenes (textures, normal maps, principled parameter sets)."""
    p = pairs(scene_id)
    rng = np.random.default_rng(scene_id)
    for m in range(H.desc_header(p.scene)["n_materials"]):
        q = H.random_bsdf_queries(pt, rng, 4000)
        q["point"] *= 100.0
        a, b = p.dev.bsdf_eval_pdf(m, q), p.ora.bsdf_eval_pdf(m, q)
        for f in ("eval", "pdf", "emitted"):
            assert H.max_rel_err(a[f], b[f]) <= BSDF_REL, f"scene {scene_id} material {m} {f}"


def test_lights_sample_pdf_match_oracle(pt, orc, ctx, pairs):
    rng = np.random.default_rng(4)
    n = 20000
    for scene_id in (3, 7):
        p = pairs(scene_id)
        o = rng.uniform(10, 540, size=(n, 3)); t = rng.uniform(size=n); u = rng.uniform(size=(n, 4))
        da, va, pa = p.dev.lights_sample_pdf(o, t, u)
        db, vb, pb = p.ora.lights_sample_pdf(o, t, u)
        assert np.array_equal(va, vb) and np.abs(da - db).max() < 1e-12 and H.max_rel_err(pa, pb) < 1e-9
    # a sphere light (sphere.rs:110-135, Q8) next to a quad light: list.rs:78-96 averages the pdf over all lights (Q10)
    w = pt.World()
    w.add_light(pt.Sphere.new_still(0.5, (0, 3, 0), pt.DiffuseLight((4, 4, 4))))
    w.add_light(pt.Quad((-1, 4, -1), (2, 0, 0), (0, 0, 2), pt.DiffuseLight((2, 2, 2))))
    w.add_object(pt.Quad((-5, 0, -5), (10, 0, 0), (0, 0, 10), pt.DiffuseBRDF((0.5, 0.5, 0.5))))
    w.build_bvh()
    scene = pt.Scene.from_world(w, pt.make_camera(8))
    dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
    o = rng.uniform(-2, 2, size=(n, 3)) * [1, 0.2, 1]; t = rng.uniform(size=n); u = rng.uniform(size=(n, 4))
    da, va, pa = dev.lights_sample_pdf(o, t, u)
    db, vb, pb = ora.lights_sample_pdf(o, t, u)
    assert np.array_equal(va, vb) and np.abs(da - db).max() < 1e-9 and H.max_rel_err(pa, pb) < 1e-7
    dev.close(); ora.close()


def test_every_light_kind_matches_oracle(pt, orc, ctx):
    """World.lights holding a sphere, quad, cuboid, mesh and instances of quad / cuboid / mesh (cuboid.rs:74-80,
    mesh.rs:122-141,213-219, instance.rs:64-75): sample + mixture pdf per query, closest hits, and a render that
    follows the oracle's paths sample for sample."""
    scene = mixed_lights_world(pt, 64)
    dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
    rng = np.random.default_rng(14)
    n = 20000
    o = rng.uniform(-2, 2, size=(n, 3)) * [1, 0.1, 1] + [0, 0.5, 0]; t = rng.uniform(size=n); u = rng.uniform(size=(n, 4))
    da, va, pa = dev.lights_sample_pdf(o, t, u)
    db, vb, pb = ora.lights_sample_pdf(o, t, u)
    assert vb.all() and np.array_equal(va, vb) and np.abs(da - db).max() < 1e-9
    assert np.array_equal(pa > 0, pb > 0) and H.max_rel_err(pa, pb) < 1e-7
    rays = ora.dump_path_rays(scene.camera, 3, 1, 4, 0, 60000)
    assert_hits_equal(pt, dev.trace_closest(rays), ora.trace_closest(rays), "mixed lights")
    # mesh.rs:122-129 samples outside its triangle half of the time; where the BSDF pdf is 0 too the weight is 0/0, so this
    # scene poisons many samples (Q32).  PT_NAN_REFERENCE walks those paths on like the reference: same segment count.
    img, st = dev.render(spp=8, seed=15, nan_policy=pt.PT_NAN_REFERENCE)
    ref, ost = ora.render(scene.camera, 8, seed=15, nan_policy=pt.PT_NAN_REFERENCE)
    assert st.paths == ost.paths and abs(int(st.segments) - int(ost.segments)) <= 64
    fin, fin_dev = np.isfinite(ref).all(axis=2), np.isfinite(img).all(axis=2)
    print(f"finite pixels: oracle {fin.mean():.3f} device {fin_dev.mean():.3f} differ {(fin != fin_dev).mean():.4f}")
    assert (fin != fin_dev).mean() < 0.01 and 0.3 < fin.mean()      # a diverged path may poison a pixel on one side only
    fin &= fin_dev
    d = np.abs(img - ref)[fin].max(axis=1)
    assert (d > 1e-4 * np.maximum(ref[fin].max(axis=1), 1.0)).mean() < 0.02
    # PT_NAN_DROP: a poisoned path ends where it is poisoned and keeps what it had gathered (include/pt_b200.h)
    img, st = dev.render(spp=8, seed=15, nan_policy=pt.PT_NAN_DROP)
    ref, ost = ora.render(scene.camera, 8, seed=15, nan_policy=pt.PT_NAN_DROP)
    d = np.abs(img - ref).max(axis=2)
    assert np.isfinite(img).all() and st.nonfinite > 100
    assert abs(int(st.nonfinite) - int(ost.nonfinite)) <= 8 and abs(int(st.segments) - int(ost.segments)) <= 64
    assert (d > 1e-4 * np.maximum(ref.max(axis=2), 1.0)).mean() < 0.02 and H.rel_rmse(img, ref) < 0.05
    dev.close(); ora.close()


# ---------------------------------------------------------------- environment importance sampling (ours)
@pytest.mark.parametrize("with_light", [False, True])
def test_env_importance_sampling_matches_oracle(pt, orc, ctx, with_light):
    """PT_RENDER_ENV_IMPORTANCE is not reference behaviour (flag-gated, off by default); the oracle restates the same
    sampler, so the device is checked sample for sample, and against the reference estimator's expectation."""
    scene = sun_world(pt, 64, with_light)
    dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
    with pytest.raises(pt.PtError):                                       # the sampler has to be built first
        dev.render(spp=1, flags=pt.PT_RENDER_ENV_IMPORTANCE)
    dev.build_env_sampler(); ora.build_env_sampler(scene.camera.env_image)
    u = np.random.default_rng(22).uniform(size=(50000, 2))
    (da, pa), (db, pb) = dev.env_sample_pdf(u), ora.env_sample_pdf(u)
    assert np.abs(da - db).max() < 1e-12 and H.max_rel_err(pa, pb) < 1e-9
    img, st = dev.render(spp=16, seed=23, nan_policy=pt.PT_NAN_DROP, flags=pt.PT_RENDER_ENV_IMPORTANCE)
    ref, ost = ora.render(scene.camera, 16, seed=23, nan_policy=pt.PT_NAN_DROP, flags=pt.PT_RENDER_ENV_IMPORTANCE)
    assert st.paths == ost.paths and abs(int(st.segments) - int(ost.segments)) <= 64
    d = np.abs(img - ref).max(axis=2)
    assert (d > 1e-4 * np.maximum(ref.max(axis=2), 1.0)).mean() < 0.02 and H.rel_rmse(img, ref) < 0.05
    # same expectation as the reference's estimator, much less noise (400 spp each, two seeds)
    base = [dev.render(spp=400, seed=s, nan_policy=pt.PT_NAN_DROP)[0] for s in (1, 2)]
    envs = [dev.render(spp=400, seed=s, nan_policy=pt.PT_NAN_DROP, flags=pt.PT_RENDER_ENV_IMPORTANCE)[0] for s in (1, 2)]
    n_base, n_env = H.rel_rmse(base[0], base[1]), H.rel_rmse(envs[0], envs[1])
    print(f"lights={with_light}: noise at 400 spp: reference estimator {n_base:.4f}, env importance sampling {n_env:.4f}")
    assert n_env < (0.9 if with_light else 0.5) * n_base                   # with the quad light on, its own noise remains
    assert abs((base[0] + base[1]).mean() - (envs[0] + envs[1]).mean()) < 0.03 * (base[0] + base[1]).mean()
    dev.close(); ora.close()


def test_env_importance_sampling_scene5_and_default_unchanged(pt, orc, pairs):
    """Scene 5 (87 MB sky): device vs oracle sample for sample with the sampler on; with the flag off the result does not
    depend on whether a sampler was built."""
    p = pairs(5, 96)
    before, _ = p.dev.render(spp=4, seed=9, nan_policy=pt.PT_NAN_DROP)
    p.dev.build_env_sampler(); p.ora.build_env_sampler(p.scene.camera.env_image)
    after, _ = p.dev.render(spp=4, seed=9, nan_policy=pt.PT_NAN_DROP)
    assert np.allclose(before, after, rtol=1e-6, atol=1e-7)
    img, st = p.dev.render(spp=8, seed=9, nan_policy=pt.PT_NAN_DROP, flags=pt.PT_RENDER_ENV_IMPORTANCE)
    ref, ost = p.ora.render(p.scene.camera, 8, seed=9, nan_policy=pt.PT_NAN_DROP, flags=pt.PT_RENDER_ENV_IMPORTANCE)
    assert abs(int(st.segments) - int(ost.segments)) <= max(64, ost.segments // 2000)
    d = np.abs(img - ref).max(axis=2)
    assert (d > 1e-4 * np.maximum(ref.max(axis=2), 1.0)).mean() < 0.02 and H.rel_rmse(img, ref) < 0.05


# ---------------------------------------------------------------- next-event estimation (ours)
def test_nee_matches_oracle(pt, orc, ctx, pairs):
    """PT_RENDER_NEE: shadow paths ride the ordinary wavefront pool.  The oracle resolves each shadow ray on the spot;
    the sums must agree sample for sample — also when the pool is so small that spawning paths are throttled — and the
    NEE image must converge to the reference estimator's with less noise."""
    scene = small_light_world(pt, 64, sphere_light=True)
    dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
    for policy in (pt.PT_NAN_DROP, pt.PT_NAN_REFERENCE):
        img, st = dev.render(spp=8, seed=41, nan_policy=policy, flags=pt.PT_RENDER_NEE)
        ref, ost = ora.render(scene.camera, 8, seed=41, nan_policy=policy, flags=pt.PT_RENDER_NEE)
        assert st.paths == ost.paths and abs(int(st.segments) - int(ost.segments)) <= max(64, ost.segments // 2000)
        fin = np.isfinite(ref).all(axis=2) & np.isfinite(img).all(axis=2)
        assert fin.mean() > 0.99
        d = np.abs(img - ref)[fin].max(axis=1)
        assert (d > 1e-4 * np.maximum(ref[fin].max(axis=1), 1.0)).mean() < 0.02
    assert st.segments > 1.5 * dev.render(spp=8, seed=41)[1].segments        # the shadow rays are traced and counted
    small, st_small = dev.render(spp=8, seed=41, nan_policy=pt.PT_NAN_REFERENCE, flags=pt.PT_RENDER_NEE, pool_paths=3000)
    assert st_small.segments == st.segments and np.allclose(np.nan_to_num(small), np.nan_to_num(img), rtol=2e-5, atol=2e-6)
    dev.close(); ora.close()
    # convergence on the consistent (quad-light) variant: same mean as the reference estimator, less noise
    scene = small_light_world(pt, 64)
    dev = ctx.upload(scene)
    base = [dev.render(spp=512, seed=s, nan_policy=pt.PT_NAN_DROP)[0] for s in (1, 2)]
    nee = [dev.render(spp=512, seed=s, nan_policy=pt.PT_NAN_DROP, flags=pt.PT_RENDER_NEE)[0] for s in (1, 2)]
    n_base, n_nee = H.rel_rmse(base[0], base[1]), H.rel_rmse(nee[0], nee[1])
    a, b = np.clip((base[0] + base[1]) / 2, 0, 0.999), np.clip((nee[0] + nee[1]) / 2, 0, 0.999)
    print(f"noise at 512 spp: reference estimator {n_base:.3f}, NEE {n_nee:.3f}; clipped means {a.mean():.4f} vs {b.mean():.4f}")
    assert n_nee < 0.7 * n_base and abs(a.mean() - b.mean()) < 0.05 * a.mean()
    dev.close()
    p = pairs(3, 96)                                                            # the Cornell box of the reference
    img, st = p.dev.render(spp=8, seed=43, nan_policy=pt.PT_NAN_DROP, flags=pt.PT_RENDER_NEE)
    ref, ost = p.ora.render(p.scene.camera, 8, seed=43, nan_policy=pt.PT_NAN_DROP, flags=pt.PT_RENDER_NEE)
    assert abs(int(st.segments) - int(ost.segments)) <= max(64, ost.segments // 2000) and H.rel_rmse(img, ref) < 0.05


# ---------------------------------------------------------------- constant-density media (ours)
def test_volumes_match_oracle(pt, orc, ctx):
    """pt_volume (the reference's volume.rs is a stub, so parity is against the oracle's statement of include/pt_b200.h):
    scatter distances of ray batches, closest / any hits in a world with fog spheres, smoke cuboids and an instanced
    medium, Beer-Lambert transmittance on the device, and a render that follows the oracle sample for sample."""
    scene = fog_world(pt, 64)
    dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
    cam = scene.camera
    rows, cols = np.divmod(np.arange(64 * 64, dtype=np.uint32), 64)
    rays = orc.camera_rays(cam, 3, rows, cols, np.zeros_like(rows), pt)
    a, b = dev.trace_closest(rays), ora.trace_closest(rays)
    assert (b["prim_kind"][b["hit"] == 1] == 6).mean() > 0.05            # a good share of the camera rays scatter in a medium
    assert_hits_equal(pt, a, b, "fog world camera rays")
    bounce = ora.dump_path_rays(cam, 5, 1, 2, 1, 60000)
    assert len(bounce) > 5000
    assert_hits_equal(pt, dev.trace_closest(bounce), ora.trace_closest(bounce), "fog world bounce rays")
    t_max = np.random.default_rng(31).uniform(0.2, 8.0, size=len(bounce))
    assert np.array_equal(dev.trace_any(bounce, t_max), ora.trace_any(bounce, t_max))
    n = 200000                                                            # Beer-Lambert on the device itself
    probe = np.zeros(n, dtype=pt.RAY_DTYPE)
    probe["origin"] = (-1.2, 1.0, 5.0); probe["direction"] = (0, 0, -1)   # through the centre of the fog sphere: chord 2, density 0.8
    h = dev.trace_closest(probe)
    assert abs((h["prim_kind"] == 6).mean() - (1 - np.exp(-0.8 * 2.0))) < 0.004
    img, st = dev.render(spp=8, seed=33, nan_policy=pt.PT_NAN_DROP)
    ref, ost = ora.render(cam, 8, seed=33, nan_policy=pt.PT_NAN_DROP)
    assert st.paths == ost.paths and abs(int(st.segments) - int(ost.segments)) <= max(64, ost.segments // 2000)
    d = np.abs(img - ref).max(axis=2)
    assert (d > 1e-4 * np.maximum(ref.max(axis=2), 1.0)).mean() < 0.02 and H.rel_rmse(img, ref) < 0.05
    dev.close(); ora.close()


# ---------------------------------------------------------------- renders
_ORACLE_RENDERS = {}


@pytest.mark.parametrize("scene_id,width,spp", [(3, 96, 8), (1, 128, 8), (7, 96, 8), (6, 128, 6), (5, 128, 8), (4, 128, 8), (6, 256, 4), (70, 192, 4)])
@pytest.mark.parametrize("policy", [0, 1])
@pytest.mark.parametrize("flags", [0, 0x4000])
def test_render_matches_oracle_sample_for_sample(pt, pairs, scene_id, width, spp, policy, flags):
    """Same Philox streams => the device follows the same paths as the oracle.  A path diverges only where a
    transcendental (CUDA vs glibc, <= 2 ulp) flips a branch or a texel, so all but a few pixels agree to fp32 accumulation.
    flags 0: the render ends in the tail megakernel (at these sizes it takes over after the first iterations);
    0x4000: wavefront iterations (per-class shade kernels) to the last path."""
    p = pairs(scene_id, width)
    img, st = p.dev.render(spp=spp, seed=21, nan_policy=policy, flags=flags)
    key = (scene_id, width, spp, policy)
    if key not in _ORACLE_RENDERS:
        _ORACLE_RENDERS[key] = p.ora.render(p.scene.camera, spp, seed=21, nan_policy=policy)
    ref, ost = _ORACLE_RENDERS[key]
    assert st.tail_paths == 0 if flags else (st.tail_paths > 0 or scene_id in (4, 5, 2))   # short-lived paths may all end before the threshold is checked
    assert st.paths == ost.paths == img.shape[0] * img.shape[1] * spp
    # the two mesh scenes at >= 65 536 paths in flight go through the two-pass traversal, like every benchmarked iteration
    assert (st.two_pass_iterations > 0) == (scene_id in (6, 70) and st.paths >= TWO_PASS_MIN)
    assert abs(int(st.segments) - int(ost.segments)) <= max(64, ost.segments // 2000)
    fin = np.isfinite(ref).all(axis=2) & np.isfinite(img).all(axis=2)
    assert np.array_equal(np.isfinite(ref).all(axis=2), np.isfinite(img).all(axis=2)) or policy == 0
    d = np.abs(img - ref)[fin].reshape(-1, 3).max(axis=1)
    scale = np.maximum(ref[fin].reshape(-1, 3).max(axis=1), 1.0)
    bad = (d > 1e-4 * scale).mean()
    assert bad < 0.02, f"{bad:.2%} of pixels differ from the oracle"
    assert H.rel_rmse(img, ref) < 0.05
    if policy == 1:
        assert abs(img.mean() - ref.mean()) < 2e-3 * max(ref.mean(), 1e-3) + 1e-6


def test_virtual_rank_split_equals_single_rank(pt, pairs):
    """spp split over G ranks (sample = g + k*G) sums to the single-rank accumulators."""
    p = pairs(3, 96)
    full, _ = p.dev.render(spp=8, seed=5, nan_policy=1)
    parts = [p.dev.render(spp=2, seed=5, nan_policy=1, sample_begin=g, sample_stride=4)[0] for g in range(4)]
    assert np.allclose(full, sum(parts) / 4.0, rtol=2e-5, atol=2e-6)
    small_pool, _ = p.dev.render(spp=8, seed=5, nan_policy=1, pool_paths=4096)      # pool size must not change the result
    assert np.allclose(full, small_pool, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("scene_id,width,spp", [(3, 96, 8), (6, 160, 8)])
def test_scheduling_knobs_do_not_change_the_image(pt, pairs, scene_id, width, spp):
    """The tail megakernel (one launch runs the last <= 64 Ki paths to their end; flag 0x4000 = the wavefront iterations
    instead), the forked shade streams and the octant grouping of survivors only reorder work: same paths, same segments,
    same iteration count, same image (up to fp32 atomic-add order)."""
    p = pairs(scene_id, width)
    base, st0 = p.dev.render(spp=spp, seed=9, nan_policy=1)
    assert 0 < st0.tail_paths <= 65536                                       # the default path did end in the megakernel
    assert p.dev.render(spp=spp, seed=9, nan_policy=1, flags=0x4000)[1].tail_paths == 0
    for flags in (0x4000, 0x2000, 0x8000, 0x100000, 0x400000, 0x400000 | 0x200000, 0x4000 | 0x2000 | 0x8000 | (5 << 16) | 0x100000):
        img, st = p.dev.render(spp=spp, seed=9, nan_policy=1, flags=flags)
        assert (st.paths, st.segments, st.iterations, st.nonfinite) == (st0.paths, st0.segments, st0.iterations, st0.nonfinite), hex(flags)
        assert np.allclose(img, base, rtol=2e-5, atol=2e-6), hex(flags)
        # flags 0x100000 / 0x200000 only mean something when the two-pass path is taken (scene 6: 115 200 paths in flight)
        # (the batched tail sizes its launches by the live count at the start of a batch, so the exact number may differ)
        assert (st.two_pass_iterations > 0) == (st0.two_pass_iterations > 0 and not flags & 0x100000), hex(flags)
    assert st0.iterations > 8 and st0.segments > st0.paths
    assert (st0.two_pass_iterations > 0) == (scene_id == 6)


def test_render_multi_equals_single_device(pt, pairs):
    """pt_render_multi (one process, one host thread per listed device, spp split, host reduce) gives pt_render's image.
    On a one-GPU box the shares run as three contexts on device 0; with more GPUs they spread over real devices."""
    p = pairs(6, 160)                                                           # 160 x 90 x 7 spp: the shares run the two-pass traversal too
    n_dev = pt.device_lib().pt_device_count()
    devices = [g % n_dev for g in range(3)]
    for kw in (dict(), dict(flags=pt.PT_RENDER_NEE), dict(sample_begin=5, sample_stride=2)):
        single, st1 = p.dev.render(spp=7, seed=4, nan_policy=1, **kw)
        multi, stm = pt.render_multi(p.scene, devices, spp=7, seed=4, nan_policy=1, **kw)
        assert (stm.paths, stm.segments, stm.nonfinite) == (st1.paths, st1.segments, st1.nonfinite), kw
        assert np.allclose(multi, single, rtol=2e-5, atol=2e-6), kw
        # the reduce runs on devices[0]: shares on other GPUs are read over peer access (NVLink) where the box allows it
        assert stm.p2p_shares <= len(set(devices)) - 1 and (n_dev == 1 or stm.p2p_shares == len([d for d in devices if d != devices[0]]))
    two, st2 = pt.render_multi(p.scene, [0] * 5, spp=2, seed=4, nan_policy=1)      # more devices than samples: two shares
    assert np.allclose(two, p.dev.render(spp=2, seed=4, nan_policy=1)[0], rtol=2e-5, atol=2e-6) and st2.paths == 2 * 160 * 90
    with pytest.raises(pt.PtError, match="device 99"):
        pt.render_multi(p.scene, [0, 99], spp=2)


def test_progressive_checkpoint_resume_and_noise_floor(pt, ctx, tmp_path):
    """batches + checkpoint/resume give the image of one pt_render call; the A/B half estimate of the noise
    predicts the relRMSE actually measured against a converged render."""
    P = __import__("importlib").import_module("pt_b200.progressive")
    scene = pt.Scene.build(3, width=96, spp=64, seed=1)
    dev = ctx.upload(scene)
    one_shot, st = dev.render(spp=64, seed=5, nan_policy=pt.PT_NAN_DROP)
    ck = str(tmp_path / "ck.npz")
    pr = P.ProgressiveRender(dev, seed=5, nan_policy=pt.PT_NAN_DROP)
    pr.advance(16).advance(16); pr.save(ck)
    pr2 = P.ProgressiveRender(dev, seed=5, nan_policy=pt.PT_NAN_DROP).load(ck)     # "after the crash"
    assert pr2.spp == 32
    pr2.advance(32)
    mean, se = pr2.result()
    assert pr2.spp == 64 and pr2.paths == st.paths
    assert np.allclose(mean, one_shot, rtol=2e-5, atol=2e-6)
    with pytest.raises(ValueError):
        P.ProgressiveRender(dev, seed=6, nan_policy=pt.PT_NAN_DROP).load(ck)         # another seed: not this render's checkpoint
    converged, _ = dev.render(spp=4096, seed=99, nan_policy=pt.PT_NAN_DROP)
    measured, predicted = H.rel_rmse(mean, converged), pr2.noise_floor()
    print(f"64 spp: relRMSE vs 4096-spp render {measured:.4f}, predicted from the A/B halves {predicted:.4f}")
    assert 0.6 * measured < predicted < 1.6 * measured
    pr2.advance(960)
    later = pr2.noise_floor()
    print(f"1024 spp: predicted {later:.4f}")
    assert pr2.spp == 1024 and later < 0.6 * predicted                              # 16x the samples (1/4 for Gaussian noise; fireflies decay slower)
    dev.close()


def test_edge_sizes_and_error_codes(pt, orc, ctx):
    """Empty / ragged / extreme arguments of the C ABI: zero samples, a 1x1 and an odd-sized image (the 8x4 tile order does
    not apply), a pool smaller than one warp's worth of pixels, max_depth 0 and 1, empty ray batches, and the error codes
    of a corrupted scene description (no exception crosses the ABI)."""
    import ctypes as C
    scene = pt.Scene.build(3, width=37, spp=4, seed=1)                     # 37 x 37: not a multiple of 8 x 4
    dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
    with pytest.raises(pt.PtError, match="sample_count"):                  # the mean of zero samples is undefined
        dev.render(spp=0)
    img, st = dev.render(spp=5, seed=3, nan_policy=pt.PT_NAN_DROP)
    ref, ost = ora.render(scene.camera, 5, seed=3, nan_policy=pt.PT_NAN_DROP)
    assert img.shape == (37, 37, 3) and st.paths == 37 * 37 * 5 and abs(int(st.segments) - int(ost.segments)) <= 64
    assert H.rel_rmse(img, ref) < 0.05
    tiny, st2 = dev.render(spp=5, seed=3, nan_policy=pt.PT_NAN_DROP, pool_paths=100)   # rounded up to one block of 128 paths
    assert np.allclose(tiny, img, rtol=2e-5, atol=2e-6) and st2.segments == st.segments and st2.iterations > 100
    for depth, bound in ((0, 0.0), (1, None)):
        cam = scene.camera_copy(max_depth=depth)
        im, s = dev.render(camera=cam, spp=2, seed=3)
        assert s.segments == (0 if depth == 0 else s.paths)
        if bound is not None:
            assert not im.any()                                             # no bounce: nothing is ever added (camera.rs:177)
    one = scene.camera_copy(image_width=1)
    im, s = dev.render(camera=one, spp=3, seed=3, nan_policy=pt.PT_NAN_DROP)
    assert im.shape == (1, 1, 3) and s.paths == 3
    assert len(dev.trace_closest(np.zeros(0, dtype=pt.RAY_DTYPE))) == 0
    assert len(dev.trace_any(np.zeros(0, dtype=pt.RAY_DTYPE), np.zeros(0))) == 0
    # corrupted descriptions: a copy of the 200-byte pt_scene_desc with one field broken
    raw = C.string_at(scene.desc, 200)
    def create(mutate):
        buf = C.create_string_buffer(raw, 200)
        mutate(buf)
        out = C.c_void_p()
        rc = ctx.lib.pt_scene_create(ctx.ptr, C.cast(buf, C.c_void_p), C.byref(out))
        msg = ctx.lib.pt_last_error().decode()
        if rc == 0:
            ctx.lib.pt_scene_destroy(out)
        return rc, msg
    rc, msg = create(lambda b: C.memmove(b, (C.c_uint32 * 1)(99), 4))                       # abi_version
    assert rc == pt.PT_ERR_INVALID and "abi_version" in msg
    rc, msg = create(lambda b: C.memmove(C.addressof(b) + 56, (C.c_uint64 * 1)(0), 8))      # textures = NULL with n_textures > 0
    assert rc == pt.PT_ERR_INVALID and "null array" in msg
    rc, msg = create(lambda b: C.memmove(C.addressof(b) + 176, (C.c_uint32 * 1)(0x7FFFFFF0), 4))  # objects_bvh_root out of range
    assert rc != 0 and msg
    rc, _ = create(lambda b: None)                                                           # the untouched copy is fine
    assert rc == 0
    dev.close(); ora.close()


def test_malformed_descriptions_are_rejected_before_use(pt, ctx):
    """Every index the uploader follows is range-checked first (no host or device out-of-bounds read through the C ABI):
    a cuboid whose six quads run past the quad table, a mesh whose triangle range does, a mesh BVH leaf that holds
    something other than the mesh's own triangles, and a bare triangle inside a top-level leaf."""
    import ctypes as C
    PTR = {"cuboids": 120, "meshes": 128, "leaf_refs": 152}

    def create(scene, table, dtype, count, mutate):
        arr = H.desc_array(scene, table, dtype, count)
        mutate(arr)
        buf = C.create_string_buffer(C.string_at(scene.desc, 200), 200)
        C.memmove(C.addressof(buf) + PTR[table], (C.c_uint64 * 1)(arr.ctypes.data), 8)
        out = C.c_void_p()
        rc = ctx.lib.pt_scene_create(ctx.ptr, C.cast(buf, C.c_void_p), C.byref(out))
        msg = ctx.lib.pt_last_error().decode()
        if rc == 0:
            ctx.lib.pt_scene_destroy(out)
        return rc, msg

    CUBOID_DT = np.dtype([("first_quad", "<u4"), ("material", "<u4"), ("a", "<f8", 3), ("b", "<f8", 3)])
    s3 = pt.Scene.build(3, width=32, spp=1, seed=1)
    hd = H.desc_header(s3)
    assert hd["n_cuboids"] == 2
    rc, msg = create(s3, "cuboids", CUBOID_DT, 2, lambda a: a["first_quad"].__setitem__(1, hd["n_quads"] - 3))
    assert rc == pt.PT_ERR_INVALID and "cuboid" in msg
    rc, msg = create(s3, "cuboids", CUBOID_DT, 2, lambda a: None)
    assert rc == 0
    s6 = pt.Scene.build(6, width=32, spp=1, seed=1)
    hd = H.desc_header(s6)
    rc, msg = create(s6, "meshes", H.MESH_DT, hd["n_meshes"], lambda a: a["first"].__setitem__(0, hd["n_triangles"] - 1))
    assert rc == pt.PT_ERR_INVALID and "mesh" in msg
    nodes = H.desc_array(s6, "nodes", H.NODE_DT, hd["n_nodes"])
    meshes = H.desc_array(s6, "meshes", H.MESH_DT, hd["n_meshes"])

    def first_leaf_ref(root):
        i = root
        while nodes[i]["left"] != 0xFFFFFFFF:
            i = int(nodes[i]["left"])
        return int(nodes[i]["first"])
    k = first_leaf_ref(int(meshes[1]["root"]))

    def put(a, i, kind, index):
        a["kind"][i] = kind; a["index"][i] = index
    rc, msg = create(s6, "leaf_refs", H.REF_DT, hd["n_leaf_refs"], lambda a: put(a, k, pt.PRIM_SPHERE, 0))
    assert rc == pt.PT_ERR_INVALID and "mesh BVH leaf" in msg                                  # not a triangle
    rc, msg = create(s6, "leaf_refs", H.REF_DT, hd["n_leaf_refs"], lambda a: put(a, k, pt.PRIM_TRIANGLE, int(meshes[0]["first"])))
    assert rc == pt.PT_ERR_INVALID and "mesh BVH leaf" in msg                                  # another mesh's triangle
    k_top = first_leaf_ref(H.desc_roots(s6)[0])
    rc, msg = create(s6, "leaf_refs", H.REF_DT, hd["n_leaf_refs"], lambda a: put(a, k_top, pt.PRIM_TRIANGLE, 0))
    assert rc == pt.PT_ERR_INVALID and "top-level" in msg
    rc, msg = create(s6, "leaf_refs", H.REF_DT, hd["n_leaf_refs"], lambda a: None)
    assert rc == 0


def test_device_sah_sweep_builds_the_same_trees(pt, ctx):
    """with a build context the O(n^2) SAH sweep of bvh.rs:54-120 runs on the device (pt_sah_sweep);
    costs are bit-identical to the host's, so the flattened BVH (nodes + leaf order) is the same byte for byte."""
    import ctypes as C
    import time

    def tree_bytes(scene):
        raw = C.string_at(scene.desc, 200)
        n_nodes, n_refs = np.frombuffer(raw, np.uint32, 2, 40)
        nodes_ptr, refs_ptr = np.frombuffer(raw, np.uint64, 2, 144)
        return C.string_at(int(nodes_ptr), int(n_nodes) * 64), C.string_at(int(refs_ptr), int(n_refs) * 8), int(n_nodes)

    for sid in (6, 70, 1):
        t0 = time.time(); host = pt.Scene.build(sid, width=64, spp=1, seed=1); t_host = time.time() - t0
        pt.set_build_context(ctx)
        try:
            t0 = time.time(); devb = pt.Scene.build(sid, width=64, spp=1, seed=1); t_dev = time.time() - t0
        finally:
            pt.set_build_context(None)
        a, b = tree_bytes(host), tree_bytes(devb)
        print(f"scene {sid}: {a[2]} BVH nodes, scene build {t_host:.2f} s on the host, {t_dev:.2f} s with the device sweep")
        assert a[0] == b[0] and a[1] == b[1]


def test_tonemap_matches_reference_formula(pt, orc, ctx):
    import torch
    x = np.array([0.0, 1.0, 4.0, 0.25, -1.0, np.nan, np.inf, 1e-6, 0.5, 0.9981], dtype=np.float32)
    x = np.resize(x, 30)
    d = torch.tensor(x, device="cuda")
    out = np.zeros(30, np.uint8)
    ctx._check(ctx.lib.pt_tonemap_rgb8(ctx.ptr, d.data_ptr(), 1.0, 10, out.ctypes.data))
    assert np.array_equal(out, orc.tonemap_rgb8(x.astype(np.float64)))


# ---------------------------------------------------------------- against the reference's own demo renders, full size
@pytest.mark.parametrize("scene_id,spp,tol", [(4, 64, 0.02), (2, 64, 0.05), (5, 64, 0.04), (6, 64, 0.09)])
def test_full_size_render_matches_reference_demo(pt, ctx, scene_id, spp, tol):
    """1920x1080 like the reference's `-q` renders; relRMSE on 8x8 cells vs demo/*.png (4000 spp, 8-bit).
    Measured: see tests/golden/README.md.  The floor is the demo's own quantisation + our noise at `spp`."""
    scene = pt.Scene.build(scene_id, width=1920, spp=spp, seed=1)
    dev = ctx.upload(scene)
    img, st = dev.render(spp=spp, seed=31, nan_policy=pt.PT_NAN_DROP)
    err, frac = H.compare_with_demo(img, scene_id)
    print(f"scene {scene_id}: relRMSE vs reference demo {err:.4f} ({frac:.0%} cells) {st.segments / st.device_ms / 1e3:.0f} Mrays/s")
    assert img.shape == (1080, 1920, 3) and frac > 0.6
    assert err < tol
    # size-independent properties at full size: every sample accounted for, no non-finite pixel, deterministic paths
    assert st.paths == 1920 * 1080 * spp and np.isfinite(img).all()
    dev.close()


# ---------------------------------------------------------------- the C++ host mirror end to end (reference: `cargo run -r -- -s 3`)
def test_cli_renders_cornell_box(pt, tmp_path):
    import subprocess
    from PIL import Image
    exe = os.path.join(os.path.dirname(pt.device_lib_path()), "..", "bin", "ptb200")
    out = subprocess.run([exe, "-s", "3", "--width", "96", "--spp", "16", "--seed", "7", "--out", str(tmp_path), "--assets", pt.ASSETS_DIR],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "rendering production" in out.stdout                       # camera.rs:101
    img = np.asarray(Image.open(tmp_path / "cornell.png"))
    assert img.shape == (96, 96, 3) and img.std() > 10
    # same scene, seed and spp through the Python path: identical bytes up to fp32 accumulation order
    scene = pt.Scene.build(3, width=96, spp=16, seed=7)
    ctx = pt.Context(0); dev = ctx.upload(scene)
    mean, _ = dev.render(spp=16, seed=7, nan_policy=pt.PT_NAN_REFERENCE)
    ref = pt.tonemap_rgb8(mean)
    assert (np.abs(img.astype(int) - ref.astype(int)) > 1).mean() < 0.01
    dev.close(); ctx.close()


# ---------------------------------------------------------------- the headline workload at full size
def test_scene6_fhd_4000spp_matches_reference_demo(pt, ctx):
    """BASELINE.json's target: scene 6 at 1920x1080 x 4000 spp (8.29 G paths) in one pt_render call, against the
    reference's own demo/scene6.png (also 4000 spp).  Size-independent properties: every path accounted for (64-bit
    counters), no non-finite pixel, and the virtual split of the sample range reproduces the same sums."""
    scene = pt.Scene.build(6, width=1920, spp=4000, seed=1)
    dev = ctx.upload(scene)
    img, st = dev.render(spp=4000, seed=77, nan_policy=pt.PT_NAN_DROP)
    err, frac = H.compare_with_demo(img, 6)
    print(f"scene 6 FHD x 4000 spp: {st.device_ms / 1e3:.2f} s, {st.segments / st.device_ms / 1e3:.0f} Mrays/s, relRMSE vs demo {err:.4f} ({frac:.0%} cells)")
    assert st.paths == 1920 * 1080 * 4000 and st.segments > st.paths and np.isfinite(img).all()
    assert err < 0.03  # measured 0.0132: the floor is the demo's 8-bit quantisation and its own 4000-spp noise
    dev.close()


def _holey_sheets_world(pt, n_sheets=5, g=9, seed=11):
    """n_sheets wavy triangle sheets with holes, stacked along z: odd ones behind a rotated Instance, even ones directly in
    World.objects.  Every camera ray enters all their boxes, so it queues three mesh visits and walks the others inline."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-1.0, 1.0, g + 1)
    w = pt.World()
    keep_alive = []
    for s in range(n_sheets):
        X, Y = np.meshgrid(xs, xs, indexing="ij")
        Z = 0.12 * np.sin(3.0 * X + s) * np.cos(2.0 * Y - s) + rng.normal(scale=0.01, size=X.shape)
        z0 = -2.0 + s
        direct = s % 2 == 0
        pos = np.stack([X, Y, Z + (z0 if direct else 0.0)], axis=-1).reshape(-1, 3).astype(np.float32)
        vid = np.arange((g + 1) * (g + 1)).reshape(g + 1, g + 1)
        tris = np.concatenate([np.stack([vid[:-1, :-1], vid[1:, :-1], vid[1:, 1:]], axis=-1).reshape(-1, 3),
                               np.stack([vid[:-1, :-1], vid[1:, 1:], vid[:-1, 1:]], axis=-1).reshape(-1, 3)])
        tris = tris[rng.uniform(size=len(tris)) > 0.4].astype(np.uint32)            # holes: deeper sheets stay visible
        assert len(tris) >= 64                                                       # large enough for the 4-wide flavour
        mat = pt.DiffuseBRDF(tuple(0.25 + 0.7 * rng.uniform(size=3))) if s != 2 else pt.MetalBRDF((0.9, 0.8, 0.6), 0.2)
        mesh = pt.TriangleMesh.from_arrays(1.6, pos, tris, mat)
        obj = mesh if direct else pt.Instance(mesh, (0.1, 0.2, 1.0), 0.3 * s, (0.05 * s, -0.03 * s, z0))
        keep_alive += [mat, mesh, obj]
        w.add_object(obj)
    w.add_object(pt.Sphere.new_still(0.5, (0.3, 0.2, -3.2), pt.DiffuseBRDF((0.8, 0.3, 0.3))))
    w.build_bvh()
    cam = pt.make_camera(96, 1.0, 4, 12, vfov=34.0, look_from=(0.2, 0.1, 7.0), look_at=(0, 0, 0), env_color=(0.7, 0.8, 1.0))
    return w, cam, keep_alive


def test_two_pass_traversal_beyond_the_reference_scenes(pt, orc, ctx):
    """The multi-pass traversals where no reference scene takes them: meshes that sit in World.objects directly, five mesh
    visits per ray (five k_mesh_enter / k_mesh_walk rounds; with the BVH kernels, flag 0x400000, more than the three visit
    queues of k_trace<DEFER> hold, so the rest is walked inline).  Sample for sample against the oracle, and against the fused
    kernel (flag 0x100000) and the grid-stride rounds (0x200000)."""
    w, cam, keep = _holey_sheets_world(pt)
    scene = pt.Scene.from_world(w, cam)
    assert H.desc_header(scene)["n_meshes"] == 5 and H.desc_header(scene)["n_instances"] == 2
    dev, ora = ctx.upload(scene), orc.OracleScene(scene.desc, pt)
    spp = 12                                                                          # 96 x 96 x 12 = 110 592 paths: above the 64 Ki floor of the two-pass path
    img, st = dev.render(spp=spp, seed=5, nan_policy=1)
    ref, ost = ora.render(cam, spp, seed=5, nan_policy=1)
    assert st.paths == ost.paths == 96 * 96 * spp and st.two_pass_iterations > 0
    assert abs(int(st.segments) - int(ost.segments)) <= max(64, ost.segments // 2000)
    d = np.abs(img - ref).reshape(-1, 3).max(axis=1)
    assert (d > 1e-4 * np.maximum(ref.reshape(-1, 3).max(axis=1), 1.0)).mean() < 0.02
    assert H.rel_rmse(img, ref) < 0.05
    # the traversal stage alone, ID for ID, on 90 000 camera rays that enter up to five meshes each: later rounds see hits that
    # earlier rounds replaced (shade class looked up again), and every ray must still join exactly one shade queue
    big = scene.camera_copy(image_width=300)
    rows, cols = np.divmod(np.arange(300 * 300, dtype=np.uint32), 300)
    rays = orc.camera_rays(big, 3, rows, cols, np.zeros_like(rows), pt)
    want = ora.trace_closest(rays)
    assert (want["prim_kind"][want["hit"] == 1] == 2).mean() > 0.5
    for flags in (0, 0x400000, 0x400000 | 0x200000, 0x100000):
        got, stq = dev.trace_closest_wavefront(rays, flags=flags)
        assert stq.queue_errors == 0 and (stq.two_pass_iterations > 0) == (flags != 0x100000), hex(flags)
        assert_hits_equal(pt, got, want, f"holey sheets, flags {flags:#x}")
    for flags in (0x100000, 0x400000, 0x400000 | 0x200000):
        other, st2 = dev.render(spp=spp, seed=5, nan_policy=1, flags=flags)
        assert (st2.paths, st2.segments) == (st.paths, st.segments), hex(flags)
        assert (st2.two_pass_iterations > 0) == (flags != 0x100000)
        assert np.allclose(other, img, rtol=2e-5, atol=2e-6), hex(flags)
    dev.close(); ora.close()
