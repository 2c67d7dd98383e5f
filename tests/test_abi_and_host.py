"""CPU tests: the C-ABI library loads and exports exactly what include/pt_b200.h declares (no compute calls),
ABI struct layouts match the Python mirrors, and the C++ host mirror's derivations (quad fields, cuboid sides,
instance matrices, bounding boxes, SAH BVH trees) equal the oracle's independent restatement bit for bit."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_match_header(pt):
    hdr = open(os.path.join(ROOT, "include", "pt_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(pt.ABI_SYMBOLS)
    lib = C.CDLL(pt.device_lib_path())
    for s in declared:
        assert getattr(lib, s) is not None
    out = subprocess.check_output(["nm", "-D", "--defined-only", pt.device_lib_path()]).decode()
    exported = sorted(set(re.findall(r" T (pt_[a-z0-9_]+)", out)))
    assert exported == declared, "the .so must export the header's entry points and nothing else named pt_*"


def test_abi_struct_sizes(pt, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "pt_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(pt_ray),sizeof(pt_hit),sizeof(pt_bsdf_query),sizeof(pt_bsdf_result),sizeof(pt_bsdf_sample_result),sizeof(pt_camera),"
                   "sizeof(pt_render_params),sizeof(pt_stats),sizeof(pt_bvh_node),sizeof(pt_quad),sizeof(pt_instance),sizeof(pt_mesh));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [pt.RAY_DTYPE.itemsize, pt.HIT_DTYPE.itemsize, pt.BSDF_QUERY_DTYPE.itemsize, pt.BSDF_RESULT_DTYPE.itemsize,
            pt.BSDF_SAMPLE_DTYPE.itemsize, C.sizeof(pt.CameraABI), C.sizeof(pt.RenderParams), C.sizeof(pt.Stats),
            H.NODE_DT.itemsize, H.QUAD_DT.itemsize, H.INST_DT.itemsize, H.MESH_DT.itemsize]
    assert got == want


def test_rust_sys_crate_matches_header(pt, tmp_path):
    """ffi/pt-b200-sys (the binding a maintainer of the reference adds; not compiled here — no Rust toolchain): it declares
    exactly the entry points of include/pt_b200.h, and the struct sizes it asserts at compile time are the C compiler's."""
    rs = open(os.path.join(ROOT, "ffi", "pt-b200-sys", "src", "lib.rs")).read()
    declared = sorted(set(re.findall(r"pub fn (pt_[a-z0-9_]+)\(", rs)))
    assert declared == sorted(pt.ABI_SYMBOLS)
    sizes = dict(re.findall(r"size_of::<(pt_[a-z0-9_]+)>\(\) == (\d+)", rs))
    assert len(sizes) >= 15
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "pt_b200.h"\nint main(){' + "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in sizes) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    assert got == sizes


def test_no_device_fails_loudly(pt):
    """Without a GPU the product must refuse, not fall back (the CPU box has no CUDA device)."""
    lib = pt.device_lib()
    if lib.pt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pt.PtError, match="no CUDA device"):
        pt.Context(0)
    with pytest.raises(pt.PtError, match="device 0: no CUDA device"):     # the single-process multi-GPU entry: same refusal,
        pt.render_multi(pt.Scene.build(3, 32, 1, 1), [0, 0], spp=2)        # carried over from the worker thread


def test_render_multi_rejects_bad_arguments(pt):
    """pt_render_multi validates before it touches a device (so this runs on the CPU box too)."""
    scene = pt.Scene.build(3, 32, 1, 1)
    with pytest.raises(pt.PtError, match="bad argument"):
        pt.render_multi(scene, [], spp=2)
    with pytest.raises(pt.PtError, match="sample_count is zero"):
        pt.render_multi(scene, [0], spp=0)


def test_product_never_touches_oracle():
    pkg = os.path.join(ROOT, "thu-acg-f2024-path-tracer_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                for token in ("oracle_py", "liboracle", "orc_", "oracle_capi", "oracle_math", "oracle_scene", "oracle_bsdf", "oracle/_build", "load_oracle"):
                    assert token not in txt, f"{f} references the oracle ({token}): the product path must not route through it"


@pytest.mark.parametrize("scene_id", [1, 3, 4, 7, 6])
def test_host_derivations_match_oracle(pt, orc, scene_id):
    scene = pt.Scene.build(scene_id, width=64, spp=1, seed=3)
    hd = H.desc_header(scene)
    ora = orc.OracleScene(scene.desc, pt)
    nodes = H.desc_array(scene, "nodes", H.NODE_DT, hd["n_nodes"])
    refs = H.desc_array(scene, "leaf_refs", H.REF_DT, hd["n_leaf_refs"])
    objects = H.desc_array(scene, "objects", H.REF_DT, hd["n_objects"])
    lights = H.desc_array(scene, "lights", H.REF_DT, hd["n_lights"])
    quads = H.desc_array(scene, "quads", H.QUAD_DT, hd["n_quads"])
    insts = H.desc_array(scene, "instances", H.INST_DT, hd["n_instances"])
    meshes = H.desc_array(scene, "meshes", H.MESH_DT, hd["n_meshes"])
    oroot, lroot = H.desc_roots(scene)
    # ---- trees: same topology, same leaf contents in the same order, bit-identical boxes
    for which, lst, root in ((0, objects, oroot), (1, lights, lroot)):
        if len(lst) == 0:
            continue
        pos = {(int(r["kind"]), int(r["index"])): i for i, r in enumerate(lst)}
        hs, hb = H.host_bvh_signature(nodes, refs, root, lambda k, i: pos[(k, i)])
        os_, ob = ora.bvh_signature(which, with_boxes=True)
        assert np.array_equal(hs, os_)
        assert np.array_equal(hb, ob[: hb.size])
        for i in range(len(lst)):  # per-item boxes are implied by the leaf boxes; check derived fields
            k, idx = int(lst[i]["kind"]), int(lst[i]["index"])
            if k == pt.PRIM_QUAD:
                qd = ora.quad_derived(which, i)
                assert np.array_equal(qd, np.concatenate([quads[idx]["w"], quads[idx]["normal"], [quads[idx]["d"]]]))
            if k == pt.OBJ_INSTANCE:
                m = ora.instance_matrices(which, i)
                assert np.array_equal(m[0], insts[idx]["transform"])
                assert np.array_equal(m[1], insts[idx]["inverse"])
                assert np.array_equal(m[2], insts[idx]["normal_matrix"])
    for mi, m in enumerate(meshes):
        hs, hb = H.host_bvh_signature(nodes, refs, int(m["root"]), lambda k, i: i - int(m["first"]))
        os_, ob = ora.bvh_signature(2 + mi, with_boxes=True)
        assert np.array_equal(hs, os_), f"mesh {mi} tree differs"
        assert np.array_equal(hb, ob[: hb.size])
    ora.close()


def test_scene_counts(pt):
    """Scene contents as SURVEY §8(d) lists them (reference src/main.rs)."""
    s3 = H.desc_header(pt.Scene.build(3, 64, 1, 1))
    assert (s3["n_objects"], s3["n_lights"], s3["n_cuboids"], s3["n_instances"], s3["n_quads"], s3["n_spheres"]) == (8, 1, 2, 2, 18, 1)
    s6 = H.desc_header(pt.Scene.build(6, 64, 1, 1))
    assert s6["n_triangles"] == 4968 + 5856 + 5804 and s6["n_meshes"] == 3 and s6["n_lights"] == 0
    s1 = H.desc_header(pt.Scene.build(1, 64, 1, 1))
    assert 400 < s1["n_spheres"] <= 488
    assert pt.Scene.build(1, 600, 1, 1).image_height() == 337 and pt.Scene.build(3, 600, 1, 1).image_height() == 600


def test_png_roundtrip(pt, tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    p = str(tmp_path / "x.png")
    pt.write_png(p, img)
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(p)), img)          # our encoder -> PIL
    back = pt.Image(path=p)                                         # our decoder (used for the baked textures)
    assert back.ptr


def test_jpeg_decoder_matches_libjpeg(pt, tmp_path):
    """host/jpeg.cpp (no Python bake for the reference's .jpg assets): baseline and progressive Huffman
    JPEG, 4:4:4 / 4:2:2 / 4:2:0 / grey, odd sizes, optimised tables, restart intervals — byte-identical to PIL (libjpeg:
    islow IDCT, fancy upsampling), and the reference's 7616x3808 progressive assets/envmap.jpg of scene 5."""
    import io
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    rng = np.random.default_rng(1)
    img = (rng.uniform(size=(45, 67, 3)) * 255).astype(np.uint8)
    img[10:30, 5:50] = (200, 30, 90)
    cases = [dict(subsampling=0), dict(subsampling=0, progressive=True), dict(subsampling=1), dict(subsampling=2),
             dict(subsampling=2, progressive=True), dict(subsampling=0, quality=30, optimize=True), dict(grey=True),
             dict(subsampling=2, restart_marker_blocks=3), dict(subsampling=0, progressive=True, restart_marker_rows=1)]
    for kw in cases:
        kw = dict(kw)
        src = Image.fromarray(img[:, :, 0] if kw.pop("grey", False) else img)
        buf = io.BytesIO()
        try:
            src.save(buf, "JPEG", quality=kw.pop("quality", 85), **kw)
        except TypeError:
            continue                                                    # an older Pillow without restart-marker options
        p = str(tmp_path / "t.jpg")
        open(p, "wb").write(buf.getvalue())
        ours = pt.Image(path=p).pixels()
        theirs = np.asarray(Image.open(p).convert("RGB"))
        assert np.array_equal(ours, theirs), kw
    env = os.path.join(pt.ASSETS_DIR, "envmap.jpg")
    assert np.array_equal(pt.Image(path=env).pixels(), pt.load_rgb8(env))
    with pytest.raises(pt.PtError):
        open(str(tmp_path / "bad.jpg"), "wb").write(b"\xff\xd8\xff\xc3\x00\x04")   # lossless SOF3: unsupported, must say so
        pt.Image(path=str(tmp_path / "bad.jpg"))


def test_reference_asset_files_load_natively(pt, tmp_path):
    """the host reads the reference's own asset files.  host/hdr.cpp (Radiance RGBE with the reference's
    clamp-and-round `to_rgb8`, src/texture.rs:62-69, Q22) must give the bytes of the shipped bake, which was made by an
    independent decoder (cv2, tools/bake_assets.py); synthetic files cover flat pixels, both RLE flavours and the
    rounding.  When the reference's assets directory is present (this container; not the GPU box), every scene built
    from it (.obj, .jpg, .hdr, bricks/*.png) must flatten to the same bytes as the scene built from the bakes."""
    import struct
    w, h = 16, 3
    rng = np.random.default_rng(3)
    rgbe = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    rgbe[..., 3] = rng.integers(120, 140, size=(h, w))
    rgbe[0, 0] = (200, 1, 0, 135)                                        # 0.5 -> 127.5 -> 128 (round half away from zero)
    rgbe[0, 1] = (255, 255, 255, 0)                                      # e == 0 -> black
    rgbe[1, 4:12] = rgbe[1, 3]                                           # a run for the RLE encoders below
    want = np.where(rgbe[..., 3:4] == 0, 0.0, rgbe[..., :3].astype(np.float32) * np.exp2(rgbe[..., 3:4].astype(np.float32) - 136))
    want = np.floor(np.clip(want, 0, 1).astype(np.float32) * np.float32(255) + np.float32(0.5)).astype(np.uint8)
    head = b"#?RADIANCE\n# made by a test\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=1.0\n\n-Y %d +X %d\n" % (h, w)
    flat = head + rgbe.tobytes()
    rle = head
    for y in range(h):
        rle += bytes([2, 2, w >> 8, w & 255])
        for ch in range(4):
            row, x = rgbe[y, :, ch], 0
            while x < w:
                n = 1
                while x + n < w and row[x + n] == row[x] and n < 127:
                    n += 1
                if n >= 3:
                    rle += bytes([128 + n, row[x]]); x += n
                else:
                    m = min(5, w - x)
                    rle += bytes([m]) + row[x:x + m].tobytes(); x += m
    old = head
    for y in range(h):
        x = 0
        while x < w:
            n = 1
            while x + n < w and (rgbe[y, x + n] == rgbe[y, x]).all():
                n += 1
            old += rgbe[y, x].tobytes() + (bytes([1, 1, 1, n - 1]) if n > 1 else b"")
            x += n
    for name, blob in [("flat", flat), ("rle", rle), ("old", old)]:
        p = str(tmp_path / (name + ".hdr"))
        open(p, "wb").write(blob)
        assert np.array_equal(pt.Image(path=p).pixels(), want), name
    for bad in [b"P6\n", head[:30], head + b"\x02\x02\x00\x11", b"#?RADIANCE\nFORMAT=32-bit_rle_xyze\n\n-Y 1 +X 1\n\0\0\0\0",
                b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n+Y 1 +X 1\n\0\0\0\0"]:
        p = str(tmp_path / "bad.hdr")
        open(p, "wb").write(bad)
        with pytest.raises(pt.PtError):
            pt.Image(path=p)
    shipped = os.path.join(pt.ASSETS_DIR, "grace_probe_latlong.hdr")
    assert np.array_equal(pt.Image(path=shipped).pixels(), pt.Image(path=os.path.join(pt.ASSETS_DIR, "grace_probe_latlong.png")).pixels())
    assert np.array_equal(pt.Image(path=os.path.join(pt.ASSETS_DIR, "earthmap.jpg")).pixels(),
                          pt.Image(path=os.path.join(pt.ASSETS_DIR, "earthmap.png")).pixels())
    ref_assets = "/root/reference/assets"
    if not os.path.isdir(ref_assets):
        return
    baked = str(tmp_path / "baked")                                      # only the bakes: forces the fallback names
    os.makedirs(baked)
    for f in os.listdir(pt.ASSETS_DIR):
        if f.endswith((".mesh", ".png")) or f == "envmap.jpg":
            os.symlink(os.path.join(pt.ASSETS_DIR, f), os.path.join(baked, f))
    for scene_id in (2, 4, 6, 7, 70):
        a, b = pt.Scene.build(scene_id, 64, 1, 1, assets_dir=ref_assets), pt.Scene.build(scene_id, 64, 1, 1, assets_dir=baked)
        ha, hb = H.desc_header(a), H.desc_header(b)
        assert ha == hb, scene_id
        for name, dt, n in [("nodes", H.NODE_DT, ha["n_nodes"]), ("leaf_refs", H.REF_DT, ha["n_leaf_refs"]),
                            ("triangles", np.dtype([("v", "<f8", 9)]), ha["n_triangles"]), ("quads", H.QUAD_DT, ha["n_quads"])]:
            assert H.desc_array(a, name, dt, n).tobytes() == H.desc_array(b, name, dt, n).tobytes(), (scene_id, name)
        ia, ib = H.desc_images(a), H.desc_images(b)
        assert len(ia) == len(ib) == ha["n_images"] and all(np.array_equal(x, y) for x, y in zip(ia, ib)), scene_id


def test_unsupported_constructs_are_rejected(pt):
    mat = pt.DiffuseBRDF((0.5, 0.5, 0.5))
    inner = pt.Instance(pt.Sphere.new_still(1.0, (0, 0, 0), mat), (0, 1, 0), 0.3, (1, 0, 0))
    with pytest.raises(pt.PtError, match="nested Instance"):
        pt.Instance(inner, (0, 1, 0), 0.1, (0, 0, 0))
