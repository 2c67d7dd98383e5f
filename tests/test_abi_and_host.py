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


def test_no_device_fails_loudly(pt):
    """Without a GPU the product must refuse, not fall back (the CPU box has no CUDA device)."""
    lib = pt.device_lib()
    if lib.pt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pt.PtError, match="no CUDA device"):
        pt.Context(0)


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


def test_unsupported_constructs_are_rejected(pt):
    mat = pt.DiffuseBRDF((0.5, 0.5, 0.5))
    inner = pt.Instance(pt.Sphere.new_still(1.0, (0, 0, 0), mat), (0, 1, 0), 0.3, (1, 0, 0))
    with pytest.raises(pt.PtError, match="nested Instance"):
        pt.Instance(inner, (0, 1, 0), 0.1, (0, 0, 0))
