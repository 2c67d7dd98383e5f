"""Shared test helpers: the relRMSE metric of SURVEY §8(d), golden-demo comparison, random query builders."""
import ctypes as C
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLIP = 0.999 ** 2  # the PNG clips sqrt(x) at 0.999 (reference camera.rs:111-113)


def rel_rmse(a, b):
    """sqrt(mean((a-b)^2 / (b^2 + 1e-2))) on linear radiance clipped like the PNG; b is the reference."""
    a = np.clip(np.nan_to_num(np.asarray(a, np.float64)), 0.0, CLIP)
    b = np.clip(np.nan_to_num(np.asarray(b, np.float64)), 0.0, CLIP)
    return float(np.sqrt(np.mean((a - b) ** 2 / (b ** 2 + 1e-2))))


def load_golden(scene):
    g = np.load(os.path.join(GOLDEN_DIR, f"demo_scene{scene}.npz"))
    return g["mean"].astype(np.float64), g["clipped"].astype(np.float64), int(g["width"]), int(g["height"]), int(g["factor"])


def compare_with_demo(mean_full, scene):
    """mean_full: our [H,W,3] linear mean radiance at the demo's resolution (1920x1080).
    Returns relRMSE over the 8x-downsampled cells that contain no clipped reference pixel."""
    ref, clipped, w, h, f = load_golden(scene)
    assert mean_full.shape[:2] == (h, w), (mean_full.shape, (h, w))
    ours = np.clip(np.nan_to_num(np.asarray(mean_full, np.float64), nan=0.0, posinf=1.0), 0.0, CLIP)
    cells = ours.reshape(h // f, f, w // f, f, 3).mean(axis=(1, 3))
    ok = clipped == 0
    d = (cells[ok] - ref[ok]) ** 2 / (ref[ok] ** 2 + 1e-2)
    return float(np.sqrt(d.mean())), float(ok.mean())


def desc_header(scene):
    """Counts at the head of pt_scene_desc (include/pt_b200.h)."""
    names = "abi n_textures n_images n_materials n_spheres n_quads n_triangles n_cuboids n_meshes n_instances n_nodes n_leaf_refs n_objects n_lights".split()
    u32 = C.cast(scene.desc, C.POINTER(C.c_uint32))
    return {n: int(u32[i]) for i, n in enumerate(names)}


_PTRS = "textures images materials spheres quads triangles tri_normals tri_uvs cuboids meshes instances nodes leaf_refs objects lights".split()
NODE_DT = np.dtype([("bmin", "<f8", 3), ("bmax", "<f8", 3), ("left", "<u4"), ("right", "<u4"), ("first", "<u4"), ("n", "<u4")])
REF_DT = np.dtype([("kind", "<u4"), ("index", "<u4")])
MESH_DT = np.dtype([("first", "<u4"), ("n", "<u4"), ("material", "<u4"), ("root", "<u4"), ("has_normals", "<u4"), ("has_uvs", "<u4")])
QUAD_DT = np.dtype([("q", "<f8", 3), ("u", "<f8", 3), ("v", "<f8", 3), ("w", "<f8", 3), ("normal", "<f8", 3), ("d", "<f8"), ("material", "<u4"), ("_pad", "<u4")])
INST_DT = np.dtype([("child_kind", "<u4"), ("child_index", "<u4"), ("axis", "<f8", 3), ("angle", "<f8"), ("translation", "<f8", 3),
                    ("transform", "<f8", 16), ("inverse", "<f8", 16), ("normal_matrix", "<f8", 16)])


def desc_array(scene, name, dtype, count):
    ptrs = C.cast(C.c_void_p(scene.desc + 56), C.POINTER(C.c_void_p))
    p = ptrs[_PTRS.index(name)]
    if not p or count == 0:
        return np.zeros(0, dtype=dtype)
    raw = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(count * dtype.itemsize,))
    return raw.view(dtype).copy()


def desc_images(scene):
    """The decoded RGB8 images of a flattened scene (pt_image: pointer, width, height)."""
    n = desc_header(scene)["n_images"]
    out = []
    for im in desc_array(scene, "images", np.dtype([("rgb", "<u8"), ("w", "<u4"), ("h", "<u4")]), n):
        raw = np.ctypeslib.as_array(C.cast(C.c_void_p(int(im["rgb"])), C.POINTER(C.c_uint8)), shape=(int(im["h"]), int(im["w"]), 3))
        out.append(raw.copy())
    return out


def desc_roots(scene):
    """(objects_bvh_root, lights_bvh_root) — the two uint32 after the 15 pointers."""
    u32 = C.cast(C.c_void_p(scene.desc + 56 + 15 * 8), C.POINTER(C.c_uint32))
    return int(u32[0]), int(u32[1])


def host_bvh_signature(nodes, refs, root, item_index_of):
    """DFS pre-order signature of a host-built tree: internal -> -1; leaf -> n, then list positions."""
    sig, boxes, stack = [], [], [root]
    while stack:
        i = stack.pop()
        n = nodes[i]
        boxes.extend(n["bmin"]); boxes.extend(n["bmax"])
        if n["left"] == 0xFFFFFFFF:
            sig.append(int(n["n"]))
            for r in refs[n["first"]: n["first"] + n["n"]]:
                sig.append(item_index_of(int(r["kind"]), int(r["index"])))
        else:
            sig.append(-1)
            stack.append(int(n["right"])); stack.append(int(n["left"]))
    return np.array(sig, dtype=np.int64), np.array(boxes)


def rand_dirs(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def random_bsdf_queries(pt, rng, n, tilt=0.2):
    q = np.zeros(n, dtype=pt.BSDF_QUERY_DTYPE)
    gn = rand_dirs(rng, n)
    sn = gn + tilt * rand_dirs(rng, n)
    sn /= np.linalg.norm(sn, axis=1, keepdims=True)
    q["geometric_normal"], q["shading_normal"] = gn, sn
    q["view_dir"], q["light_dir"] = rand_dirs(rng, n), rand_dirs(rng, n)
    q["point"] = rng.uniform(-5, 5, size=(n, 3))
    q["u"], q["v"] = rng.uniform(size=n), rng.uniform(size=n)
    q["front_face"] = rng.integers(0, 2, size=n)
    return q


def max_rel_err(x, y):
    """Max relative error; identical non-finite values count as equal."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    fin = np.isfinite(x) & np.isfinite(y)
    same_nf = (~fin) & ((np.isnan(x) & np.isnan(y)) | (x == y))
    d = np.abs(np.where(fin, x - y, 0.0))
    s = np.maximum(np.abs(np.where(fin, y, 1.0)), 1e-300)
    r = np.where(fin, d / s, np.where(same_nf, 0.0, np.inf))
    r = np.where(fin & (d < 1e-290), 0.0, r)
    return float(r.max()) if r.size else 0.0


def ulp_diff(a, b):
    return np.abs(np.asarray(a, np.float64).view(np.int64) - np.asarray(b, np.float64).view(np.int64))
