"""pt_b200 — thin ctypes layer over the two in-tree native libraries.

  lib/libptb200.so       CUDA kernels + the C ABI of include/pt_b200.h (the product)
  lib/libptb200_host.so  C++ host mirror of the reference's scene API (src/main.rs, src/hittable/*.rs
                         constructors, the SAH BVH build of src/hittable/bvh.rs)

Python is plumbing only (tests, bench, multi-GPU launch); no rendering or intersection code lives here,
and nothing in this package touches oracle/.  If the CUDA library is missing the import of the device
API fails loudly — there is no CPU fallback.

Names follow the reference: World, Sphere.new_still/new_moving, Quad, Cuboid, Instance, TriangleMesh,
DiffuseBRDF, MetalBRDF, GlassBSDF, PrincipledBSDF, DiffuseLight, SheenBRDF, ClearcoatBRDF, MixBxDf,
SolidTexture, CheckerTexture, ImageTexture, Camera.render.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
ASSETS_DIR = os.path.join(REPO_ROOT, "assets")
PT_NONE = 0xFFFFFFFF
PT_OK, PT_ERR_INVALID, PT_ERR_CUDA, PT_ERR_UNSUPPORTED, PT_ERR_NO_DEVICE = 0, -1, -2, -3, -4  # pt_status
PT_NAN_REFERENCE, PT_NAN_DROP = 0, 1
PT_RENDER_ENV_IMPORTANCE = 0x1  # pt_render_params.flags: environment-map importance sampling (not reference behaviour)
PT_RENDER_NEE = 0x2             # next-event estimation with MIS and shadow rays (not reference behaviour)
PRIM_SPHERE, PRIM_QUAD, PRIM_TRIANGLE, OBJ_CUBOID, OBJ_MESH, OBJ_INSTANCE = range(6)


# ---------------------------------------------------------------------------------------------- ABI structs
class Vec3(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]


class CameraABI(C.Structure):  # pt_camera
    _fields_ = [("aspect_ratio", C.c_double), ("image_width", C.c_uint32), ("samples_per_pixel", C.c_uint32),
                ("max_depth", C.c_uint32), ("env_is_map", C.c_uint32), ("vfov", C.c_double), ("look_from", Vec3),
                ("look_at", Vec3), ("vup", Vec3), ("blur_strength", C.c_double), ("focal_length", C.c_double),
                ("defocus_angle", C.c_double), ("env_color", Vec3), ("env_image", C.c_uint32), ("_pad", C.c_uint32)]


class RenderParams(C.Structure):  # pt_render_params
    _fields_ = [("seed", C.c_uint64), ("sample_begin", C.c_uint32), ("sample_count", C.c_uint32),
                ("sample_stride", C.c_uint32), ("nan_policy", C.c_uint32), ("pool_paths", C.c_uint32), ("flags", C.c_uint32)]


class Stats(C.Structure):  # pt_stats
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("nonfinite", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("iterations", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("device_ms", C.c_float),
                ("trace_ms", C.c_float), ("shade_ms", C.c_float), ("raygen_ms", C.c_float),
                ("node_pairs", C.c_uint64), ("ref_boxes", C.c_uint64), ("prim_tests", C.c_uint64),
                ("two_pass_iterations", C.c_uint32), ("queue_errors", C.c_uint32), ("p2p_shares", C.c_uint32), ("tail_paths", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


RAY_DTYPE = np.dtype([("origin", "<f8", 3), ("direction", "<f8", 3), ("time", "<f8")])
HIT_DTYPE = np.dtype([("t", "<f8"), ("u", "<f8"), ("v", "<f8"), ("point", "<f8", 3), ("geometric_normal", "<f8", 3),
                      ("shading_normal", "<f8", 3), ("hit", "<u4"), ("prim_kind", "<u4"), ("prim_index", "<u4"),
                      ("instance", "<u4"), ("material", "<u4"), ("front_face", "<u4"), ("is_light", "<u4"), ("work", "<u4")])
BSDF_QUERY_DTYPE = np.dtype([("view_dir", "<f8", 3), ("light_dir", "<f8", 3), ("point", "<f8", 3), ("geometric_normal", "<f8", 3),
                             ("shading_normal", "<f8", 3), ("u", "<f8"), ("v", "<f8"), ("front_face", "<u4"), ("_pad", "<u4")])
BSDF_RESULT_DTYPE = np.dtype([("eval", "<f8", 3), ("pdf", "<f8"), ("emitted", "<f8", 3), ("_pad", "<f8")])
BSDF_SAMPLE_DTYPE = np.dtype([("dir", "<f8", 3), ("valid", "<u4"), ("n_uniforms", "<u4")])
assert RAY_DTYPE.itemsize == 56 and HIT_DTYPE.itemsize == 128 and BSDF_QUERY_DTYPE.itemsize == 144
assert BSDF_RESULT_DTYPE.itemsize == 64 and BSDF_SAMPLE_DTYPE.itemsize == 32

# every symbol include/pt_b200.h declares (tests check the .so exports exactly these)
ABI_SYMBOLS = ["pt_ctx_create", "pt_ctx_destroy", "pt_ctx_set_stream", "pt_ctx_set_profiling", "pt_last_error", "pt_device_count",
               "pt_scene_create", "pt_scene_destroy", "pt_scene_device_bytes", "pt_camera_image_height", "pt_render_accumulate",
               "pt_render", "pt_tonemap_rgb8", "pt_trace_closest", "pt_trace_any", "pt_bsdf_eval_pdf", "pt_bsdf_sample",
               "pt_camera_rays", "pt_lights_sample_pdf", "pt_scene_build_env_sampler", "pt_env_sample_pdf", "pt_sah_sweep",
               "pt_render_multi", "pt_render_multi_release", "pt_trace_closest_wavefront", "pt_trace_camera_wavefront", "pt_debug_histograms", "pt_debug_stage_ms", "pt_debug_div_check"]


class PtError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------------------------------------- library loading
_dev = None
_host = None


_LIBDIR = os.environ.get("PT_B200_LIBDIR", os.path.join(_HERE, "lib"))  # override: A/B builds while tuning


def device_lib_path():
    return os.path.join(_LIBDIR, "libptb200.so")


def host_lib_path():
    return os.path.join(_LIBDIR, "libptb200_host.so")


def device_lib():
    """The CUDA library. Missing => hard error (no fallback path exists)."""
    global _dev
    if _dev is None:
        p = device_lib_path()
        if not os.path.exists(p):
            raise PtError(f"{p} is missing: build it with __graft_entry__.build() (nvcc, sm_100a); there is no CPU fallback")
        lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
        lib.pt_last_error.restype = C.c_char_p
        lib.pt_scene_device_bytes.restype = C.c_uint64
        lib.pt_camera_image_height.restype = C.c_uint32
        lib.pt_scene_device_bytes.argtypes = [C.c_void_p]
        lib.pt_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.pt_ctx_destroy.argtypes = [C.c_void_p]
        lib.pt_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.pt_ctx_set_profiling.argtypes = [C.c_void_p, C.c_int]
        lib.pt_scene_create.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.pt_scene_destroy.argtypes = [C.c_void_p]
        lib.pt_camera_image_height.argtypes = [C.POINTER(CameraABI)]
        lib.pt_render_accumulate.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CameraABI), C.POINTER(RenderParams), C.c_void_p, C.POINTER(Stats)]
        lib.pt_render.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CameraABI), C.POINTER(RenderParams), C.c_void_p, C.POINTER(Stats)]
        lib.pt_tonemap_rgb8.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint32, C.c_void_p]
        lib.pt_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p]
        lib.pt_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        lib.pt_trace_closest_wavefront.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_uint32, C.c_void_p, C.POINTER(Stats)]
        lib.pt_bsdf_eval_pdf.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.pt_bsdf_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pt_camera_rays.argtypes = [C.c_void_p, C.POINTER(CameraABI), C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pt_lights_sample_pdf.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pt_scene_build_env_sampler.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        lib.pt_env_sample_pdf.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pt_sah_sweep.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pt_render_multi.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(CameraABI), C.c_void_p, C.c_void_p, C.c_void_p]
        _dev = lib
    return _dev


def host_lib():
    global _host
    if _host is None:
        p = host_lib_path()  # opens libptb200.so itself, and only when a device entry point is called (host/device_api.cpp)
        if not os.path.exists(p):
            raise PtError(f"{p} is missing: build it with __graft_entry__.build()")
        lib = C.CDLL(p)
        lib.pth_last_error.restype = C.c_char_p
        vp = C.c_void_p
        d3 = C.POINTER(C.c_double)
        sig = {
            "pth_solid_texture": [C.c_double] * 3, "pth_solid_scalar": [C.c_double], "pth_checker_texture": [C.c_double, vp, vp],
            "pth_image": [vp, C.c_uint32, C.c_uint32], "pth_image_load": [C.c_char_p], "pth_image_texture": [vp],
            "pth_diffuse": [vp, vp], "pth_metal": [vp, vp], "pth_glass": [vp, vp, C.c_double], "pth_principled": [vp, d3],
            "pth_diffuse_light": [vp], "pth_sheen": [C.c_double] * 4, "pth_clearcoat": [C.c_double], "pth_mix": [C.c_double, vp, vp],
            "pth_sphere_still": [C.c_double, d3, vp], "pth_sphere_moving": [C.c_double, d3, d3, vp], "pth_quad": [d3, d3, d3, vp],
            "pth_cuboid": [d3, d3, vp], "pth_mesh_load": [C.c_char_p, C.c_double, vp],
            "pth_mesh_from_arrays": [C.c_double, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp],
            "pth_instance": [vp, d3, C.c_double, d3], "pth_volume": [vp, C.c_double, vp],
            "pth_world_new": [], "pth_scene_from_world": [vp, C.POINTER(CameraABI), vp],
            "pth_scene_build": [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.c_char_p, vp, C.c_uint32, C.c_uint32],
            "pth_scene_desc": [vp], "pth_scene_camera": [vp],
        }
        for name, args in sig.items():
            f = getattr(lib, name)
            f.restype = vp
            f.argtypes = args
        lib.pth_scene_camera.restype = C.POINTER(CameraABI)
        lib.pth_scene_output_name.restype = C.c_char_p
        lib.pth_scene_output_name.argtypes = [vp]
        for name in ["pth_world_add_object", "pth_world_add_light"]:
            getattr(lib, name).argtypes = [vp, vp]
            getattr(lib, name).restype = None
        for name in ["pth_world_build_bvh", "pth_world_free", "pth_scene_free", "pth_set_build_context"]:
            getattr(lib, name).argtypes = [vp]
            getattr(lib, name).restype = None
        lib.pth_scene_render.argtypes = [vp, C.c_char_p, C.c_uint64, C.c_int, C.c_uint32, C.c_int, C.POINTER(Stats)]
        lib.pth_write_png.argtypes = [C.c_char_p, vp, C.c_uint32, C.c_uint32]
        _host = lib
    return _host


def _h(ptr):
    if not ptr:
        raise PtError("host: " + host_lib().pth_last_error().decode())
    return ptr


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


# ---------------------------------------------------------------------------------------------- host mirror
class _Handle:
    def __init__(self, ptr, keep=()):
        self.ptr = _h(ptr)
        self._keep = keep  # buffers the native side may still reference


class SolidTexture(_Handle):
    def __init__(self, value):
        L = host_lib()
        if np.isscalar(value):
            super().__init__(L.pth_solid_scalar(float(value)))
        else:
            super().__init__(L.pth_solid_texture(*[float(x) for x in value]))


class CheckerTexture(_Handle):
    def __init__(self, scale, tex1, tex2):
        super().__init__(host_lib().pth_checker_texture(float(scale), tex1.ptr, tex2.ptr), (tex1, tex2))


class Image(_Handle):
    """RGB8 image as image::to_rgb8() yields it (reference src/texture.rs:62-69)."""

    def __init__(self, rgb=None, path=None):
        L = host_lib()
        if path is not None:
            super().__init__(L.pth_image_load(path.encode()))
        else:
            a = np.ascontiguousarray(rgb, dtype=np.uint8)
            assert a.ndim == 3 and a.shape[2] == 3
            super().__init__(L.pth_image(_ptr(a), a.shape[1], a.shape[0]))
        if not self.ptr:
            raise PtError(L.pth_last_error().decode())

    def pixels(self):
        """The decoded RGB8 pixels as a numpy array [H, W, 3] (a copy)."""
        L = host_lib()
        w, h = C.c_uint32(), C.c_uint32()
        L.pth_image_pixels.restype = C.POINTER(C.c_uint8)
        L.pth_image_pixels.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        p = L.pth_image_pixels(self.ptr, C.byref(w), C.byref(h))
        return np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()


def load_rgb8(path):
    """Decode an asset to RGB8 with PIL (independent of the C++ host's own PNG / JPEG decoders; used by tests and bakes)."""
    from PIL import Image as PILImage
    PILImage.MAX_IMAGE_PIXELS = None
    return np.asarray(PILImage.open(path).convert("RGB"), dtype=np.uint8)


class ImageTexture(_Handle):
    def __init__(self, image):
        super().__init__(host_lib().pth_image_texture(image.ptr), (image,))


def _tex(t):
    return t if isinstance(t, _Handle) else SolidTexture(t)


class DiffuseBRDF(_Handle):
    def __init__(self, base_color, normal_map=None):
        t = _tex(base_color)
        super().__init__(host_lib().pth_diffuse(t.ptr, normal_map.ptr if normal_map else None), (t, normal_map))


class MetalBRDF(_Handle):
    def __init__(self, base_color, roughness):
        t, r = _tex(base_color), _tex(roughness)
        super().__init__(host_lib().pth_metal(t.ptr, r.ptr), (t, r))


class GlassBSDF(_Handle):
    def __init__(self, base_color=(1, 1, 1), roughness=0.001, ior=1.5):
        t, r = _tex(base_color), _tex(roughness)
        super().__init__(host_lib().pth_glass(t.ptr, r.ptr, float(ior)), (t, r))

    @staticmethod
    def basic(ior):
        return GlassBSDF((1, 1, 1), 0.001, ior)


class PrincipledBSDF(_Handle):
    def __init__(self, base_color, metallic, roughness, subsurface, specular, specular_tint, ior, spec_trans, sheen,
                 sheen_tint, clearcoat, clearcoat_gloss):
        t = _tex(base_color)
        p = (C.c_double * 11)(metallic, roughness, subsurface, specular, specular_tint, ior, spec_trans, sheen, sheen_tint,
                              clearcoat, clearcoat_gloss)
        super().__init__(host_lib().pth_principled(t.ptr, p), (t,))


class DiffuseLight(_Handle):
    def __init__(self, emission):
        t = _tex(emission)
        super().__init__(host_lib().pth_diffuse_light(t.ptr), (t,))


class SheenBRDF(_Handle):
    def __init__(self, base_color, sheen_tint):
        super().__init__(host_lib().pth_sheen(*[float(x) for x in base_color], float(sheen_tint)))


class ClearcoatBRDF(_Handle):
    def __init__(self, clearcoat_gloss):
        super().__init__(host_lib().pth_clearcoat(float(clearcoat_gloss)))


class MixBxDf(_Handle):
    def __init__(self, t, bxdf1, bxdf2):
        super().__init__(host_lib().pth_mix(float(t), bxdf1.ptr, bxdf2.ptr), (bxdf1, bxdf2))


class Sphere(_Handle):
    @staticmethod
    def new_still(radius, position, material):
        return Sphere(host_lib().pth_sphere_still(float(radius), _d3(position), material.ptr), (material,))

    @staticmethod
    def new_moving(radius, position1, position2, material):
        return Sphere(host_lib().pth_sphere_moving(float(radius), _d3(position1), _d3(position2), material.ptr), (material,))


class Quad(_Handle):
    def __init__(self, q, u, v, material):
        super().__init__(host_lib().pth_quad(_d3(q), _d3(u), _d3(v), material.ptr), (material,))


class Cuboid(_Handle):
    def __init__(self, a, b, material):
        super().__init__(host_lib().pth_cuboid(_d3(a), _d3(b), material.ptr), (material,))


class TriangleMesh(_Handle):
    @staticmethod
    def from_obj(scale, path, material):
        return TriangleMesh(host_lib().pth_mesh_load(path.encode(), float(scale), material.ptr), (material,))

    @staticmethod
    def from_arrays(scale, positions, indices, material, texcoords=None, normals=None):
        pos = np.ascontiguousarray(positions, np.float32).ravel()
        idx = np.ascontiguousarray(indices, np.uint32).ravel()
        tex = None if texcoords is None else np.ascontiguousarray(texcoords, np.float32).ravel()
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float32).ravel()
        return TriangleMesh(host_lib().pth_mesh_from_arrays(float(scale), _ptr(pos), pos.size, _ptr(idx), idx.size,
                                                            _ptr(tex) if tex is not None else None, 0 if tex is None else tex.size,
                                                            _ptr(nrm) if nrm is not None else None, 0 if nrm is None else nrm.size,
                                                            material.ptr), (material,))


class Instance(_Handle):
    def __init__(self, obj, axis, angle, translation):
        super().__init__(host_lib().pth_instance(obj.ptr, _d3(axis), float(angle), _d3(translation)), (obj,))


class HomogeneousVolume(_Handle):
    """volume.rs:15-41 (a stub in the reference; semantics in include/pt_b200.h): constant-density medium inside a sphere
    or cuboid boundary with an isotropic phase function of the given albedo (colour or texture)."""

    def __init__(self, boundary, density, albedo):
        t = _tex(albedo)
        p = host_lib().pth_volume(boundary.ptr, float(density), t.ptr)
        if not p:
            raise PtError(host_lib().pth_last_error().decode())
        super().__init__(p, (boundary, t))


class World(_Handle):
    def __init__(self):
        super().__init__(host_lib().pth_world_new())
        self._objs = []

    def add_object(self, h):
        self._objs.append(h)
        host_lib().pth_world_add_object(self.ptr, h.ptr)

    def add_light(self, h):
        self._objs.append(h)
        host_lib().pth_world_add_light(self.ptr, h.ptr)

    def build_bvh(self):
        host_lib().pth_world_build_bvh(self.ptr)


def set_build_context(ctx):
    """BVH builds of this thread price the splits of large nodes on the device (pt_sah_sweep): same trees, much faster.
    Pass None to go back to the host-only build."""
    host_lib().pth_set_build_context(ctx.ptr if ctx is not None else None)


def make_camera(image_width, aspect_ratio=1.0, samples_per_pixel=1, max_depth=50, vfov=40.0, look_from=(0, 0, 0), look_at=(0, 0, -1),
                vup=(0, 1, 0), blur_strength=0.5, focal_length=10.0, defocus_angle=0.0, env_color=(0, 0, 0), env_is_map=False):
    c = CameraABI()
    c.aspect_ratio, c.image_width, c.samples_per_pixel, c.max_depth = aspect_ratio, image_width, samples_per_pixel, max_depth
    c.vfov, c.blur_strength, c.focal_length, c.defocus_angle = vfov, blur_strength, focal_length, defocus_angle
    c.look_from, c.look_at, c.vup, c.env_color = Vec3(*look_from), Vec3(*look_at), Vec3(*vup), Vec3(*env_color)
    c.env_is_map, c.env_image = int(env_is_map), PT_NONE
    return c


class Scene:
    """A flattened scene (pt_scene_desc + pt_camera) owned by the host library."""

    def __init__(self, ptr, keep=()):
        self.ptr = _h(ptr)
        self._keep = keep
        L = host_lib()
        self.desc = L.pth_scene_desc(self.ptr)  # const pt_scene_desc*
        self.camera = L.pth_scene_camera(self.ptr).contents
        self.output_name = L.pth_scene_output_name(self.ptr).decode()

    @staticmethod
    def from_world(world, camera, env_image=None):
        return Scene(host_lib().pth_scene_from_world(world.ptr, C.byref(camera), env_image.ptr if env_image else None), (world, env_image))

    @staticmethod
    def build(scene_id, width=600, spp=100, seed=1, assets_dir=ASSETS_DIR):
        """One of the reference's scenes (src/main.rs:635-644); 70 = our mesh variant of scene 7."""
        env = None  # an RGB8 array here would override scene 5's assets/envmap.jpg (decoded natively by host/jpeg.cpp)
        ptr = host_lib().pth_scene_build(scene_id, width, spp, seed, assets_dir.encode(), _ptr(env) if env is not None else None,
                                         0 if env is None else env.shape[1], 0 if env is None else env.shape[0])
        return Scene(ptr)

    def camera_copy(self, **overrides):
        c = CameraABI.from_buffer_copy(self.camera)
        for k, v in overrides.items():
            setattr(c, k, v)
        return c

    def image_height(self, camera=None):
        cam = camera if camera is not None else self.camera
        return int(int(cam.image_width) / cam.aspect_ratio)

    def __del__(self):
        try:
            host_lib().pth_scene_free(self.ptr)
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------- device API
class Context:
    """pt_ctx on one CUDA device."""

    def __init__(self, device=0):
        self.lib = device_lib()
        self.ptr = C.c_void_p()
        self._check(self.lib.pt_ctx_create(device, C.byref(self.ptr)))

    def _check(self, rc):
        if rc != 0:
            raise PtError(f"pt_b200 error {rc}: {self.lib.pt_last_error().decode()}")

    def set_stream(self, cuda_stream):
        """Launch on an external stream (e.g. torch.cuda.current_stream().cuda_stream).  Handle 0 is the legacy
        default stream, which the C ABI spells cudaStreamLegacy (0x1) because NULL there means "the ctx's own"."""
        self._check(self.lib.pt_ctx_set_stream(self.ptr, C.c_void_p(cuda_stream if cuda_stream else 1)))

    def set_profiling(self, level):
        """0 off, 1 per-stage event times, 2 also the traversal work counters (slower kernel variants)."""
        self._check(self.lib.pt_ctx_set_profiling(self.ptr, int(level)))

    def upload(self, scene):
        return DeviceScene(self, scene)

    def tonemap_rgb8(self, d_accum_ptr, scale, height, width):
        """pt_tonemap_rgb8: sqrt-gamma + 8-bit quantisation (camera.rs:109-114,128-130) of `scale * accum` on the device;
        d_accum_ptr = device pointer of H*W*3 fp32 radiance sums.  Returns uint8 [H, W, 3] in host memory."""
        out = np.zeros((height, width, 3), dtype=np.uint8)
        self._check(self.lib.pt_tonemap_rgb8(self.ptr, C.c_void_p(d_accum_ptr), float(scale), height * width, _ptr(out)))
        return out

    def stage_ms(self, reset=True):
        """pt_debug_stage_ms: per-kernel-family times of the traversal stage (profiling level >= 1) since the last reset."""
        out = np.zeros(16, dtype=np.float64)
        self.lib.pt_debug_stage_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self._check(self.lib.pt_debug_stage_ms(self.ptr, _ptr(out), int(reset)))
        names = ("top_old", "top_new", "mesh_enter", "mesh_walk", "bvh", "generate", "tail", "mesh_multi", "miss", "light", "diffuse", "metal", "glass", "principled", "other")
        return {k: v for k, v in zip(names, out.tolist()) if k != "-"}

    def div_check(self, n, seed=1):
        """pt_debug_div_check: quotients of the shared-reciprocal DVec3 / f64 that differ from the `/` operator's bits (must be 0)."""
        out = C.c_uint64(0)
        self.lib.pt_debug_div_check.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        self._check(self.lib.pt_debug_div_check(self.ptr, int(n), int(seed), C.byref(out)))
        return int(out.value)

    def histograms(self, reset=True):
        """pt_debug_histograms: [8, 64] counters of the profiling-level-2 kernel variants since the last reset."""
        out = np.zeros((8, 64), dtype=np.uint64)
        self.lib.pt_debug_histograms.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self._check(self.lib.pt_debug_histograms(self.ptr, _ptr(out), int(reset)))
        return out

    def close(self):
        if self.ptr:
            self.lib.pt_ctx_destroy(self.ptr)
            self.ptr = None


class DeviceScene:
    def __init__(self, ctx, scene):
        self.ctx, self.host_scene = ctx, scene
        self.ptr = C.c_void_p()
        ctx._check(ctx.lib.pt_scene_create(ctx.ptr, scene.desc, C.byref(self.ptr)))
        self.device_bytes = ctx.lib.pt_scene_device_bytes(self.ptr)

    def close(self):
        if self.ptr:
            self.ctx.lib.pt_scene_destroy(self.ptr)
            self.ptr = None

    # World::intersect_all for a batch of rays
    def trace_closest(self, rays, t_min=1e-3):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        self.ctx._check(self.ctx.lib.pt_trace_closest(self.ctx.ptr, self.ptr, rays.shape[0], _ptr(rays), t_min, _ptr(hits)))
        return hits

    def trace_closest_wavefront(self, rays, t_min=1e-3, flags=0):
        """World::intersect_all through the render's own traversal stage (pt_trace_closest_wavefront); returns (hits, stats)."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        st = Stats()
        self.ctx._check(self.ctx.lib.pt_trace_closest_wavefront(self.ctx.ptr, self.ptr, rays.shape[0], _ptr(rays), t_min, flags, _ptr(hits), C.byref(st)))
        return hits, st

    def trace_camera_wavefront(self, camera=None, seed=1, sample=0, flags=0):
        """pt_trace_camera_wavefront: (camera rays, their closest hits, stats) of one sample per pixel through the render's
        own start-of-path traversal stage."""
        cam = camera if camera is not None else self.host_scene.camera
        n = self.ctx.lib.pt_camera_image_height(C.byref(cam)) * cam.image_width
        rays, hits, st = np.zeros(n, dtype=RAY_DTYPE), np.zeros(n, dtype=HIT_DTYPE), Stats()
        self.ctx.lib.pt_trace_camera_wavefront.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CameraABI), C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        self.ctx._check(self.ctx.lib.pt_trace_camera_wavefront(self.ctx.ptr, self.ptr, C.byref(cam), seed, sample, flags, _ptr(rays), _ptr(hits), C.byref(st)))
        return rays, hits, st

    def trace_any(self, rays, t_max, t_min=1e-3):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        t_max = np.ascontiguousarray(t_max, dtype=np.float64)
        out = np.zeros(rays.shape[0], dtype=np.uint8)
        self.ctx._check(self.ctx.lib.pt_trace_any(self.ctx.ptr, self.ptr, rays.shape[0], _ptr(rays), t_min, _ptr(t_max), _ptr(out)))
        return out

    def bsdf_eval_pdf(self, material, queries):
        q = np.ascontiguousarray(queries, dtype=BSDF_QUERY_DTYPE)
        out = np.zeros(q.shape[0], dtype=BSDF_RESULT_DTYPE)
        self.ctx._check(self.ctx.lib.pt_bsdf_eval_pdf(self.ctx.ptr, self.ptr, material, q.shape[0], _ptr(q), _ptr(out)))
        return out

    def bsdf_sample(self, material, queries, uniforms8):
        q = np.ascontiguousarray(queries, dtype=BSDF_QUERY_DTYPE)
        u = np.ascontiguousarray(uniforms8, dtype=np.float64).reshape(q.shape[0], 8)
        out = np.zeros(q.shape[0], dtype=BSDF_SAMPLE_DTYPE)
        self.ctx._check(self.ctx.lib.pt_bsdf_sample(self.ctx.ptr, self.ptr, material, q.shape[0], _ptr(q), _ptr(u), _ptr(out)))
        return out

    def lights_sample_pdf(self, origins, times, uniforms4):
        o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        n = o.shape[0]
        t = np.ascontiguousarray(times, dtype=np.float64)
        u = np.ascontiguousarray(uniforms4, dtype=np.float64).reshape(n, 4)
        d, valid, pdf = np.zeros((n, 3)), np.zeros(n, np.uint32), np.zeros(n)
        self.ctx._check(self.ctx.lib.pt_lights_sample_pdf(self.ctx.ptr, self.ptr, n, _ptr(o), _ptr(t), _ptr(u), _ptr(d), _ptr(valid), _ptr(pdf)))
        return d, valid, pdf

    def build_env_sampler(self, image=None, max_rows=0, max_cols=0):
        """pt_scene_build_env_sampler for `image` (default: the camera's environment map); needed by PT_RENDER_ENV_IMPORTANCE."""
        image = self.host_scene.camera.env_image if image is None else image
        self.ctx._check(self.ctx.lib.pt_scene_build_env_sampler(self.ptr, image, max_rows, max_cols))

    def env_sample_pdf(self, uniforms2):
        u = np.ascontiguousarray(uniforms2, dtype=np.float64).reshape(-1, 2)
        d, pdf = np.zeros((u.shape[0], 3)), np.zeros(u.shape[0])
        self.ctx._check(self.ctx.lib.pt_env_sample_pdf(self.ctx.ptr, self.ptr, u.shape[0], _ptr(u), _ptr(d), _ptr(pdf)))
        return d, pdf

    def params(self, spp, seed=1, sample_begin=0, sample_stride=1, nan_policy=PT_NAN_REFERENCE, pool_paths=0, flags=0):
        return RenderParams(seed, sample_begin, spp, sample_stride, nan_policy, pool_paths, flags)

    def render(self, camera=None, spp=None, **kw):
        """Camera::render minus the PNG: mean radiance as float32 [H, W, 3] in host memory, plus stats."""
        cam = camera if camera is not None else self.host_scene.camera
        spp = spp if spp is not None else cam.samples_per_pixel
        h = self.ctx.lib.pt_camera_image_height(C.byref(cam))
        out = np.zeros((h, cam.image_width, 3), dtype=np.float32)
        st = Stats()
        p = self.params(spp, **kw)
        self.ctx._check(self.ctx.lib.pt_render(self.ctx.ptr, self.ptr, C.byref(cam), C.byref(p), _ptr(out), C.byref(st)))
        return out, st

    def render_accumulate(self, d_accum_ptr, camera=None, spp=None, **kw):
        """Adds radiance sums into a caller-owned device buffer (e.g. a torch CUDA tensor's data_ptr())."""
        cam = camera if camera is not None else self.host_scene.camera
        spp = spp if spp is not None else cam.samples_per_pixel
        st = Stats()
        p = self.params(spp, **kw)
        self.ctx._check(self.ctx.lib.pt_render_accumulate(self.ctx.ptr, self.ptr, C.byref(cam), C.byref(p), C.c_void_p(d_accum_ptr), C.byref(st)))
        return st


def render_multi(scene, devices, camera=None, spp=None, seed=1, sample_begin=0, sample_stride=1, nan_policy=PT_NAN_REFERENCE, pool_paths=0, flags=0):
    """pt_render_multi: Camera::render on several GPUs of this one process (one host thread per listed device)."""
    lib = device_lib()
    cam = camera if camera is not None else scene.camera
    spp = spp if spp is not None else cam.samples_per_pixel
    h = lib.pt_camera_image_height(C.byref(cam))
    out = np.zeros((h, cam.image_width, 3), dtype=np.float32)
    st = Stats()
    p = RenderParams(seed, sample_begin, spp, sample_stride, nan_policy, pool_paths, flags)
    devs = (C.c_int * len(devices))(*devices)
    rc = lib.pt_render_multi(len(devices), devs, scene.desc, C.byref(cam), C.byref(p), _ptr(out), C.byref(st))
    if rc != 0:
        raise PtError(f"pt_b200 error {rc}: {lib.pt_last_error().decode()}")
    return out, st


def camera_rays(ctx, camera, seed, rows, cols, samples):
    rows, cols, samples = (np.ascontiguousarray(a, dtype=np.uint32) for a in (rows, cols, samples))
    out = np.zeros(rows.shape[0], dtype=RAY_DTYPE)
    ctx._check(ctx.lib.pt_camera_rays(ctx.ptr, C.byref(camera), seed, rows.shape[0], _ptr(rows), _ptr(cols), _ptr(samples), _ptr(out)))
    return out


def tonemap_rgb8(mean):
    """camera.rs:109-114,128-130 on a host array of mean radiance."""
    g = np.sqrt(np.maximum(np.nan_to_num(np.asarray(mean, np.float64), nan=0.0, posinf=np.inf, neginf=0.0), 0.0))
    return (np.clip(g, 0.0, 0.999) * 256.0).astype(np.uint8)


def write_png(path, rgb8):
    a = np.ascontiguousarray(rgb8, dtype=np.uint8)
    if host_lib().pth_write_png(path.encode(), _ptr(a), a.shape[1], a.shape[0]) != 0:
        raise PtError("cannot write " + path)
