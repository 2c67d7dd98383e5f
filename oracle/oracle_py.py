"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_build/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (thu-acg-f2024-path-tracer_b200/) never does.
PARITY UNPINNED except via the reference's demo/*.png (tests/golden/).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


class OrcStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("boxes", C.c_uint64), ("spheres", C.c_uint64), ("quads", C.c_uint64),
                ("triangles", C.c_uint64), ("instances", C.c_uint64), ("nonfinite", C.c_uint64), ("seconds", C.c_double), ("threads", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_scene_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_bvh_signature.restype = C.c_int64
        L.orc_bvh_signature.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        L.orc_dump_path_rays.restype = C.c_int64
        L.orc_dump_path_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int64]
        L.orc_camera_image_height.restype = C.c_uint32
        L.orc_camera_image_height.argtypes = [C.c_void_p]
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p]
        L.orc_trace_any.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_bsdf_eval_pdf.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_bsdf_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_lights_sample_pdf.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_camera_rays.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(OrcStats)]
        L.orc_tonemap_rgb8.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_scene_build_env_sampler.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_env_sample_pdf.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_item_bbox.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
        L.orc_quad_derived.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
        L.orc_instance_matrices.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleScene:
    """The restated reference World built from the same pt_scene_desc the device library consumes."""

    def __init__(self, desc_ptr, dtypes):
        self.L = lib()
        self.dt = dtypes  # the ABI numpy dtypes (module pt_b200), passed in so oracle/ imports nothing from the product
        self.ptr = C.c_void_p()
        if self.L.orc_scene_create(desc_ptr, C.byref(self.ptr)) != 0:
            raise RuntimeError("oracle: " + self.L.orc_last_error().decode())

    def close(self):
        if self.ptr:
            self.L.orc_scene_destroy(self.ptr)
            self.ptr = None

    def trace_closest(self, rays, t_min=1e-3):
        rays = np.ascontiguousarray(rays, dtype=self.dt.RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=self.dt.HIT_DTYPE)
        self.L.orc_trace_closest(self.ptr, rays.shape[0], _ptr(rays), t_min, _ptr(hits))
        return hits

    def trace_any(self, rays, t_max, t_min=1e-3):
        rays = np.ascontiguousarray(rays, dtype=self.dt.RAY_DTYPE)
        t_max = np.ascontiguousarray(t_max, dtype=np.float64)
        out = np.zeros(rays.shape[0], dtype=np.uint8)
        self.L.orc_trace_any(self.ptr, rays.shape[0], _ptr(rays), t_min, _ptr(t_max), _ptr(out))
        return out

    def bsdf_eval_pdf(self, material, queries):
        q = np.ascontiguousarray(queries, dtype=self.dt.BSDF_QUERY_DTYPE)
        out = np.zeros(q.shape[0], dtype=self.dt.BSDF_RESULT_DTYPE)
        assert self.L.orc_bsdf_eval_pdf(self.ptr, material, q.shape[0], _ptr(q), _ptr(out)) == 0
        return out

    def bsdf_sample(self, material, queries, uniforms8):
        q = np.ascontiguousarray(queries, dtype=self.dt.BSDF_QUERY_DTYPE)
        u = np.ascontiguousarray(uniforms8, dtype=np.float64).reshape(q.shape[0], 8)
        out = np.zeros(q.shape[0], dtype=self.dt.BSDF_SAMPLE_DTYPE)
        assert self.L.orc_bsdf_sample(self.ptr, material, q.shape[0], _ptr(q), _ptr(u), _ptr(out)) == 0
        return out

    def lights_sample_pdf(self, origins, times, uniforms4):
        o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        n = o.shape[0]
        t = np.ascontiguousarray(times, dtype=np.float64)
        u = np.ascontiguousarray(uniforms4, dtype=np.float64).reshape(n, 4)
        d, valid, pdf = np.zeros((n, 3)), np.zeros(n, np.uint32), np.zeros(n)
        self.L.orc_lights_sample_pdf(self.ptr, n, _ptr(o), _ptr(t), _ptr(u), _ptr(d), _ptr(valid), _ptr(pdf))
        return d, valid, pdf

    def build_env_sampler(self, image, max_rows=0, max_cols=0):
        assert self.L.orc_scene_build_env_sampler(self.ptr, image, max_rows, max_cols) == 0, self.L.orc_last_error().decode()

    def env_sample_pdf(self, uniforms2):
        u = np.ascontiguousarray(uniforms2, dtype=np.float64).reshape(-1, 2)
        d, pdf = np.zeros((u.shape[0], 3)), np.zeros(u.shape[0])
        assert self.L.orc_env_sample_pdf(self.ptr, u.shape[0], _ptr(u), _ptr(d), _ptr(pdf)) == 0
        return d, pdf

    def render(self, camera, spp, seed=1, sample_begin=0, sample_stride=1, nan_policy=0, threads=0, flags=0):
        """Camera::render minus the PNG.  Returns (mean radiance float64 [H,W,3], OrcStats)."""
        params = self.dt.RenderParams(seed, sample_begin, spp, sample_stride, nan_policy, 0, flags)
        h = self.L.orc_camera_image_height(C.byref(camera))
        out = np.zeros((h, camera.image_width, 3), dtype=np.float64)
        st = OrcStats()
        assert self.L.orc_render(self.ptr, C.byref(camera), C.byref(params), threads, _ptr(out), C.byref(st)) == 0, self.L.orc_last_error().decode()
        return out, st

    def dump_path_rays(self, camera, seed, pixel_step, spp, min_bounce, cap):
        out = np.zeros(cap, dtype=self.dt.RAY_DTYPE)
        n = self.L.orc_dump_path_rays(self.ptr, C.byref(camera), seed, pixel_step, spp, min_bounce, _ptr(out), cap)
        return out[:n].copy()

    def bvh_signature(self, which, with_boxes=False):
        n = self.L.orc_bvh_signature(self.ptr, which, None, 0, None, 0)
        sig = np.zeros(max(n, 1), dtype=np.int64)
        boxes = np.zeros(6 * max(n, 1), dtype=np.float64)
        self.L.orc_bvh_signature(self.ptr, which, _ptr(sig), n, _ptr(boxes) if with_boxes else None, boxes.size)
        return (sig[:n], boxes) if with_boxes else sig[:n]

    def item_bbox(self, which, i):
        out = np.zeros(6)
        assert self.L.orc_item_bbox(self.ptr, which, i, _ptr(out)) == 0
        return out

    def quad_derived(self, which, i):
        out = np.zeros(7)
        return out if self.L.orc_quad_derived(self.ptr, which, i, _ptr(out)) == 0 else None

    def instance_matrices(self, which, i):
        out = np.zeros(48)
        return out.reshape(3, 16) if self.L.orc_instance_matrices(self.ptr, which, i, _ptr(out)) == 0 else None


def camera_rays(camera, seed, rows, cols, samples, dtypes):
    rows, cols, samples = (np.ascontiguousarray(a, dtype=np.uint32) for a in (rows, cols, samples))
    out = np.zeros(rows.shape[0], dtype=dtypes.RAY_DTYPE)
    lib().orc_camera_rays(C.byref(camera), seed, rows.shape[0], _ptr(rows), _ptr(cols), _ptr(samples), _ptr(out))
    return out


def tonemap_rgb8(mean):
    a = np.ascontiguousarray(mean, dtype=np.float64)
    out = np.zeros(a.shape, dtype=np.uint8)
    lib().orc_tonemap_rgb8(_ptr(a), a.size, _ptr(out))
    return out


def num_threads():
    return lib().orc_num_threads()
