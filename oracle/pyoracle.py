"""ctypes view of oracle/libgi_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (gpugi_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgi_oracle.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


if not os.path.exists(LIB_PATH):
    build()
lib = C.CDLL(LIB_PATH)

vp, u32, i32, f32 = C.c_void_p, C.c_uint32, C.c_int32, C.c_float

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("t_max", "<f4"), ("direction", "<f4", 3), ("flags", "<u4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("leaf_tri", "<i4"), ("bary_y", "<f4"), ("bary_z", "<f4")])
LIGHT_SAMPLE_DTYPE = np.dtype([("position", "<f4", 3), ("normal_phi", "<f4"), ("intensity", "<f4", 3), ("normal_theta_cos", "<f4")])
COUNTER_NAMES = ["closest_rays", "any_rays", "box_tests", "tri_tests", "camera_paths", "light_paths", "shaded_vertices", "nee", "connections", "splats",
                 "stored_light_vertices"]


class OrCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in COUNTER_NAMES]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n in COUNTER_NAMES}


class OrRegion(C.Structure):
    _fields_ = [("x0", i32), ("y0", i32), ("x1", i32), ("y1", i32)]


lib.or_num_threads.restype = C.c_int
lib.or_wanghash.restype = u32
lib.or_wanghash.argtypes = [u32]
lib.or_xorshift32.restype = u32
lib.or_xorshift32.argtypes = [u32]
lib.or_xorshift64.restype = C.c_uint64
lib.or_xorshift64.argtypes = [C.c_uint64]
lib.or_xorshift_f.restype = C.c_float
lib.or_xorshift_f.argtypes = [C.POINTER(C.c_uint64)]
lib.or_trace_batch.restype = C.c_int
lib.or_trace_batch.argtypes = [vp, vp, u32, vp, vp, vp]
lib.or_intersect_box.restype = C.c_int
lib.or_intersect_box.argtypes = [vp, vp, vp, vp, C.POINTER(f32), C.POINTER(f32)]
lib.or_intersect_triangle.restype = C.c_int
lib.or_intersect_triangle.argtypes = [vp, vp, vp, vp, vp, C.POINTER(f32), vp, vp]
lib.or_render_whitted.restype = C.c_int
lib.or_render_whitted.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, u32, C.c_int, vp, vp]
lib.or_set_environment_cube.restype = None
lib.or_set_environment_cube.argtypes = [C.c_int, vp]
lib.or_render_pt.restype = C.c_int
lib.or_render_pt.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, u32, C.c_int, vp, vp, vp, vp]
lib.or_render_lpt.restype = C.c_int
lib.or_render_lpt.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, u32, C.c_int, C.c_int, u32, u32, C.c_int, vp, vp]
lib.or_render_bpt.restype = C.c_int
lib.or_render_bpt.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, u32, C.c_int, C.c_int, u32, vp, C.c_int, vp, vp, vp, vp]
lib.or_render_bpt_with_cache.restype = C.c_int
lib.or_render_bpt_with_cache.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, u32, C.c_int, C.c_int, u32, vp, vp, u32, vp, vp, vp, vp]
lib.or_camera.restype = None
lib.or_camera.argtypes = [vp, vp, vp, f32, f32, C.c_int, C.c_int, vp]
lib.or_light_sampler_create.restype = vp
lib.or_light_sampler_create.argtypes = [vp, vp, vp, u32, vp, vp, u32]
lib.or_light_sampler_destroy.restype = None
lib.or_light_sampler_destroy.argtypes = [vp]
lib.or_light_sampler_seed.restype = None
lib.or_light_sampler_seed.argtypes = [vp, C.c_uint64]
lib.or_light_sampler_num_light_triangles.restype = u32
lib.or_light_sampler_num_light_triangles.argtypes = [vp]
lib.or_light_sampler_generate.restype = None
lib.or_light_sampler_generate.argtypes = [vp, vp, u32, f32]
lib.or_mse.restype = C.c_double
lib.or_mse.argtypes = [vp, vp, C.c_int, C.c_int, u32]


def _p(a):
    return a.ctypes.data_as(vp)


def _f3(v):
    return np.ascontiguousarray(v, dtype=np.float32)


lib.or_set_num_threads.restype = None
lib.or_set_num_threads.argtypes = [C.c_int]


def set_num_threads(n):
    lib.or_set_num_threads(int(n))


def num_threads():
    return int(lib.or_num_threads())


def wanghash(x):
    return int(lib.or_wanghash(x & 0xFFFFFFFF))


def intersect_box(o, d, bmin, bmax):
    o, d, bmin, bmax = _f3(o), _f3(d), _f3(bmin), _f3(bmax)
    first, last = f32(), f32()
    hit = lib.or_intersect_box(_p(o), _p(d), _p(bmin), _p(bmax), C.byref(first), C.byref(last))
    return bool(hit), first.value, last.value


def intersect_triangle(o, d, p0, p1, p2):
    o, d, p0, p1, p2 = _f3(o), _f3(d), _f3(p0), _f3(p1), _f3(p2)
    t = f32()
    bary = np.zeros(3, dtype=np.float32)
    nrm = np.zeros(3, dtype=np.float32)
    hit = lib.or_intersect_triangle(_p(o), _p(d), _p(p0), _p(p1), _p(p2), C.byref(t), _p(bary), _p(nrm))
    return bool(hit), t.value, bary, nrm


def trace_batch(scene_desc, rays, want_counts=False):
    """scene_desc: a ctypes GiSceneDesc (layout of include/gpugi_b200.h)."""
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
    nodes = np.zeros(rays.shape[0], dtype=np.uint32) if want_counts else None
    tris = np.zeros(rays.shape[0], dtype=np.uint32) if want_counts else None
    lib.or_trace_batch(C.addressof(scene_desc), _p(rays), rays.shape[0], _p(hits), _p(nodes) if want_counts else None, _p(tris) if want_counts else None)
    if want_counts:
        return hits, nodes, tris
    return hits


def camera(pos, lookat, yfov, width, height, up=(0.0, 1.0, 0.0)):
    """-> 13 floats in GiCamera layout."""
    out = np.zeros(13, dtype=np.float32)
    pos, lookat, up = _f3(pos), _f3(lookat), _f3(up)
    lib.or_camera(_p(pos), _p(lookat), _p(up), float(yfov), float(width) / float(height), int(width), int(height), _p(out))
    return out


def _cam_ptr(cam):
    if isinstance(cam, np.ndarray):
        return _p(cam)
    return C.addressof(cam)


def _region(region):
    if region is None:
        return None
    r = OrRegion(*[int(v) for v in region])
    return r


def render_pt(scene_desc, cam, w, h, light_samples, frame_seed, max_len=16, env_rgb=None, region=None, accum=None):
    ls = np.ascontiguousarray(light_samples, dtype=LIGHT_SAMPLE_DTYPE)
    if accum is None:
        accum = np.zeros((h, w, 4), dtype=np.float32)
    cnt = OrCounters()
    env = _f3(env_rgb) if env_rgb is not None else None
    reg = _region(region)
    lib.or_render_pt(C.addressof(scene_desc), _cam_ptr(cam), w, h, _p(ls), ls.shape[0], frame_seed & 0xFFFFFFFF, max_len, _p(env) if env is not None else None,
                     C.addressof(reg) if reg is not None else None, _p(accum), C.addressof(cnt))
    return accum, cnt.as_dict()


def render_whitted(scene_desc, cam, w, h, light_samples, frame_seed, max_len=16, accum=None):
    ls = np.ascontiguousarray(light_samples, dtype=LIGHT_SAMPLE_DTYPE)
    if accum is None:
        accum = np.zeros((h, w, 4), dtype=np.float32)
    cnt = OrCounters()
    lib.or_render_whitted(C.addressof(scene_desc), _cam_ptr(cam), w, h, _p(ls), ls.shape[0], frame_seed & 0xFFFFFFFF, max_len, _p(accum), C.addressof(cnt))
    return accum, cnt.as_dict()


def set_environment_cube(faces):
    """faces: None, or 6 (size, size, 4) uint8 sRGB arrays in the order -X,+X,-Y,+Y,-Z,+Z, used by the following render_pt calls."""
    global _env_keep
    if faces is None:
        lib.or_set_environment_cube(0, None)
        _env_keep = None
        return
    arrs = [np.ascontiguousarray(f, dtype=np.uint8) for f in faces]
    ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
    _env_keep = (arrs, ptrs)
    lib.or_set_environment_cube(arrs[0].shape[0], ptrs)


_env_keep = None


lib.or_set_variant.restype = None
lib.or_set_variant.argtypes = [C.c_int, C.c_int]


def set_variant(reconstruction_filter=0, chained_mis=False):
    """Shader #define variants for the render calls that follow: RECONSTRUCTION_FILTER 0-3 (screenreproject.glsl:14-17), CHAINED_MIS."""
    lib.or_set_variant(int(reconstruction_filter), 1 if chained_mis else 0)


def render_lpt(scene_desc, cam, w, h, light_samples, frame_seed, num_rays_per_light_sample, max_len=16, first_path=0, num_paths=0, show_light_caches=True,
               accum=None):
    ls = np.ascontiguousarray(light_samples, dtype=LIGHT_SAMPLE_DTYPE)
    if accum is None:
        accum = np.zeros((h, w, 4), dtype=np.float32)
    cnt = OrCounters()
    lib.or_render_lpt(C.addressof(scene_desc), _cam_ptr(cam), w, h, _p(ls), ls.shape[0], frame_seed & 0xFFFFFFFF, max_len, int(num_rays_per_light_sample),
                      int(first_path), int(num_paths), 1 if show_light_caches else 0, _p(accum), C.addressof(cnt))
    return accum, cnt.as_dict()


def render_bpt(scene_desc, cam, w, h, light_samples, frame_seed, num_rays_per_light_sample, capacity, max_len=16, warmup=False, region=None, accum=None,
               return_cache=False):
    ls = np.ascontiguousarray(light_samples, dtype=LIGHT_SAMPLE_DTYPE)
    if accum is None:
        accum = np.zeros((h, w, 4), dtype=np.float32)
    cnt = OrCounters()
    counts = np.zeros(2, dtype=np.uint32)
    cache = None if warmup else np.zeros(int(capacity) * 16, dtype=np.float32)
    reg = _region(region)
    lib.or_render_bpt(C.addressof(scene_desc), _cam_ptr(cam), w, h, _p(ls), ls.shape[0], frame_seed & 0xFFFFFFFF, max_len, int(num_rays_per_light_sample),
                      int(capacity), _p(cache) if cache is not None else None, 1 if warmup else 0, C.addressof(reg) if reg is not None else None, _p(accum),
                      _p(counts), C.addressof(cnt))
    if return_cache:
        return accum, cnt.as_dict(), (int(counts[0]), int(counts[1])), cache.reshape(-1, 16)
    return accum, cnt.as_dict(), (int(counts[0]), int(counts[1]))


def render_bpt_with_cache(scene_desc, cam, w, h, light_samples, frame_seed, num_rays_per_light_sample, capacity, eye_cache, eye_cache_entries, max_len=16,
                          region=None, accum=None):
    """One BPT iteration whose eye pass reads `eye_cache` (capacity x 16 floats, LightCacheEntry layout) with `eye_cache_entries` as its
    NumLightCacheEntries counter -- e.g. the cache another implementation's light pass wrote."""
    ls = np.ascontiguousarray(light_samples, dtype=LIGHT_SAMPLE_DTYPE)
    if accum is None:
        accum = np.zeros((h, w, 4), dtype=np.float32)
    cnt = OrCounters()
    counts = np.zeros(2, dtype=np.uint32)
    own = np.zeros(int(capacity) * 16, dtype=np.float32)
    given = np.ascontiguousarray(eye_cache, dtype=np.float32).reshape(-1)
    assert given.size >= int(capacity) * 16
    reg = _region(region)
    lib.or_render_bpt_with_cache(C.addressof(scene_desc), _cam_ptr(cam), w, h, _p(ls), ls.shape[0], frame_seed & 0xFFFFFFFF, max_len, int(num_rays_per_light_sample),
                                 int(capacity), _p(own), _p(given), int(eye_cache_entries), C.addressof(reg) if reg is not None else None, _p(accum), _p(counts),
                                 C.addressof(cnt))
    return accum, cnt.as_dict(), (int(counts[0]), int(counts[1]))


class LightSampler:
    """Restatement of Scene::LoadLightSources + ComputePointLightTable + LightSampler::GenerateRandomSamples."""

    def __init__(self, positions, triangles, tri_materials, emissivity, point_lights):
        self._keep = [np.ascontiguousarray(positions, dtype=np.float32), np.ascontiguousarray(triangles, dtype=np.uint32),
                      np.ascontiguousarray(tri_materials, dtype=np.uint32), np.ascontiguousarray(emissivity, dtype=np.float32),
                      np.ascontiguousarray(point_lights, dtype=np.float32).reshape(-1, 6)]
        k = self._keep
        self._h = vp(lib.or_light_sampler_create(_p(k[0]), _p(k[1]), _p(k[2]), k[1].shape[0], _p(k[3]), _p(k[4]) if k[4].size else None, k[4].shape[0]))

    def __del__(self):
        if getattr(self, "_h", None):
            lib.or_light_sampler_destroy(self._h)
            self._h = None

    def seed(self, s):
        lib.or_light_sampler_seed(self._h, s)

    def num_light_triangles(self):
        return int(lib.or_light_sampler_num_light_triangles(self._h))

    def generate(self, n, position_bias=1.0e-4):
        out = np.zeros(n, dtype=LIGHT_SAMPLE_DTYPE)
        lib.or_light_sampler_generate(self._h, _p(out), n, float(position_bias))
        return out


def mse(accum_rgba, ref_rgb, divisor):
    a = np.ascontiguousarray(accum_rgba, dtype=np.float32)
    r = np.ascontiguousarray(ref_rgb, dtype=np.float32)
    return float(lib.or_mse(_p(a), _p(r), a.shape[1], a.shape[0], int(divisor)))
