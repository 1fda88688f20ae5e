"""ctypes view of tests/emul/libgi_emul.so (host build of the device logic) -- tests only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GI_EMUL_VARIANT=bin2 selects the build with binary inner nodes (WF_WIDE4=0) -- same tests, other build switch
LIB_PATH = os.path.join(_HERE, "libgi_emul_%s.so" % os.environ["GI_EMUL_VARIANT"] if os.environ.get("GI_EMUL_VARIANT") else "libgi_emul.so")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


build()
lib = C.CDLL(LIB_PATH)
vp, u32 = C.c_void_p, C.c_uint32

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("t_max", "<f4"), ("direction", "<f4", 3), ("flags", "<u4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("leaf_tri", "<i4"), ("bary_y", "<f4"), ("bary_z", "<f4")])
LIGHT_SAMPLE_DTYPE = np.dtype([("position", "<f4", 3), ("normal_phi", "<f4"), ("intensity", "<f4", 3), ("normal_theta_cos", "<f4")])


class EmCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ["closest_rays", "any_rays", "camera_paths", "light_paths", "shaded_vertices", "splats", "fallback_rays",
                                          "node_tests", "tri_tests"]] + [("light_cache_entries", u32), ("light_path_length_sum", u32)]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class IterationParams(C.Structure):
    _fields_ = [("renderer", C.c_int32), ("frame_seed", u32), ("max_path_length", C.c_int32), ("num_rays_per_light_sample", C.c_int32),
                ("light_cache_capacity", C.c_int32), ("warmup", C.c_int32), ("reserved", C.c_int32 * 2)]


lib.em_trace.restype = C.c_int
lib.em_trace.argtypes = [vp, vp, u32, vp, C.c_int, C.POINTER(C.c_int), vp]
lib.em_render.restype = C.c_int
lib.em_render.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp]
lib.em_render_wavefront.restype = C.c_int
lib.em_render_wavefront.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp]
lib.em_bidir_eye.restype = C.c_int
lib.em_bidir_eye.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, u32, vp, vp]
lib.em_slab_check.restype = C.c_int
lib.em_slab_check.argtypes = [vp, vp, u32, u32, vp]


lib.em_set_variant.restype = None
lib.em_set_variant.argtypes = [C.c_int, C.c_int]
lib.em_set_topology.restype = None
lib.em_set_topology.argtypes = [C.c_int, u32]
lib.em_topology_applied.restype = C.c_int
lib.em_topology_applied.argtypes = []
TOPOLOGY_REFERENCE, TOPOLOGY_MEDIAN, TOPOLOGY_RANDOM = 0, 1, 2


def set_variant(reconstruction_filter=0, chained_mis=False):
    """GI_OPT_RECONSTRUCTION_FILTER / GI_OPT_CHAINED_MIS of every following render call (the one-thread-per-path forms carry CHAINED_MIS)."""
    lib.em_set_variant(int(reconstruction_filter), 1 if chained_mis else 0)


def set_topology(kind, seed=1):
    """Tree the ordered walks of the following calls run on (gi_layout.h applyLeafTree): the reference's, a median split of its leaves,
    or a random balanced pairing of them."""
    lib.em_set_topology(int(kind), int(seed))


lib.em_set_tree.restype = None
lib.em_set_tree.argtypes = [u32, vp, vp, vp, u32]


def set_tree(left, right, leaf_of, root):
    """Ordered walks of the following calls run on this tree over the reference's leaves (node ids: leaf positions 0..L-1, inner L..2L-2)."""
    left = np.ascontiguousarray(left, dtype=np.uint32); right = np.ascontiguousarray(right, dtype=np.uint32)
    leaf_of = np.ascontiguousarray(leaf_of, dtype=np.uint32)
    lib.em_set_tree(len(leaf_of), _p(left), _p(right), _p(leaf_of), int(root))


def topology_applied():
    return bool(lib.em_topology_applied())


def _p(a):
    return a.ctypes.data_as(vp)


def trace(scene_desc, rays, mode=0):
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
    cnt = EmCounters()
    fast = C.c_int()
    st = lib.em_trace(C.addressof(scene_desc), _p(rays), rays.shape[0], _p(hits), mode, C.byref(fast), C.addressof(cnt))
    if st != 0:
        raise RuntimeError("em_trace rejected the scene")
    d = cnt.as_dict()
    d["fast_valid"] = fast.value
    return hits, d


def render(scene_desc, cam, w, h, light_samples, renderer, frame_seed, max_len=16, num_rays_per_light_sample=0, capacity=0, warmup=False, mode=0,
           env_rgb=None, show_light_caches=False, accum=None, cache=None):
    ls = np.ascontiguousarray(light_samples, dtype=LIGHT_SAMPLE_DTYPE)
    if accum is None:
        accum = np.zeros((h, w, 4), dtype=np.float32)
    p = IterationParams(renderer=renderer, frame_seed=frame_seed & 0xFFFFFFFF, max_path_length=max_len, num_rays_per_light_sample=num_rays_per_light_sample,
                        light_cache_capacity=capacity, warmup=1 if warmup else 0)
    if cache is None:
        cache = np.zeros(max(int(capacity), 1) * 16, dtype=np.float32)
    cnt = EmCounters()
    env = np.ascontiguousarray(env_rgb, dtype=np.float32) if env_rgb is not None else None
    camp = _p(cam) if isinstance(cam, np.ndarray) else C.addressof(cam)
    st = lib.em_render(C.addressof(scene_desc), camp, w, h, _p(ls), ls.shape[0], C.addressof(p), mode, _p(env) if env is not None else None,
                       1 if show_light_caches else 0, _p(cache), _p(accum), C.addressof(cnt))
    if st != 0:
        raise RuntimeError("em_render failed: %d" % st)
    return accum, cnt.as_dict()


def slab_check(scene_desc, rays, stride=64):
    """Conservativeness of the quantised inner-node test vs the exact-zone box test (see em_slab_check)."""
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    out = np.zeros(5, dtype=np.uint64)
    st = lib.em_slab_check(C.addressof(scene_desc), _p(rays), rays.shape[0], stride, _p(out))
    if st != 0:
        raise RuntimeError("em_slab_check failed: %d" % st)
    return dict(zip(["tested", "exact_accepts", "conservative_accepts", "violations", "out_of_domain"], (int(x) for x in out)))


def render_wavefront(scene_desc, cam, w, h, light_samples, renderer, frame_seed, max_len=16, num_rays_per_light_sample=0, capacity=0, warmup=False, mode=0,
                     phase=0, cache=None, cache_count=0, accum=None):
    """The wavefront schedule of LPT (renderer 1) / BPT (renderer 2) run sequentially with the product's lane functions (gi_shade.h).
    phase 0: whole iteration, 1: BPT light pass only, 2: BPT eye pass only on (cache, cache_count).  Returns (accum, counters, cache, count)."""
    ls = np.ascontiguousarray(light_samples, dtype=LIGHT_SAMPLE_DTYPE)
    if accum is None:
        accum = np.zeros((h, w, 4), dtype=np.float32)
    p = IterationParams(renderer=renderer, frame_seed=frame_seed & 0xFFFFFFFF, max_path_length=max_len, num_rays_per_light_sample=num_rays_per_light_sample,
                        light_cache_capacity=capacity, warmup=1 if warmup else 0)
    if cache is None:
        cache = np.zeros((max(int(capacity), 1), 16), dtype=np.float32)
    cnt = EmCounters()
    count = u32(cache_count)
    camp = _p(cam) if isinstance(cam, np.ndarray) else C.addressof(cam)
    st = lib.em_render_wavefront(C.addressof(scene_desc), camp, w, h, _p(ls), ls.shape[0], C.addressof(p), mode, phase, _p(cache), C.addressof(count), _p(accum),
                                 C.addressof(cnt))
    if st != 0:
        raise RuntimeError("em_render_wavefront failed: %d" % st)
    return accum, cnt.as_dict(), cache, int(count.value)


def bidir_eye(scene_desc, cam, w, h, frame_seed, cache, cache_count, capacity, max_len=16, num_rays_per_light_sample=64, mode=0):
    """k_bidir_eye (megakernel form) alone on a given light cache."""
    accum = np.zeros((h, w, 4), dtype=np.float32)
    p = IterationParams(renderer=2, frame_seed=frame_seed & 0xFFFFFFFF, max_path_length=max_len, num_rays_per_light_sample=num_rays_per_light_sample,
                        light_cache_capacity=capacity, warmup=0)
    cnt = EmCounters()
    camp = _p(cam) if isinstance(cam, np.ndarray) else C.addressof(cam)
    st = lib.em_bidir_eye(C.addressof(scene_desc), camp, w, h, C.addressof(p), mode, _p(cache), cache_count, _p(accum), C.addressof(cnt))
    if st != 0:
        raise RuntimeError("em_bidir_eye failed: %d" % st)
    return accum, cnt.as_dict()
