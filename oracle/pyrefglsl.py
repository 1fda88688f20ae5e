"""ctypes binding of oracle/_ref/libref_glsl_*.so -- TEST INFRASTRUCTURE: the reference's own GLSL shader source compiled as C++
(oracle/ref_glsl/, recipe in oracle/Makefile).  Present where /root/reference is (this container) and on the GPU box (prebuilt, shipped).
Used by tests/ (the pin of the oracle and of the CUDA path) and by bench.py's cpu_baseline / --impl reference legs (kind "reference")."""
import ctypes as C
import os

import numpy as np

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
TRACE_LIB = os.path.join(_REF, "libref_glsl_trace.so")
PT_LIB = os.path.join(_REF, "libref_glsl_pt.so")


LPT_LIB = os.path.join(_REF, "libref_glsl_lpt.so")
LPT_SAVE_LIB = os.path.join(_REF, "libref_glsl_lpt_save.so")
SHOW_LIB = os.path.join(_REF, "libref_glsl_show.so")
BPT_LIB = os.path.join(_REF, "libref_glsl_bpt.so")
WHITTED_LIB = os.path.join(_REF, "libref_glsl_whitted.so")


def available():
    names = ["lpt_f1", "lpt_f2", "lpt_f3", "lpt_save_cmis", "bpt_cmis"]
    return all(os.path.exists(p) for p in [TRACE_LIB, PT_LIB, LPT_LIB, LPT_SAVE_LIB, SHOW_LIB, BPT_LIB, WHITTED_LIB] + [os.path.join(_REF, "libref_glsl_%s.so" % n) for n in names])


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _Lib:
    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.refglsl_set_scene.restype = C.c_int
        self.lib.refglsl_set_scene.argtypes = [C.c_void_p]
        self.lib.refglsl_max_path_length.restype = C.c_int
        self.scene = None

    def set_scene(self, scene):
        """scene: gpugi_b200.Scene (kept alive: the shader's buffers point into its arrays)"""
        self.scene = scene
        self._desc = scene.desc   # (a fresh structure per access: keep THIS one alive; the arrays belong to the scene)
        st = self.lib.refglsl_set_scene(C.byref(self._desc))
        if st != 0:
            raise RuntimeError("refglsl_set_scene failed: %d" % st)


class Trace(_Lib):
    """TraceRay / TraceRayAnyHit of gpugi/shader/traceray.glsl (HIT_INDEX_OUTPUT) on GiRay batches."""

    def __init__(self):
        super().__init__(TRACE_LIB)
        self.lib.refglsl_trace.restype = C.c_int
        self.lib.refglsl_trace.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]

    def trace(self, g, rays):
        hits = np.zeros(rays.shape[0], dtype=g.HIT_DTYPE)
        st = self.lib.refglsl_trace(_p(rays), rays.shape[0], _p(hits))
        if st != 0:
            raise RuntimeError("refglsl_trace failed: %d" % st)
        return hits


class PathTracer(_Lib):
    """main() of gpugi/shader/pathtracer.comp, one invocation per pixel of a region; MAX_PATHLENGTH is compiled in (16)."""

    def __init__(self, lib=None):
        super().__init__(lib or PT_LIB)
        self.lib.refglsl_render_pt.restype = C.c_int
        self.lib.refglsl_render_pt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        self.max_path_length = self.lib.refglsl_max_path_length()

    def render(self, cam, width, height, samples, frame_seed, accum=None, region=None):
        """cam: the 13 floats (GiCamera layout) the oracle's camera() returns; samples: GiLightSample array; accum: (H, W, 4) float32 sum image."""
        if accum is None:
            accum = np.zeros((height, width, 4), dtype=np.float32)
        x0, y0, x1, y1 = region if region else (0, 0, width, height)
        samples = np.ascontiguousarray(samples)
        cam = np.ascontiguousarray(cam, dtype=np.float32)   # 13 floats in GiCamera layout (oracle.pyoracle.camera)
        st = self.lib.refglsl_render_pt(_p(cam), width, height, _p(samples), samples.shape[0], frame_seed, _p(accum), x0, y0, x1, y1)
        if st != 0:
            raise RuntimeError("refglsl_render_pt failed: %d" % st)
        return accum


class WhittedRayTracer(PathTracer):
    """main() of gpugi/shader/whittedraytracer.comp (STOP_ON_DIFFUSE_BOUNCE, MAX_PATHLENGTH 16)."""

    def __init__(self):
        super().__init__(WHITTED_LIB)


class LightTracer(_Lib):
    """main() of gpugi/shader/lighttracer.comp, invocations in ascending order on one thread; save=True: the SAVE_LIGHT_CACHE build."""

    def __init__(self, save=False, reconstruction_filter=0, chained_mis=False):
        name = "lpt_save_cmis" if chained_mis else "lpt_save" if save else ("lpt_f%d" % reconstruction_filter if reconstruction_filter else "lpt")
        super().__init__(os.path.join(_REF, "libref_glsl_%s.so" % name))
        self.save = save or chained_mis
        self.lib.refglsl_render_lpt.restype = C.c_int
        self.lib.refglsl_render_lpt.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_uint32, C.c_uint32,
                                                C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        self.lib.refglsl_light_cache_entry_size.restype = C.c_int
        if save:
            assert self.lib.refglsl_light_cache_entry_size() == 64

    def render(self, cam, width, height, samples, frame_seed, num_rays_per_light_sample, accum=None, first=0, count=None, capacity=0, parallel=False):
        """parallel: all host threads (throughput runs; the order of float atomics and light-cache slots is then scheduling dependent)
        -> accum, or (accum, cache (capacity x 16 floats), NumLightCacheEntries, LightPathLengthSum) for the SAVE_LIGHT_CACHE build"""
        if accum is None:
            accum = np.zeros((height, width, 4), dtype=np.float32)
        samples = np.ascontiguousarray(samples)
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        total = samples.shape[0] * num_rays_per_light_sample // 64 * 64   # whole work groups of 64 (lightpathtracer.cpp:48)
        if count is None:
            count = total - first
        cache = np.zeros((max(capacity, 1), 16), dtype=np.float32)
        counters = np.array([samples.shape[0] * num_rays_per_light_sample, 0], dtype=np.int32)   # resetcounter.comp
        st = self.lib.refglsl_render_lpt(_p(cam), width, height, _p(samples), samples.shape[0], frame_seed, num_rays_per_light_sample, _p(accum), first, count,
                                         _p(cache), capacity, _p(counters), 1 if parallel else 0)
        if st != 0:
            raise RuntimeError("refglsl_render_lpt failed: %d" % st)
        if self.save:
            return accum, cache, int(counters[0]), int(counters[1])
        return accum


class ShowLightCaches(_Lib):
    """main() of gpugi/shader/showlightcaches.comp, one invocation per initial light sample."""

    def __init__(self):
        super().__init__(SHOW_LIB)
        self.lib.refglsl_show_light_caches.restype = C.c_int
        self.lib.refglsl_show_light_caches.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]

    def render(self, cam, width, height, samples, accum):
        samples = np.ascontiguousarray(samples)
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        st = self.lib.refglsl_show_light_caches(_p(cam), width, height, _p(samples), samples.shape[0], _p(accum))
        if st != 0:
            raise RuntimeError("refglsl_show_light_caches failed: %d" % st)
        return accum


class BidirEyePass(_Lib):
    """main() of gpugi/shader/bidirpathtracer/pathtracer_bidir.comp over a pixel region, reading a given light cache."""

    def __init__(self, chained_mis=False):
        super().__init__(os.path.join(_REF, "libref_glsl_bpt_cmis.so") if chained_mis else BPT_LIB)
        self.lib.refglsl_render_bpt_eye.restype = C.c_int
        self.lib.refglsl_render_bpt_eye.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                    C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]

    def render(self, cam, width, height, samples, frame_seed, num_rays_per_light_sample, cache, capacity, num_entries, accum=None, region=None,
               average_light_path_length=0):
        if accum is None:
            accum = np.zeros((height, width, 4), dtype=np.float32)
        x0, y0, x1, y1 = region if region else (0, 0, width, height)
        samples = np.ascontiguousarray(samples)
        cam = np.ascontiguousarray(cam, dtype=np.float32)
        cache = np.ascontiguousarray(cache, dtype=np.float32)
        st = self.lib.refglsl_render_bpt_eye(_p(cam), width, height, _p(samples), samples.shape[0], frame_seed, num_rays_per_light_sample, _p(cache), capacity,
                                             num_entries, average_light_path_length, _p(accum), x0, y0, x1, y1)
        if st != 0:
            raise RuntimeError("refglsl_render_bpt_eye failed: %d" % st)
        return accum
