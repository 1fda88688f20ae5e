"""ctypes binding of oracle/_ref/libref_glsl_*.so -- TEST INFRASTRUCTURE: the reference's own GLSL shader source compiled as C++
(oracle/ref_glsl/, recipe in oracle/Makefile).  Present where /root/reference is (this container) and on the GPU box (prebuilt, shipped)."""
import ctypes as C
import os

import numpy as np

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
TRACE_LIB = os.path.join(_REF, "libref_glsl_trace.so")
PT_LIB = os.path.join(_REF, "libref_glsl_pt.so")


def available():
    return os.path.exists(TRACE_LIB) and os.path.exists(PT_LIB)


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

    def __init__(self):
        super().__init__(PT_LIB)
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
