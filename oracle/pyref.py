"""ctypes view of oracle/_ref/libref_bim.so -- the REFERENCE'S OWN dependencies/bim + dependencies/epsilon, compiled where they lie
(oracle/Makefile).  TEST INFRASTRUCTURE ONLY: used by tests/ and tests/golden/make_golden_bim.py to pin SURVEY 8 row a-15 and the
epsilon intersection KATs against the reference's object code.  The library exists where /root/reference does (this container) and
travels prebuilt to the GPU box; `available()` says whether it is here.  The product package never imports this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_bim.so")

NODE_DTYPE = np.dtype([("first_child", "<u4"), ("escape", "<u4")])
OBOX_DTYPE = np.dtype([("center", "<f4", 3), ("half_sides", "<f4", 3), ("orientation", "<f4", 4)])

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        vp, u32 = C.c_void_p, C.c_uint32
        _lib.ref_bim_build.restype = vp
        _lib.ref_bim_build.argtypes = [vp, vp, vp, u32, vp, vp, u32, C.c_char_p, C.POINTER(C.c_char_p), u32, C.c_int, C.c_int]
        _lib.ref_bim_store.restype = C.c_int
        _lib.ref_bim_store.argtypes = [vp, C.c_char_p, C.c_char_p]
        _lib.ref_bim_load.restype = vp
        _lib.ref_bim_load.argtypes = [C.c_char_p, C.c_int]
        _lib.ref_bim_free.restype = None
        _lib.ref_bim_free.argtypes = [vp]
        _lib.ref_bim_counts.restype = None
        _lib.ref_bim_counts.argtypes = [vp, vp]
        _lib.ref_bim_get.restype = C.c_uint64
        _lib.ref_bim_get.argtypes = [vp, C.c_int, vp]
        _lib.ref_bim_material_name.restype = C.c_char_p
        _lib.ref_bim_material_name.argtypes = [vp, u32]
        _lib.ref_ei_ray_box.restype = C.c_int
        _lib.ref_ei_ray_box.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _lib.ref_ei_ray_triangle.restype = C.c_int
        _lib.ref_ei_ray_triangle.argtypes = [vp, vp, vp, vp, vp, C.POINTER(C.c_float), vp]
        _lib.ref_ei_camera_params.restype = None
        _lib.ref_ei_camera_params.argtypes = [vp, vp, vp, C.c_float, C.c_float, vp, vp, vp]
        _lib.ref_ei_light_sample_point.restype = None
        _lib.ref_ei_light_sample_point.argtypes = [vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float, vp, vp]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefModel:
    """A bim::BinaryModel held by the reference's code."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference bim: load / build failed")
        self._h = C.c_void_p(handle)

    @classmethod
    def build(cls, positions, normals, texcoords, triangles, tri_materials, material_names, method="sah", obox=False, env_json=None):
        """tobim's call sequence on caller-provided arrays (ref_bim_shim.cpp)."""
        pos = np.ascontiguousarray(positions, dtype=np.float32)
        nrm = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32)
        uv = None if texcoords is None else np.ascontiguousarray(texcoords, dtype=np.float32)
        tri = np.ascontiguousarray(triangles, dtype=np.uint32)
        mat = None if tri_materials is None else np.ascontiguousarray(tri_materials, dtype=np.uint32)
        names = (C.c_char_p * len(material_names))(*[n.encode() for n in material_names])
        h = lib().ref_bim_build(_p(pos), _p(nrm), _p(uv), pos.shape[0], _p(tri), _p(mat), tri.shape[0], os.fsencode(env_json) if env_json else None, names,
                                len(material_names), {"sah": 0, "kd": 1}[method], 1 if obox else 0)
        return cls(h)

    @classmethod
    def load(cls, json_path, obox=False):
        return cls(lib().ref_bim_load(os.fsencode(json_path), 1 if obox else 0))

    def store(self, json_path, bim_path):
        lib().ref_bim_store(self._h, os.fsencode(json_path), os.fsencode(bim_path))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_bim_free(self._h)
            self._h = None

    def counts(self):
        c = np.zeros(10, dtype=np.uint32)
        lib().ref_bim_counts(self._h, _p(c))
        keys = ["vertices", "triangles", "nodes", "leaf_slots", "tree_levels", "materials", "has_normals", "has_texcoords", "has_oboxes", "accelerator"]
        return dict(zip(keys, [int(v) for v in c]))

    def _get(self, which, dtype, shape_tail=()):
        n = int(lib().ref_bim_get(self._h, which, None))
        buf = np.zeros(n, dtype=np.uint8)
        if n:
            lib().ref_bim_get(self._h, which, _p(buf))
        a = buf.view(dtype)
        return a.reshape((-1,) + shape_tail) if shape_tail else a

    def arrays(self):
        """Every array the reference's Scene uploads (scene.cpp:75-130), as the reference's chunk holds them."""
        c = self.counts()
        out = {"positions": self._get(0, np.float32, (3,)), "triangles": self._get(3, np.uint32, (3,)), "triangle_materials": self._get(4, np.uint32),
               "hierarchy": self._get(5, NODE_DTYPE), "parents": self._get(6, np.uint32), "leaves": self._get(7, np.uint32, (4,)),
               "aaboxes": self._get(8, np.float32, (6,)), "bounding_box": self._get(10, np.float32), "tree_levels": np.uint32(c["tree_levels"])}
        if c["has_normals"]:
            out["normals"] = self._get(1, np.float32, (3,))
        if c["has_texcoords"]:
            out["texcoords"] = self._get(2, np.float32, (2,))
        return out

    def oboxes(self):
        return self._get(9, OBOX_DTYPE)

    def material_names(self):
        return [lib().ref_bim_material_name(self._h, i).decode() for i in range(self.counts()["materials"])]


def _f3(v):
    return np.ascontiguousarray(v, dtype=np.float32)


def ei_ray_box(o, d, bmin, bmax):
    """(ei::intersects(Ray, Box), distance, distanceExit, ei::intersects(Ray, Box, distance, distanceExit)) -- dependencies/epsilon/src/3dintersection.cpp:326-349, 467-"""
    o, d, bmin, bmax = _f3(o), _f3(d), _f3(bmin), _f3(bmax)
    t0, t1 = C.c_float(), C.c_float()
    hit = lib().ref_ei_ray_box(_p(o), _p(d), _p(bmin), _p(bmax), C.byref(t0), C.byref(t1))
    return bool(hit & 1), t0.value, t1.value, bool(hit & 2)


def ei_ray_triangle(o, d, p0, p1, p2):
    """ei::intersects(Ray, Triangle, distance, barycentric)"""
    o, d, p0, p1, p2 = _f3(o), _f3(d), _f3(p0), _f3(p1), _f3(p2)
    t = C.c_float()
    bary = np.zeros(3, dtype=np.float32)
    hit = lib().ref_ei_ray_triangle(_p(o), _p(d), _p(p0), _p(p1), _p(p2), C.byref(t), _p(bary))
    return bool(hit), t.value, bary


def ei_camera_params(pos, lookat, up, yfov, aspect):
    """Camera::ComputeCameraParams (camera.cpp:76-85) on epsilon's own normalize / cross -> (U, V, W)"""
    u, v, w = (np.zeros(3, dtype=np.float32) for _ in range(3))
    lib().ref_ei_camera_params(_p(_f3(pos)), _p(_f3(lookat)), _p(_f3(up)), float(yfov), float(aspect), _p(u), _p(v), _p(w))
    return u, v, w


def ei_light_sample_point(v0, v1, v2, alpha, beta, gamma, bias):
    """position and normal of an area-light sample (lightsampler.cpp:62-69) on epsilon's own vector arithmetic"""
    p, n = np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.float32)
    lib().ref_ei_light_sample_point(_p(_f3(v0)), _p(_f3(v1)), _p(_f3(v2)), float(alpha), float(beta), float(gamma), float(bias), _p(p), _p(n))
    return p, n


# ---- bvhmake's build methods (oracle/_ref/libref_bvhmake.so: buildmethods/kdtree.cpp, sweep.cpp + fitmethods/aaboxfit.cpp of the reference)
BVHMAKE_LIB_PATH = os.path.join(_HERE, "_ref", "libref_bvhmake.so")
_bvhmake = None


def bvhmake_available():
    return os.path.exists(BVHMAKE_LIB_PATH)


def bvhmake_build(positions, triangles, tri_materials, method="kdtree"):
    """BVHBuilder::BuildBVH with the reference's own build method -> (nodes {left, right} in pool order, boxes (n, 6), leaves (m, 8, 4) uint32)."""
    global _bvhmake
    if _bvhmake is None:
        _bvhmake = C.CDLL(BVHMAKE_LIB_PATH)
        _bvhmake.ref_bvhmake_build.restype = C.c_int
        _bvhmake.ref_bvhmake_build.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    pos = np.ascontiguousarray(positions, dtype=np.float32)
    tri = np.ascontiguousarray(triangles, dtype=np.uint32)
    mat = np.ascontiguousarray(tri_materials, dtype=np.uint32)
    t = tri.shape[0]
    cap_nodes, cap_leaves = max(8, 4 * t // 8), max(8, 2 * t // 8)
    lr = np.zeros((cap_nodes, 2), dtype=np.uint32)
    boxes = np.zeros((cap_nodes, 6), dtype=np.float32)
    leaves = np.zeros((cap_leaves, 8, 4), dtype=np.uint32)
    n, m = C.c_uint32(), C.c_uint32()
    st = _bvhmake.ref_bvhmake_build(_p(pos), pos.shape[0], _p(tri), _p(mat), t, {"kdtree": 0, "sweep": 1}[method], _p(lr), _p(boxes), _p(leaves), C.byref(n), C.byref(m))
    if st != 0:
        raise RuntimeError("reference bvhmake build failed")
    return lr[: n.value], boxes[: n.value], leaves[: m.value]


# ---- stb_image as vendored by the reference (oracle/_ref/libref_stb.so)
STB_LIB_PATH = os.path.join(_HERE, "_ref", "libref_stb.so")
_stb = None


def stb_available():
    return os.path.exists(STB_LIB_PATH)


def stbi_load(path):
    """stbi_load(path, &w, &h, &n, 4) -> ((h, w, 4) uint8, n) or None."""
    global _stb
    if _stb is None:
        _stb = C.CDLL(STB_LIB_PATH)
        _stb.ref_stbi_load.restype = C.c_int
        _stb.ref_stbi_load.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    w, h, n = C.c_int(), C.c_int(), C.c_int()
    if _stb.ref_stbi_load(os.fsencode(path), None, 0, C.byref(w), C.byref(h), C.byref(n)) != 0:
        return None
    out = np.zeros((h.value, w.value, 4), dtype=np.uint8)
    _stb.ref_stbi_load(os.fsencode(path), _p(out), w.value * h.value, C.byref(w), C.byref(h), C.byref(n))
    return out, n.value
