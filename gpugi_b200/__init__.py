"""gpugi_b200 -- B200-native light-transport hot path behind gpugi's Renderer API.

Python is only glue here: these classes are thin ctypes views of
  * the C ABI of include/gpugi_b200.h            (class Device)
  * the C++ host layer mirroring the reference   (classes Scene, RendererSystem, LightSampler)
The light transport itself runs in the CUDA kernels of csrc/gi_device.cu; there is no Python or
CPU implementation of it in this package.
"""
import ctypes as C
import json
import os

import numpy as np

from . import _capi as capi
from ._capi import (GI_RENDERER_PT, GI_RENDERER_LPT, GI_RENDERER_BPT, GI_RENDERER_WHITTED, GI_TRAVERSAL_FAST, GI_TRAVERSAL_REFERENCE,
                    GI_OPT_TRAVERSAL, GI_OPT_TEXTURE_FILTER, GI_OPT_SCHEDULER, GI_OPT_WF_REFILL_THRESHOLD, GI_OPT_WF_MIN_INNER,
                    GI_OPT_WF_BLOCKS_PER_SM, GI_OPT_WF_SHADE_BLOCKS_PER_SM, GI_OPT_RECONSTRUCTION_FILTER, GI_OPT_CHAINED_MIS, GI_OPT_COUNT_STEPS, GI_OPT_KERNEL_TIMING, GI_OPT_STREAMING, GI_OPT_DEVICE_BUILDER, GI_BUILDER_LBVH, GI_BUILDER_PLOC, GI_OPT_UPLOAD_TOPOLOGY, GI_OPT_PLOC_RADIUS, GI_TOPOLOGY_REFERENCE, GI_TOPOLOGY_PLOC, GI_SCHEDULER_WAVEFRONT, GI_SCHEDULER_MEGAKERNEL, RAY_MAX)

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("t_max", "<f4"), ("direction", "<f4", 3), ("flags", "<u4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("leaf_tri", "<i4"), ("bary_y", "<f4"), ("bary_z", "<f4")])
LIGHT_SAMPLE_DTYPE = np.dtype([("position", "<f4", 3), ("normal_phi", "<f4"), ("intensity", "<f4", 3), ("normal_theta_cos", "<f4")])
NODE_DTYPE = np.dtype([("min", "<f4", 3), ("first_child", "<u4"), ("max", "<f4", 3), ("escape", "<u4")])
TRIANGLE_DTYPE = np.dtype([("v", "<u4", 3), ("material", "<u4")])

RENDERERS = {"PT": GI_RENDERER_PT, "LPT": GI_RENDERER_LPT, "BPT": GI_RENDERER_BPT, "WPT": GI_RENDERER_WHITTED}


class GpugiError(RuntimeError):
    pass


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def wang_hash(x):
    """utilities/random.cpp:5-13"""
    return int(capi.host.gih_wanghash(C.c_uint32(x & 0xFFFFFFFF)))


def set_path_length(n):
    """GlobalConfig parameter "pathLength" (application.cpp:123); latched by a renderer in SetScene."""
    capi.host.gih_set_path_length(int(n))


def camera_params(pos, lookat, yfov, width, height, up=(0.0, 1.0, 0.0)):
    """Camera::ComputeCameraParams + PixelArea (camera.cpp:76-85, renderersystem.cpp:208) -> GiCamera."""
    out = capi.GiCamera()
    capi.host.gih_camera_params(_f3(pos), _f3(lookat), _f3(up), float(yfov), float(width) / float(height), int(width), int(height), C.byref(out))
    return out


# ---------------------------------------------------------------------------------------------------
class Scene:
    """Host mirror of gpugi's Scene (gpugi/scene/scene.hpp)."""

    def __init__(self, handle):
        if not handle:
            raise GpugiError("Scene: " + capi.host.gih_last_error().decode())
        self._h = C.c_void_p(handle)

    @classmethod
    def load(cls, json_path, bvh_type=0):
        """Scene(file, bvhType): JSON environment file + .bim (scene.cpp:15-59)."""
        return cls(capi.host.gih_scene_load(os.fsencode(json_path), int(bvh_type)))

    @classmethod
    def from_mesh(cls, positions, normals, texcoords, triangles, tri_materials, materials, texture_dir="", build_method="sah", bvh_type=0):
        """Synthetic scene: mesh arrays + material dict (JSON 'materials' object, order = material index).
        bvh_type 1: oriented boxes (the reference's second accelerator, GlobalConfig "bvhType").
        build_method: "sah" (bim_sahspilt.cpp sweep, host), "kd" (bim_kdtree.cpp median, host), "lbvh" / "ploc" (gi_build_hierarchy: on the GPU, Karras radix tree / PLOC clustering)."""
        positions = np.ascontiguousarray(positions, dtype=np.float32)
        normals = np.ascontiguousarray(normals, dtype=np.float32)
        texcoords = np.ascontiguousarray(texcoords, dtype=np.float32)
        triangles = np.ascontiguousarray(triangles, dtype=np.uint32)
        tri_materials = np.ascontiguousarray(tri_materials, dtype=np.uint32)
        assert positions.shape == normals.shape and positions.shape[1] == 3 and texcoords.shape == (positions.shape[0], 2)
        assert triangles.shape[1] == 3 and tri_materials.shape[0] == triangles.shape[0]
        env = json.dumps({"scene": "", "accelerator": "obox" if bvh_type else "aabox", "materials": materials, "lights": {}, "cameras": {}})
        h = capi.host.gih_scene_from_mesh_bvh(_vp(positions), _vp(normals), _vp(texcoords), positions.shape[0], _vp(triangles), _vp(tri_materials),
                                              triangles.shape[0], env.encode(), os.fsencode(texture_dir), {"sah": 0, "kd": 1, "lbvh": 2, "ploc": 3}[build_method], int(bvh_type))
        if not h:
            raise GpugiError("Scene.from_mesh: " + capi.host.gih_last_error().decode())
        return cls(h)

    def __del__(self):
        if getattr(self, "_h", None):
            capi.host.gih_scene_destroy(self._h)
            self._h = None

    def store(self, json_path, bim_name):
        """Write <json_path> and the .bim next to it (BinaryModel::store*)."""
        if capi.host.gih_scene_store(self._h, os.fsencode(json_path), os.fsencode(bim_name)) != 0:
            raise GpugiError(capi.host.gih_last_error().decode())

    def add_point_light(self, position, intensity):
        """Scene::AddPointLight ("addPointLight" command, application.cpp:158-163)."""
        capi.host.gih_scene_add_point_light(self._h, _f3(position), _f3(intensity))

    @property
    def desc(self):
        d = capi.GiSceneDesc()
        capi.host.gih_scene_desc(self._h, C.byref(d))
        return d

    @property
    def info(self):
        i = capi.GihSceneInfo()
        capi.host.gih_scene_info(self._h, C.byref(i))
        return i

    # numpy views (copies) of the arrays the reference would upload
    def nodes(self):
        d = self.desc
        return np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_uint8)), shape=(d.num_nodes * 32,)).view(NODE_DTYPE).copy()

    def obox_nodes(self):
        """TreeNode<ei::OBox> array (bvh_type 1), else None."""
        d = self.desc
        if not d.obox_nodes:
            return None
        dt = np.dtype([("center", "<f4", 3), ("first_child", "<u4"), ("sides_half", "<f4", 3), ("escape", "<u4"), ("rotation_inv", "<f4", 4)])
        return np.ctypeslib.as_array(C.cast(d.obox_nodes, C.POINTER(C.c_uint8)), shape=(d.num_nodes * 48,)).view(dt).copy()

    def leaf_triangles(self):
        d = self.desc
        return np.ctypeslib.as_array(C.cast(d.leaf_triangles, C.POINTER(C.c_uint8)), shape=(d.num_leaf_triangles * 16,)).view(TRIANGLE_DTYPE).copy()

    def positions(self):
        d = self.desc
        return np.ctypeslib.as_array(d.positions, shape=(d.num_vertices, 3)).copy()

    def vertex_info(self):
        d = self.desc
        return np.ctypeslib.as_array(C.cast(d.vertex_info, C.POINTER(C.c_float)), shape=(d.num_vertices, 4)).copy()

    def triangles(self):
        i = self.info
        return np.ctypeslib.as_array(i.triangles, shape=(i.num_triangles, 3)).copy()

    def triangle_materials(self):
        i = self.info
        return np.ctypeslib.as_array(i.triangle_materials, shape=(i.num_triangles,)).copy()

    def parents(self):
        i = self.info
        return np.ctypeslib.as_array(i.parents, shape=(i.num_nodes,)).copy()

    def emissivity(self):
        out = np.zeros((self.info.num_materials, 3), dtype=np.float32)
        capi.host.gih_scene_emissivity(self._h, _vp(out))
        return out

    def point_lights(self):
        out = np.zeros((self.info.num_point_lights, 6), dtype=np.float32)
        if out.size:
            capi.host.gih_scene_point_lights(self._h, _vp(out))
        return out


RAW_NODE_DTYPE = np.dtype([("parent", "<u4"), ("first_child", "<u4"), ("escape", "<u4")])
RAW_VERTEX_DTYPE = np.dtype([("position", "<f4", 3), ("normal", "<f4", 3), ("texcoord", "<f4", 2)])
RAW_LEAF_DTYPE = np.dtype([("triangles", [("v", "<u4", 3), ("material", "<u4")], 8)])


class RawScene:
    """bvhmake's `.rawscene` container + its kdtree / sweep / lds builders over 8-triangle leaves (host/bvhmake.hpp)."""

    def __init__(self, handle):
        if not handle:
            raise GpugiError("RawScene: " + capi.host.gih_last_error().decode())
        self._h = C.c_void_p(handle)

    @classmethod
    def build(cls, positions, normals, texcoords, triangles, tri_materials, material_names, method="kdtree"):
        positions = np.ascontiguousarray(positions, dtype=np.float32)
        normals = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32)
        texcoords = None if texcoords is None else np.ascontiguousarray(texcoords, dtype=np.float32)
        triangles = np.ascontiguousarray(triangles, dtype=np.uint32)
        tri_materials = None if tri_materials is None else np.ascontiguousarray(tri_materials, dtype=np.uint32)
        names = (C.c_char_p * len(material_names))(*[n.encode() for n in material_names])
        return cls(capi.host.gih_rawscene_build(_vp(positions), _vp(normals) if normals is not None else None, _vp(texcoords) if texcoords is not None else None,
                                                positions.shape[0], _vp(triangles), _vp(tri_materials) if tri_materials is not None else None, triangles.shape[0],
                                                names, len(material_names), {"kdtree": 0, "sweep": 1, "lds": 2}[method]))

    @classmethod
    def load(cls, path):
        return cls(capi.host.gih_rawscene_load(os.fsencode(path)))

    def store(self, path):
        if capi.host.gih_rawscene_store(self._h, os.fsencode(path)) != 0:
            raise GpugiError(capi.host.gih_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            capi.host.gih_rawscene_free(self._h)
            self._h = None

    def _get(self, which, dtype):
        n = int(capi.host.gih_rawscene_get(self._h, which, None))
        buf = np.zeros(n, dtype=np.uint8)
        if n:
            capi.host.gih_rawscene_get(self._h, which, _vp(buf))
        return buf.view(dtype)

    def material_names(self):
        return [bytes(r).split(b"\0")[0].decode() for r in self._get(0, np.uint8).reshape(-1, 32)]

    def vertices(self):
        return self._get(1, RAW_VERTEX_DTYPE)

    def boxes(self):
        return self._get(2, np.float32).reshape(-1, 6)

    def nodes(self):
        return self._get(3, RAW_NODE_DTYPE)

    def leaves(self):
        return self._get(4, RAW_LEAF_DTYPE)

    def to_scene(self, materials, texture_dir=""):
        """Scene with the file's hierarchy, every 8-triangle leaf expanded into 2-triangle leaves; materials: dict name -> definition."""
        env = json.dumps({"scene": "", "accelerator": "aabox", "materials": materials, "lights": {}, "cameras": {}})
        return Scene(capi.host.gih_scene_from_rawscene(self._h, env.encode(), os.fsencode(texture_dir)))


class LightSampler:
    """LightSampler (gpugi/scene/lightsampler.cpp) on its own, for host-logic tests."""

    def __init__(self, scene):
        self._scene = scene
        self._h = C.c_void_p(capi.host.gih_light_sampler_create(scene._h))

    def __del__(self):
        if getattr(self, "_h", None):
            capi.host.gih_light_sampler_destroy(self._h)
            self._h = None

    def seed(self, s):
        capi.host.gih_light_sampler_seed(self._h, C.c_uint64(s))

    def generate(self, n, position_bias=1.0e-4):
        out = np.zeros(n, dtype=LIGHT_SAMPLE_DTYPE)
        capi.host.gih_light_sampler_generate(self._h, _vp(out), n, float(position_bias))
        return out


# ---------------------------------------------------------------------------------------------------
class Device:
    """Raw view of a GiContext (include/gpugi_b200.h).  Owns it unless borrowed from a RendererSystem."""

    def __init__(self, device=0, _borrowed=None):
        self._owned = _borrowed is None
        if _borrowed is not None:
            self._c = C.c_void_p(_borrowed)
            return
        ctx = C.c_void_p()
        st = capi.dev.gi_create(int(device), C.byref(ctx))
        if st != capi.GI_OK:
            raise GpugiError("gi_create: " + capi.dev.gi_last_error(None).decode())
        self._c = ctx

    def close(self):
        if self._owned and getattr(self, "_c", None):
            capi.dev.gi_destroy(self._c)
        self._c = None

    __del__ = close

    def _check(self, st, what):
        if st != capi.GI_OK:
            raise GpugiError("%s failed (%d): %s" % (what, st, capi.dev.gi_last_error(self._c).decode()))

    def set_stream(self, cuda_stream_ptr):
        self._check(capi.dev.gi_set_stream(self._c, C.c_void_p(cuda_stream_ptr)), "gi_set_stream")

    def set_option(self, option, value):
        self._check(capi.dev.gi_set_option(self._c, option, value), "gi_set_option")

    def upload_scene(self, scene):
        d = scene.desc
        self._check(capi.dev.gi_upload_scene(self._c, C.byref(d)), "gi_upload_scene")

    def set_camera(self, cam):
        self._check(capi.dev.gi_set_camera(self._c, C.byref(cam)), "gi_set_camera")

    def set_screen_size(self, w, h):
        self._w, self._h = int(w), int(h)
        self._check(capi.dev.gi_set_screen_size(self._c, int(w), int(h)), "gi_set_screen_size")

    def set_light_samples(self, samples):
        samples = np.ascontiguousarray(samples, dtype=LIGHT_SAMPLE_DTYPE)
        self._check(capi.dev.gi_set_light_samples(self._c, C.cast(_vp(samples), C.POINTER(capi.GiLightSample)), samples.shape[0]), "gi_set_light_samples")

    def set_light_samples_device(self, d_ptr, n):
        """gi_set_light_samples_device: samples already resident in device memory (no copy)."""
        self._check(capi.dev.gi_set_light_samples_device(self._c, C.c_void_p(d_ptr), int(n)), "gi_set_light_samples_device")

    def generate_light_samples(self, rng_state, n, position_bias=1.0e-4):
        """gi_generate_light_samples: the next iteration's set drawn on the device (needs gi_set_light_tables, which
        RendererSystem.SetDeviceLightSampling uploads); returns the advanced xorshift64* state."""
        st = C.c_uint64(int(rng_state))
        self._check(capi.dev.gi_generate_light_samples(self._c, C.byref(st), int(n), float(position_bias)), "gi_generate_light_samples")
        return int(st.value)

    def get_light_samples(self, n):
        out = np.zeros(int(n), dtype=LIGHT_SAMPLE_DTYPE)
        self._check(capi.dev.gi_get_light_samples(self._c, C.cast(_vp(out), C.POINTER(capi.GiLightSample)), int(n)), "gi_get_light_samples")
        return out

    def get_light_cache(self, capacity):
        """gi_get_light_cache -> (entries as (capacity, 16) float32, NumLightCacheEntries counter)."""
        out = np.zeros((int(capacity), 16), dtype=np.float32)
        n = C.c_uint32()
        self._check(capi.dev.gi_get_light_cache(self._c, _vp(out), int(capacity), C.byref(n)), "gi_get_light_cache")
        return out, int(n.value)

    def set_environment_constant(self, srgb8):
        """1x1 cube with the same sRGB8 colour on all faces (the reference default is black)."""
        face = np.array([srgb8[0], srgb8[1], srgb8[2], 255], dtype=np.uint8)
        arr = (C.c_void_p * 6)(*[face.ctypes.data] * 6)
        self._check(capi.dev.gi_set_environment_map(self._c, 1, arr), "gi_set_environment_map")

    def set_environment_cube(self, faces):
        """6 (size, size, 4) uint8 sRGB faces in the order -X,+X,-Y,+Y,-Z,+Z (RendererSystem::SetEnvironmentMap, renderersystem.cpp:173-186)."""
        arrs = [np.ascontiguousarray(f, dtype=np.uint8) for f in faces]
        size = arrs[0].shape[0]
        assert all(a.shape == (size, size, 4) for a in arrs) and len(arrs) == 6
        ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
        self._check(capi.dev.gi_set_environment_map(self._c, size, ptrs), "gi_set_environment_map")

    def clear(self):
        self._check(capi.dev.gi_clear_accumulator(self._c), "gi_clear_accumulator")

    def alloc_light_cache(self, capacity):
        self._check(capi.dev.gi_alloc_light_cache(self._c, int(capacity)), "gi_alloc_light_cache")

    def build_hierarchy(self, positions, triangles, tri_materials=None):
        """gi_build_hierarchy: LBVH on the device in the reference's layout -> (nodes, leaf_triangles, parents, num_tree_levels)."""
        positions = np.ascontiguousarray(positions, dtype=np.float32)
        triangles = np.ascontiguousarray(triangles, dtype=np.uint32)
        mats = None if tri_materials is None else np.ascontiguousarray(tri_materials, dtype=np.uint32)
        T = triangles.shape[0]
        nodes = np.zeros(max(2 * T - 1, 1), dtype=NODE_DTYPE)
        leaves = np.zeros(max(2 * T, 2), dtype=TRIANGLE_DTYPE)
        parents = np.zeros(max(2 * T - 1, 1), dtype=np.uint32)
        n, lv = C.c_uint32(), C.c_uint32()
        self._check(capi.dev.gi_build_hierarchy(self._c, _vp(positions), positions.shape[0], _vp(triangles), _vp(mats) if mats is not None else None, T,
                                                _vp(nodes), _vp(leaves), _vp(parents), C.byref(n), C.byref(lv)), "gi_build_hierarchy")
        return nodes[: n.value], leaves[: n.value + 1], parents[: n.value], lv.value

    def render_iteration(self, renderer, frame_seed, max_path_length=16, num_rays_per_light_sample=0, light_cache_capacity=0, warmup=0):
        p = capi.GiIterationParams(renderer=int(renderer), frame_seed=int(frame_seed) & 0xFFFFFFFF, max_path_length=int(max_path_length),
                                   num_rays_per_light_sample=int(num_rays_per_light_sample), light_cache_capacity=int(light_cache_capacity),
                                   warmup=int(warmup))
        self._check(capi.dev.gi_render_iteration(self._c, C.byref(p)), "gi_render_iteration")

    def render_iterations(self, renderer, frame_seeds, max_path_length=16, d_light_samples=0, samples_per_iteration=0, num_rays_per_light_sample=0,
                          light_cache_capacity=0):
        """gi_render_iterations: a batch of iterations in one call (PT: overlapped wave by wave, result identical to one by one)."""
        seeds = np.ascontiguousarray(frame_seeds, dtype=np.uint32)
        p = capi.GiIterationParams(renderer=int(renderer), frame_seed=0, max_path_length=int(max_path_length),
                                   num_rays_per_light_sample=int(num_rays_per_light_sample), light_cache_capacity=int(light_cache_capacity), warmup=0)
        self._check(capi.dev.gi_render_iterations(self._c, C.byref(p), C.cast(_vp(seeds), C.POINTER(C.c_uint32)), seeds.shape[0],
                                                  C.c_void_p(d_light_samples or None), int(samples_per_iteration)), "gi_render_iterations")

    def show_light_caches(self):
        self._check(capi.dev.gi_show_light_caches(self._c), "gi_show_light_caches")

    def trace(self, rays):
        """gi_trace: host ray batch in, host hits out."""
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        self._check(capi.dev.gi_trace(self._c, _vp(rays), rays.shape[0], _vp(hits)), "gi_trace")
        return hits

    def trace_device(self, d_rays_ptr, n, d_hits_ptr):
        self._check(capi.dev.gi_trace_device(self._c, C.c_void_p(d_rays_ptr), int(n), C.c_void_p(d_hits_ptr)), "gi_trace_device")

    def synchronize(self):
        self._check(capi.dev.gi_synchronize(self._c), "gi_synchronize")

    def read_accumulator(self, w=None, h=None, out=None):
        w = w or self._w
        h = h or self._h
        if out is None:
            out = np.zeros((h, w, 4), dtype=np.float32)
        self._check(capi.dev.gi_read_accumulator(self._c, _vp(out)), "gi_read_accumulator")
        return out

    def read_accumulator_async(self, out):
        """gi_read_accumulator_async: snapshot + copy on a second stream, no drain; returns the iterations still in flight.
        `out` (pinned for a truly asynchronous copy) is complete after wait_reads()."""
        n = C.c_uint32()
        self._check(capi.dev.gi_read_accumulator_async(self._c, _vp(out), C.byref(n)), "gi_read_accumulator_async")
        return n.value

    def wait_reads(self):
        self._check(capi.dev.gi_wait_reads(self._c), "gi_wait_reads")

    # ---- multi-GPU read-out (gi_comm_*): one process per GPU, see parallel.connect_ranks for the rendezvous
    def comm_create(self, rank, world):
        """gi_comm_create -> this rank's 64-byte exchange-block handle (bytes)."""
        h = (C.c_ubyte * 64)()
        self._check(capi.dev.gi_comm_create(self._c, int(rank), int(world), h), "gi_comm_create")
        return bytes(h)

    def comm_connect(self, handles):
        """gi_comm_connect: handles = the 64-byte handles of all ranks, in rank order."""
        buf = (C.c_ubyte * (64 * len(handles))).from_buffer_copy(b"".join(handles))
        self._check(capi.dev.gi_comm_connect(self._c, buf), "gi_comm_connect")

    def comm_destroy(self):
        self._check(capi.dev.gi_comm_destroy(self._c), "gi_comm_destroy")

    def read_accumulator_reduced(self, out_rows):
        """gi_read_accumulator_reduced (collective): sum over all ranks, this rank's rows (comm_slice) into out_rows; drains + waits."""
        self._check(capi.dev.gi_read_accumulator_reduced(self._c, _vp(out_rows)), "gi_read_accumulator_reduced")
        return out_rows

    def read_accumulator_reduced_async(self, out_rows):
        n = C.c_uint32()
        self._check(capi.dev.gi_read_accumulator_reduced_async(self._c, _vp(out_rows), C.byref(n)), "gi_read_accumulator_reduced_async")
        return n.value

    def accumulator_device(self):
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        self._check(capi.dev.gi_accumulator_device(self._c, C.byref(ptr), C.byref(nbytes)), "gi_accumulator_device")
        return ptr.value, nbytes.value

    def compute_mse(self, reference_rgb, divisor):
        ref = np.ascontiguousarray(reference_rgb, dtype=np.float32)
        out = C.c_float()
        self._check(capi.dev.gi_compute_mse(self._c, _vp(ref), int(divisor), C.byref(out)), "gi_compute_mse")
        return out.value

    def counters(self):
        c = capi.GiCounters()
        self._check(capi.dev.gi_get_counters(self._c, C.byref(c)), "gi_get_counters")
        return {name: getattr(c, name) for name, _ in capi.GiCounters._fields_}

    def reset_counters(self):
        self._check(capi.dev.gi_reset_counters(self._c), "gi_reset_counters")

    def kernel_launches(self):
        return int(capi.dev.gi_kernel_launches(self._c))

    def scene_topology(self):
        """gi_scene_topology: (GI_TOPOLOGY_* of the uploaded scene's ordered layout, ms the re-clustering added to the upload)."""
        ms = C.c_double(0.0)
        return int(capi.dev.gi_scene_topology(self._c, C.byref(ms))), float(ms.value)

    def scene_resident_bytes(self):
        """gi_scene_resident_bytes: {inner_records, leaf_records, reference_nodes, shading_records} in bytes."""
        out = (C.c_uint64 * 4)()
        self._check(capi.dev.gi_scene_resident_bytes(self._c, out), "gi_scene_resident_bytes")
        return dict(zip(("inner_records", "leaf_records", "reference_nodes", "shading_records"), (int(v) for v in out)))

    def kernel_times(self):
        """gi_get_kernel_times: {class: (ms, launches)} collected while GI_OPT_KERNEL_TIMING is on."""
        t = capi.GiKernelTimes()
        self._check(capi.dev.gi_get_kernel_times(self._c, C.byref(t)), "gi_get_kernel_times")
        names = ["trace", "shade", "generate", "megakernel", "other"]
        return {n: (float(t.ms[i]), int(t.launches[i])) for i, n in enumerate(names)}


# ---------------------------------------------------------------------------------------------------
class RendererSystem:
    """Host mirror of gpugi's RendererSystem + active Renderer (renderersystem.hpp:27-144)."""

    def __init__(self, device=0):
        h = capi.host.gih_rs_create(int(device))
        if not h:
            raise GpugiError("RendererSystem: " + capi.host.gih_last_error().decode())
        self._h = C.c_void_p(h)
        self._scene = None
        self._size = None
        self.device = Device(_borrowed=capi.host.gih_rs_device(self._h))

    def close(self):
        if getattr(self, "_h", None):
            capi.host.gih_rs_destroy(self._h)
            self._h = None

    __del__ = close

    def SetScene(self, scene):
        self._scene = scene
        capi.host.gih_rs_set_scene(self._h, scene._h)

    def SetScreenSize(self, w, h):
        self._size = (int(w), int(h))
        self.device._w, self.device._h = int(w), int(h)
        capi.host.gih_rs_set_screen_size(self._h, int(w), int(h))

    def SetCamera(self, pos, lookat, yfov):
        capi.host.gih_rs_set_camera(self._h, _f3(pos), _f3(lookat), float(yfov))

    def SetRenderer(self, name):
        if capi.host.gih_rs_set_renderer(self._h, RENDERERS[name] if isinstance(name, str) else int(name)) != 0:
            raise GpugiError(capi.host.gih_last_error().decode())

    def GetRendererName(self):
        return capi.host.gih_rs_renderer_name(self._h).decode()

    def SetIterationShard(self, rank, world):
        capi.host.gih_rs_set_shard(self._h, int(rank), int(world))

    def SetSynchronousDraw(self, sync):
        capi.host.gih_rs_set_synchronous(self._h, 1 if sync else 0)

    def Draw(self):
        capi.host.gih_rs_draw(self._h)

    def GetIterationCount(self):
        return int(capi.host.gih_rs_iteration_count(self._h))

    def GetRenderTime(self):
        return int(capi.host.gih_rs_render_time_ms(self._h))

    def GetFrameSeed(self):
        return int(capi.host.gih_rs_frame_seed(self._h))

    def GetGlobalIterationIndex(self):
        return int(capi.host.gih_rs_global_iteration(self._h))

    def GetNumInitialLightSamples(self):
        return int(capi.host.gih_rs_num_light_samples(self._h))

    def GetInitialLightSamples(self):
        out = np.zeros(self.GetNumInitialLightSamples(), dtype=LIGHT_SAMPLE_DTYPE)
        if out.size:
            capi.host.gih_rs_light_samples(self._h, _vp(out))
        return out

    def GetCamera(self):
        out = capi.GiCamera()
        capi.host.gih_rs_camera(self._h, C.byref(out))
        return out

    def GetRendererInfo(self):
        a, b, c = C.c_int32(), C.c_uint32(), C.c_int32()
        capi.host.gih_rs_renderer_info(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"num_rays_per_light_sample": a.value, "light_cache_capacity": b.value, "average_light_path_length": c.value}

    def SetStreaming(self, streaming):
        capi.host.gih_rs_set_streaming(self._h, 1 if streaming else 0)

    def SetDeviceLightSampling(self, on_device):
        capi.host.gih_rs_set_device_light_sampling(self._h, 1 if on_device else 0)

    def ReadBackbufferAsync(self, out):
        """Progressive display read: never stalls or drains the renderer.  Returns the number of submitted iterations that
        are not in the image yet (<= pathLength + 1 while streaming, else 0); `out` is complete after WaitReads()."""
        n = C.c_uint32()
        if capi.host.gih_rs_read_backbuffer_async(self._h, _vp(out), C.byref(n)) != 0:
            raise GpugiError("ReadBackbufferAsync: " + capi.dev.gi_last_error(self.device._c).decode())
        return n.value

    def WaitReads(self):
        if capi.host.gih_rs_wait_reads(self._h) != 0:
            raise GpugiError("WaitReads: " + capi.dev.gi_last_error(self.device._c).decode())

    def ReadBackbufferReduced(self, out_rows):
        """Multi-GPU: sum of all ranks' backbuffers, this rank's rows (comm_slice) -- collective, drains and waits."""
        if capi.host.gih_rs_read_backbuffer_reduced(self._h, _vp(out_rows)) != 0:
            raise GpugiError("ReadBackbufferReduced: " + capi.dev.gi_last_error(self.device._c).decode())
        return out_rows

    def ReadBackbufferReducedAsync(self, out_rows):
        n = C.c_uint32()
        if capi.host.gih_rs_read_backbuffer_reduced_async(self._h, _vp(out_rows), C.byref(n)) != 0:
            raise GpugiError("ReadBackbufferReducedAsync: " + capi.dev.gi_last_error(self.device._c).decode())
        return n.value

    def ReadBackbuffer(self, out=None):
        w, h = self._size
        if out is None:
            out = np.zeros((h, w, 4), dtype=np.float32)
        if capi.host.gih_rs_read_backbuffer(self._h, _vp(out)) != 0:
            raise GpugiError("ReadBackbuffer: " + capi.dev.gi_last_error(self.device._c).decode())
        return out


class Application:
    """Headless Application of the host layer (gpugi/application.cpp): runs the reference's command scripts (scenes/*.txt syntax)."""

    def __init__(self, script_path="", device=0):
        self._h = capi.host.gih_app_create(os.fsencode(script_path) if script_path else None, int(device))
        if not self._h:
            raise GpugiError("Application: " + capi.host.gih_last_error().decode())

    def close(self):
        if self._h:
            capi.host.gih_app_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def Command(self, line):
        capi.host.gih_app_command(self._h, line.encode())

    def Run(self, max_iterations=0):
        capi.host.gih_app_run(self._h, int(max_iterations))

    def GetIterationCount(self):
        return int(capi.host.gih_app_iteration_count(self._h))

    def HasShutDown(self):
        return bool(capi.host.gih_app_has_shut_down(self._h))

    def GetLastMSE(self):
        return float(capi.host.gih_app_last_mse(self._h))

    def GetLastScreenshot(self):
        return capi.host.gih_app_last_screenshot(self._h).decode()


def comm_slice(height, rank, world):
    """gi_comm_slice: (first_row, num_rows) of the summed image that `rank` of `world` reduces and copies out."""
    a, b = C.c_uint32(), C.c_uint32()
    if capi.dev.gi_comm_slice(int(height), int(rank), int(world), C.byref(a), C.byref(b)) != capi.GI_OK:
        raise GpugiError("gi_comm_slice: bad arguments")
    return a.value, b.value


def config_parse(line):
    """One script line through ScriptProcessing::ParseCommand; True if the processor is asked to stop (wait / pause commands)."""
    return bool(capi.host.gih_config_parse(line.encode()))


def config_get(name):
    n = capi.host.gih_config_count(name.encode())
    if n < 0:
        return None
    return [capi.host.gih_config_get(name.encode(), i).decode() for i in range(n)]


def load_image(path):
    """Texture file decoding of the host layer (host/imageio.cpp; the reference: stb_image) -> (h, w, 4) uint8, top row first."""
    w, h = C.c_int(), C.c_int()
    if capi.host.gih_load_image(os.fsencode(path), None, 0, C.byref(w), C.byref(h)) != 0:
        raise GpugiError("load_image: " + capi.host.gih_last_error().decode())
    out = np.zeros((h.value, w.value, 4), dtype=np.uint8)
    if capi.host.gih_load_image(os.fsencode(path), _vp(out), w.value * h.value, C.byref(w), C.byref(h)) != 0:
        raise GpugiError("load_image: " + capi.host.gih_last_error().decode())
    return out


def write_pfm(path, rgba, divisor=1):
    """WritePfm (imageut/hdrimage.cpp:6-33)."""
    rgba = np.ascontiguousarray(rgba, dtype=np.float32)
    if capi.host.gih_write_pfm(os.fsencode(path), _vp(rgba), rgba.shape[1], rgba.shape[0], int(divisor)) != 0:
        raise GpugiError("WritePfm failed: " + str(path))


def load_pfm(path):
    """LoadPfm (imageut/hdrimage.cpp:35-60) -> (h, w, 3) float32."""
    w, h = C.c_int(), C.c_int()
    probe = np.zeros(3, dtype=np.float32)
    st = capi.host.gih_load_pfm(os.fsencode(path), _vp(probe), 0, C.byref(w), C.byref(h))
    if st == 1:
        raise GpugiError("LoadPfm failed: " + str(path))
    out = np.zeros((h.value, w.value, 3), dtype=np.float32)
    if capi.host.gih_load_pfm(os.fsencode(path), _vp(out), w.value * h.value, C.byref(w), C.byref(h)) != 0:
        raise GpugiError("LoadPfm failed: " + str(path))
    return out
