"""ctypes declarations for the two in-tree shared libraries.

libgpugi_b200.so       -- CUDA kernels + the C ABI of include/gpugi_b200.h (the product's hot path)
libgpugi_b200_host.so  -- C++ host classes with the reference's surface + include/gpugi_b200_host.h

There is no fallback of any kind: if the libraries are missing the import fails loudly, and every
compute entry point fails when no sm_100 device is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPUGI_B200_LIBDIR: directory holding an A/B build of BOTH libraries (tools/gpu/variants.sh writes gpugi_b200/variants/<tag>/); the
# product libraries next to this file are never overwritten by an experiment.
_LIBDIR = os.environ.get("GPUGI_B200_LIBDIR") or _HERE
DEVICE_LIB_PATH = os.path.join(_LIBDIR, "libgpugi_b200.so")
HOST_LIB_PATH = os.path.join(_LIBDIR, "libgpugi_b200_host.so")


class GpugiBuildError(ImportError):
    pass


def _load(path):
    if not os.path.exists(path):
        raise GpugiBuildError(
            "%s is missing: run `make -C gpugi_b200` (or __graft_entry__.build()). "
            "gpugi_b200 has no CPU/Python fallback for the CUDA hot path." % path)
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


dev = _load(DEVICE_LIB_PATH)
host = _load(HOST_LIB_PATH)

# ------------------------------------------------------------------ PODs (include/gpugi_b200.h)
f32, u32, i32, u64 = C.c_float, C.c_uint32, C.c_int32, C.c_uint64


class GiNode(C.Structure):
    _fields_ = [("min", f32 * 3), ("first_child", u32), ("max", f32 * 3), ("escape", u32)]


class GiTriangle(C.Structure):
    _fields_ = [("v", u32 * 3), ("material", u32)]


class GiVertexInfo(C.Structure):
    _fields_ = [("normal_phi", f32), ("normal_z", f32), ("texcoord", f32 * 2)]


class GiLightSample(C.Structure):
    _fields_ = [("position", f32 * 3), ("normal_phi", f32), ("intensity", f32 * 3), ("normal_theta_cos", f32)]


class GiTexture(C.Structure):
    _fields_ = [("width", i32), ("height", i32), ("format", i32), ("reserved", i32), ("pixels", C.c_void_p)]


class GiMaterial(C.Structure):
    _fields_ = [("diffuse_tex", i32), ("opacity_tex", i32), ("reflectiveness_tex", i32), ("emissivity_tex", i32),
                ("fresnel0", f32 * 3), ("refraction_index_avg", f32), ("fresnel1", f32 * 3), ("padding", f32)]


class GiOBoxNode(C.Structure):
    _fields_ = [("center", f32 * 3), ("first_child", u32), ("sides_half", f32 * 3), ("escape", u32), ("rotation_inv", f32 * 4)]


class GiSceneDesc(C.Structure):
    _fields_ = [("nodes", C.POINTER(GiNode)), ("num_nodes", u32),
                ("leaf_triangles", C.POINTER(GiTriangle)), ("num_leaf_triangles", u32),
                ("positions", C.POINTER(f32)), ("vertex_info", C.POINTER(GiVertexInfo)), ("num_vertices", u32),
                ("materials", C.POINTER(GiMaterial)), ("num_materials", u32),
                ("textures", C.POINTER(GiTexture)), ("num_textures", u32),
                ("num_tree_levels", u32), ("obox_nodes", C.POINTER(GiOBoxNode))]


class GiCamera(C.Structure):
    _fields_ = [("position", f32 * 3), ("u", f32 * 3), ("v", f32 * 3), ("w", f32 * 3), ("pixel_area", f32)]


class GiRay(C.Structure):
    _fields_ = [("origin", f32 * 3), ("t_max", f32), ("direction", f32 * 3), ("flags", u32)]


class GiHit(C.Structure):
    _fields_ = [("t", f32), ("leaf_tri", i32), ("bary_y", f32), ("bary_z", f32)]


class GiIterationParams(C.Structure):
    _fields_ = [("renderer", i32), ("frame_seed", u32), ("max_path_length", i32), ("num_rays_per_light_sample", i32),
                ("light_cache_capacity", i32), ("warmup", i32), ("reserved", i32 * 2)]


class GiCounters(C.Structure):
    _fields_ = [("closest_rays", u64), ("any_rays", u64), ("camera_paths", u64), ("light_paths", u64),
                ("shaded_vertices", u64), ("splats", u64), ("exact_fallback_rays", u64), ("node_tests", u64),
                ("triangle_tests", u64), ("light_cache_entries", u32), ("light_path_length_sum", u32), ("inner_node_steps", u64), ("leaf_steps", u64)]


class GiKernelTimes(C.Structure):
    _fields_ = [("ms", C.c_double * 5), ("launches", u64 * 5)]


class GihSceneInfo(C.Structure):
    _fields_ = [("num_triangles", u32), ("num_vertices", u32), ("num_nodes", u32), ("num_leaf_triangles", u32),
                ("num_materials", u32), ("num_tree_levels", u32), ("num_light_triangles", u32), ("num_point_lights", u32),
                ("triangles", C.POINTER(u32)), ("triangle_materials", C.POINTER(u32)), ("parents", C.POINTER(u32)),
                ("bbox_min", f32 * 3), ("bbox_max", f32 * 3),
                ("light_area_sum", f32), ("total_area_light_flux", f32), ("total_point_light_flux", f32)]


GI_OK = 0
GI_RENDERER_PT, GI_RENDERER_LPT, GI_RENDERER_BPT, GI_RENDERER_WHITTED = 0, 1, 2, 3
GI_TRAVERSAL_FAST, GI_TRAVERSAL_REFERENCE = 0, 1
GI_OPT_TRAVERSAL, GI_OPT_TEXTURE_FILTER, GI_OPT_SCHEDULER = 0, 1, 2
GI_OPT_WF_REFILL_THRESHOLD, GI_OPT_WF_MIN_INNER, GI_OPT_WF_BLOCKS_PER_SM, GI_OPT_COUNT_STEPS = 3, 4, 6, 7
GI_OPT_KERNEL_TIMING = 8
GI_OPT_STREAMING = 9
GI_OPT_DEVICE_BUILDER = 10
GI_OPT_WF_SHADE_BLOCKS_PER_SM = 13
GI_OPT_RECONSTRUCTION_FILTER = 14
GI_OPT_CHAINED_MIS = 15
GI_OPT_UPLOAD_TOPOLOGY = 11
GI_OPT_PLOC_RADIUS = 12
GI_TOPOLOGY_REFERENCE, GI_TOPOLOGY_PLOC = 0, 1
GI_BUILDER_LBVH, GI_BUILDER_PLOC = 0, 1
GI_KERNEL_TRACE, GI_KERNEL_SHADE, GI_KERNEL_GENERATE, GI_KERNEL_MEGAKERNEL, GI_KERNEL_OTHER, GI_KERNEL_CLASS_COUNT = 0, 1, 2, 3, 4, 5
GI_SCHEDULER_WAVEFRONT, GI_SCHEDULER_MEGAKERNEL = 0, 1
GI_TEX_SRGB8_ALPHA8, GI_TEX_RGBA8, GI_TEX_RGBA32F = 0, 1, 2
RAY_MAX = 3.402823466e+38

P = C.POINTER
vp = C.c_void_p

# every symbol declared in include/gpugi_b200.h: (restype, argtypes)
DEVICE_API = {
    "gi_version": (C.c_int, []),
    "gi_create": (C.c_int, [C.c_int, P(vp)]),
    "gi_destroy": (None, [vp]),
    "gi_last_error": (C.c_char_p, [vp]),
    "gi_set_stream": (C.c_int, [vp, vp]),
    "gi_set_option": (C.c_int, [vp, C.c_int, C.c_int]),
    "gi_upload_scene": (C.c_int, [vp, P(GiSceneDesc)]),
    "gi_set_environment_map": (C.c_int, [vp, C.c_int, P(vp)]),
    "gi_set_camera": (C.c_int, [vp, P(GiCamera)]),
    "gi_set_screen_size": (C.c_int, [vp, C.c_int, C.c_int]),
    "gi_set_light_samples": (C.c_int, [vp, P(GiLightSample), u32]),
    "gi_set_light_samples_device": (C.c_int, [vp, vp, u32]),
    "gi_clear_accumulator": (C.c_int, [vp]),
    "gi_alloc_light_cache": (C.c_int, [vp, u32]),
    "gi_render_iteration": (C.c_int, [vp, P(GiIterationParams)]),
    "gi_render_iterations": (C.c_int, [vp, P(GiIterationParams), P(u32), u32, vp, u32]),
    "gi_show_light_caches": (C.c_int, [vp]),
    "gi_trace": (C.c_int, [vp, vp, u32, vp]),
    "gi_trace_device": (C.c_int, [vp, vp, u32, vp]),
    "gi_synchronize": (C.c_int, [vp]),
    "gi_read_accumulator": (C.c_int, [vp, vp]),
    "gi_read_accumulator_async": (C.c_int, [vp, vp, P(u32)]),
    "gi_wait_reads": (C.c_int, [vp]),
    "gi_comm_slice": (C.c_int, [C.c_int, C.c_int, C.c_int, P(u32), P(u32)]),
    "gi_comm_create": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "gi_comm_connect": (C.c_int, [vp, vp]),
    "gi_comm_destroy": (C.c_int, [vp]),
    "gi_read_accumulator_reduced": (C.c_int, [vp, vp]),
    "gi_read_accumulator_reduced_async": (C.c_int, [vp, vp, P(u32)]),
    "gi_set_light_tables": (C.c_int, [vp, vp]),
    "gi_generate_light_samples": (C.c_int, [vp, P(u64), u32, f32]),
    "gi_get_light_samples": (C.c_int, [vp, vp, u32]),
    "gi_get_light_cache": (C.c_int, [vp, vp, u32, P(u32)]),
    "gi_build_hierarchy": (C.c_int, [vp, vp, u32, vp, vp, u32, vp, vp, vp, P(u32), P(u32)]),
    "gi_accumulator_device": (C.c_int, [vp, P(vp), P(C.c_size_t)]),
    "gi_compute_mse": (C.c_int, [vp, vp, u32, P(f32)]),
    "gi_get_counters": (C.c_int, [vp, P(GiCounters)]),
    "gi_reset_counters": (C.c_int, [vp]),
    "gi_kernel_launches": (u64, [vp]),
    "gi_scene_topology": (C.c_int, [vp, P(C.c_double)]),
    "gi_scene_resident_bytes": (C.c_int, [vp, P(C.c_uint64)]),
    "gi_get_kernel_times": (C.c_int, [vp, P(GiKernelTimes)]),
}

HOST_API = {
    "gih_last_error": (C.c_char_p, []),
    "gih_wanghash": (u32, [u32]),
    "gih_xorshift32": (u32, [u32]),
    "gih_xorshift64": (C.c_uint64, [C.c_uint64]),
    "gih_xorshift_f": (f32, [P(C.c_uint64)]),
    "gih_xorshift_d": (C.c_double, [P(C.c_uint64)]),
    "gih_set_path_length": (None, [C.c_int]),
    "gih_scene_load": (vp, [C.c_char_p, C.c_int]),
    "gih_scene_from_mesh": (vp, [vp, vp, vp, u32, vp, vp, u32, C.c_char_p, C.c_char_p, C.c_int]),
    "gih_scene_from_mesh_bvh": (vp, [vp, vp, vp, u32, vp, vp, u32, C.c_char_p, C.c_char_p, C.c_int, C.c_int]),
    "gih_scene_destroy": (None, [vp]),
    "gih_rawscene_build": (vp, [vp, vp, vp, u32, vp, vp, u32, P(C.c_char_p), u32, C.c_int]),
    "gih_rawscene_load": (vp, [C.c_char_p]),
    "gih_rawscene_store": (C.c_int, [vp, C.c_char_p]),
    "gih_rawscene_free": (None, [vp]),
    "gih_rawscene_get": (u64, [vp, C.c_int, vp]),
    "gih_scene_from_rawscene": (vp, [vp, C.c_char_p, C.c_char_p]),
    "gih_scene_store": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
    "gih_scene_desc": (None, [vp, P(GiSceneDesc)]),
    "gih_scene_info": (None, [vp, P(GihSceneInfo)]),
    "gih_scene_add_point_light": (None, [vp, P(f32), P(f32)]),
    "gih_scene_emissivity": (None, [vp, vp]),
    "gih_scene_point_lights": (None, [vp, vp]),
    "gih_light_sampler_create": (vp, [vp]),
    "gih_light_sampler_destroy": (None, [vp]),
    "gih_light_sampler_seed": (None, [vp, u64]),
    "gih_light_sampler_generate": (None, [vp, vp, u32, f32]),
    "gih_camera_params": (None, [P(f32), P(f32), P(f32), f32, f32, C.c_int, C.c_int, P(GiCamera)]),
    "gih_rs_create": (vp, [C.c_int]),
    "gih_rs_destroy": (None, [vp]),
    "gih_rs_device": (vp, [vp]),
    "gih_rs_set_scene": (None, [vp, vp]),
    "gih_rs_set_screen_size": (None, [vp, C.c_int, C.c_int]),
    "gih_rs_set_camera": (None, [vp, P(f32), P(f32), f32]),
    "gih_rs_set_renderer": (C.c_int, [vp, C.c_int]),
    "gih_rs_renderer_name": (C.c_char_p, [vp]),
    "gih_rs_set_shard": (None, [vp, C.c_uint, C.c_uint]),
    "gih_rs_set_synchronous": (None, [vp, C.c_int]),
    "gih_rs_draw": (None, [vp]),
    "gih_rs_iteration_count": (u64, [vp]),
    "gih_rs_render_time_ms": (u64, [vp]),
    "gih_rs_frame_seed": (u32, [vp]),
    "gih_rs_global_iteration": (u64, [vp]),
    "gih_rs_read_backbuffer": (C.c_int, [vp, vp]),
    "gih_rs_set_streaming": (None, [vp, C.c_int]),
    "gih_rs_set_device_light_sampling": (None, [vp, C.c_int]),
    "gih_rs_read_backbuffer_async": (C.c_int, [vp, vp, P(u32)]),
    "gih_rs_wait_reads": (C.c_int, [vp]),
    "gih_rs_read_backbuffer_reduced": (C.c_int, [vp, vp]),
    "gih_rs_read_backbuffer_reduced_async": (C.c_int, [vp, vp, P(u32)]),
    "gih_rs_num_light_samples": (u32, [vp]),
    "gih_rs_set_num_light_samples": (None, [vp, u32]),
    "gih_rs_light_samples": (None, [vp, vp]),
    "gih_rs_camera": (None, [vp, P(GiCamera)]),
    "gih_rs_renderer_info": (None, [vp, P(i32), P(u32), P(i32)]),
    "gih_app_create": (vp, [C.c_char_p, C.c_int]),
    "gih_app_destroy": (None, [vp]),
    "gih_app_command": (None, [vp, C.c_char_p]),
    "gih_app_run": (None, [vp, u64]),
    "gih_app_iteration_count": (u64, [vp]),
    "gih_app_has_shut_down": (C.c_int, [vp]),
    "gih_app_last_mse": (f32, [vp]),
    "gih_app_last_screenshot": (C.c_char_p, [vp]),
    "gih_app_device": (vp, [vp]),
    "gih_config_parse": (C.c_int, [C.c_char_p]),
    "gih_config_add": (C.c_int, [C.c_char_p, C.c_char_p]),
    "gih_config_remove": (None, [C.c_char_p]),
    "gih_config_count": (C.c_int, [C.c_char_p]),
    "gih_config_get": (C.c_char_p, [C.c_char_p, C.c_int]),
    "gih_load_image": (C.c_int, [C.c_char_p, vp, C.c_int, P(C.c_int), P(C.c_int)]),
    "gih_write_pfm": (C.c_int, [C.c_char_p, vp, C.c_int, C.c_int, u32]),
    "gih_load_pfm": (C.c_int, [C.c_char_p, vp, C.c_int, P(C.c_int), P(C.c_int)]),
}


def _bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)  # AttributeError => the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args


_bind(dev, DEVICE_API)
_bind(host, HOST_API)
