/*
 * gpugi_b200.h -- C ABI of the B200-native light-transport hot path.
 *
 * This is the drop-in boundary for gpugi's `Renderer` seam.  The reference has no
 * FFI layer; the seam is the abstract C++ class `Renderer`
 * (gpugi/renderer/renderer.hpp:25-51) plus the services of `RendererSystem`
 * (gpugi/renderer/renderersystem.hpp:37-144).  Every entry point below cites the
 * reference interface it replaces.  Conventions:
 *   - plain C, opaque handle, `int` status (0 = GI_OK), no exceptions, no torch types;
 *   - all pointer arguments are HOST memory and are copied during the call unless
 *     the name says `_device`;
 *   - one host thread per context (the reference runs everything on one GL thread,
 *     scriptprocessing.cpp:174-181); render calls are asynchronous on the context's
 *     CUDA stream, `gi_synchronize` / read-back calls are the sync points
 *     (the reference's `glFinish`, renderersystem.cpp:249-253).
 *   - there is NO CPU fallback: every compute entry point fails with GI_ERR_CUDA
 *     when no sm_100 device is usable.
 */
#ifndef GPUGI_B200_H
#define GPUGI_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GI_API __attribute__((visibility("default")))

typedef struct GiContext GiContext;   /* opaque; owns device buffers, textures, stream */

/* ------------------------------------------------------------------ status */
enum {
	GI_OK = 0,
	GI_ERR_INVALID = 1,   /* bad argument / call order                         */
	GI_ERR_CUDA = 2,      /* CUDA runtime error, text in gi_last_error()       */
	GI_ERR_NOSCENE = 3,   /* render/trace before gi_upload_scene               */
	GI_ERR_CAPACITY = 4   /* light cache / queue capacity exceeded             */
};

/* ------------------------------------------------------------------ scene PODs
 * Byte-identical to what the reference uploads into its four texture buffers. */

/* Scene::TreeNode<ei::Box>, gpugi/scene/scene.hpp:31-38; consumed by
 * FetchIntersectBoxNode, gpugi/shader/intersectiontests.glsl:100-111.
 * first_child: MSB set => leaf, low 31 bits = leaf index (2 triangles per leaf).
 * escape: next node in preorder when this subtree is skipped; 0 terminates.   */
typedef struct GiNode {
	float    min[3];
	uint32_t first_child;
	float    max[3];
	uint32_t escape;
} GiNode;

/* Scene::Triangle, scene.hpp:63-67 == bim leaf UVec4 (bim_sahspilt.cpp:44-46).
 * An unused leaf slot is all zero (v[0]==v[1] marks it, traceray.glsl:127).   */
typedef struct GiTriangle {
	uint32_t v[3];
	uint32_t material;
} GiTriangle;

/* Scene::VertexInfo, scene.hpp:71-75: (atan2(n.y,n.x), n.z, u, v).            */
typedef struct GiVertexInfo {
	float normal_phi;
	float normal_z;
	float texcoord[2];
} GiVertexInfo;

/* LightSampler::LightSample, gpugi/scene/lightsampler.hpp:19-25.
 * normal_theta_cos > 1.5 marks an omnidirectional point light
 * (lightsource.glsl:25, lightsampler.cpp:92).                                 */
typedef struct GiLightSample {
	float position[3];
	float normal_phi;
	float intensity[3];
	float normal_theta_cos;
} GiLightSample;

/* Texture payload formats (replace the GL formats of scene.cpp:196,206-242). */
enum {
	GI_TEX_SRGB8_ALPHA8 = 0, /* 4 x u8, sRGB decode on RGB (file textures, texture2d.cpp:31-52) */
	GI_TEX_RGBA8        = 1, /* 4 x u8 linear unorm (1x1 RGB8 constants, alpha = 255)           */
	GI_TEX_RGBA32F      = 2  /* 4 x f32 (1x1 RGB32F / RGBA32F constants, alpha = 1 for RGB)      */
};

/* One 2D texture, level 0 only.  Sampler state is fixed to the reference's
 * (min NEAREST / mag LINEAR / REPEAT, scene.cpp:16-21) => bilinear + wrap.   */
typedef struct GiTexture {
	int32_t     width;
	int32_t     height;
	int32_t     format;   /* GI_TEX_*                                          */
	int32_t     reserved;
	const void* pixels;   /* width*height texels, row 0 first                  */
} GiTexture;

/* Scene::Material, scene.hpp:79-89, with the 64-bit bindless GL handles
 * replaced by indices into GiSceneDesc::textures.                            */
typedef struct GiMaterial {
	int32_t diffuse_tex;
	int32_t opacity_tex;
	int32_t reflectiveness_tex;
	int32_t emissivity_tex;
	float   fresnel0[3];
	float   refraction_index_avg;
	float   fresnel1[3];
	float   padding;
} GiMaterial;

/* Everything Scene::UploadGeometry / UploadHierarchy / LoadMaterial hand to GL
 * (scene.cpp:75-130,132-180) and RendererSystem::SetScene binds
 * (renderersystem.cpp:132-152).                                              */
/* Scene::TreeNode<ei::OBox> (scene.hpp:40-49) as assembled by Scene::UploadHierarchy for bvhType = 1 (scene.cpp:107-118): three 16-byte
 * texels per node, read by FetchIntersectOBoxNode (intersectiontests.glsl:114-128). */
typedef struct GiOBoxNode {
	float    center[3];
	uint32_t first_child;      /* same coding as GiNode */
	float    sides_half[3];
	uint32_t escape;
	float    rotation_inv[4];  /* conjugate(orientation), (x, y, z, w) */
} GiOBoxNode;

typedef struct GiSceneDesc {
	const GiNode*       nodes;
	uint32_t            num_nodes;
	const GiTriangle*   leaf_triangles;      /* 2 per leaf                      */
	uint32_t            num_leaf_triangles;
	const float*        positions;           /* 3 floats per vertex             */
	const GiVertexInfo* vertex_info;
	uint32_t            num_vertices;
	const GiMaterial*   materials;
	uint32_t            num_materials;
	const GiTexture*    textures;
	uint32_t            num_textures;
	uint32_t            num_tree_levels;     /* Chunk::getNumTreeLevels, 0 = unknown */
	const GiOBoxNode*   obox_nodes;          /* bvhType = 1 (OBOX_BVH): num_nodes oriented boxes, tested INSTEAD of the AABoxes of `nodes`
	                                            (which stay mandatory: upload validation and the leaf records use them); the walk is then
	                                            always the reference-order one.  NULL = bvhType 0. */
} GiSceneDesc;

/* Camera UBO, gpugi/shader/globaluniforms.glsl:7-15, filled by
 * RendererSystem::SetCamera, renderersystem.cpp:196-213.                     */
typedef struct GiCamera {
	float position[3];
	float u[3];
	float v[3];
	float w[3];
	float pixel_area;
} GiCamera;

/* Ray batch records for gi_trace (parity + Mrays/s).                         */
typedef struct GiRay {
	float origin[3];
	float t_max;          /* closest-hit: initial rayLength (RAY_MAX = 3.402823466e38 for "infinite");
	                         any-hit: the finite rayLength                     */
	float direction[3];
	uint32_t flags;       /* bit 0: any-hit query                              */
} GiRay;

typedef struct GiHit {
	float   t;            /* closest-hit distance; t_max unchanged if missed   */
	int32_t leaf_tri;     /* HIT_INDEX_OUTPUT .y: index into leaf_triangles, -1 = miss (traceray.glsl:153-156).
	                         any-hit: 1 = occluded, 0 = free                   */
	float   bary_y;       /* barycentricCoord.y / .z (x = 1-(z+y))             */
	float   bary_z;
} GiHit;

/* Renderer selection = Application::SwitchRenderer cases 0,1,2 (application.cpp:233-241). */
enum { GI_RENDERER_PT = 0, GI_RENDERER_LPT = 1, GI_RENDERER_BPT = 2,
       GI_RENDERER_WHITTED = 3 /* whittedraytracer.comp; one thread per pixel, the reference's launch shape, in either scheduler */ };

/* gi_trace / gi_set_option modes */
enum {
	GI_TRAVERSAL_FAST = 0,      /* ordered short-stack traversal + exact fallback (default)      */
	GI_TRAVERSAL_REFERENCE = 1  /* stackless escape-pointer walk in the reference's visit order  */
};
enum {
	GI_OPT_TRAVERSAL = 0,       /* GI_TRAVERSAL_*                                                */
	GI_OPT_TEXTURE_FILTER = 1,  /* 0 = hardware bilinear (9-bit weights), 1 = fp32 bilinear      */
	GI_OPT_SCHEDULER = 2,       /* GI_SCHEDULER_*                                                */
	/* tuning of the persistent wavefront trace kernel (defaults are the measured optimum) */
	GI_OPT_WF_REFILL_THRESHOLD = 3, /* idle lanes per warp that trigger a refill from the ray queue  */
	GI_OPT_WF_MIN_INNER = 4,        /* keep taking inner-node steps while this many of 32 active lanes are at inner nodes */
	GI_OPT_WF_BLOCKS_PER_SM = 6,    /* resident trace blocks per SM (0 = occupancy maximum)          */
	GI_OPT_WF_SHADE_BLOCKS_PER_SM = 13, /* resident shade blocks per SM (0 = occupancy maximum)          */
	GI_OPT_RECONSTRUCTION_FILTER = 14,  /* RECONSTRUCTION_FILTER of screenreproject.glsl:14-17,186-223 for every camera-connection splat (LPT, the BPT light
	                                       pass): 0 = box (default, the reference's build), 1 = Mitchell-Netravali 1/3 1/3, 2 = cone, 3 = B-spline; 5 x 5 pixels
	                                       around the rounded position, weighted by the distance to the unrounded one */
	GI_OPT_CHAINED_MIS = 15,            /* 1 = CHAINED_MIS (bidirpathtracer/chained_mis.glsl; hooks lighttracer.comp:62-69,89-92,141-144,165-181,
	                                       screenreproject.glsl:160-170, pathtracer_bidir.comp:49-52,85-89,187-199,236-242): quality-weighted MIS along the
	                                       light and eye paths of the bidirectional path tracer.  Off in the reference's build (bidirectionalpathtracer.cpp:142).
	                                       BPT iterations then run one thread per path (the eye path carries g_mis[MAX_PATHLENGTH]); max_path_length <= 64. */
	GI_OPT_COUNT_STEPS = 7,         /* 1 = the wavefront walk counts its box / triangle tests (GiCounters::node_tests,
	                                   triangle_tests); off by default: two counters less in the hot loop.  The megakernel and
	                                   GI_TRAVERSAL_REFERENCE always count. */
	GI_OPT_KERNEL_TIMING = 8,       /* 1 = bracket every kernel launch of gi_render_iteration with CUDA events on the launching
	                                   stream and sum the durations per kernel class (gi_get_kernel_times); setting the option
	                                   clears the sums.  The iteration call then synchronises -- measurement passes only.
	                                   Replaces the reference's glFinish + clock() pair (renderersystem.cpp:249-255). */
	GI_OPT_STREAMING = 9,           /* 1 = consecutive gi_render_iteration(GI_RENDERER_PT, _LPT or _BPT) calls form a STREAM: each call submits one
	                                   wave (generate + trace + shade) of a run that stays open, so the bounces of up to
	                                   max_path_length + 1 iterations share every wave exactly as in gi_render_iterations -- the
	                                   reference's one-Draw()-per-iteration loop (renderersystem.cpp:246-256) without its
	                                   per-iteration tail.  Any call that reads or redefines state (gi_read_accumulator,
	                                   gi_synchronize, gi_set_camera, gi_clear_accumulator, gi_get_counters, ...) drains the
	                                   stream first, after which the accumulator equals the one of the same iterations rendered
	                                   one by one, bit for bit (PT; BPT: as for a pipelined batch, up to the scheduling-dependent order of
	                                   light-cache slots, and 2 x max_path_length + 1 iterations deep over a ring of light caches; one
	                                   iteration at a time if the ring does not fit the device memory; LPT: up to the order of the float
	                                   atomics inside a pixel).  gi_read_accumulator_async and, in an LPT stream, gi_show_light_caches do
	                                   NOT drain.  LPT / BPT streams collect the splats of every iteration in flight in that iteration's own
	                                   image (a ring of max_path_length + 1, BPT: 2 x max_path_length + 2, RGBA32F images) and add it to the
	                                   accumulator when the iteration is complete.  Light samples: an iteration for which no set was
	                                   installed since the previous one renders with the previous set, as outside a stream. */
	GI_OPT_DEVICE_BUILDER = 10,     /* GI_BUILDER_*: what gi_build_hierarchy builds */
	GI_OPT_PLOC_RADIUS = 12,        /* neighbour window (1..64, default 4) of the clustering passes behind GI_TOPOLOGY_PLOC: tree quality vs upload time */
	GI_OPT_UPLOAD_TOPOLOGY = 11     /* GI_TOPOLOGY_*: the tree the ORDERED walks of the next gi_upload_scene run on.  Results do not depend on it
	                                   (bit-exact either way, see GI_TOPOLOGY_PLOC); refNodes / the reference-order walk are never touched. */
};
enum {
	GI_TOPOLOGY_REFERENCE = 0,  /* the bim hierarchy's own topology (Scene::UploadHierarchy, scene.cpp:95-118)                                  */
	GI_TOPOLOGY_PLOC = 1        /* (default) the reference's LEAVES -- same triangle pairs, same leaf boxes, same ids -- re-clustered on the device
	                               by PLOC at upload.  The reference reaches a leaf iff the slab test of the leaf node's own box passes (the test is
	                               monotone in the box and bim's boxes are nested), so any tree over those leaves finds the same candidates; ties
	                               and near-ties go to the reference-order re-trace as before.  Falls back to GI_TOPOLOGY_REFERENCE by itself when
	                               the input's boxes are not nested. */
};
enum {
	GI_BUILDER_LBVH = 0,        /* Karras' binary radix tree over the Morton-ordered leaves: fastest build, Morton-bit splits       */
	GI_BUILDER_PLOC = 1         /* parallel locally-ordered clustering over the same leaves (default): close to the host SAH sweep   */
};
/* kernel classes of gi_get_kernel_times */
enum {
	GI_KERNEL_TRACE = 0,     /* persistent closest-hit / any-hit traversal waves (k_wf_trace)       */
	GI_KERNEL_SHADE = 1,     /* shading / bounce waves (k_wf_shade_*)                                */
	GI_KERNEL_GENERATE = 2,  /* ray generation (k_wf_generate*)                                      */
	GI_KERNEL_MEGAKERNEL = 3,/* one-thread-per-path kernels (k_pathtrace, k_lighttrace, k_bidir_eye) */
	GI_KERNEL_OTHER = 4,
	GI_KERNEL_CLASS_COUNT = 5
};
typedef struct GiKernelTimes {
	double   ms[GI_KERNEL_CLASS_COUNT];       /* summed CUDA-event durations since GI_OPT_KERNEL_TIMING was set */
	uint64_t launches[GI_KERNEL_CLASS_COUNT];
} GiKernelTimes;
enum {
	GI_SCHEDULER_WAVEFRONT = 0, /* HBM ray queues, persistent trace warps, ballot-compacted waves (default; FAST traversal) */
	GI_SCHEDULER_MEGAKERNEL = 1 /* one thread per path, the reference's launch shape               */
};

/* Per-iteration parameters: the compile-time #defines and small UBOs of the three renderers. */
typedef struct GiIterationParams {
	int32_t  renderer;                  /* GI_RENDERER_*                                          */
	uint32_t frame_seed;                /* PerIteration.FrameSeed = WangHash(iteration), renderersystem.cpp:101 */
	int32_t  max_path_length;           /* MAX_PATHLENGTH, pathtracer.cpp:61                      */
	int32_t  num_rays_per_light_sample; /* LightPathTrace.NumRaysPerLightSample, lightpathtracer.cpp:34 */
	int32_t  light_cache_capacity;      /* LightPathTrace.LightCacheCapacity, bidirectionalpathtracer.cpp:112-126 */
	int32_t  warmup;                    /* 1 = SAVE_LIGHT_CACHE_WARMUP light pass only (bidirectionalpathtracer.cpp:76-108) */
	int32_t  reserved[2];
} GiIterationParams;

/* Work counters (TRACERAY_DEBUG_VARS-style, stdheader.glsl:17-22) + BPT counters
 * (lightcache.glsl:27-37).  Cumulative since the last gi_reset_counters.      */
typedef struct GiCounters {
	uint64_t closest_rays;       /* TraceRay calls                                         */
	uint64_t any_rays;           /* TraceRayAnyHit calls                                   */
	uint64_t camera_paths;       /* eye paths started                                      */
	uint64_t light_paths;        /* light paths started                                    */
	uint64_t shaded_vertices;    /* GetTriangleHitInfo+SampleMaterialData evaluations       */
	uint64_t splats;             /* WriteAtomic calls                                      */
	uint64_t exact_fallback_rays;/* fast traversal rays re-traced in reference order       */
	uint64_t node_tests;         /* box tests actually executed by the device traversal    */
	uint64_t triangle_tests;     /* triangle tests actually executed                       */
	uint32_t light_cache_entries;/* NumLightCacheEntries after the last light pass         */
	uint32_t light_path_length_sum; /* LightPathLengthSum (warm-up)                        */
	/* with GI_OPT_COUNT_STEPS, the record loads of the wavefront walk (DESIGN.md "Data layout"): what the kernel's own layout touches */
	uint64_t inner_node_steps;   /* inner-node records fetched: 96 B each (one 4-wide node = 4 of the node_tests)                   */
	uint64_t leaf_steps;         /* leaf records fetched: 96 B each (both triangles + the leaf's own box = 1 of the node_tests)     */
} GiCounters;

/* ------------------------------------------------------------------ lifecycle */

/* Library/ABI version (major*100 + minor). */
GI_API int gi_version(void);

/* Replaces RendererSystem::RendererSystem (renderersystem.cpp:23-48): creates the
 * context on CUDA device `device` (one process per GPU, LOCAL_RANK).  Fails with
 * GI_ERR_CUDA if the device is missing or is not sm_100 -- there is no CPU path. */
GI_API int gi_create(int device, GiContext** out_ctx);

/* Replaces RendererSystem::~RendererSystem (renderersystem.cpp:50-55) and Scene::~Scene. */
GI_API void gi_destroy(GiContext* ctx);

/* Text of the last error on this context (ctx may be NULL: last gi_create error).
 * Replaces LOG_ERROR + GL debug output (scene.cpp:31,42; outputwindow.cpp:68-70). */
GI_API const char* gi_last_error(const GiContext* ctx);

/* Launch all following work on `cuda_stream` (a cudaStream_t; NULL = the context's own
 * stream).  Lets the caller time with events on its own stream. */
GI_API int gi_set_stream(GiContext* ctx, void* cuda_stream);

/* GI_OPT_* switches; the shader-#define equivalents that are not per-iteration. */
GI_API int gi_set_option(GiContext* ctx, int option, int value);

/* ------------------------------------------------------------------ scene / camera / screen */

/* Replaces Scene::UploadGeometry + UploadHierarchy + LoadMaterial (scene.cpp:75-242)
 * and RendererSystem::SetScene (renderersystem.cpp:132-168): copies the bim hierarchy,
 * re-lays it out for the device (DESIGN.md "HBM layout"), creates texture objects,
 * resets the iteration state and clears the accumulator. */
GI_API int gi_upload_scene(GiContext* ctx, const GiSceneDesc* scene);

/* Replaces RendererSystem::SetEnvironmentMap (renderersystem.cpp:173-186) /
 * Pathtracer::SetEnvironmentMap (pathtracer.cpp:41-45).  faces = -X,+X,-Y,+Y,-Z,+Z,
 * each size*size RGBA8 sRGB; faces == NULL restores the 1x1 black default
 * (renderersystem.cpp:39-47). */
GI_API int gi_set_environment_map(GiContext* ctx, int size, const uint8_t* const* faces);

/* Replaces RendererSystem::SetCamera (renderersystem.cpp:188-219): uploads the camera
 * UBO values and clears the accumulator. */
GI_API int gi_set_camera(GiContext* ctx, const GiCamera* camera);

/* Replaces RendererSystem::SetScreenSize + Renderer::SetScreenSize
 * (renderersystem.cpp:221-233; pathtracer.cpp:36-39): (re)allocates the RGBA32F sum
 * accumulator and clears it. */
GI_API int gi_set_screen_size(GiContext* ctx, int width, int height);

/* Replaces the light-sample half of RendererSystem::PerIterationBufferUpdate
 * (renderersystem.cpp:106-111) and SetNumInitialLightSamples (:118-130): the next
 * iteration -- and every later one, until another set is installed -- uses these `n`
 * initial light samples.  Copied at call time.  (Inside a GI_OPT_STREAMING stream every
 * iteration in flight keeps the set it started with.) */
GI_API int gi_set_light_samples(GiContext* ctx, const GiLightSample* samples, uint32_t n);

/* Same, but the samples already live in device memory (caller-owned, must stay valid until the
 * iterations that use them have finished).  No copy: the persistently mapped light-sample buffer of
 * renderersystem.cpp:121-126 without the host hop. */
GI_API int gi_set_light_samples_device(GiContext* ctx, const GiLightSample* d_samples, uint32_t n);

/* Replaces gl::Texture2D::ClearToZero on the backbuffer (renderersystem.hpp:79). */
GI_API int gi_clear_accumulator(GiContext* ctx);

/* Allocates the BPT light cache for `capacity` 64-byte entries
 * (bidirectionalpathtracer.cpp:116-131). */
GI_API int gi_alloc_light_cache(GiContext* ctx, uint32_t capacity);

/* ------------------------------------------------------------------ hot path */

/* Replaces Renderer::Draw (renderer.hpp:42) of Pathtracer (pathtracer.cpp:47-50),
 * LightPathtracer (lightpathtracer.cpp:45-48, without the show-light-cache pass) and
 * BidirectionalPathtracer (bidirectionalpathtracer.cpp:190-200; with params->warmup the
 * warm-up pass of :90-93).  Adds one sample pass into the accumulator.  Asynchronous. */
GI_API int gi_render_iteration(GiContext* ctx, const GiIterationParams* params);

/* `count` iterations in one call: the application's render loop (application.cpp: `while(...) { Update(); Draw(); }`,
 * script command `waitIterations N`, scriptprocessing.cpp) handed to the device as a whole.  frame_seeds[k] is
 * PerIteration.FrameSeed of the k-th iteration (params->frame_seed is ignored); d_light_samples (DEVICE memory, count x
 * samples_per_iteration records, iteration-major; must stay valid until the call's work has finished) holds the initial
 * light samples of every iteration, or NULL to use the set installed by gi_set_light_samples for all of them.
 * PT + wavefront scheduler: the iterations are overlapped wave by wave (bounce b of iteration k runs together with
 * bounce b+1 of iteration k-1, ...), which removes the per-iteration tail of nearly empty bounces; the accumulator
 * receives every iteration's radiance per pixel in iteration order, so the result is bit-identical to `count` calls of
 * gi_render_iteration.  Other renderers / schedulers: the iterations run one after the other. */
GI_API int gi_render_iterations(GiContext* ctx, const GiIterationParams* params, const uint32_t* frame_seeds, uint32_t count,
                                const GiLightSample* d_light_samples, uint32_t samples_per_iteration);

/* Replaces RendererSystem::DispatchShowLightCacheShader (renderersystem.cpp:277-281):
 * splats the directly visible initial light samples (showlightcaches.comp).  In an LPT stream: those of the iteration
 * submitted last, into that iteration's splat image. */
GI_API int gi_show_light_caches(GiContext* ctx);

/* Replaces TraceRay / TraceRayAnyHit (traceray.glsl:19-180) for a host ray batch:
 * copies rays in, traces, copies hits out (synchronous). */
GI_API int gi_trace(GiContext* ctx, const GiRay* rays, uint32_t n, GiHit* hits);

/* Same, on device-resident buffers (asynchronous on the context stream). */
GI_API int gi_trace_device(GiContext* ctx, const GiRay* d_rays, uint32_t n, GiHit* d_hits);

/* Device-side light sampling: what LightSampler::GenerateRandomSamples (lightsampler.cpp:18-95) needs from the scene,
 * uploaded once per scene / point-light change (Scene::LoadLightSources, ComputePointLightTable, scene.cpp:244-307). */
typedef struct GiLightTriangle { float v0[3], v1[3], v2[3], luminance[3]; } GiLightTriangle;
typedef struct GiLightTables {
	const GiLightTriangle* light_triangles;   /* emissive triangles                                       */
	const float*           summed_area;       /* normalised running sum of their areas (lower_bound key)   */
	uint32_t               num_light_triangles;
	float                  light_area_sum;
	const float*           point_positions;   /* 3 floats per point light                                  */
	const float*           point_intensities; /* 3 floats per point light                                  */
	const float*           point_summed_flux; /* normalised running sum of log(1 + 4 pi luminance)         */
	uint32_t               num_point_lights;
	float                  total_area_flux;
	float                  total_point_flux;
} GiLightTables;
GI_API int gi_set_light_tables(GiContext* ctx, const GiLightTables* tables);

/* LightSampler::GenerateRandomSamples on the device: the `n` light samples of the NEXT iteration, written straight into the context's
 * light-sample ring (what gi_set_light_samples does with host data).  rng_state is the sampler's xorshift64* state (utilities/random.cpp:
 * 25-43, seed 123 at start); it is advanced exactly as the host sampler would advance it (the recurrence is sequential: one thread
 * draws, n threads build the samples).  Same records as the host sampler up to the last bit of atan2. */
GI_API int gi_generate_light_samples(GiContext* ctx, uint64_t* rng_state, uint32_t n, float position_bias);
/* The current light-sample set (host or device generated), for inspection. */
GI_API int gi_get_light_samples(GiContext* ctx, GiLightSample* out, uint32_t n);

/* The light cache as the last non-streamed BPT iteration left it (LightCacheEntry records of 64 bytes, lightcache.glsl:1-16: the SSBO of
 * bidirectionalpathtracer.cpp:116-131) and its NumLightCacheEntries counter, for inspection: copies min(counter, capacity, max_entries)
 * entries to `entries` (host) and stores the counter (which may exceed the capacity: later vertices were dropped) in *num_entries. */
GI_API int gi_get_light_cache(GiContext* ctx, void* entries, uint32_t max_entries, uint32_t* num_entries);

/* Device-side replacement of Chunk::buildHierarchy + remapNodePointers + computeBVHAABoxes (bim_chunk.cpp:227-242, bim_sahspilt.cpp:27-151,
 * bim_hierarchy.cpp:8-50,238-252) for a mesh that comes without a hierarchy: an LBVH (Morton order of the triangle centroids, consecutive
 * pairs per leaf, Karras' binary radix tree, bottom-up boxes) written in the REFERENCE'S layout -- preorder GiNode array with child / escape
 * links, 2 slots per leaf (unused slot all-zero), parents, tree levels -- so that gi_upload_scene, both walks and the .bim writer take it
 * unchanged.  Host buffers in and out, sized for the worst case of one triangle per leaf: out_nodes / out_parents 2*T-1 entries,
 * out_leaf_triangles 2*T; *out_num_nodes = N tells how much was written (N nodes, N+1 leaf slots).
 * triangle_materials may be NULL (material 0). */
GI_API int gi_build_hierarchy(GiContext* ctx, const float* positions, uint32_t num_vertices, const uint32_t* triangles, const uint32_t* triangle_materials,
                              uint32_t num_triangles, GiNode* out_nodes, GiTriangle* out_leaf_triangles, uint32_t* out_parents, uint32_t* out_num_nodes,
                              uint32_t* out_num_tree_levels);

/* Replaces glFinish (renderersystem.cpp:249,253). */
GI_API int gi_synchronize(GiContext* ctx);

/* Which GI_TOPOLOGY_* the uploaded scene's ordered layout stands on (-1: no scene); *build_ms (may be NULL) = time the re-clustering
 * added to gi_upload_scene. */
GI_API int gi_scene_topology(const GiContext* ctx, double* build_ms);

/* Bytes of the uploaded scene that stay resident for the kernels (no reference counterpart: the GL buffers of scene.cpp:75-130 in this
 * path's own layout): out[0] inner-node records of the ordered walk, out[1] leaf records, out[2] the reference-order node array (exact
 * re-traces), out[3] leaf triangles + vertex info + material records (shading; textures not included).  Used by bench.py's roofline. */
GI_API int gi_scene_resident_bytes(const GiContext* ctx, uint64_t out[4]);

/* ------------------------------------------------------------------ read-out */

/* Replaces glGetTextureImage(RGBA32F) in Application::SaveImage (application.cpp:411):
 * copies the width*height*4 float SUM accumulator to host (synchronous).  Division by
 * the iteration count stays with the caller (hdrimage.cpp:21, displayHDR.frag:12). */
GI_API int gi_read_accumulator(GiContext* ctx, float* rgba);

/* Progressive display read-out (the reference shows the backbuffer every frame, application.cpp:336-343, without ever
 * stalling the renderer): snapshots the accumulator in stream order and copies it to `rgba` on a second stream; returns at
 * once.  `rgba` (pinned memory for a truly asynchronous copy) is valid after the next gi_wait_reads / gi_read_accumulator_async
 * -but-one; at most two reads are in flight.  With GI_OPT_STREAMING the snapshot holds every submitted iteration except the
 * last *iterations_in_flight ones (PT: <= max_path_length + 1; LPT: <= max_path_length; BPT: <= 2 x max_path_length + 1), whose paths are still
 * bouncing; it never holds a partial iteration: per-path radiance sums wait in the path state, light-path splats (LPT, and the
 * camera connections of a BPT light pass) in the iteration's own splat image, and both join the accumulator together when the
 * iteration is complete.  Does not drain the stream. */
GI_API int gi_read_accumulator_async(GiContext* ctx, float* rgba, uint32_t* iterations_in_flight);

/* Blocks until every gi_read_accumulator_async copy has landed. */
GI_API int gi_wait_reads(GiContext* ctx);

/* ------------------------------------------------------------------ multi-GPU read-out
 * One process per GPU; rank r of W renders the iterations k = r (mod W) (FrameSeed = WangHash(k), renderersystem.cpp:101) into its own
 * accumulator, and the accumulators only meet at read-out: Application::SaveImage's glGetTextureImage of ONE backbuffer (application.cpp:411,
 * renderersystem.cpp:224) becomes a reduce-scatter over NVLink peer memory in which every rank sums and copies out 1/W of the image
 * (csrc/gi_comm.cuh).  The ranks' exchange blocks are plain device allocations shared through CUDA IPC handles; HOW the 64-byte handles
 * travel between the processes is the caller's business (MPI, a pipe, torch.distributed -- gpugi_b200/parallel.py uses all_gather). */
typedef struct GiCommHandle { unsigned char opaque[64]; } GiCommHandle;

/* Rows [*first_row, *first_row + *num_rows) of a `height`-row image belong to `rank` of `world`.  Pure function, no context. */
GI_API int gi_comm_slice(int height, int rank, int world, uint32_t* first_row, uint32_t* num_rows);

/* Allocates this rank's exchange block (two snapshots of the current screen size + flags) and returns its handle.  Call after
 * gi_set_screen_size; a later size change needs a new communicator.  world == 1 is allowed (no peer, the handle is unused). */
GI_API int gi_comm_create(GiContext* ctx, int rank, int world, GiCommHandle* out_handle);

/* Maps the peers' exchange blocks: handles[r] = the handle rank r got from gi_comm_create (world entries; the own entry is ignored). */
GI_API int gi_comm_connect(GiContext* ctx, const GiCommHandle* handles);

/* COLLECTIVE (every rank issues the same sequence of reduced reads): the sum over all ranks of the accumulators, this rank's rows only,
 * width * num_rows * 4 floats to `rgba_rows` (host memory).  Summation runs in rank order on every rank: the result equals the sequential
 * sum rank 0 + rank 1 + ... of the per-rank images bit for bit.  Drains an open stream first and returns when the rows are on the host. */
GI_API int gi_read_accumulator_reduced(GiContext* ctx, float* rgba_rows);

/* The same without draining or waiting (the multi-GPU gi_read_accumulator_async): snapshot in stream order, reduce + copy on the second
 * stream while the next waves run; `rgba_rows` (pinned) is valid after gi_wait_reads, which also reports a peer that never arrived.
 * *iterations_in_flight as for gi_read_accumulator_async (the same on every rank when all ranks submit in lockstep). */
GI_API int gi_read_accumulator_reduced_async(GiContext* ctx, float* rgba_rows, uint32_t* iterations_in_flight);

/* Unmaps the peers and frees the exchange block (also done by gi_destroy). */
GI_API int gi_comm_destroy(GiContext* ctx);

/* Device address of the accumulator (for callers that bring their own collective, e.g. an NCCL all-reduce). */
GI_API int gi_accumulator_device(GiContext* ctx, void** device_ptr, size_t* num_bytes);

/* Replaces TextureMSE::ComputeCurrentMSE (texturemse.cpp:39-66, sqerror.frag:16-17):
 * mean over pixels of sum_rgb (acc/divisor - ref)^2, ref = width*height*3 floats. */
GI_API int gi_compute_mse(GiContext* ctx, const float* reference_rgb, uint32_t divisor, float* mse);

/* Work counters (see GiCounters). */
GI_API int gi_get_counters(GiContext* ctx, GiCounters* out);
GI_API int gi_reset_counters(GiContext* ctx);

/* Number of CUDA kernel launches issued by this context since creation. */
GI_API uint64_t gi_kernel_launches(const GiContext* ctx);

/* Per-kernel-class device time collected while GI_OPT_KERNEL_TIMING is on (RendererSystem::GetRenderTime's role,
 * renderersystem.hpp:98, per kernel instead of per Draw). */
GI_API int gi_get_kernel_times(GiContext* ctx, GiKernelTimes* out);

#ifdef __cplusplus
}
#endif
#endif /* GPUGI_B200_H */
