/*
 * gpugi_b200_host.h -- flat C view of the C++ host layer (gpugi_b200/host/): the classes that keep
 * the reference's API surface (Scene, Camera, LightSampler, RendererSystem, Pathtracer,
 * LightPathtracer, BidirectionalPathtracer, PFM I/O).  C++ embedders include the .hpp files
 * directly; this header exists so that Python (ctypes) can drive the very same objects.
 * The device hot path is behind include/gpugi_b200.h; nothing here computes light transport.
 */
#ifndef GPUGI_B200_HOST_H
#define GPUGI_B200_HOST_H

#include "gpugi_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct GihScene GihScene;                   /* std::shared_ptr<Scene>           */
typedef struct GihRendererSystem GihRendererSystem; /* RendererSystem + active renderer */
typedef struct GihLightSampler GihLightSampler;

typedef struct GihSceneInfo {
	uint32_t num_triangles, num_vertices, num_nodes, num_leaf_triangles, num_materials, num_tree_levels;
	uint32_t num_light_triangles, num_point_lights;
	const uint32_t* triangles;          /* chunk triangle list, 3 indices each (not the leaf list)  */
	const uint32_t* triangle_materials;
	const uint32_t* parents;            /* HIERARCHY_PARENTS                                         */
	float bbox_min[3], bbox_max[3];
	float light_area_sum, total_area_light_flux, total_point_light_flux;
} GihSceneInfo;

GI_API const char* gih_last_error(void);
GI_API uint32_t gih_wanghash(uint32_t x);                         /* utilities/random.cpp:5-13     */
GI_API uint32_t gih_xorshift32(uint32_t x);                       /* utilities/random.cpp:15-22    */
GI_API uint64_t gih_xorshift64(uint64_t x);                       /* utilities/random.cpp:25-31    */
GI_API float gih_xorshift_f(uint64_t* state);                     /* utilities/random.cpp:39-43 (the light sampler's generator) */
GI_API double gih_xorshift_d(uint64_t* state);                    /* utilities/random.cpp:33-37    */
GI_API void gih_set_path_length(int n);                           /* GlobalConfig "pathLength"      */

/* Scene::Scene(file, bvhType), scene.cpp:15-59.  NULL on failure (gih_last_error).                */
GI_API GihScene* gih_scene_load(const char* json_path, int bvh_type);
/* Synthetic scenes: mesh + environment JSON text -> hierarchy build (0 SAH, 1 kd) -> Scene.       */
GI_API GihScene* gih_scene_from_mesh(const float* positions, const float* normals, const float* texcoords, uint32_t num_vertices,
                                     const uint32_t* triangles, const uint32_t* tri_materials, uint32_t num_triangles, const char* env_json,
                                     const char* texture_dir, int build_method);
/* same with bvh_type: 0 = AABoxes, 1 = OBoxes (GlobalConfig "bvhType", application.cpp:136; oriented boxes are fitted by Chunk::computeBVHOBoxes) */
GI_API GihScene* gih_scene_from_mesh_bvh(const float* positions, const float* normals, const float* texcoords, uint32_t num_vertices,
                                     const uint32_t* triangles, const uint32_t* tri_materials, uint32_t num_triangles, const char* env_json,
                                     const char* texture_dir, int build_method, int bvh_type);
GI_API void gih_scene_destroy(GihScene* s);
/* BinaryModel::storeEnvironmentFile/storeBinaryHeader/storeChunk: writes <json_path> and <dir>/<bim_name>. */
GI_API int gih_scene_store(GihScene* s, const char* json_path, const char* bim_name);
GI_API void gih_scene_desc(GihScene* s, GiSceneDesc* out);
GI_API void gih_scene_info(GihScene* s, GihSceneInfo* out);
GI_API void gih_scene_add_point_light(GihScene* s, const float pos[3], const float intensity[3]); /* application.cpp:158-163 */
GI_API void gih_scene_emissivity(GihScene* s, float* out_rgb /* num_materials*3 */);
GI_API void gih_scene_point_lights(GihScene* s, float* out_pos_intensity /* num_point_lights*6 */);

/* ---- bvhmake's legacy pipeline (host/bvhmake.hpp): the `.rawscene` container (bvhmake/filedef.hpp:11-89) and the three builders
 * over 8-triangle leaves (bvhmake/buildmethods/{kdtree,sweep,lds}.cpp) ---- */
typedef struct GihRawScene GihRawScene;
/* BVHBuilder::BuildBVH on a mesh: method 0 kdtree, 1 sweep, 2 lds; material_names[i] = materialref entry i (<= 31 characters) */
GI_API GihRawScene* gih_rawscene_build(const float* positions, const float* normals, const float* texcoords, uint32_t num_vertices, const uint32_t* triangles,
                                       const uint32_t* tri_materials, uint32_t num_triangles, const char* const* material_names, uint32_t num_materials, int method);
GI_API GihRawScene* gih_rawscene_load(const char* path);                 /* NULL on failure (gih_last_error) */
GI_API int gih_rawscene_store(GihRawScene* r, const char* path);
GI_API void gih_rawscene_free(GihRawScene* r);
/* which: 0 materialref (32 B) 1 vertices (32 B) 2 bounding_aabox (24 B) 3 hierarchy {parent, firstChild, escape} (12 B) 4 triangles = leaves (128 B);
 * copies to dst when it is not NULL and returns the size in bytes */
GI_API uint64_t gih_rawscene_get(GihRawScene* r, int which, void* dst);
/* The raw scene as a Scene: the file's hierarchy with its 8-triangle leaves expanded into subtrees of 2-triangle leaves; env_json = the material file's text */
GI_API GihScene* gih_scene_from_rawscene(GihRawScene* r, const char* env_json, const char* texture_dir);

/* LightSampler (lightsampler.cpp) and Camera::ComputeCameraParams (camera.cpp:76-85) on their own. */
GI_API GihLightSampler* gih_light_sampler_create(GihScene* s);
GI_API void gih_light_sampler_destroy(GihLightSampler* ls);
GI_API void gih_light_sampler_seed(GihLightSampler* ls, uint64_t seed);
GI_API void gih_light_sampler_generate(GihLightSampler* ls, GiLightSample* dst, uint32_t n, float position_bias);
GI_API void gih_camera_params(const float pos[3], const float lookat[3], const float up[3], float yfov_deg, float aspect, int w, int h, GiCamera* out);

/* RendererSystem (renderersystem.hpp:27-144) */
GI_API GihRendererSystem* gih_rs_create(int device);
GI_API void gih_rs_destroy(GihRendererSystem* rs);
GI_API GiContext* gih_rs_device(GihRendererSystem* rs);
GI_API void gih_rs_set_scene(GihRendererSystem* rs, GihScene* s);
GI_API void gih_rs_set_screen_size(GihRendererSystem* rs, int w, int h);     /* "resolution" command */
GI_API void gih_rs_set_camera(GihRendererSystem* rs, const float pos[3], const float lookat[3], float yfov_deg);
GI_API int gih_rs_set_renderer(GihRendererSystem* rs, int renderer);         /* "renderer" command: 0 PT, 1 LPT, 2 BPT */
GI_API const char* gih_rs_renderer_name(GihRendererSystem* rs);
GI_API void gih_rs_set_shard(GihRendererSystem* rs, unsigned rank, unsigned world);
GI_API void gih_rs_set_synchronous(GihRendererSystem* rs, int sync);
GI_API void gih_rs_draw(GihRendererSystem* rs);
GI_API uint64_t gih_rs_iteration_count(GihRendererSystem* rs);
GI_API uint64_t gih_rs_render_time_ms(GihRendererSystem* rs);
GI_API uint32_t gih_rs_frame_seed(GihRendererSystem* rs);
GI_API uint64_t gih_rs_global_iteration(GihRendererSystem* rs);
GI_API int gih_rs_read_backbuffer(GihRendererSystem* rs, float* rgba);
GI_API void gih_rs_set_streaming(GihRendererSystem* rs, int streaming);       /* RendererSystem::SetStreaming (GI_OPT_STREAMING) */
GI_API void gih_rs_set_device_light_sampling(GihRendererSystem* rs, int on_device); /* RendererSystem::SetDeviceLightSampling (gi_generate_light_samples) */
GI_API int gih_rs_read_backbuffer_async(GihRendererSystem* rs, float* rgba, uint32_t* iterations_in_flight); /* Texture2D::ReadImageAsync */
GI_API int gih_rs_wait_reads(GihRendererSystem* rs);
/* multi-GPU: Texture2D::ReadImageReduced / ReadImageReducedAsync = gi_read_accumulator_reduced[_async] on the backbuffer (collective; the
   communicator is set up on the device context: gi_comm_create / gi_comm_connect through gih_rs_device) */
GI_API int gih_rs_read_backbuffer_reduced(GihRendererSystem* rs, float* rgba_rows);
GI_API int gih_rs_read_backbuffer_reduced_async(GihRendererSystem* rs, float* rgba_rows, uint32_t* iterations_in_flight);
GI_API uint32_t gih_rs_num_light_samples(GihRendererSystem* rs);
GI_API void gih_rs_set_num_light_samples(GihRendererSystem* rs, uint32_t n);   /* RendererSystem::SetNumInitialLightSamples (renderersystem.hpp:124) */
GI_API void gih_rs_light_samples(GihRendererSystem* rs, GiLightSample* out);
GI_API void gih_rs_camera(GihRendererSystem* rs, GiCamera* out);
GI_API void gih_rs_renderer_info(GihRendererSystem* rs, int32_t* num_rays_per_light_sample, uint32_t* light_cache_capacity, int32_t* avg_light_path_length);

/* ---- script front end: Application (gpugi/application.cpp:77-193, 287-344) + ScriptProcessing (gpugi/control/scriptprocessing.cpp) ---- */
typedef struct GihApplication GihApplication;
GI_API GihApplication* gih_app_create(const char* script_path, int cuda_device);   /* main.cpp:41 Application(argc, argv); NULL on failure */
GI_API void gih_app_destroy(GihApplication* app);                                  /* saves a screenshot unless "exit" was run (application.cpp:196-198) */
GI_API void gih_app_command(GihApplication* app, const char* line);                /* a console line (scriptprocessing.cpp:29-42) */
GI_API void gih_app_run(GihApplication* app, uint64_t max_iterations);             /* Application::Run; 0 = until "exit" or the script ran out */
GI_API uint64_t gih_app_iteration_count(GihApplication* app);
GI_API int gih_app_has_shut_down(GihApplication* app);
GI_API float gih_app_last_mse(GihApplication* app);
GI_API const char* gih_app_last_screenshot(GihApplication* app);
GI_API GiContext* gih_app_device(GihApplication* app);
/* GlobalConfig (gpugi/control/globalconfig.hpp) through the script syntax "name = v0, v1" */
GI_API int gih_config_parse(const char* line);
GI_API int gih_config_add(const char* name, const char* comma_separated_defaults);
GI_API void gih_config_remove(const char* name);
GI_API int gih_config_count(const char* name);                                     /* -1: no such parameter */
GI_API const char* gih_config_get(const char* name, int index);

/* Texture file decoding (host/imageio.hpp; the reference: stbi_load(file, &w, &h, &n, 4), texture2d.cpp:31-52): width * height RGBA8 texels, top row
 * first.  capacity_pixels == 0 only queries the size.  0 = ok, 1 = cannot decode (gih_last_error), 2 = buffer too small. */
GI_API int gih_load_image(const char* path, unsigned char* rgba, int capacity_pixels, int* w, int* h);

/* WritePfm / LoadPfm (imageut/hdrimage.cpp:6-60) */
GI_API int gih_write_pfm(const char* path, const float* rgba, int w, int h, uint32_t divisor);
GI_API int gih_load_pfm(const char* path, float* rgb, int capacity_pixels, int* w, int* h);

#ifdef __cplusplus
}
#endif
#endif
