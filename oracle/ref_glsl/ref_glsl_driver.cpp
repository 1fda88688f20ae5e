// ref_glsl_driver.cpp -- TEST INFRASTRUCTURE.  Appended by oracle/Makefile to the C++ text glsl2cpp.py makes of one of the reference's shaders
// (same translation unit: it sees the shader's uniforms, buffers and functions as plain globals) and exports a small C interface over it:
// the scene goes in as the GiSceneDesc the product and the oracle take, rays / images come out.  tests/ref_glsl.py binds it.
//   REFGLSL_TRACE   (trace_batch.comp)   refglsl_trace: TraceRay / TraceRayAnyHit of traceray.glsl on a ray batch
//   REFGLSL_PT      (pathtracer.comp)    refglsl_render_pt: one dispatch of the reference's path tracer main() over a pixel region
//   REFGLSL_WHITTED (whittedraytracer.comp) the same entry point over the Whitted tracer's main()
//   REFGLSL_LPT     (lighttracer.comp, optionally -DSAVE_LIGHT_CACHE [-DSAVE_LIGHT_CACHE_WARMUP])  refglsl_render_lpt: light paths first .. first + count
//   REFGLSL_SHOW    (showlightcaches.comp)  refglsl_show_light_caches
//   REFGLSL_BPT     (bidirpathtracer/pathtracer_bidir.comp)  refglsl_render_bpt_eye: the eye pass over a pixel region, reading a given light cache
// Invocations run in ascending order on ONE thread wherever their order is visible in the result (atomic float adds, light-cache slots).
#include "gpugi_b200.h"

#include <vector>

GlslTexture* g_glslTextures = nullptr;
thread_local uvec2 gl_GlobalInvocationID_xy;
thread_local uint gl_GlobalInvocationID_x = 0u;

#define REFGLSL_API extern "C" __attribute__((visibility("default")))

namespace {
std::vector<GlslTexture> s_textures;
std::vector<std::vector<float>> s_texels;
float s_srgb[256];

// What the GL does BEFORE the shader sees a texel (not shader code): sRGB8_ALPHA8 decode (GL 4.5 section 8.24), unorm8 / 255, float as is --
// the definition the oracle and the product use for their texel tables.
void decodeTexture(const GiTexture& t, std::vector<float>& out)
{
	const size_t n = size_t(t.width) * size_t(t.height);
	out.resize(4 * n);
	for (size_t i = 0; i < n; ++i)
	{
		if (t.format == GI_TEX_RGBA32F)
			for (int c = 0; c < 4; ++c) out[4 * i + c] = static_cast<const float*>(t.pixels)[4 * i + c];
		else
		{
			const uint8_t* p = static_cast<const uint8_t*>(t.pixels) + 4 * i;
			for (int c = 0; c < 3; ++c) out[4 * i + c] = t.format == GI_TEX_SRGB8_ALPHA8 ? s_srgb[p[c]] : float(p[c]) / 255.0f;
			out[4 * i + 3] = float(p[3]) / 255.0f;
		}
	}
}
}

REFGLSL_API int refglsl_max_path_length(void)
{
#ifdef MAX_PATHLENGTH
	return MAX_PATHLENGTH;
#else
	return 0;
#endif
}

// Scene::UploadGeometry / UploadHierarchy / LoadMaterial + RendererSystem::SetScene (scene.cpp:75-242, renderersystem.cpp:138-152): the four
// texture buffers and the material UBO, from the arrays of the GiSceneDesc (which must stay alive).
REFGLSL_API int refglsl_set_scene(const GiSceneDesc* d)
{
	if (!d || d->num_materials > 256u) return 1;
	for (int i = 0; i < 256; ++i)
	{
		const float c = float(i) / 255.0f;
		s_srgb[i] = (c <= 0.04045f) ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f);
	}
	HierachyBuffer.data = reinterpret_cast<const float*>(d->nodes); HierachyBuffer.comps = 4;
	TriangleBuffer.data = reinterpret_cast<const int*>(d->leaf_triangles);
	VertexPositionBuffer.data = d->positions; VertexPositionBuffer.comps = 3;   // GL_RGB32F (scene.cpp:84)
	VertexInfoBuffer.data = reinterpret_cast<const float*>(d->vertex_info); VertexInfoBuffer.comps = 4;
	s_textures.assign(d->num_textures, GlslTexture());
	s_texels.assign(d->num_textures, std::vector<float>());
	for (uint32_t i = 0; i < d->num_textures; ++i)
	{
		decodeTexture(d->textures[i], s_texels[i]);
		s_textures[i].w = d->textures[i].width; s_textures[i].h = d->textures[i].height; s_textures[i].rgba = s_texels[i].data();
	}
	g_glslTextures = s_textures.data();
	for (uint32_t i = 0; i < d->num_materials; ++i)
	{
		const GiMaterial& m = d->materials[i];
		Materials[i].diffuseTexHandle = uvec2(uint(m.diffuse_tex), 0u);
		Materials[i].opacityTexHandle = uvec2(uint(m.opacity_tex), 0u);
		Materials[i].reflectivenessTexHandle = uvec2(uint(m.reflectiveness_tex), 0u);
		Materials[i].emissivityTexHandle = uvec2(uint(m.emissivity_tex), 0u);
		Materials[i].Fresnel0 = vec3(m.fresnel0[0], m.fresnel0[1], m.fresnel0[2]);
		Materials[i].RefractionIndexAvg = m.refraction_index_avg;
		Materials[i].Fresnel1 = vec3(m.fresnel1[0], m.fresnel1[1], m.fresnel1[2]);
		Materials[i].padding = m.padding;
	}
	return 0;
}

#ifdef REFGLSL_TRACE
// GiRay batch through the reference's TraceRay (closest hit, HIT_INDEX_OUTPUT: leaf_tri = _hitIndex.y) / TraceRayAnyHit (flags bit 0)
REFGLSL_API int refglsl_trace(const GiRay* rays, uint32_t n, GiHit* hits)
{
#pragma omp parallel for schedule(dynamic, 256)
	for (int64_t i = 0; i < int64_t(n); ++i)
	{
		Ray ray;
		ray.Origin = vec3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
		ray.Direction = vec3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
		GiHit h = {rays[i].t_max, -1, 0.0f, 0.0f};
		if (rays[i].flags & 1u)
			h.leaf_tri = TraceRayAnyHit(ray, rays[i].t_max) ? 1 : 0;
		else
		{
			float rayLength = rays[i].t_max;
			ivec2 hitIndex; vec3 bary; Triangle tri; vec3 triNormal;
			TraceRay(ray, rayLength, hitIndex, bary, tri, triNormal);
			h.t = rayLength;
			if (uint(hitIndex.y) != 0xFFFFFFFFu) { h.leaf_tri = hitIndex.y; h.bary_y = bary.y; h.bary_z = bary.z; }
		}
		hits[i] = h;
	}
	return 0;
}
#endif

#if defined(REFGLSL_PT) || defined(REFGLSL_WHITTED)
// One Pathtracer::Draw (pathtracer.cpp:47-50: glDispatchCompute(W / 8, H / 8, 1) of pathtracer.comp) restricted to the invocations of a pixel
// region; uniforms as RendererSystem::PerIterationBufferUpdate / SetCamera / SetScreenSize fill them.  accum = the RGBA32F backbuffer (in / out).
REFGLSL_API int refglsl_render_pt(const GiCamera* cam, int width, int height, const GiLightSample* samples, int num_samples, uint32_t frame_seed,
                                  float* accum, int x0, int y0, int x1, int y1)
{
	BackbufferSize = ivec2(width, height);
	NumInitialLightSamples = num_samples;
	CameraU = vec3(cam->u[0], cam->u[1], cam->u[2]);
	CameraV = vec3(cam->v[0], cam->v[1], cam->v[2]);
	CameraW = vec3(cam->w[0], cam->w[1], cam->w[2]);
	CameraPosition = vec3(cam->position[0], cam->position[1], cam->position[2]);
	PixelArea = cam->pixel_area;
	FrameSeed = frame_seed;
	InitialLightSampleBuffer.data = reinterpret_cast<const float*>(samples); InitialLightSampleBuffer.comps = 4;
#ifdef REFGLSL_PT
	EnvironmentMap.constant = vec4(0.0f);
#endif
	OutputTexture.rgba = accum; OutputTexture.w = width; OutputTexture.h = height;
	const int gw = (width / 8) * 8, gh = (height / 8) * 8;   // whole work groups only (integer division, pathtracer.cpp:50)
	x1 = x1 < gw ? x1 : gw; y1 = y1 < gh ? y1 : gh;
#pragma omp parallel for schedule(dynamic, 1)
	for (int y = y0; y < y1; ++y)
		for (int x = x0; x < x1; ++x)
		{
			gl_GlobalInvocationID_xy = uvec2(uint(x), uint(y));
			gl_GlobalInvocationID_x = uint(x);
			shader_main();
		}
	return 0;
}
#endif

#ifndef REFGLSL_TRACE
namespace {
void setFrame(const GiCamera* cam, int width, int height, const GiLightSample* samples, int num_samples, uint32_t frame_seed, float* accum)
{
	BackbufferSize = ivec2(width, height);
	NumInitialLightSamples = num_samples;
	CameraU = vec3(cam->u[0], cam->u[1], cam->u[2]);
	CameraV = vec3(cam->v[0], cam->v[1], cam->v[2]);
	CameraW = vec3(cam->w[0], cam->w[1], cam->w[2]);
	CameraPosition = vec3(cam->position[0], cam->position[1], cam->position[2]);
	PixelArea = cam->pixel_area;
	FrameSeed = frame_seed;
	InitialLightSampleBuffer.data = reinterpret_cast<const float*>(samples); InitialLightSampleBuffer.comps = 4;
	OutputTexture.rgba = accum; OutputTexture.w = width; OutputTexture.h = height;
}
}
#endif

#ifdef REFGLSL_LPT
// LightPathtracer::Draw / the light pass of BidirectionalPathtracer::Draw (lightpathtracer.cpp:45-51, bidirectionalpathtracer.cpp:182-195): invocations
// first .. first + count of lighttracer.comp's main().  SAVE_LIGHT_CACHE builds: `cache` = LightCacheEntry array (64 bytes each, capacity entries),
// counters[0] = NumLightCacheEntries (in: what resetcounter.comp set, out: after the pass), counters[1] = LightPathLengthSum.
REFGLSL_API int refglsl_render_lpt(const GiCamera* cam, int width, int height, const GiLightSample* samples, int num_samples, uint32_t frame_seed,
                                   int num_rays_per_light_sample, float* accum, uint32_t first, uint32_t count, void* cache, int capacity, int* counters, int parallel)
{
	setFrame(cam, width, height, samples, num_samples, frame_seed, accum);
	NumRaysPerLightSample = num_rays_per_light_sample;
#ifdef SAVE_LIGHT_CACHE
	LightRayPixelWeight = 3.14159265358979f * 2.0f / float(num_rays_per_light_sample);   // bidirectionalpathtracer.cpp:34 (not read by these shaders)
	LightCacheCapacity = capacity;
	AverageLightPathLength = 0;
	LightCacheEntries = static_cast<LightCacheEntry*>(cache);
	NumLightCacheEntries = counters[0];
	LightPathLengthSum = counters[1];
#endif
	// parallel != 0: throughput runs on all host threads (the atomics of the shim are atomic); the order of the float adds and of the
	// light-cache slots is then scheduling dependent, as on a GPU
#pragma omp parallel for schedule(dynamic, 64) if (parallel)
	for (int64_t i = first; i < int64_t(first) + int64_t(count); ++i)
	{
		gl_GlobalInvocationID_x = uint(i);
		gl_GlobalInvocationID_xy = uvec2(uint(i), 0u);
		shader_main();
	}
#ifdef SAVE_LIGHT_CACHE
	counters[0] = NumLightCacheEntries;
	counters[1] = LightPathLengthSum;
#endif
	return 0;
}
REFGLSL_API int refglsl_light_cache_entry_size(void)
{
#ifdef SAVE_LIGHT_CACHE
	return int(sizeof(LightCacheEntry));
#else
	return 0;
#endif
}
#endif

#ifdef REFGLSL_SHOW
// the second dispatch of LightPathtracer::Draw (lightpathtracer.cpp:49-50): showlightcaches.comp, one invocation per initial light sample
REFGLSL_API int refglsl_show_light_caches(const GiCamera* cam, int width, int height, const GiLightSample* samples, int num_samples, float* accum)
{
	setFrame(cam, width, height, samples, num_samples, 0u, accum);
	for (int i = 0; i < num_samples; ++i)
	{
		gl_GlobalInvocationID_x = uint(i);
		gl_GlobalInvocationID_xy = uvec2(uint(i), 0u);
		shader_main();
	}
	return 0;
}
#endif

#ifdef REFGLSL_BPT
// the eye pass of BidirectionalPathtracer::Draw (bidirectionalpathtracer.cpp:196-201): pathtracer_bidir.comp over a pixel region, with the
// LightPathTrace block and the light cache as the light pass left them
REFGLSL_API int refglsl_render_bpt_eye(const GiCamera* cam, int width, int height, const GiLightSample* samples, int num_samples, uint32_t frame_seed,
                                       int num_rays_per_light_sample, const void* cache, int capacity, int num_entries, int average_light_path_length,
                                       float* accum, int x0, int y0, int x1, int y1)
{
	setFrame(cam, width, height, samples, num_samples, frame_seed, accum);
	NumRaysPerLightSample = num_rays_per_light_sample;
	LightRayPixelWeight = 3.14159265358979f * 2.0f / float(num_rays_per_light_sample);   // bidirectionalpathtracer.cpp:34 (not read by these shaders)
	LightCacheCapacity = capacity;
	AverageLightPathLength = average_light_path_length;
	LightCacheEntries = static_cast<LightCacheEntry*>(const_cast<void*>(cache));
	NumLightCacheEntries = num_entries;
	const int gw = (width / 8) * 8, gh = (height / 8) * 8;
	x1 = x1 < gw ? x1 : gw; y1 = y1 < gh ? y1 : gh;
#pragma omp parallel for schedule(dynamic, 1)
	for (int y = y0; y < y1; ++y)
		for (int x = x0; x < x1; ++x)
		{
			gl_GlobalInvocationID_xy = uvec2(uint(x), uint(y));
			gl_GlobalInvocationID_x = uint(x);
			shader_main();
		}
	return 0;
}
#endif
