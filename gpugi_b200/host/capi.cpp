// capi.cpp -- flat C entry points over the C++ host classes (Scene, RendererSystem, the three
// renderers, PFM I/O) so that Python (ctypes) tests and bench.py can drive the same objects a C++
// embedder would.  Not part of the device C ABI (include/gpugi_b200.h); see include/gpugi_b200_host.h.
#include "../../include/gpugi_b200_host.h"

#include "bidirectionalpathtracer.hpp"
#include "bvhmake.hpp"
#include "imageio.hpp"
#include "globalconfig.hpp"
#include "hdrimage.hpp"
#include "lightpathtracer.hpp"
#include "logger.hpp"
#include "pathtracer.hpp"
#include "whittedraytracer.hpp"
#include "random.hpp"
#include "application.hpp"
#include "renderersystem.hpp"
#include "scriptprocessing.hpp"
#include "scene.hpp"

#include <cstring>
#include <vector>
#include <memory>
#include <stdexcept>

using gi::bim::BinaryModel;
using gi::bim::Chunk;

struct GihScene { std::shared_ptr<Scene> scene; };
struct GihRendererSystem { std::unique_ptr<RendererSystem> rs; };

extern "C" {

const char* gih_last_error(void) { return gi::LastHostError().c_str(); }

uint32_t gih_wanghash(uint32_t x) { return WangHash(x); }
// the host generators of utilities/random.cpp:15-43, exported for the known-answer tests against the reference's own object code
uint32_t gih_xorshift32(uint32_t x) { return Xorshift(x); }
uint64_t gih_xorshift64(uint64_t x) { return Xorshift(x); }
float gih_xorshift_f(uint64_t* state) { return XorshiftF(*state); }
double gih_xorshift_d(uint64_t* state) { return XorshiftD(*state); }

void gih_set_path_length(int n) { GlobalConfig::SetParameter("pathLength", {Variant(n)}); }

GihScene* gih_scene_load(const char* json_path, int bvh_type)
{
	gi::LastHostError().clear();
	std::unique_ptr<GihScene> h(new GihScene());
	h->scene = std::make_shared<Scene>(json_path, bvh_type == 0 ? ei::Types3D::BOX : ei::Types3D::OBOX);
	if (!h->scene->IsValid()) return nullptr;
	return h.release();
}

GihScene* gih_scene_from_mesh(const float* positions, const float* normals, const float* texcoords, uint32_t num_vertices, const uint32_t* triangles,
                              const uint32_t* tri_materials, uint32_t num_triangles, const char* env_json, const char* texture_dir, int build_method)
{
	return gih_scene_from_mesh_bvh(positions, normals, texcoords, num_vertices, triangles, tri_materials, num_triangles, env_json, texture_dir, build_method, 0);
}

GihScene* gih_scene_from_mesh_bvh(const float* positions, const float* normals, const float* texcoords, uint32_t num_vertices, const uint32_t* triangles,
                                  const uint32_t* tri_materials, uint32_t num_triangles, const char* env_json, const char* texture_dir, int build_method, int bvh_type)
{
	gi::LastHostError().clear();
	std::unique_ptr<BinaryModel> model(new BinaryModel());
	std::string err;
	if (!model->loadEnvironmentText(env_json, &err)) { LOG_ERROR("materials JSON: " << err); return nullptr; }
	model->chunk().setMesh(reinterpret_cast<const gi::bim::Vec3*>(positions), reinterpret_cast<const gi::bim::Vec3*>(normals),
	                       reinterpret_cast<const gi::bim::Vec2*>(texcoords), num_vertices, reinterpret_cast<const gi::bim::UVec3*>(triangles), tri_materials,
	                       num_triangles);
	model->useJsonMaterialOrder(); // TRIANGLE_MAT indices follow the order of the JSON "materials" object
	if (build_method == 2 || build_method == 3)
	{
		// device-side LBVH (gi_build_hierarchy): same layout as the host builders, built on the GPU
		GiContext* ctx = nullptr;
		if (gi_create(0, &ctx) != GI_OK) { LOG_ERROR("device hierarchy build: " << gi_last_error(nullptr)); return nullptr; }
		gi_set_option(ctx, GI_OPT_DEVICE_BUILDER, build_method == 2 ? GI_BUILDER_LBVH : GI_BUILDER_PLOC);
		const uint32_t N = 2u * num_triangles - 1u; // worst case: one triangle per leaf
		std::vector<GiNode> nodes(N);
		std::vector<GiTriangle> leaves(size_t(num_triangles) * 2);
		std::vector<uint32_t> parents(N);
		uint32_t numNodes = 0, levels = 0;
		int st = gi_build_hierarchy(ctx, positions, num_vertices, triangles, tri_materials, num_triangles, nodes.data(), leaves.data(), parents.data(), &numNodes, &levels);
		if (st != GI_OK) { LOG_ERROR("device hierarchy build: " << gi_last_error(ctx)); gi_destroy(ctx); return nullptr; }
		gi_destroy(ctx);
		std::vector<gi::bim::Node> hn(numNodes);
		std::vector<gi::bim::Box> hb(numNodes);
		for (uint32_t i = 0; i < numNodes; ++i)
		{
			hn[i].firstChild = nodes[i].first_child; hn[i].escape = nodes[i].escape;
			hb[i].min = {nodes[i].min[0], nodes[i].min[1], nodes[i].min[2]}; hb[i].max = {nodes[i].max[0], nodes[i].max[1], nodes[i].max[2]};
		}
		static_assert(sizeof(GiTriangle) == sizeof(gi::bim::UVec4), "leaf slot layout");
		model->chunk().setHierarchy(hn.data(), parents.data(), hb.data(), numNodes, reinterpret_cast<const gi::bim::UVec4*>(leaves.data()), numNodes + 1u, levels);
	}
	else
	{
		model->chunk().buildHierarchy(build_method == 1 ? Chunk::BuildMethod::KD_TREE : Chunk::BuildMethod::SAH);
		model->chunk().computeBVHAABoxes();
	}
	std::unique_ptr<GihScene> h(new GihScene());
	h->scene = std::make_shared<Scene>(std::move(model), texture_dir ? texture_dir : "", bvh_type == 0 ? ei::Types3D::BOX : ei::Types3D::OBOX);
	if (!h->scene->IsValid()) return nullptr;
	return h.release();
}

void gih_scene_destroy(GihScene* s) { delete s; }

// ---- bvhmake's legacy pipeline (host/bvhmake.hpp)
struct GihRawScene { gi::bvhmake::RawScene raw; };

GihRawScene* gih_rawscene_build(const float* positions, const float* normals, const float* texcoords, uint32_t num_vertices, const uint32_t* triangles,
                                const uint32_t* tri_materials, uint32_t num_triangles, const char* const* material_names, uint32_t num_materials, int method)
{
	gi::LastHostError().clear();
	if (!positions || !triangles || num_vertices == 0 || num_triangles == 0 || method < 0 || method > 2) { LOG_ERROR("gih_rawscene_build: bad arguments"); return nullptr; }
	std::unique_ptr<GihRawScene> h(new GihRawScene());
	gi::bvhmake::RawScene& r = h->raw;
	r.vertices.resize(num_vertices);
	for (uint32_t i = 0; i < num_vertices; ++i)
	{
		gi::bvhmake::Vertex& v = r.vertices[i];
		v.position = {positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]};
		v.normal = normals ? gi::bim::Vec3{normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]} : gi::bim::Vec3{0.0f, 1.0f, 0.0f};
		v.texcoord = texcoords ? gi::bim::Vec2{texcoords[2 * i], texcoords[2 * i + 1]} : gi::bim::Vec2{0.0f, 0.0f};
	}
	std::vector<gi::bvhmake::Triangle> tris(num_triangles);
	for (uint32_t t = 0; t < num_triangles; ++t)
	{
		for (int k = 0; k < 3; ++k)
		{
			if (triangles[3 * t + k] >= num_vertices) { LOG_ERROR("gih_rawscene_build: vertex index out of range"); return nullptr; }
			tris[t].vertices[k] = triangles[3 * t + k];
		}
		tris[t].material = tri_materials ? tri_materials[t] : 0u;
	}
	r.materials.resize(num_materials);
	for (uint32_t i = 0; i < num_materials; ++i)
	{
		std::memset(r.materials[i].material, 0, 32);
		std::strncpy(r.materials[i].material, material_names[i], 31);
	}
	gi::bvhmake::BuildHierarchy(r, tris, method == 0 ? gi::bvhmake::BuildMethod::KDTREE : (method == 1 ? gi::bvhmake::BuildMethod::SWEEP : gi::bvhmake::BuildMethod::LDS));
	return h.release();
}

GihRawScene* gih_rawscene_load(const char* path)
{
	gi::LastHostError().clear();
	std::unique_ptr<GihRawScene> h(new GihRawScene());
	std::string err;
	if (!gi::bvhmake::LoadRawScene(path, h->raw, &err)) { LOG_ERROR(err); return nullptr; }
	return h.release();
}

int gih_rawscene_store(GihRawScene* r, const char* path)
{
	std::string err;
	if (!gi::bvhmake::StoreRawScene(path, r->raw, &err)) { LOG_ERROR(err); return 1; }
	return 0;
}

void gih_rawscene_free(GihRawScene* r) { delete r; }

uint64_t gih_rawscene_get(GihRawScene* r, int which, void* dst)
{
	const void* src = nullptr; uint64_t n = 0;
	switch (which)
	{
	case 0: src = r->raw.materials.data(); n = r->raw.materials.size() * sizeof(gi::bvhmake::MaterialRef); break;
	case 1: src = r->raw.vertices.data(); n = r->raw.vertices.size() * sizeof(gi::bvhmake::Vertex); break;
	case 2: src = r->raw.boxes.data(); n = r->raw.boxes.size() * sizeof(gi::bim::Box); break;
	case 3: src = r->raw.nodes.data(); n = r->raw.nodes.size() * sizeof(gi::bvhmake::Node); break;
	case 4: src = r->raw.leaves.data(); n = r->raw.leaves.size() * sizeof(gi::bvhmake::Leaf); break;
	}
	if (dst && n) std::memcpy(dst, src, n);
	return n;
}

GihScene* gih_scene_from_rawscene(GihRawScene* r, const char* env_json, const char* texture_dir)
{
	gi::LastHostError().clear();
	std::unique_ptr<BinaryModel> model(new BinaryModel());
	std::string err;
	if (!gi::bvhmake::ToModel(r->raw, env_json ? env_json : "{}", *model, &err)) { LOG_ERROR("raw scene: " << err); return nullptr; }
	std::unique_ptr<GihScene> h(new GihScene());
	h->scene = std::make_shared<Scene>(std::move(model), texture_dir ? texture_dir : "", ei::Types3D::BOX);
	if (!h->scene->IsValid()) return nullptr;
	return h.release();
}

int gih_scene_store(GihScene* s, const char* json_path, const char* bim_name)
{
	std::string err;
	if (!s->scene->GetModel().store(json_path, bim_name, &err)) { LOG_ERROR(err); return 1; }
	return 0;
}

void gih_scene_desc(GihScene* s, GiSceneDesc* out) { *out = s->scene->GetSceneDesc(); }

void gih_scene_add_point_light(GihScene* s, const float pos[3], const float intensity[3])
{
	Scene::PointLight l;
	l.position = ei::Vec3(pos[0], pos[1], pos[2]);
	l.intensity = ei::Vec3(intensity[0], intensity[1], intensity[2]);
	s->scene->AddPointLight(l);
}

void gih_scene_info(GihScene* s, GihSceneInfo* out)
{
	const Scene& sc = *s->scene;
	const Chunk& c = sc.GetModel().chunk();
	out->num_triangles = c.getNumTriangles();
	out->num_vertices = c.getNumVertices();
	out->num_nodes = c.getNumNodes();
	out->num_leaf_triangles = sc.GetNumLeafTriangles();
	out->num_materials = sc.GetNumMaterials();
	out->num_tree_levels = c.getNumTreeLevels();
	out->num_light_triangles = sc.GetNumLightTriangles();
	out->num_point_lights = uint32_t(sc.GetPointLights().size());
	out->triangles = reinterpret_cast<const uint32_t*>(c.getTriangles());
	out->triangle_materials = c.getTriangleMaterials();
	out->parents = c.getHierarchyParents();
	ei::Box b = sc.GetBoundingBox();
	out->bbox_min[0] = b.min.x; out->bbox_min[1] = b.min.y; out->bbox_min[2] = b.min.z;
	out->bbox_max[0] = b.max.x; out->bbox_max[1] = b.max.y; out->bbox_max[2] = b.max.z;
	out->light_area_sum = sc.GetLightAreaSum();
	out->total_area_light_flux = sc.GetTotalAreaLightFlux();
	out->total_point_light_flux = sc.GetTotalPointLightFlux();
}

void gih_scene_emissivity(GihScene* s, float* out_rgb)
{
	const std::vector<ei::Vec3>& e = s->scene->GetConstantEmissivity();
	for (size_t i = 0; i < e.size(); ++i) { out_rgb[3 * i] = e[i].x; out_rgb[3 * i + 1] = e[i].y; out_rgb[3 * i + 2] = e[i].z; }
}

void gih_scene_point_lights(GihScene* s, float* out_pos_intensity)
{
	const auto& l = s->scene->GetPointLights();
	for (size_t i = 0; i < l.size(); ++i)
	{
		float* o = out_pos_intensity + 6 * i;
		o[0] = l[i].position.x; o[1] = l[i].position.y; o[2] = l[i].position.z;
		o[3] = l[i].intensity.x; o[4] = l[i].intensity.y; o[5] = l[i].intensity.z;
	}
}

// ---- standalone light sampler (host code under test against the oracle's restatement)
struct GihLightSampler { LightSampler sampler; };
GihLightSampler* gih_light_sampler_create(GihScene* s)
{
	GihLightSampler* ls = new GihLightSampler();
	ls->sampler.SetScene(s->scene);
	return ls;
}
void gih_light_sampler_destroy(GihLightSampler* ls) { delete ls; }
void gih_light_sampler_seed(GihLightSampler* ls, uint64_t seed) { ls->sampler.SetSeed(seed); }
void gih_light_sampler_generate(GihLightSampler* ls, GiLightSample* dst, uint32_t n, float bias) { ls->sampler.GenerateRandomSamples(dst, n, bias); }

void gih_camera_params(const float pos[3], const float lookat[3], const float up[3], float yfov, float aspect, int w, int h, GiCamera* out)
{
	Camera cam(ei::Vec3(pos[0], pos[1], pos[2]), ei::Vec3(lookat[0], lookat[1], lookat[2]), aspect, yfov, ei::Vec3(up[0], up[1], up[2]));
	ei::Vec3 U, V, W;
	cam.ComputeCameraParams(U, V, W);
	out->position[0] = pos[0]; out->position[1] = pos[1]; out->position[2] = pos[2];
	out->u[0] = U.x; out->u[1] = U.y; out->u[2] = U.z;
	out->v[0] = V.x; out->v[1] = V.y; out->v[2] = V.z;
	out->w[0] = W.x; out->w[1] = W.y; out->w[2] = W.z;
	out->pixel_area = ei::len(U) * ei::len(V) * 4.0f / (w * h);
}

// ---- RendererSystem
GihRendererSystem* gih_rs_create(int device)
{
	gi::LastHostError().clear();
	try
	{
		std::unique_ptr<GihRendererSystem> h(new GihRendererSystem());
		h->rs.reset(new RendererSystem(device));
		return h.release();
	}
	catch (const std::exception& e)
	{
		LOG_ERROR(e.what());
		return nullptr;
	}
}
void gih_rs_destroy(GihRendererSystem* h) { delete h; }
GiContext* gih_rs_device(GihRendererSystem* h) { return h->rs->GetDevice(); }
void gih_rs_set_scene(GihRendererSystem* h, GihScene* s) { h->rs->SetScene(s->scene); }
void gih_rs_set_screen_size(GihRendererSystem* h, int w, int hgt)
{
	h->rs->SetScreenSize(ei::IVec2(w, hgt));
	// application.cpp:128-131: the aspect ratio follows the resolution
	Camera cam = h->rs->GetCamera();
	cam.SetAspectRatio(float(w) / float(hgt));
	h->rs->SetCamera(cam);
}
void gih_rs_set_camera(GihRendererSystem* h, const float pos[3], const float lookat[3], float yfov)
{
	Camera cam = h->rs->GetCamera();
	cam.SetPosition(ei::Vec3(pos[0], pos[1], pos[2]));
	cam.SetLookAt(ei::Vec3(lookat[0], lookat[1], lookat[2]));
	cam.SetVFov(yfov);
	h->rs->SetCamera(cam);
}
int gih_rs_set_renderer(GihRendererSystem* h, int renderer)
{
	switch (renderer)
	{
	case GI_RENDERER_PT: h->rs->SetRenderer<Pathtracer>(); return 0;
	case GI_RENDERER_LPT: h->rs->SetRenderer<LightPathtracer>(); return 0;
	case GI_RENDERER_BPT: h->rs->SetRenderer<BidirectionalPathtracer>(); return 0;
	case GI_RENDERER_WHITTED: h->rs->SetRenderer<WhittedRayTracer>(); return 0;
	}
	LOG_ERROR("Unknown renderer index " << renderer << " (0 PT, 1 LPT, 2 BPT, 3 Whitted)");
	return 1;
}
const char* gih_rs_renderer_name(GihRendererSystem* h)
{
	static thread_local std::string name;
	name = h->rs->GetActiveRenderer() ? h->rs->GetActiveRenderer()->GetName() : "";
	return name.c_str();
}
void gih_rs_set_shard(GihRendererSystem* h, unsigned rank, unsigned world) { h->rs->SetIterationShard(rank, world); }
void gih_rs_set_synchronous(GihRendererSystem* h, int sync) { h->rs->SetSynchronousDraw(sync != 0); }
void gih_rs_draw(GihRendererSystem* h) { h->rs->Draw(); }
uint64_t gih_rs_iteration_count(GihRendererSystem* h) { return h->rs->GetIterationCount(); }
uint64_t gih_rs_render_time_ms(GihRendererSystem* h) { return h->rs->GetRenderTime(); }
uint32_t gih_rs_frame_seed(GihRendererSystem* h) { return h->rs->GetFrameSeed(); }
uint64_t gih_rs_global_iteration(GihRendererSystem* h) { return h->rs->GetGlobalIterationIndex(); }
int gih_rs_read_backbuffer(GihRendererSystem* h, float* rgba) { return h->rs->GetBackbuffer().ReadImage(rgba) ? 0 : 1; }
void gih_rs_set_streaming(GihRendererSystem* h, int streaming) { h->rs->SetStreaming(streaming != 0); }
void gih_rs_set_device_light_sampling(GihRendererSystem* h, int on_device) { h->rs->SetDeviceLightSampling(on_device != 0); }
int gih_rs_read_backbuffer_async(GihRendererSystem* h, float* rgba, uint32_t* iterations_in_flight)
{
	return h->rs->GetBackbuffer().ReadImageAsync(rgba, iterations_in_flight) ? 0 : 1;
}
int gih_rs_wait_reads(GihRendererSystem* h) { return h->rs->GetBackbuffer().WaitReads() ? 0 : 1; }
int gih_rs_read_backbuffer_reduced(GihRendererSystem* h, float* rgba_rows) { return h->rs->GetBackbuffer().ReadImageReduced(rgba_rows) ? 0 : 1; }
int gih_rs_read_backbuffer_reduced_async(GihRendererSystem* h, float* rgba_rows, uint32_t* iterations_in_flight)
{
	return h->rs->GetBackbuffer().ReadImageReducedAsync(rgba_rows, iterations_in_flight) ? 0 : 1;
}
uint32_t gih_rs_num_light_samples(GihRendererSystem* h) { return h->rs->GetNumInitialLightSamples(); }
void gih_rs_set_num_light_samples(GihRendererSystem* h, uint32_t n) { h->rs->SetNumInitialLightSamples(n); }
void gih_rs_light_samples(GihRendererSystem* h, GiLightSample* out)
{
	const auto& v = h->rs->GetInitialLightSamples();
	if (!v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(GiLightSample));
}
void gih_rs_camera(GihRendererSystem* h, GiCamera* out)
{
	const Camera& cam = h->rs->GetCamera();
	float p[3] = {cam.GetPosition().x, cam.GetPosition().y, cam.GetPosition().z};
	float l[3] = {cam.GetLookAt().x, cam.GetLookAt().y, cam.GetLookAt().z};
	float u[3] = {cam.GetUp().x, cam.GetUp().y, cam.GetUp().z};
	gih_camera_params(p, l, u, cam.GetVFov(), cam.GetAspectRatio(), h->rs->GetBackbuffer().GetWidth(), h->rs->GetBackbuffer().GetHeight(), out);
}
void gih_rs_renderer_info(GihRendererSystem* h, int32_t* num_rays_per_light_sample, uint32_t* light_cache_capacity, int32_t* avg_light_path_length)
{
	*num_rays_per_light_sample = 0; *light_cache_capacity = 0; *avg_light_path_length = 0;
	Renderer* r = h->rs->GetActiveRenderer();
	if (auto* l = dynamic_cast<LightPathtracer*>(r)) *num_rays_per_light_sample = l->GetNumRaysPerLightSample();
	if (auto* b = dynamic_cast<BidirectionalPathtracer*>(r))
	{
		*num_rays_per_light_sample = b->GetNumRaysPerLightSample();
		*light_cache_capacity = b->GetLightCacheCapacity();
		*avg_light_path_length = b->GetAverageLightPathLength();
	}
}

// ---- script front end (application.hpp, scriptprocessing.hpp)
struct GihApplication { std::unique_ptr<Application> app; };

GihApplication* gih_app_create(const char* script_path, int cuda_device)
{
	gi::LastHostError().clear();
	try
	{
		std::unique_ptr<GihApplication> h(new GihApplication());
		h->app.reset(new Application(script_path ? script_path : "", cuda_device));
		return h.release();
	}
	catch (const std::exception& e)
	{
		LOG_ERROR("gih_app_create: " << e.what());
		return nullptr;
	}
}
void gih_app_destroy(GihApplication* h) { delete h; }
void gih_app_command(GihApplication* h, const char* line) { h->app->Command(line); }
void gih_app_run(GihApplication* h, uint64_t max_iterations) { h->app->Run(static_cast<unsigned long>(max_iterations)); }
uint64_t gih_app_iteration_count(GihApplication* h) { return h->app->GetRendererSystem().GetIterationCount(); }
int gih_app_has_shut_down(GihApplication* h) { return h->app->HasShutDown() ? 1 : 0; }
float gih_app_last_mse(GihApplication* h) { return h->app->GetLastMSE(); }
const char* gih_app_last_screenshot(GihApplication* h) { return h->app->GetLastScreenshot().c_str(); }
GiContext* gih_app_device(GihApplication* h) { return h->app->GetRendererSystem().GetDevice(); }

// One command line through ScriptProcessing::ParseCommand (script-file rules) against the global parameter table; returns 1 if the
// command asks the processor to stop (waitSeconds / waitIterations / a registered pause command).
int gih_config_parse(const char* line)
{
	static ScriptProcessing processor;
	bool stop = false;
	processor.ParseCommand(line, true, stop);
	return stop ? 1 : 0;
}
int gih_config_add(const char* name, const char* comma_separated_defaults)
{
	try
	{
		GlobalConfig::ParameterType v;
		std::string s = comma_separated_defaults ? comma_separated_defaults : "";
		size_t a = 0;
		while (!s.empty() && a <= s.size())
		{
			size_t b = s.find(',', a);
			v.push_back(Variant(s.substr(a, b == std::string::npos ? std::string::npos : b - a)));
			if (b == std::string::npos) break;
			a = b + 1;
		}
		GlobalConfig::AddParameter(name, v);
		return 0;
	}
	catch (const std::invalid_argument& e) { LOG_ERROR(e.what()); return 1; }
}
void gih_config_remove(const char* name) { GlobalConfig::RemoveParameter(name); }
int gih_config_count(const char* name)
{
	try { return int(GlobalConfig::GetParameter(name).size()); }
	catch (const std::invalid_argument&) { return -1; }
}
const char* gih_config_get(const char* name, int index)
{
	static thread_local std::string out;
	try
	{
		GlobalConfig::ParameterType v = GlobalConfig::GetParameter(name);
		if (index < 0 || size_t(index) >= v.size()) return nullptr;
		out = v[size_t(index)].As<std::string>();
		return out.c_str();
	}
	catch (const std::invalid_argument&) { return nullptr; }
}

// ---- PFM
int gih_load_image(const char* path, unsigned char* rgba, int capacity_pixels, int* w, int* h)
{
	gi::LastHostError().clear();
	std::vector<unsigned char> px;
	std::string err;
	int iw = 0, ih = 0;
	if (!gi::LoadImageRGBA8(path, iw, ih, px, &err)) { LOG_ERROR(err); return 1; }
	if (w) *w = iw;
	if (h) *h = ih;
	if (capacity_pixels == 0) return 0;
	if (size_t(capacity_pixels) < size_t(iw) * ih || !rgba) return 2;
	std::memcpy(rgba, px.data(), px.size());
	return 0;
}

int gih_write_pfm(const char* path, const float* rgba, int w, int h, uint32_t divisor)
{
	return WritePfm(reinterpret_cast<const ei::Vec4*>(rgba), ei::IVec2(w, h), path, divisor) ? 0 : 1;
}
int gih_load_pfm(const char* path, float* rgb, int capacity_pixels, int* w, int* h)
{
	ei::IVec2 size;
	std::unique_ptr<ei::Vec3[]> img = LoadPfm(path, size);
	if (!img) return 1;
	*w = size.x; *h = size.y;
	if (size.x * size.y > capacity_pixels) return 2;
	std::memcpy(rgb, img.get(), size_t(size.x) * size.y * sizeof(ei::Vec3));
	return 0;
}

} // extern "C"
