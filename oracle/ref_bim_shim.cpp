// ref_bim_shim.cpp -- C-linkage doors into the REFERENCE'S OWN dependencies/bim + dependencies/epsilon, compiled where they lie
// (oracle/Makefile -> oracle/_ref/libref_bim.so).  TEST INFRASTRUCTURE ONLY: tests/ and tests/golden/make_golden_bim.py call it to pin
// SURVEY 8 row a-15 (bim file format, SAH / kd hierarchy build, pointer remap, AABoxes) and the epsilon intersection routines against
// the reference's object code.  Nothing under gpugi_b200/ links or loads it.
//
// ref_bim_build follows the call sequence of the reference's converter (dependencies/bim/tools/tobim/src/tobim.cpp:300-343) without
// Assimp: materials and mesh arrays come from the caller.  tobim's removeRedundantVertices / computeTangentSpace are skipped (the
// synthetic meshes are already indexed and carry normals), everything else is the converter's order of calls.
#include "bim/bim.hpp"
#include <ei/3dintersection.hpp>

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {
struct RefModel
{
	std::unique_ptr<bim::BinaryModel> model;
	bim::Chunk* chunk = nullptr;
};
}

// method: 0 = Chunk::BuildMethod::SAH, 1 = KD_TREE.  obox: also computeBVHOBoxes (brute force, small meshes only).
// env_json_in: environment file to take materials from (may be NULL); material_names: TRIANGLE_MAT index -> name.
REF_API void* ref_bim_build(const float* positions, const float* normals, const float* texcoords, uint32_t num_vertices, const uint32_t* triangles,
                            const uint32_t* triangle_materials, uint32_t num_triangles, const char* env_json_in, const char* const* material_names,
                            uint32_t num_materials, int method, int obox)
{
	using namespace bim;
	Property::Val props = Property::Val(Property::POSITION | Property::TRIANGLE_IDX | Property::TRIANGLE_MAT);
	if (normals) props = Property::Val(props | Property::NORMAL);
	if (texcoords) props = Property::Val(props | Property::TEXCOORD0);
	RefModel* r = new RefModel;
	r->model.reset(new BinaryModel(props, ei::IVec3(1)));
	BinaryModel& model = *r->model;
	if (env_json_in) model.loadEnvironmentFile(env_json_in);
	model.setNumTrianglesPerLeaf(2);
	for (uint32_t i = 0; i < num_materials; ++i)
	{
		if (!model.getMaterial(std::string(material_names[i]))) model.addMaterial(Material(material_names[i], "legacy"));
		model.getUniqueMaterialIndex(material_names[i]);
	}
	model.makeChunkResident(ei::IVec3(0));
	Chunk& chunk = *model.getChunk(ei::IVec3(0));
	r->chunk = &chunk;
	for (uint32_t v = 0; v < num_vertices; ++v)
	{
		Chunk::FullVertex fv;
		fv.position = ei::Vec3(positions[3 * v], positions[3 * v + 1], positions[3 * v + 2]);
		if (normals) fv.normal = ei::Vec3(normals[3 * v], normals[3 * v + 1], normals[3 * v + 2]);
		if (texcoords) fv.texCoord0 = ei::Vec2(texcoords[2 * v], texcoords[2 * v + 1]);
		chunk.addVertex(fv);
	}
	for (uint32_t t = 0; t < num_triangles; ++t)
		chunk.addTriangle(ei::UVec3(triangles[3 * t], triangles[3 * t + 1], triangles[3 * t + 2]), triangle_materials ? triangle_materials[t] : 0u);
	model.refreshBoundingBox();
	chunk.buildHierarchy(method == 1 ? Chunk::BuildMethod::KD_TREE : Chunk::BuildMethod::SAH);
	chunk.computeBVHAABoxes();
	if (obox) chunk.computeBVHOBoxes();
	if (obox) model.setAccelerator(Property::OBOX_BVH);
	model.setAccelerator(Property::AABOX_BVH);
	return r;
}

// BinaryModel::storeEnvironmentFile + storeBinaryHeader + storeChunk (tobim.cpp:340-343)
REF_API int ref_bim_store(void* h, const char* json_path, const char* bim_path)
{
	RefModel* r = static_cast<RefModel*>(h);
	r->model->storeEnvironmentFile(json_path, bim_path);
	r->model->storeBinaryHeader(bim_path);
	r->model->storeChunk(bim_path, ei::IVec3(0));
	return 0;
}

// BinaryModel::load + makeChunkResident(0), as Scene::Scene does (gpugi/scene/scene.cpp:22-48)
REF_API void* ref_bim_load(const char* json_path, int obox)
{
	using namespace bim;
	RefModel* r = new RefModel;
	r->model.reset(new BinaryModel());
	const Property::Val required = Property::Val(Property::POSITION | Property::NORMAL | Property::TRIANGLE_IDX | Property::TRIANGLE_MAT | Property::HIERARCHY |
	                                             (obox ? Property::OBOX_BVH : Property::AABOX_BVH));
	if (!r->model->load(json_path, required, Property::Val(Property::TEXCOORD0 | Property::AABOX_BVH | Property::OBOX_BVH), false)) { delete r; return nullptr; }
	r->model->makeChunkResident(ei::IVec3(0));
	r->chunk = r->model->getChunk(ei::IVec3(0));
	return r;
}

REF_API void ref_bim_free(void* h) { delete static_cast<RefModel*>(h); }

// counts: [0] vertices [1] triangles [2] nodes [3] leaf slots (UVec4 entries) [4] tree levels [5] used materials [6] has normals [7] has texcoords [8] has oboxes [9] accelerator
REF_API void ref_bim_counts(void* h, uint32_t* out)
{
	RefModel* r = static_cast<RefModel*>(h);
	const bim::Chunk& c = *r->chunk;
	out[0] = c.getNumVertices(); out[1] = c.getNumTriangles(); out[2] = c.getNumNodes(); out[3] = c.getNumLeafNodes() * 2u; out[4] = c.getNumTreeLevels();
	out[5] = r->model->getNumUsedMaterials(); out[6] = c.getNormals() ? 1u : 0u; out[7] = c.getTexCoords0() ? 1u : 0u;
	out[8] = (r->model->getAccelerator() == bim::Property::OBOX_BVH || c.getHierarchyOBoxes()) ? 1u : 0u; out[9] = uint32_t(r->model->getAccelerator());
}

// which: 0 positions (12 B) 1 normals (12) 2 texcoords (8) 3 triangles (12) 4 triangle materials (4) 5 hierarchy {firstChild, escape} (8) 6 parents (4)
//        7 leaves UVec4 (16) 8 AABoxes (24) 9 OBoxes (40) 10 model bounding box (24).  Returns the number of bytes copied.
REF_API uint64_t ref_bim_get(void* h, int which, void* dst)
{
	RefModel* r = static_cast<RefModel*>(h);
	const bim::Chunk& c = *r->chunk;
	const void* src = nullptr; uint64_t n = 0;
	switch (which)
	{
	case 0: src = c.getPositions(); n = uint64_t(c.getNumVertices()) * 12; break;
	case 1: src = c.getNormals(); n = src ? uint64_t(c.getNumVertices()) * 12 : 0; break;
	case 2: src = c.getTexCoords0(); n = src ? uint64_t(c.getNumVertices()) * 8 : 0; break;
	case 3: src = c.getTriangles(); n = uint64_t(c.getNumTriangles()) * 12; break;
	case 4: src = c.getTriangleMaterials(); n = src ? uint64_t(c.getNumTriangles()) * 4 : 0; break;
	case 5: src = c.getHierarchy(); n = uint64_t(c.getNumNodes()) * 8; break;
	case 6: src = c.getHierarchyParents(); n = src ? uint64_t(c.getNumNodes()) * 4 : 0; break;
	case 7: src = c.getLeafNodes(); n = uint64_t(c.getNumLeafNodes()) * 2 * 16; break;
	case 8: src = c.getHierarchyAABoxes(); n = uint64_t(c.getNumNodes()) * 24; break;
	case 9: src = c.getHierarchyOBoxes(); n = uint64_t(c.getNumNodes()) * sizeof(ei::OBox); break;
	case 10: src = &r->model->getBoundingBox(); n = 24; break;
	}
	if (src && n && dst) std::memcpy(dst, src, n);
	return n;
}

// name of the material behind TRIANGLE_MAT index i (BinaryModel::getMaterial(uint), bim.hpp:76)
REF_API const char* ref_bim_material_name(void* h, uint32_t i)
{
	RefModel* r = static_cast<RefModel*>(h);
	const bim::Material* m = r->model->getMaterial(i);
	return m ? m->getName().c_str() : "";
}

// ---- epsilon's own ray / box / triangle routines (dependencies/epsilon/src/3dintersection.cpp:326-520), live instead of transcribed KATs
REF_API int ref_ei_ray_box(const float* o, const float* d, const float* bmin, const float* bmax, float* distance, float* distance_exit)
{
	ei::Ray ray(ei::Vec3(o[0], o[1], o[2]), ei::Vec3(d[0], d[1], d[2]));
	ei::Box box(ei::Vec3(bmin[0], bmin[1], bmin[2]), ei::Vec3(bmax[0], bmax[1], bmax[2]));
	// bit 0: intersects(ray, box) (3dintersection.cpp:326-349; its commented-out body :343-348 is the GLSL slab test verbatim);
	// bit 1: the (distance, distanceExit) overload, which has no `tmax >= 0` test after the z slab and therefore also reports boxes that lie
	// entirely behind the origin along z only (distance 0, negative exit)
	float t0 = 0.0f, t1 = 0.0f;
	const bool hit = ei::intersects(ray, box);
	const bool hitD = ei::intersects(ray, box, t0, t1);
	if (distance) *distance = t0;
	if (distance_exit) *distance_exit = t1;
	return (hit ? 1 : 0) | (hitD ? 2 : 0);
}

REF_API int ref_ei_ray_triangle(const float* o, const float* d, const float* p0, const float* p1, const float* p2, float* distance, float* bary)
{
	ei::Ray ray(ei::Vec3(o[0], o[1], o[2]), ei::Vec3(d[0], d[1], d[2]));
	ei::Triangle tri(ei::Vec3(p0[0], p0[1], p0[2]), ei::Vec3(p1[0], p1[1], p1[2]), ei::Vec3(p2[0], p2[1], p2[2]));
	float t = 0.0f; ei::Vec3 b(0.0f);
	const bool hit = ei::intersects(ray, tri, t, b);
	if (distance) *distance = t;
	if (bary) { bary[0] = b.x; bary[1] = b.y; bary[2] = b.z; }
	return hit ? 1 : 0;
}

// ---- epsilon's vector arithmetic under the two host formulas that feed the device (restated HERE, on epsilon's own types and functions, so
// that the rounding of ei::normalize / ei::cross / operator* is the reference's): Camera::ComputeCameraParams (gpugi/camera/camera.cpp:76-85;
// the file itself does not compile under GCC: utilities/variant.hpp needs MSVC's one-phase lookup) and the light-sample normal + position of
// LightSampler::GenerateRandomSamples (gpugi/scene/lightsampler.cpp:62-72).
REF_API void ref_ei_camera_params(const float* pos, const float* lookat, const float* up, float yfov, float aspect, float* u, float* v, float* w)
{
	const ei::Vec3 position(pos[0], pos[1], pos[2]), look(lookat[0], lookat[1], lookat[2]), upv(up[0], up[1], up[2]);
	ei::Vec3 cameraW = ei::normalize(look - position);
	ei::Vec3 cameraU = ei::normalize(ei::cross(cameraW, upv));
	ei::Vec3 cameraV = ei::normalize(ei::cross(cameraU, cameraW));
	float f = tanf(yfov / 2.0f * (ei::PI / 180.0f));
	cameraU *= f * aspect;
	cameraV *= f;
	for (int k = 0; k < 3; ++k) { u[k] = cameraU[k]; v[k] = cameraV[k]; w[k] = cameraW[k]; }
}

REF_API void ref_ei_light_sample_point(const float* v0, const float* v1, const float* v2, float alpha, float beta, float gamma, float bias, float* position, float* normal)
{
	const ei::Vec3 a(v0[0], v0[1], v0[2]), b(v1[0], v1[1], v1[2]), c(v2[0], v2[1], v2[2]);
	ei::Vec3 p = a * alpha + b * beta + c * gamma;
	ei::Vec3 n = ei::normalize(ei::cross(a - b, a - c));
	p += n * bias;
	for (int k = 0; k < 3; ++k) { position[k] = p[k]; normal[k] = n[k]; }
}
