// scene.cpp -- host mirror of gpugi/scene/scene.cpp.  Produces, on the CPU, exactly the arrays the
// reference uploads to GL, in the layout gi_upload_scene() expects.
#include "scene.hpp"
#include "imageio.hpp"
#include "logger.hpp"

#include <cmath>
#include <cstring>
#include <fstream>

using namespace gi::bim;

namespace gi { std::string& LastHostError() { static std::string s; return s; } }

Scene::Scene(const std::string& _file, ei::Types3D _bvhType) : m_model(new BinaryModel()), m_bvhType(_bvhType)
{
	size_t p = _file.find_last_of("/\\");
	m_sourceDirectory = p == std::string::npos ? std::string() : _file.substr(0, p + 1);
	{
		std::ifstream probe(_file);
		if (!probe) LOG_ERROR("No scene description file '" + _file + "' found.");
	}
	std::string err;
	if (!m_model->load(_file, &err))
	{
		LOG_ERROR("Failed to load scene " + _file + ": " + err);
		return;
	}
	Init();
}

Scene::Scene(std::unique_ptr<BinaryModel> _model, const std::string& _sourceDirectory, ei::Types3D _bvhType)
	: m_model(std::move(_model)), m_sourceDirectory(_sourceDirectory), m_bvhType(_bvhType)
{
	if (!m_sourceDirectory.empty() && m_sourceDirectory.back() != '/') m_sourceDirectory += '/';
	Init();
}

Scene::~Scene() {}

void Scene::Init()
{
	if (m_bvhType != ei::Types3D::BOX && m_bvhType != ei::Types3D::OBOX)
	{
		LOG_ERROR("Only AABox (bvhType=0) and OBox (bvhType=1) hierarchies exist (scene.cpp:95-118).");
		return;
	}
	Chunk& chunk = m_model->chunk();
	if (chunk.getNumTriangles() == 0) { LOG_ERROR("Scene has no triangles."); return; }
	// The reference expects the hierarchy in the file (built by tobim); build what is missing.
	if (!(chunk.getProperties() & Property::HIERARCHY) || chunk.getNumNodes() == 0)
		chunk.buildHierarchy(Chunk::BuildMethod::SAH);
	if (!(chunk.getProperties() & Property::AABOX_BVH) || chunk.getHierarchyAABoxes() == nullptr)
		chunk.computeBVHAABoxes();
	if (m_bvhType == ei::Types3D::OBOX && (!(chunk.getProperties() & Property::OBOX_BVH) || chunk.getHierarchyOBoxes() == nullptr))
		chunk.computeBVHOBoxes();
	UploadGeometry();
	UploadHierarchy(m_bvhType);
	for (uint32_t i = 0; i < m_model->getNumUsedMaterials(); ++i)
	{
		const Material* m = m_model->getMaterial(i);
		if (m) LoadMaterial(*m);
		else { LOG_ERROR("Material " << i << " is referenced by the geometry but not defined."); LoadMaterial(Material("missing", "legacy")); }
	}
	LoadLightSources();
	m_valid = true;
}

const char* Scene::GetBvhTypeDefineString() const
{
	switch (m_bvhType)
	{
	case ei::Types3D::BOX: return "#define AABOX_BVH\n";
	case ei::Types3D::OBOX: return "#define OBOX_BVH\n";
	}
	return nullptr;
}

ei::Box Scene::GetBoundingBox() const
{
	const gi::bim::Box& b = m_model->chunk().getBoundingBox();
	ei::Box r;
	r.min = ei::Vec3(b.min.x, b.min.y, b.min.z);
	r.max = ei::Vec3(b.max.x, b.max.y, b.max.z);
	return r;
}

// scene.cpp:75-90
void Scene::UploadGeometry()
{
	const Chunk& chunk = m_model->chunk();
	const uint32_t nv = chunk.getNumVertices();
	m_vertexInfo.resize(nv);
	m_vertexPositions.resize(size_t(nv) * 3);
	for (uint32_t v = 0; v < nv; ++v)
	{
		const gi::bim::Vec3 n = chunk.getNormals() ? chunk.getNormals()[v] : gi::bim::Vec3{0.0f, 0.0f, 0.0f};
		m_vertexInfo[v].normal_phi = atan2f(n.y, n.x);
		m_vertexInfo[v].normal_z = n.z;
		m_vertexInfo[v].texcoord[0] = chunk.getTexCoords0() ? chunk.getTexCoords0()[v].x : 0.0f;
		m_vertexInfo[v].texcoord[1] = chunk.getTexCoords0() ? chunk.getTexCoords0()[v].y : 0.0f;
		m_vertexPositions[3 * size_t(v)] = chunk.getPositions()[v].x;
		m_vertexPositions[3 * size_t(v) + 1] = chunk.getPositions()[v].y;
		m_vertexPositions[3 * size_t(v) + 2] = chunk.getPositions()[v].z;
	}
	// straight copy of the leaf array (scene.cpp:89)
	const uint32_t nl = chunk.getNumLeafNodes() * chunk.getNumTrianglesPerLeaf();
	m_leafTriangles.resize(nl);
	for (uint32_t i = 0; i < nl; ++i)
	{
		const UVec4& t = chunk.getLeafNodes()[i];
		m_leafTriangles[i] = GiTriangle{{t.x, t.y, t.z}, t.w};
	}
}

// scene.cpp:92-130 (AABox branch; written through resize())
void Scene::UploadHierarchy(ei::Types3D _bvhType)
{
	const Chunk& chunk = m_model->chunk();
	m_oboxHierarchy.clear();
	if (_bvhType == ei::Types3D::OBOX)
	{
		// scene.cpp:107-118: center, sidesHalf, rotationInv = conjugate(orientation) + the tree pointers
		m_oboxHierarchy.resize(chunk.getNumNodes());
		for (uint32_t i = 0; i < chunk.getNumNodes(); ++i)
		{
			const gi::bim::OBox& b = chunk.getHierarchyOBoxes()[i];
			GiOBoxNode& n = m_oboxHierarchy[i];
			n.center[0] = b.center.x; n.center[1] = b.center.y; n.center[2] = b.center.z;
			n.sides_half[0] = b.halfSides.x; n.sides_half[1] = b.halfSides.y; n.sides_half[2] = b.halfSides.z;
			n.rotation_inv[0] = -b.orientation.i; n.rotation_inv[1] = -b.orientation.j; n.rotation_inv[2] = -b.orientation.k; n.rotation_inv[3] = b.orientation.r;
			n.escape = chunk.getHierarchy()[i].escape;
			n.first_child = chunk.getHierarchy()[i].firstChild;
		}
	}
	m_hierarchy.resize(chunk.getNumNodes());
	for (uint32_t i = 0; i < chunk.getNumNodes(); ++i)
	{
		const gi::bim::Box& b = chunk.getHierarchyAABoxes()[i];
		GiNode& n = m_hierarchy[i];
		n.min[0] = b.min.x; n.min[1] = b.min.y; n.min[2] = b.min.z;
		n.max[0] = b.max.x; n.max[1] = b.max.y; n.max[2] = b.max.z;
		n.escape = chunk.getHierarchy()[i].escape;
		n.first_child = chunk.getHierarchy()[i].firstChild;
	}
}

// scene.cpp:132-180
void Scene::LoadMaterial(const Material& _material)
{
	GiMaterial mat;
	std::memset(&mat, 0, sizeof(mat));
	ei::Vec3 emissivity(0.0f);
	mat.refraction_index_avg = _material.get("refractionIndex", 1.5f);
	gi::bim::Vec3 f0 = _material.get("specularColor", gi::bim::Vec3{0.03f, 0.03f, 0.03f});
	mat.fresnel0[0] = f0.x; mat.fresnel0[1] = f0.y; mat.fresnel0[2] = f0.z;
	for (int i = 0; i < 3; ++i) mat.fresnel1[i] = 1.0f - mat.fresnel0[i];
	float roughness = _material.get("roughness", 1.0f);
	float exponent = 1.0f / (roughness * roughness + 1e-20f);
	if (_material.getTexture("albedo")) mat.diffuse_tex = GetBindlessHandle(*_material.getTexture("albedo"));
	else
	{
		gi::bim::Vec3 d = _material.get("albedo", gi::bim::Vec3{0.5f, 0.5f, 0.5f});
		mat.diffuse_tex = GetBindlessHandle(ei::Vec3(d.x, d.y, d.z), false);
	}
	if (_material.getTexture("opacity")) mat.opacity_tex = GetBindlessHandle(*_material.getTexture("opacity"));
	else mat.opacity_tex = GetBindlessHandle(ei::Vec3(_material.get("opacity", 1.0f)), false);
	if (_material.getTexture("reflectiveness")) mat.reflectiveness_tex = GetBindlessHandle(*_material.getTexture("reflectiveness"));
	else
	{
		std::array<float, 4> r = _material.get("reflectiveness", std::array<float, 4>{1.0f, 1.0f, 1.0f, exponent});
		mat.reflectiveness_tex = GetBindlessHandle(ei::Vec4(r[0], r[1], r[2], r[3]), true);
	}
	if (_material.getTexture("emissivity")) mat.emissivity_tex = GetBindlessHandle(*_material.getTexture("emissivity"));
	else
	{
		gi::bim::Vec3 e = _material.get("emissivity", gi::bim::Vec3{0.0f, 0.0f, 0.0f});
		emissivity = ei::Vec3(e.x, e.y, e.z);
		mat.emissivity_tex = GetBindlessHandle(emissivity, true);
	}
	m_materials.push_back(mat);
	m_emissivity.push_back(emissivity);
}

int32_t Scene::AddTexture(const std::string& key, int w, int h, int format, std::vector<unsigned char>&& bytes)
{
	std::unique_ptr<TextureStorage> st(new TextureStorage());
	st->bytes = std::move(bytes);
	GiTexture t;
	t.width = w; t.height = h; t.format = format; t.reserved = 0; t.pixels = st->bytes.data();
	m_textures.push_back(t);
	m_textureStorage.push_back(std::move(st));
	int32_t idx = int32_t(m_textures.size() - 1);
	m_textureIndex[key] = idx;
	return idx;
}

// scene.cpp:182-199: file texture -> SRGB8_ALPHA8, one level (texture2d.cpp:31-52: stbi_load(file, ..., 4)).  The decoder is imageio.cpp
// (PNG, JPEG, TGA, BMP, PNM, GITX); a file that cannot be read becomes a 1 x 1 mid-grey texture, with an error in the log.
int32_t Scene::GetBindlessHandle(const std::string& _name)
{
	auto it = m_textureIndex.find(_name);
	if (it != m_textureIndex.end()) return it->second;
	const std::string path = m_sourceDirectory + _name;
	std::vector<unsigned char> rgba;
	int w = 0, h = 0;
	std::string err;
	if (!gi::LoadImageRGBA8(path, w, h, rgba, &err))
	{
		LOG_ERROR("Error loading texture " + err + "; using mid grey.");
		w = h = 1;
		rgba = {128, 128, 128, 255};
	}
	return AddTexture(_name, w, h, GI_TEX_SRGB8_ALPHA8, std::move(rgba));
}

namespace {
inline unsigned char toUnorm8(float f)
{ // GL 4.5 spec 2.3.5.2 float -> unorm8
	f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);
	return static_cast<unsigned char>(f * 255.0f + 0.5f);
}
}

// scene.cpp:201-220: 1x1 RGB8 (ldr) or RGB32F (hdr) constant
int32_t Scene::GetBindlessHandle(const ei::Vec3& _data, bool _hdr)
{
	std::string genName = "gen" + std::to_string(_data.x) + '_' + std::to_string(_data.y) + '_' + std::to_string(_data.z) + (_hdr ? "_1" : "_0");
	auto it = m_textureIndex.find(genName);
	if (it != m_textureIndex.end()) return it->second;
	if (_hdr)
	{
		float px[4] = {_data.x, _data.y, _data.z, 1.0f};
		std::vector<unsigned char> bytes(sizeof(px));
		std::memcpy(bytes.data(), px, sizeof(px));
		return AddTexture(genName, 1, 1, GI_TEX_RGBA32F, std::move(bytes));
	}
	std::vector<unsigned char> bytes = {toUnorm8(_data.x), toUnorm8(_data.y), toUnorm8(_data.z), 255};
	return AddTexture(genName, 1, 1, GI_TEX_RGBA8, std::move(bytes));
}

// scene.cpp:222-242: 1x1 RGBA8 / RGBA32F constant
int32_t Scene::GetBindlessHandle(const ei::Vec4& _data, bool _hdr)
{
	std::string genName = "gen" + std::to_string(_data.x) + '_' + std::to_string(_data.y) + '_' + std::to_string(_data.z) + '_' + std::to_string(_data.w) +
	                      (_hdr ? "_1" : "_0");
	auto it = m_textureIndex.find(genName);
	if (it != m_textureIndex.end()) return it->second;
	if (_hdr)
	{
		float px[4] = {_data.x, _data.y, _data.z, _data.w};
		std::vector<unsigned char> bytes(sizeof(px));
		std::memcpy(bytes.data(), px, sizeof(px));
		return AddTexture(genName, 1, 1, GI_TEX_RGBA32F, std::move(bytes));
	}
	std::vector<unsigned char> bytes = {toUnorm8(_data.x), toUnorm8(_data.y), toUnorm8(_data.z), toUnorm8(_data.w)};
	return AddTexture(genName, 1, 1, GI_TEX_RGBA8, std::move(bytes));
}

// scene.cpp:244-293
void Scene::LoadLightSources()
{
	++m_lightTablesVersion;
	const Chunk& chunk = m_model->chunk();
	m_totalAreaLightFlux = 0.0f;
	m_lightTriangles.clear();
	m_lightSummedArea.clear();
	float sum = 0.0f;
	for (uint32_t i = 0; i < chunk.getNumTriangles(); ++i)
	{
		const UVec3& tri = chunk.getTriangles()[i];
		uint32_t matIdx = chunk.getTriangleMaterials() ? chunk.getTriangleMaterials()[i] : 0u;
		if (matIdx >= m_emissivity.size()) continue;
		if (tri.x != tri.y && (m_emissivity[matIdx] != ei::Vec3(0.0f)))
		{
			LightTriangle lightSource;
			lightSource.luminance = m_emissivity[matIdx];
			const gi::bim::Vec3* P = chunk.getPositions();
			lightSource.triangle.v0 = ei::Vec3(P[tri.x].x, P[tri.x].y, P[tri.x].z);
			lightSource.triangle.v1 = ei::Vec3(P[tri.y].x, P[tri.y].y, P[tri.y].z);
			lightSource.triangle.v2 = ei::Vec3(P[tri.z].x, P[tri.z].y, P[tri.z].z);
			m_lightTriangles.push_back(lightSource);
			float area = ei::surface(lightSource.triangle);
			m_totalAreaLightFlux += dot(ei::Vec3(0.2126f, 0.7152f, 0.0722f), lightSource.luminance) * area * ei::PI;
			sum += area;
			m_lightSummedArea.push_back(sum);
		}
	}
	if (!m_lightSummedArea.empty())
	{
		m_lightAreaSum = m_lightSummedArea.back();
		for (size_t i = 0; i < m_lightSummedArea.size(); ++i)
			m_lightSummedArea[i] /= m_lightAreaSum;
	}
}

// scene.cpp:295-307
void Scene::ComputePointLightTable()
{
	++m_lightTablesVersion;
	m_totalPointLightFlux = 0.0f;
	m_pointLightSummedFlux.clear();
	for (auto& light : m_pointLights)
	{
		float flux = dot(ei::Vec3(0.2126f, 0.7152f, 0.0722f), light.intensity) * 4 * ei::PI;
		m_totalPointLightFlux += logf(1.0f + flux);
		m_pointLightSummedFlux.push_back(m_totalPointLightFlux);
	}
	for (size_t i = 0; i < m_pointLightSummedFlux.size(); ++i)
		m_pointLightSummedFlux[i] /= m_totalPointLightFlux;
}

bool Scene::SetPointLight(size_t _index, const PointLight& _light)
{
	if (_index >= m_pointLights.size()) return false;
	m_pointLights[_index] = _light;
	ComputePointLightTable();
	return true;
}

GiSceneDesc Scene::GetSceneDesc() const
{
	GiSceneDesc d;
	std::memset(&d, 0, sizeof(d));
	d.nodes = m_hierarchy.data(); d.num_nodes = uint32_t(m_hierarchy.size());
	d.leaf_triangles = m_leafTriangles.data(); d.num_leaf_triangles = uint32_t(m_leafTriangles.size());
	d.positions = m_vertexPositions.data(); d.vertex_info = m_vertexInfo.data(); d.num_vertices = uint32_t(m_vertexInfo.size());
	d.materials = m_materials.data(); d.num_materials = uint32_t(m_materials.size());
	d.textures = m_textures.data(); d.num_textures = uint32_t(m_textures.size());
	d.num_tree_levels = m_valid ? m_model->chunk().getNumTreeLevels() : 0;
	d.obox_nodes = m_oboxHierarchy.empty() ? nullptr : m_oboxHierarchy.data();
	return d;
}
