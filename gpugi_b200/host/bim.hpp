// bim.hpp -- scene container compatible with dependencies/bim of the reference:
// the JSON environment file (materials) + the chunked, zlib-deflated `.bim` binary, and the
// CPU hierarchy build that produces the firstChild/escape preorder layout the tracer walks.
//
// Written from the format description, not from the reference sources:
//   section header / types      dependencies/bim/src/bim_io.cpp:36-56, include/bim/chunk.hpp:9-37
//   load / makeChunkResident    bim_io.cpp:78-166, 238-317
//   storeBinaryHeader / Chunk   bim_io.cpp:174-203, 341-436
//   SAH sweep build             bim_sahspilt.cpp:27-151
//   pointer remap, AABoxes      bim_hierarchy.cpp:8-50, 238-252
// Only what the in-scope renderers consume is implemented (positions, normals, texcoord0,
// triangles, triangle materials, hierarchy, parents, leaves, AABoxes); other sections are skipped
// on load.  One chunk per file (the reference's Scene only ever loads chunk (0,0,0), scene.cpp:47-48).
#pragma once

#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace gi {
namespace bim {

struct Vec2 { float x, y; };
struct Vec3 { float x, y, z; };
struct UVec3 { uint32_t x, y, z; };
struct UVec4 { uint32_t x, y, z, w; };
struct Box { Vec3 min, max; };
struct Quaternion { float i, j, k, r; };                                   // ei::TQuaternion: imaginary part first (vector.hpp:867)
struct OBox { Vec3 center; Vec3 halfSides; Quaternion orientation; };      // ei::OBox (3dtypes.hpp:126-131), 40 bytes, the OBOX_BVH section element

// chunk.hpp:9-37 (subset + the extra section ids of bim_io.cpp:44-49)
namespace Property {
enum Val : uint32_t {
	DONT_CARE = 0,
	POSITION = 0x00000001,
	NORMAL = 0x00000002,
	TEXCOORD0 = 0x00000020,
	TRIANGLE_IDX = 0x00010000,
	TRIANGLE_MAT = 0x00020000,
	AABOX_BVH = 0x01000000,
	OBOX_BVH = 0x02000000,
	HIERARCHY = 0x08000000,
	NDF_SGGX = 0x10000000,
};
}
enum : uint32_t {
	META_SECTION = 0x0,
	CHUNK_SECTION = 0x3,
	MATERIAL_REFERENCE = 0x5,
	CHUNK_META_SECTION = 0x6,
	HIERARCHY_PARENTS = 0x08000001,
	HIERARCHY_LEAVES = 0x08000002,
};

// chunk.hpp:39-44
struct Node {
	uint32_t firstChild; // MSB set: leaf, low bits = leaf index
	uint32_t escape;     // next node in preorder if the subtree is skipped, 0 = end
};

// include/bim/material.hpp + bim_material.cpp: named float vectors and texture names.
class Material {
public:
	struct MultiValue { std::array<float, 4> values; int numComponents; };

	Material() = default;
	Material(std::string name, std::string type) : m_name(std::move(name)), m_type(std::move(type)) {}

	const std::string& getName() const { return m_name; }
	const std::string& getType() const { return m_type; }
	void setType(std::string t) { m_type = std::move(t); }

	/// nullptr if there is no texture of that name (bim_material.cpp:17-23)
	const std::string* getTexture(const std::string& name) const;
	/// Value lookups fill missing components from the default (bim_material.cpp:25-74)
	float get(const std::string& name, float def) const;
	Vec3 get(const std::string& name, const Vec3& def) const;
	std::array<float, 4> get(const std::string& name, const std::array<float, 4>& def) const;
	bool has(const std::string& name) const;

	void set(const std::string& name, const float* v, int n);
	void setTexture(const std::string& name, std::string file) { m_textureNames[name] = std::move(file); }

	const std::map<std::string, MultiValue>& values() const { return m_values; }
	const std::map<std::string, std::string>& textures() const { return m_textureNames; }

private:
	std::string m_name, m_type;
	std::map<std::string, MultiValue> m_values;
	std::map<std::string, std::string> m_textureNames;
};

class BinaryModel;

class Chunk {
public:
	enum class BuildMethod { KD_TREE, SAH };

	uint32_t getNumVertices() const { return uint32_t(m_positions.size()); }
	const Vec3* getPositions() const { return m_positions.data(); }
	const Vec3* getNormals() const { return m_normals.empty() ? nullptr : m_normals.data(); }
	const Vec2* getTexCoords0() const { return m_texCoords0.empty() ? nullptr : m_texCoords0.data(); }
	uint32_t getNumTriangles() const { return uint32_t(m_triangles.size()); }
	const UVec3* getTriangles() const { return m_triangles.data(); }
	const uint32_t* getTriangleMaterials() const { return m_triangleMaterials.empty() ? nullptr : m_triangleMaterials.data(); }
	uint32_t getNumNodes() const { return uint32_t(m_hierarchy.size()); }
	uint32_t getNumTreeLevels() const { return m_numTreeLevels; }
	uint32_t getNumLeafNodes() const { return uint32_t(m_hierarchyLeaves.size() / m_numTrianglesPerLeaf); }
	const Node* getHierarchy() const { return m_hierarchy.data(); }
	const uint32_t* getHierarchyParents() const { return m_hierarchyParents.data(); }
	const Box* getHierarchyAABoxes() const { return m_aaBoxes.empty() ? nullptr : m_aaBoxes.data(); }
	const OBox* getHierarchyOBoxes() const { return m_oBoxes.empty() ? nullptr : m_oBoxes.data(); }
	const UVec4* getLeafNodes() const { return m_hierarchyLeaves.data(); }
	uint32_t getNumTrianglesPerLeaf() const { return m_numTrianglesPerLeaf; }
	const Box& getBoundingBox() const { return m_boundingBox; }
	uint32_t getProperties() const { return m_properties; }

	/// Replace the mesh (invalidates any hierarchy).  normals / texcoords may be null.
	void setMesh(const Vec3* positions, const Vec3* normals, const Vec2* texcoords, uint32_t numVertices, const UVec3* triangles,
	             const uint32_t* triangleMaterials, uint32_t numTriangles);

	/// Chunk::buildHierarchy (bim_chunk.cpp:227-242)
	void buildHierarchy(BuildMethod method);
	/// Install a hierarchy built elsewhere (the device LBVH, gi_build_hierarchy) in the chunk's own format: nodes in preorder with
	/// child / escape links, parents, 2 slots per leaf, boxes, tree levels.
	void setHierarchy(const Node* nodes, const uint32_t* parents, const Box* boxes, uint32_t numNodes, const UVec4* leaves, uint32_t numLeafSlots, uint32_t numTreeLevels);
	/// Chunk::computeBVHAABoxes (bim_hierarchy.cpp:45-50)
	void computeBVHAABoxes();
	/// Chunk::computeBVHOBoxes (bim_hierarchy.cpp:54-96): one oriented box per node around all vertices below it, same recursion.
	/// The reference fits each box by brute force over all vertex triples (ei::OBox(points, n), O(n^4) on the convex set); here the
	/// orientation is the principal-axis frame of the vertices and the extent is ei::OBox(orientation, points, n) (3dtypes.cpp:205-225).
	/// Files written by the reference's tools carry their own boxes and are used as stored.
	void computeBVHOBoxes();
	/// Index ranges and array sizes of a chunk read from a file (the part of Scene::SanityCheck, scene.hpp:191, that keeps the host code
	/// inside its arrays): vertex indices, child / escape / parent links (preorder: a child and an escape target lie behind their node),
	/// leaf slots, one box and one parent per node.  false + *why for a file that breaks one of them.
	bool validate(std::string* why = nullptr) const;

private:
	friend class BinaryModel;
	uint32_t m_properties = 0;
	Box m_boundingBox = {{1e10f, 1e10f, 1e10f}, {-1e10f, -1e10f, -1e10f}};
	std::vector<Vec3> m_positions, m_normals;
	std::vector<Vec2> m_texCoords0;
	std::vector<UVec3> m_triangles;
	std::vector<uint32_t> m_triangleMaterials;
	std::vector<Node> m_hierarchy;
	std::vector<uint32_t> m_hierarchyParents;
	std::vector<UVec4> m_hierarchyLeaves;
	std::vector<Box> m_aaBoxes;
	std::vector<OBox> m_oBoxes;
	uint32_t m_numTrianglesPerLeaf = 2; // bim.cpp:14; BinaryModel::setNumTrianglesPerLeaf is a no-op (bim.hpp:91)
	uint32_t m_numTreeLevels = 0;

	void buildBVH_SAHsplit();
	void buildBVH_kdtree();
	uint32_t remapNodePointers(uint32_t self, uint32_t parent, uint32_t escape);
};

class BinaryModel {
public:
	/// BinaryModel::load (bim_io.cpp:78-166) + makeChunkResident(0) (:238-317).
	/// envFile: JSON with "scene" (bim file, relative), "accelerator", "materials".
	bool load(const std::string& envFile, std::string* error = nullptr);
	/// Parse only the JSON (materials); bim_io.cpp:168-171.
	bool loadEnvironmentFile(const std::string& envFile, std::string* error = nullptr);
	/// Same, from JSON text already in memory.
	bool loadEnvironmentText(const std::string& jsonText, std::string* error = nullptr);
	/// storeEnvironmentFile + storeBinaryHeader + storeChunk (bim_io.cpp:174-203, 363-436).
	bool store(const std::string& envFile, const std::string& bimFileName, std::string* error = nullptr) const;

	Chunk& chunk() { return m_chunk; }
	const Chunk& chunk() const { return m_chunk; }

	/// Material by TRIANGLE_MAT index (bim.hpp:76-78); nullptr if the name is unknown.
	const Material* getMaterial(uint32_t index) const;
	uint32_t getNumUsedMaterials() const { return uint32_t(m_materialIndirection.size()); }
	/// Appends to the indirection table and the material map; returns the TRIANGLE_MAT index.
	uint32_t addMaterial(const Material& m);
	/// Use the order of the JSON "materials" object as the TRIANGLE_MAT index table
	/// (what load() does for files without a MATERIAL_REFERENCE section, bim_io.cpp:158-163).
	void useJsonMaterialOrder() { m_materialIndirection = m_materialOrder; }
	/// Appends a material of the environment file to the TRIANGLE_MAT index table by name (BinaryModel::getUniqueMaterialIndex, bim.cpp:39-55);
	/// false if there is no such material.
	bool useMaterialName(const std::string& name)
	{
		if (m_materials.find(name) == m_materials.end()) return false;
		m_materialIndirection.push_back(name);
		return true;
	}
	const Box& getBoundingBox() const { return m_boundingBox; }
	uint32_t getNumTrianglesPerLeaf() const { return 2; }
	uint32_t getAccelerator() const { return m_accelerator; }

private:
	Chunk m_chunk;
	std::unordered_map<std::string, Material> m_materials;
	std::vector<std::string> m_materialIndirection;
	std::vector<std::string> m_materialOrder; // JSON order, used when the file has no MATERIAL_REFERENCE
	Box m_boundingBox = {{1e10f, 1e10f, 1e10f}, {-1e10f, -1e10f, -1e10f}};
	uint32_t m_accelerator = Property::DONT_CARE;
	std::string m_bimFile;
};

} // namespace bim
} // namespace gi
