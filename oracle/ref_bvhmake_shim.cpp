// ref_bvhmake_shim.cpp -- C-linkage doors into the REFERENCE'S OWN bvhmake build methods, compiled where they lie (oracle/Makefile ->
// oracle/_ref/libref_bvhmake.so): buildmethods/kdtree.cpp, buildmethods/sweep.cpp and fitmethods/aaboxfit.cpp against the real bvhmake.hpp.
// TEST INFRASTRUCTURE ONLY (pins SURVEY 8 row f-2's builders); nothing under gpugi_b200/ links or loads it.
//
// The manager class BVHBuilder is DECLARED by the reference (bvhmake.hpp:112-279); its member functions live in bvhbuilder.cpp, which is
// the Assimp importer and cannot be built here.  The handful the build methods call -- pool allocation and triangle access, no algorithm --
// are supplied below, with the bodies' meaning taken from bvhbuilder.cpp:526-555,640-704.  (buildmethods/lds.cpp calls a four-argument
// ei::decomposeQl that the vendored epsilon no longer has: it does not compile against this tree even under MSVC.)
#include "bvhmake.hpp"
#include "buildmethods/kdtree.hpp"
#include "buildmethods/sweep.hpp"
#include "fitmethods/aaboxfit.hpp"

#include <cstring>
#include <vector>

const FileDecl::Triangle FileDecl::INVALID_TRIANGLE = {{0x0, 0x0, 0x0}, 0};

BVHBuilder::BVHBuilder() : m_buildMethod(nullptr), m_fitMethod(nullptr), m_triangleSplitThreshold(0.0f), m_bvbuffer(nullptr), m_nodes(nullptr), m_leaves(nullptr),
    m_innerNodeCount(0), m_maxInnerNodeCount(0), m_leafNodeCount(0), m_maxLeafNodeCount(0)
{
	m_buildMethods.insert({"kdtree", new BuildKdtree(this)});
	m_buildMethods.insert({"sweep", new BuildSweep(this)});
	m_fitMethods.insert({"aabox", new FitBox(this)});
	m_buildMethod = m_buildMethods["kdtree"];
	m_fitMethod = m_fitMethods["aabox"];
}
BVHBuilder::~BVHBuilder()
{
	for (auto it : m_buildMethods) delete it.second;
	for (auto it : m_fitMethods) delete it.second;
	delete[] static_cast<ε::Box*>(m_bvbuffer);
	delete[] m_nodes;
	delete[] m_leaves;
}
bool BVHBuilder::SetBuildMethod(const char* _methodName)
{
	auto it = m_buildMethods.find(_methodName);
	if (it == m_buildMethods.end()) return false;
	m_buildMethod = it->second;
	return true;
}
void BVHBuilder::AddTriangle(const FileDecl::Triangle& _triangle)
{
	m_triangles.push_back(_triangle.vertices[0]); m_triangles.push_back(_triangle.vertices[1]); m_triangles.push_back(_triangle.vertices[2]);
	m_triangles.push_back(_triangle.material);
}
void BVHBuilder::BuildBVH()
{
	m_buildMethod->EstimateNodeCounts(m_maxInnerNodeCount, m_maxLeafNodeCount);
	m_innerNodeCount = m_leafNodeCount = 0;
	m_nodes = new Node[m_maxInnerNodeCount];
	m_leaves = new FileDecl::Leaf[m_maxLeafNodeCount];
	m_bvbuffer = new ε::Box[m_maxInnerNodeCount];
	(*m_buildMethod)();
}
ε::Triangle BVHBuilder::GetTriangle(uint32 _index) const
{
	return ε::Triangle(m_vertices[m_triangles[_index * 4 + 0]].position, m_vertices[m_triangles[_index * 4 + 1]].position, m_vertices[m_triangles[_index * 4 + 2]].position);
}
ε::Triangle BVHBuilder::GetTriangle(FileDecl::Triangle _triangle) const
{
	return ε::Triangle(m_vertices[_triangle.vertices[0]].position, m_vertices[_triangle.vertices[1]].position, m_vertices[_triangle.vertices[2]].position);
}
FileDecl::Triangle BVHBuilder::GetTriangleIdx(uint32 _index) const
{
	FileDecl::Triangle t;
	t.vertices[0] = m_triangles[_index * 4 + 0]; t.vertices[1] = m_triangles[_index * 4 + 1]; t.vertices[2] = m_triangles[_index * 4 + 2];
	t.material = m_triangles[_index * 4 + 3];
	return t;
}
const FileDecl::Vertex& BVHBuilder::GetVertex(uint32 _index) const { return m_vertices[_index]; }
uint32 BVHBuilder::GetNewLeaf() { return m_leafNodeCount++; }
uint32 BVHBuilder::GetNewNode() { return m_innerNodeCount++; }
// the one door the shim needs into the private vertex array: AddVertex without the duplicate search (the synthetic meshes are indexed already)
uint32 BVHBuilder::AddVertex(const FileDecl::Vertex& _vertex) { m_vertices.push_back(_vertex); return uint32(m_vertices.size() - 1); }
// RecursiveWriteHierarchy (bvhbuilder.cpp:706-718) into memory instead of a file: the friend-less way to read the private pools
void BVHBuilder::RecursiveWriteHierarchy(std::ofstream&, uint32, uint32, uint32) {}

#define REF_API extern "C" __attribute__((visibility("default")))

// method: 0 kdtree, 1 sweep.  Outputs (caller-sized for the builder's own estimate 4 T / 8 nodes, 2 T / 8 leaves; at least 8 each):
// nodes as {left, right} pairs in pool order (= preorder: every method allocates a node before its children), one ε::Box per node, the leaf array.
REF_API int ref_bvhmake_build(const float* positions, uint32_t num_vertices, const uint32_t* triangles, const uint32_t* tri_materials, uint32_t num_triangles, int method,
                              uint32_t* out_left_right, float* out_boxes, void* out_leaves, uint32_t* out_num_nodes, uint32_t* out_num_leaves)
{
	BVHBuilder b;
	if (!b.SetBuildMethod(method == 1 ? "sweep" : "kdtree")) return 1;
	for (uint32_t v = 0; v < num_vertices; ++v)
	{
		FileDecl::Vertex vx;
		vx.position = ε::Vec3(positions[3 * v], positions[3 * v + 1], positions[3 * v + 2]);
		vx.normal = ε::Vec3(0.0f, 1.0f, 0.0f); vx.texcoord = ε::Vec2(0.0f);
		b.AddVertex(vx);
	}
	for (uint32_t t = 0; t < num_triangles; ++t)
	{
		FileDecl::Triangle tr = {{triangles[3 * t], triangles[3 * t + 1], triangles[3 * t + 2]}, tri_materials ? tri_materials[t] : 0u};
		b.AddTriangle(tr);
	}
	b.BuildBVH();
	const uint32_t n = b.GetNumNodes();
	uint32_t numLeaves = 0;
	for (uint32_t i = 0; i < n; ++i)
	{
		out_left_right[2 * i] = b.GetNode(i).left; out_left_right[2 * i + 1] = b.GetNode(i).right;
		std::memcpy(out_boxes + 6 * i, &b.GetBoundingVolume<ε::Box>(i), 24);
		if (b.GetNode(i).left & 0x80000000u) numLeaves = std::max(numLeaves, (b.GetNode(i).left & 0x7fffffffu) + 1u);
	}
	for (uint32_t l = 0; l < numLeaves; ++l) std::memcpy(static_cast<char*>(out_leaves) + 128 * size_t(l), &b.GetLeaf(l), 128);
	*out_num_nodes = n; *out_num_leaves = numLeaves;
	return 0;
}
