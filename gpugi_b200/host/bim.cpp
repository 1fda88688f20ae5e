// bim.cpp -- see bim.hpp.  File format + CPU hierarchy build of the reference's scene container.
#include "bim.hpp"
#include "json.hpp"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <memory>

namespace gi {
namespace bim {

// ------------------------------------------------------------------ small math
namespace {
inline Vec3 add(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 sub(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 vmin(Vec3 a, Vec3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline Vec3 vmax(Vec3 a, Vec3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline float dot3(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross3(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Box triBox(Vec3 a, Vec3 b, Vec3 c) { return {vmin(vmin(a, b), c), vmax(vmax(a, b), c)}; }
inline Box uniteBox(const Box& a, const Box& b) { return {vmin(a.min, b.min), vmax(a.max, b.max)}; }
// ei::surface(Box), epsilon 3dtypes.inl:387-391
inline float surfaceArea(const Box& b)
{
	Vec3 s = sub(b.max, b.min);
	return 2.0f * (s.x * s.y + s.x * s.z + s.y * s.z);
}
inline float comp(const Vec3& v, int d) { return d == 0 ? v.x : (d == 1 ? v.y : v.z); }
inline float& comp(Vec3& v, int d) { return d == 0 ? v.x : (d == 1 ? v.y : v.z); }

struct Mat3 { float m[9]; }; // row major
inline Vec3 row(const Mat3& a, int r) { return {a.m[3 * r], a.m[3 * r + 1], a.m[3 * r + 2]}; }
inline Vec3 perpendicular(Vec3 v) // ei::perpendicular, vector.hpp:1878-1883
{
	return std::fabs(v.z) < std::fabs(v.x) ? Vec3{-v.y, v.x, 0.0f} : Vec3{0.0f, -v.z, v.y};
}
inline Vec3 normalized(Vec3 v)
{
	float l = std::sqrt(dot3(v, v));
	return {v.x / l, v.y / l, v.z / l};
}

// Closed-form spectral decomposition of a symmetric 3x3 matrix, eigenvalues descending, rows of Q
// are the eigenvectors.  Follows ei::decomposeQl (epsilon vector.hpp:2371-2457): trigonometric
// eigenvalues in double, eigenvectors from the better conditioned of two cross products.
void decomposeQl(const Mat3& A, Vec3 Q[3], Vec3& lambda)
{
	const float* a = A.m;
	double q = (double(a[0]) + double(a[4]) + double(a[8])) / 3.0;
	double p1 = double(a[1]) * double(a[1]) + double(a[2]) * double(a[2]) + double(a[5]) * double(a[5]);
	if (p1 <= q * 1.0e-15)
	{
		// (near) diagonal: sorting network 0-2, 0-1, 1-2 on the diagonal
		int i0 = 0, i1 = 1, i2 = 2;
		if (a[0] < a[8]) { i0 = 2; i2 = 0; }
		if (a[i0 * 4] < a[4]) { i1 = i0; i0 = 1; }
		if (a[i1 * 4] < a[i2 * 4]) { int t = i2; i2 = i1; i1 = t; }
		float l[3];
		l[i0] = a[0]; l[i1] = a[4]; l[i2] = a[8];
		lambda = {l[0], l[1], l[2]};
		Q[i0] = {1.0f, 0.0f, 0.0f};
		Q[i1] = {0.0f, 1.0f, 0.0f};
		Q[i2] = {0.0f, 0.0f, 1.0f};
		return;
	}
	double d0 = a[0] - q, d1 = a[4] - q, d2 = a[8] - q;
	double p2 = d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * p1;
	double p = std::sqrt(p2 / 6.0);
	double B[9];
	for (int i = 0; i < 9; ++i) B[i] = (double(a[i]) - ((i % 4 == 0) ? q : 0.0)) / p;
	double det2 = (B[4] * (B[0] * B[8] - B[2] * B[2]) + B[1] * B[5] * B[2] * 2.0 - B[1] * B[1] * B[8] - B[0] * B[5] * B[5]) / 2.0;
	double phi = std::acos(std::min(1.0, std::max(-1.0, det2))) / 3.0;
	lambda.x = float(q + p * 2.0 * std::cos(phi));
	lambda.z = float(q + p * 2.0 * std::cos(phi + 2 * 3.14159265358979323846 / 3.0));
	lambda.y = float(3 * q - lambda.x - lambda.z);

	auto shifted = [&](float l) {
		Mat3 M = A;
		M.m[0] -= l; M.m[4] -= l; M.m[8] -= l;
		return M;
	};
	bool deferQ0 = false;
	Mat3 M = shifted(lambda.x);
	Vec3 c1 = cross3(row(M, 0), row(M, 1));
	Vec3 c2 = cross3(row(M, 0), row(M, 2));
	float l1 = dot3(c1, c1), l2 = dot3(c2, c2);
	if (l1 + l2 == 0.0f) deferQ0 = true;
	else if (l1 > l2) { float s = std::sqrt(l1); Q[0] = {c1.x / s, c1.y / s, c1.z / s}; }
	else { float s = std::sqrt(l2); Q[0] = {c2.x / s, c2.y / s, c2.z / s}; }
	M = shifted(lambda.z);
	c1 = cross3(row(M, 1), row(M, 0));
	c2 = cross3(row(M, 2), row(M, 0));
	l1 = dot3(c1, c1); l2 = dot3(c2, c2);
	if (l1 + l2 == 0.0f) Q[2] = deferQ0 ? Vec3{0.0f, 0.0f, 1.0f} : normalized(perpendicular(Q[0]));
	else if (l1 > l2) { float s = std::sqrt(l1); Q[2] = {c1.x / s, c1.y / s, c1.z / s}; }
	else { float s = std::sqrt(l2); Q[2] = {c2.x / s, c2.y / s, c2.z / s}; }
	if (deferQ0) Q[0] = normalized(perpendicular(Q[2]));
	Q[1] = cross3(Q[0], Q[2]);
}
} // namespace

// ------------------------------------------------------------------ Material
const std::string* Material::getTexture(const std::string& name) const
{
	auto it = m_textureNames.find(name);
	return it == m_textureNames.end() ? nullptr : &it->second;
}
float Material::get(const std::string& name, float def) const
{
	auto it = m_values.find(name);
	return it == m_values.end() ? def : it->second.values[0];
}
Vec3 Material::get(const std::string& name, const Vec3& def) const
{
	auto it = m_values.find(name);
	if (it == m_values.end()) return def;
	float r[3] = {def.x, def.y, def.z};
	for (int i = 0; i < 3 && i < it->second.numComponents; ++i) r[i] = it->second.values[i];
	return {r[0], r[1], r[2]};
}
std::array<float, 4> Material::get(const std::string& name, const std::array<float, 4>& def) const
{
	auto it = m_values.find(name);
	if (it == m_values.end()) return def;
	std::array<float, 4> r = def;
	for (int i = 0; i < 4 && i < it->second.numComponents; ++i) r[i] = it->second.values[i];
	return r;
}
bool Material::has(const std::string& name) const { return m_values.count(name) || m_textureNames.count(name); }
void Material::set(const std::string& name, const float* v, int n)
{
	MultiValue mv{{0.0f, 0.0f, 0.0f, 0.0f}, std::min(n, 4)};
	for (int i = 0; i < mv.numComponents; ++i) mv.values[i] = v[i];
	m_values[name] = mv;
}

// ------------------------------------------------------------------ Chunk: mesh + hierarchy
void Chunk::setMesh(const Vec3* positions, const Vec3* normals, const Vec2* texcoords, uint32_t numVertices, const UVec3* triangles,
                    const uint32_t* triangleMaterials, uint32_t numTriangles)
{
	m_positions.assign(positions, positions + numVertices);
	m_properties = Property::POSITION | Property::TRIANGLE_IDX;
	if (normals) { m_normals.assign(normals, normals + numVertices); m_properties |= Property::NORMAL; } else m_normals.clear();
	if (texcoords) { m_texCoords0.assign(texcoords, texcoords + numVertices); m_properties |= Property::TEXCOORD0; } else m_texCoords0.clear();
	m_triangles.assign(triangles, triangles + numTriangles);
	if (triangleMaterials) { m_triangleMaterials.assign(triangleMaterials, triangleMaterials + numTriangles); m_properties |= Property::TRIANGLE_MAT; }
	else m_triangleMaterials.clear();
	m_hierarchy.clear(); m_hierarchyParents.clear(); m_hierarchyLeaves.clear(); m_aaBoxes.clear(); m_oBoxes.clear();
	m_numTreeLevels = 0;
	m_boundingBox = {{1e10f, 1e10f, 1e10f}, {-1e10f, -1e10f, -1e10f}};
	for (const Vec3& p : m_positions) { m_boundingBox.min = vmin(m_boundingBox.min, p); m_boundingBox.max = vmax(m_boundingBox.max, p); }
}

namespace {
// Shared by both builders: nodes are appended in preorder, the two children are parked in
// firstChild / escape until remapNodePointers() turns them into child / skip links.
struct BuildContext
{
	std::vector<Node>& nodes;
	std::vector<uint32_t>& parents;
	std::vector<UVec4>& leaves;
	const std::vector<Vec3>& positions;
	const std::vector<UVec3>& triangles;
	const std::vector<uint32_t>& materials;
	uint32_t trisPerLeaf;

	uint32_t newNode()
	{
		uint32_t idx = uint32_t(nodes.size());
		nodes.push_back({0u, 0u});
		parents.push_back(0u);
		return idx;
	}
	// bim_sahspilt.cpp:36-51 / bim_kdtree.cpp:55-72: leaf slots beyond the range stay all-zero.
	void makeLeaf(uint32_t node, const uint32_t* ids, uint32_t count)
	{
		size_t first = leaves.size();
		leaves.resize(first + trisPerLeaf, UVec4{0u, 0u, 0u, 0u});
		for (uint32_t i = 0; i < count; ++i)
		{
			const UVec3& t = triangles[ids[i]];
			leaves[first + i] = {t.x, t.y, t.z, materials.empty() ? 0u : materials[ids[i]]};
		}
		nodes[node].firstChild = 0x80000000u | uint32_t(first / trisPerLeaf);
	}
	Box boxOf(uint32_t tri) const
	{
		const UVec3& t = triangles[tri];
		return triBox(positions[t.x], positions[t.y], positions[t.z]);
	}
};

struct Center { Vec3 p; float key; };

// SAH sweep along the principal axis of the centroids (bim_sahspilt.cpp:27-126).
uint32_t buildSAH(BuildContext& ctx, uint32_t* ids, Center* centers, uint32_t lo, uint32_t hi)
{
	uint32_t node = ctx.newNode();
	if (hi - lo < ctx.trisPerLeaf)
	{
		ctx.makeLeaf(node, ids + lo, hi - lo + 1);
		return node;
	}
	// covariance of the centroids (two passes)
	Vec3 mean = centers[ids[lo]].p;
	for (uint32_t i = lo + 1; i <= hi; ++i) mean = add(mean, centers[ids[i]].p);
	float n = float(hi - lo + 1);
	mean = {mean.x / n, mean.y / n, mean.z / n};
	Mat3 cov = {{0, 0, 0, 0, 0, 0, 0, 0, 0}};
	for (uint32_t i = lo; i <= hi; ++i)
	{
		Vec3 e = sub(centers[ids[i]].p, mean);
		cov.m[0] += e.x * e.x; cov.m[1] += e.x * e.y; cov.m[2] += e.x * e.z;
		cov.m[3] += e.x * e.y; cov.m[4] += e.y * e.y; cov.m[5] += e.y * e.z;
		cov.m[6] += e.x * e.z; cov.m[7] += e.y * e.z; cov.m[8] += e.z * e.z;
	}
	float nm1 = float(hi - lo);
	for (float& v : cov.m) v /= nm1;
	Vec3 Q[3], lambda;
	decomposeQl(cov, Q, lambda);
	Vec3 splitDir;
	if (lambda.x > lambda.y && lambda.x > lambda.z) splitDir = Q[0];
	else if (lambda.y > lambda.z) splitDir = Q[1];
	else splitDir = Q[2];

	for (uint32_t i = lo; i <= hi; ++i) centers[ids[i]].key = dot3(centers[ids[i]].p, splitDir);
	std::sort(ids + lo, ids + hi + 1, [&](uint32_t a, uint32_t b) { return centers[a].key < centers[b].key; });

	// cost[i] = SA(left i+1 tris)*(i+1) + SA(right rest)*(count-i-1), sweeping from both ends
	uint32_t count = hi - lo; // number of split positions
	std::unique_ptr<float[]> leftCost(new float[count]), rightCost(new float[count]);
	Box lb = ctx.boxOf(ids[lo]), rb = ctx.boxOf(ids[hi]);
	leftCost[0] = surfaceArea(lb) * 1;
	rightCost[count - 1] = surfaceArea(rb) * 1;
	for (uint32_t i = lo + 1; i < hi; ++i)
	{
		lb = uniteBox(lb, ctx.boxOf(ids[i]));
		rb = uniteBox(rb, ctx.boxOf(ids[hi - i + lo]));
		leftCost[i - lo] = surfaceArea(lb) * int(i - lo + 1);
		rightCost[hi - i - 1] = surfaceArea(rb) * int(i - lo + 1);
	}
	uint32_t m = lo;
	float minCost = std::numeric_limits<float>::infinity();
	for (uint32_t i = 0; i < count; ++i)
	{
		float c = leftCost[i] + rightCost[i];
		if (c < minCost) { minCost = c; m = i + lo; }
	}
	leftCost.reset(); rightCost.reset();
	uint32_t l = buildSAH(ctx, ids, centers, lo, m);
	uint32_t r = buildSAH(ctx, ids, centers, m + 1, hi);
	ctx.nodes[node].firstChild = l;
	ctx.nodes[node].escape = r;
	return node;
}

// kd median split on the widest centroid extent (bim_kdtree.cpp:43-112).  `sorted[d]` hold the
// triangle ids ordered by centroid coordinate d; a split keeps the other two arrays partitioned.
void partitionStable(uint32_t* list, const Vec3* centers, uint32_t size, int dim, float plane)
{
	std::vector<uint32_t> tmp(size);
	uint32_t l = 0, r = (size + 1) / 2;
	for (uint32_t i = 0; i < size; ++i)
	{
		if (comp(centers[list[i]], dim) <= plane) tmp[l++] = list[i];
		else tmp[r++] = list[i];
	}
	std::memcpy(list, tmp.data(), size * sizeof(uint32_t));
}
uint32_t buildKD(BuildContext& ctx, uint32_t* sorted[3], Vec3* centers, uint32_t lo, uint32_t hi)
{
	uint32_t node = ctx.newNode();
	if (hi - lo < ctx.trisPerLeaf)
	{
		ctx.makeLeaf(node, sorted[0] + lo, hi - lo + 1);
		return node;
	}
	Vec3 w = {centers[sorted[0][hi]].x - centers[sorted[0][lo]].x, centers[sorted[1][hi]].y - centers[sorted[1][lo]].y,
	          centers[sorted[2][hi]].z - centers[sorted[2][lo]].z};
	int dim = 0;
	if (w.y > w.x && w.y > w.z) dim = 1;
	if (w.z > w.x && w.z > w.y) dim = 2;
	int co1 = (dim + 1) % 3, co2 = (dim + 2) % 3;
	uint32_t m = (lo + hi) / 2;
	float plane = comp(centers[sorted[dim][m]], dim);
	// equal coordinates right of the median are pushed over the plane for the partition, then restored
	uint32_t bumped = 0;
	while (m + bumped < hi && comp(centers[sorted[dim][m + 1 + bumped]], dim) == plane)
	{
		comp(centers[sorted[dim][m + 1 + bumped]], dim) += 1.0f;
		++bumped;
	}
	partitionStable(sorted[co1] + lo, centers, hi - lo + 1, dim, plane);
	partitionStable(sorted[co2] + lo, centers, hi - lo + 1, dim, plane);
	for (uint32_t i = m + 1; i < m + 1 + bumped; ++i) comp(centers[sorted[dim][i]], dim) = plane;
	uint32_t l = buildKD(ctx, sorted, centers, lo, m);
	uint32_t r = buildKD(ctx, sorted, centers, m + 1, hi);
	ctx.nodes[node].firstChild = l;
	ctx.nodes[node].escape = r;
	return node;
}
} // namespace

void Chunk::buildBVH_SAHsplit()
{
	uint32_t n = getNumTriangles();
	std::vector<Center> centers(n);
	std::vector<uint32_t> ids(n);
	for (uint32_t i = 0; i < n; ++i)
	{
		ids[i] = i;
		const UVec3& t = m_triangles[i];
		Vec3 s = add(add(m_positions[t.x], m_positions[t.y]), m_positions[t.z]);
		centers[i] = {{s.x / 3.0f, s.y / 3.0f, s.z / 3.0f}, 0.0f};
	}
	m_hierarchy.reserve(size_t(n) * 2);
	m_hierarchyParents.reserve(size_t(n) * 2);
	BuildContext ctx = {m_hierarchy, m_hierarchyParents, m_hierarchyLeaves, m_positions, m_triangles, m_triangleMaterials, m_numTrianglesPerLeaf};
	buildSAH(ctx, ids.data(), centers.data(), 0, n - 1);
}

void Chunk::buildBVH_kdtree()
{
	uint32_t n = getNumTriangles();
	std::vector<Vec3> centers(n);
	std::vector<uint32_t> s0(n), s1(n), s2(n);
	for (uint32_t i = 0; i < n; ++i)
	{
		s0[i] = s1[i] = s2[i] = i;
		const UVec3& t = m_triangles[i];
		Vec3 s = add(add(m_positions[t.x], m_positions[t.y]), m_positions[t.z]);
		centers[i] = {s.x / 3.0f, s.y / 3.0f, s.z / 3.0f};
	}
	std::sort(s0.begin(), s0.end(), [&](uint32_t a, uint32_t b) { return centers[a].x < centers[b].x; });
	std::sort(s1.begin(), s1.end(), [&](uint32_t a, uint32_t b) { return centers[a].y < centers[b].y; });
	std::sort(s2.begin(), s2.end(), [&](uint32_t a, uint32_t b) { return centers[a].z < centers[b].z; });
	m_hierarchy.reserve(size_t(n) * 2);
	m_hierarchyParents.reserve(size_t(n) * 2);
	BuildContext ctx = {m_hierarchy, m_hierarchyParents, m_hierarchyLeaves, m_positions, m_triangles, m_triangleMaterials, m_numTrianglesPerLeaf};
	uint32_t* sorted[3] = {s0.data(), s1.data(), s2.data()};
	buildKD(ctx, sorted, centers.data(), 0, n - 1);
}

// bim_hierarchy.cpp:238-252.  Iterative (explicit stack) so very deep SAH trees cannot overflow the
// call stack; returns the number of tree levels.
uint32_t Chunk::remapNodePointers(uint32_t root, uint32_t rootParent, uint32_t rootEscape)
{
	struct Item { uint32_t node, parent, escape, depth; };
	std::vector<Item> stack;
	stack.push_back({root, rootParent, rootEscape, 1});
	uint32_t levels = 0;
	while (!stack.empty())
	{
		Item it = stack.back();
		stack.pop_back();
		uint32_t left = m_hierarchy[it.node].firstChild;
		uint32_t right = m_hierarchy[it.node].escape;
		m_hierarchyParents[it.node] = it.parent;
		m_hierarchy[it.node].escape = it.escape;
		levels = std::max(levels, it.depth);
		if (!(left & 0x80000000u))
		{
			stack.push_back({right, it.node, it.escape, it.depth + 1});
			stack.push_back({left, it.node, right, it.depth + 1});
		}
	}
	return levels;
}

void Chunk::buildHierarchy(BuildMethod method)
{
	m_hierarchy.clear(); m_hierarchyParents.clear(); m_hierarchyLeaves.clear(); m_aaBoxes.clear(); m_oBoxes.clear();
	m_properties &= ~uint32_t(Property::HIERARCHY | Property::AABOX_BVH | Property::OBOX_BVH); // boxes of another tree are void
	m_numTrianglesPerLeaf = 2;
	if (m_triangles.empty()) return;
	if (method == BuildMethod::KD_TREE) buildBVH_kdtree();
	else buildBVH_SAHsplit();
	m_numTreeLevels = remapNodePointers(0, 0, 0);
	m_properties |= Property::HIERARCHY;
}

void Chunk::setHierarchy(const Node* nodes, const uint32_t* parents, const Box* boxes, uint32_t numNodes, const UVec4* leaves, uint32_t numLeafSlots, uint32_t numTreeLevels)
{
	m_numTrianglesPerLeaf = 2;
	m_hierarchy.assign(nodes, nodes + numNodes);
	m_hierarchyParents.assign(parents, parents + numNodes);
	m_aaBoxes.assign(boxes, boxes + numNodes);
	m_hierarchyLeaves.assign(leaves, leaves + numLeafSlots);
	m_oBoxes.clear();
	m_numTreeLevels = numTreeLevels;
	m_properties |= Property::HIERARCHY | Property::AABOX_BVH;
	m_properties &= ~uint32_t(Property::OBOX_BVH);
}

bool Chunk::validate(std::string* why) const
{
	auto bad = [&](const std::string& text) { if (why) *why = text; return false; };
	const size_t numVertices = m_positions.size(), numNodes = m_hierarchy.size();
	if (m_numTrianglesPerLeaf != 2) return bad("only 2 triangles per leaf are supported (scenedata.glsl:5)");
	for (const UVec3& t : m_triangles)
		if (t.x >= numVertices || t.y >= numVertices || t.z >= numVertices) return bad("a triangle references a missing vertex");
	if (!m_normals.empty() && m_normals.size() != numVertices) return bad("normal count differs from the vertex count");
	if (!m_texCoords0.empty() && m_texCoords0.size() != numVertices) return bad("texture coordinate count differs from the vertex count");
	if (!m_triangleMaterials.empty() && m_triangleMaterials.size() != m_triangles.size()) return bad("material index count differs from the triangle count");
	if (numNodes == 0) return true; // no hierarchy in the file: Scene::Init builds one
	if (m_hierarchyParents.size() != numNodes) return bad("hierarchy without one parent per node");
	if (m_hierarchyLeaves.empty() || m_hierarchyLeaves.size() % m_numTrianglesPerLeaf) return bad("hierarchy without leaf slots");
	if (!m_aaBoxes.empty() && m_aaBoxes.size() != numNodes) return bad("box count differs from the node count");
	if (!m_oBoxes.empty() && m_oBoxes.size() != numNodes) return bad("oriented-box count differs from the node count");
	const size_t numLeaves = m_hierarchyLeaves.size() / m_numTrianglesPerLeaf;
	for (size_t i = 0; i < numNodes; ++i)
	{
		const Node& n = m_hierarchy[i];
		const uint32_t child = n.firstChild & 0x7fffffffu;
		if (n.firstChild & 0x80000000u) { if (child >= numLeaves) return bad("a node references a missing leaf"); }
		else if (child >= numNodes || child <= i) return bad("a node's first child does not lie behind it (preorder)");
		if (n.escape >= numNodes || (n.escape != 0 && n.escape <= i)) return bad("a node's escape pointer does not lie behind it (preorder)");
		if (i != 0 && m_hierarchyParents[i] >= i) return bad("a node's parent does not lie in front of it (preorder)");
	}
	for (const UVec4& t : m_hierarchyLeaves)
		if (t.x != t.y && (t.x >= numVertices || t.y >= numVertices || t.z >= numVertices)) return bad("a leaf triangle references a missing vertex");
	if (numVertices == 0) return bad("hierarchy without vertices");
	return true;
}

// bim_hierarchy.cpp:8-50: leaf box = union of its (valid) triangles, inner box = union of children.
// Nodes are in preorder, so a reverse sweep sees every child before its parent.
void Chunk::computeBVHAABoxes()
{
	m_aaBoxes.assign(m_hierarchy.size(), Box{{0, 0, 0}, {0, 0, 0}});
	std::vector<uint8_t> seen(m_hierarchy.size(), 0);
	for (size_t i = m_hierarchy.size(); i-- > 0;)
	{
		uint32_t child = m_hierarchy[i].firstChild;
		if (child & 0x80000000u)
		{
			size_t leaf = size_t(child & 0x7fffffffu) * m_numTrianglesPerLeaf;
			const UVec4& t0 = m_hierarchyLeaves[leaf];
			Box b = triBox(m_positions[t0.x], m_positions[t0.y], m_positions[t0.z]);
			for (uint32_t k = 1; k < m_numTrianglesPerLeaf; ++k)
			{
				const UVec4& t = m_hierarchyLeaves[leaf + k];
				if (t.x == t.y) break;
				b = uniteBox(triBox(m_positions[t.x], m_positions[t.y], m_positions[t.z]), b);
			}
			m_aaBoxes[i] = b;
		}
		else
		{
			// children: firstChild, then follow escape links while the parent is this node
			Box b = m_aaBoxes[child];
			uint32_t c = m_hierarchy[child].escape;
			while (c != 0 && m_hierarchyParents[c] == i)
			{
				b = uniteBox(m_aaBoxes[c], b);
				c = m_hierarchy[c].escape;
			}
			m_aaBoxes[i] = b;
		}
	}
	m_properties |= Property::AABOX_BVH;
}

// ---- oriented boxes
namespace {
// rotate x by the unit quaternion q (the same formula the traversal uses, intersectiontests.glsl:82-88)
inline Vec3 qrotate(const Vec3& x, const Quaternion& q)
{
	Vec3 t = {q.j * x.z - q.k * x.y, q.k * x.x - q.i * x.z, q.i * x.y - q.j * x.x};
	return Vec3{x.x + 2.0f * (q.r * t.x + q.j * t.z - q.k * t.y), x.y + 2.0f * (q.r * t.y + q.k * t.x - q.i * t.z), x.z + 2.0f * (q.r * t.z + q.i * t.y - q.j * t.x)};
}
inline Quaternion qconj(const Quaternion& q) { return Quaternion{-q.i, -q.j, -q.k, q.r}; }

// unit quaternion of the rotation whose matrix has the orthonormal, right-handed rows a0, a1, a2 (world -> box frame is conj of it)
Quaternion quaternionFromAxes(const double a[3][3])
{
	// columns of the box->world rotation are the axes: R = [a0 a1 a2]
	const double m00 = a[0][0], m01 = a[1][0], m02 = a[2][0], m10 = a[0][1], m11 = a[1][1], m12 = a[2][1], m20 = a[0][2], m21 = a[1][2], m22 = a[2][2];
	double qw, qx, qy, qz;
	const double tr = m00 + m11 + m22;
	if (tr > 0) { double s = std::sqrt(tr + 1.0) * 2; qw = 0.25 * s; qx = (m21 - m12) / s; qy = (m02 - m20) / s; qz = (m10 - m01) / s; }
	else if (m00 > m11 && m00 > m22) { double s = std::sqrt(1.0 + m00 - m11 - m22) * 2; qw = (m21 - m12) / s; qx = 0.25 * s; qy = (m01 + m10) / s; qz = (m02 + m20) / s; }
	else if (m11 > m22) { double s = std::sqrt(1.0 + m11 - m00 - m22) * 2; qw = (m02 - m20) / s; qx = (m01 + m10) / s; qy = 0.25 * s; qz = (m12 + m21) / s; }
	else { double s = std::sqrt(1.0 + m22 - m00 - m11) * 2; qw = (m10 - m01) / s; qx = (m02 + m20) / s; qy = (m12 + m21) / s; qz = 0.25 * s; }
	const double n = std::sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
	return Quaternion{float(qx / n), float(qy / n), float(qz / n), float(qw / n)};
}

// principal-axis frame of a point set (Jacobi eigen decomposition of the covariance), rows = axes, right-handed
void principalFrame(const Vec3* p, size_t n, double axes[3][3])
{
	double mean[3] = {0, 0, 0};
	for (size_t i = 0; i < n; ++i) { mean[0] += p[i].x; mean[1] += p[i].y; mean[2] += p[i].z; }
	for (double& m : mean) m /= double(n);
	double c[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
	for (size_t i = 0; i < n; ++i)
	{
		const double d[3] = {p[i].x - mean[0], p[i].y - mean[1], p[i].z - mean[2]};
		for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) c[r][k] += d[r] * d[k];
	}
	double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
	for (int sweep = 0; sweep < 32; ++sweep)
	{
		double off = std::fabs(c[0][1]) + std::fabs(c[0][2]) + std::fabs(c[1][2]);
		if (off < 1e-30) break;
		for (int a = 0; a < 2; ++a) for (int b = a + 1; b < 3; ++b)
		{
			if (std::fabs(c[a][b]) < 1e-300) continue;
			const double theta = (c[b][b] - c[a][a]) / (2.0 * c[a][b]);
			const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
			const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
			for (int k = 0; k < 3; ++k) { const double ca = c[k][a], cb = c[k][b]; c[k][a] = cs * ca - sn * cb; c[k][b] = sn * ca + cs * cb; }
			for (int k = 0; k < 3; ++k) { const double ca = c[a][k], cb = c[b][k]; c[a][k] = cs * ca - sn * cb; c[b][k] = sn * ca + cs * cb; }
			for (int k = 0; k < 3; ++k) { const double va = v[k][a], vb = v[k][b]; v[k][a] = cs * va - sn * vb; v[k][b] = sn * va + cs * vb; }
		}
	}
	for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) axes[a][k] = v[k][a]; // eigenvectors are the columns of v
	// right-handed: a2 = a0 x a1
	axes[2][0] = axes[0][1] * axes[1][2] - axes[0][2] * axes[1][1];
	axes[2][1] = axes[0][2] * axes[1][0] - axes[0][0] * axes[1][2];
	axes[2][2] = axes[0][0] * axes[1][1] - axes[0][1] * axes[1][0];
}
}

void Chunk::computeBVHOBoxes()
{
	m_oBoxes.assign(m_hierarchy.size(), OBox{{0, 0, 0}, {0, 0, 0}, {0, 0, 0, 1}});
	if (m_hierarchy.empty()) return;
	// bim_hierarchy.cpp:54-85: depth-first, every node fits its box to the vertices collected below it (explicit stack instead of recursion)
	std::vector<Vec3> points;
	struct Frame { uint32_t node; size_t firstPoint; uint32_t nextChild; };
	std::vector<Frame> stack;
	stack.push_back(Frame{0u, 0u, 0xffffffffu});
	while (!stack.empty())
	{
		Frame& f = stack.back();
		const uint32_t child = m_hierarchy[f.node].firstChild;
		bool fit = false;
		if (child & 0x80000000u)
		{
			size_t leafIdx = size_t(child & 0x7fffffffu) * m_numTrianglesPerLeaf;
			for (uint32_t i = 0; i < m_numTrianglesPerLeaf; ++i)
			{
				const UVec4& t = m_hierarchyLeaves[leafIdx + i];
				if (t.x == t.y) break;
				points.push_back(m_positions[t.x]); points.push_back(m_positions[t.y]); points.push_back(m_positions[t.z]);
			}
			fit = true;
		}
		else if (f.nextChild == 0xffffffffu)
		{
			f.nextChild = m_hierarchy[child].escape;
			stack.push_back(Frame{child, points.size(), 0xffffffffu});
		}
		else if (f.nextChild != 0 && m_hierarchyParents[f.nextChild] == f.node)
		{
			const uint32_t c = f.nextChild;
			f.nextChild = m_hierarchy[c].escape;
			stack.push_back(Frame{c, points.size(), 0xffffffffu});
		}
		else
			fit = true;
		if (!fit) continue;
		const Frame done = stack.back();
		stack.pop_back();
		const Vec3* p = points.data() + done.firstPoint;
		const size_t n = points.size() - done.firstPoint;
		double axes[3][3];
		principalFrame(p, n, axes);
		OBox box;
		box.orientation = quaternionFromAxes(axes);
		// ei::OBox(orientation, points, n), 3dtypes.cpp:205-225, in the arithmetic of the traversal; padded by a few ulp so that
		// the float rotation of a contained vertex can never land outside
		const Quaternion inv = qconj(box.orientation);
		Vec3 mn = qrotate(p[0], inv), mx = mn;
		for (size_t i = 1; i < n; ++i)
		{
			const Vec3 q = qrotate(p[i], inv);
			mn = Vec3{std::min(mn.x, q.x), std::min(mn.y, q.y), std::min(mn.z, q.z)};
			mx = Vec3{std::max(mx.x, q.x), std::max(mx.y, q.y), std::max(mx.z, q.z)};
		}
		const Vec3 mid = {(mn.x + mx.x) * 0.5f, (mn.y + mx.y) * 0.5f, (mn.z + mx.z) * 0.5f};
		box.center = qrotate(mid, box.orientation);
		const float scale = std::max(std::max(std::fabs(mn.x), std::fabs(mx.x)), std::max(std::max(std::fabs(mn.y), std::fabs(mx.y)), std::max(std::fabs(mn.z), std::fabs(mx.z))));
		const float pad = 8.0f * 1.1920929e-07f * scale + 1e-30f;
		box.halfSides = Vec3{(mx.x - mn.x) * 0.5f + pad, (mx.y - mn.y) * 0.5f + pad, (mx.z - mn.z) * 0.5f + pad};
		m_oBoxes[done.node] = box;
		// (the reference reduces the collected points to their convex set before going up; the plain list gives the same box here)
	}
	m_properties |= Property::OBOX_BVH;
}

// ------------------------------------------------------------------ file format
namespace {
#pragma pack(push, 1)
struct SectionHeader { uint32_t type; uint32_t pad; uint64_t size; uint64_t uncompressedSize; }; // 24 B, bim_io.cpp:36-41
struct MetaSection { int32_t numChunks[3]; Box boundingBox; uint32_t numTrianglesPerLeaf; };       // 40 B, :51-56
struct ChunkMetaSection { Box boundingBox; uint32_t numTrianglesPerLeaf; uint32_t numTreeLevels; }; // 32 B, :58-63
#pragma pack(pop)
static_assert(sizeof(SectionHeader) == 24, "SectionHeader layout");
static_assert(sizeof(MetaSection) == 40, "MetaSection layout");
static_assert(sizeof(ChunkMetaSection) == 32, "ChunkMetaSection layout");

std::string directoryOf(const std::string& file)
{
	size_t p = file.find_last_of("/\\");
	return p == std::string::npos ? std::string() : file.substr(0, p + 1);
}

// bytes between the read position and the end of the file (a section of a corrupt file must not claim more)
uint64_t bytesLeft(std::ifstream& f)
{
	const std::streamoff here = f.tellg();
	if (here < 0) return 0;
	f.seekg(0, std::ios::end);
	const std::streamoff end = f.tellg();
	f.seekg(here, std::ios::beg);
	return end > here ? uint64_t(end - here) : 0;
}

template <typename T>
bool readSection(std::ifstream& f, const SectionHeader& h, std::vector<T>& out)
{
	if (h.size > bytesLeft(f)) return false;
	if (h.uncompressedSize)
	{
		// deflate expands at most 1032 : 1
		if (h.uncompressedSize % sizeof(T) || h.uncompressedSize / 1032 > h.size + 16) return false;
		std::vector<unsigned char> packed(h.size);
		f.read(reinterpret_cast<char*>(packed.data()), std::streamsize(h.size));
		out.assign(h.uncompressedSize / sizeof(T), T());
		uLongf len = uLongf(h.uncompressedSize);
		int r = uncompress(reinterpret_cast<Bytef*>(out.data()), &len, packed.data(), uLong(h.size));
		return r == Z_OK && len == h.uncompressedSize;
	}
	if (h.size % sizeof(T)) return false;
	out.assign(h.size / sizeof(T), T());
	f.read(reinterpret_cast<char*>(out.data()), std::streamsize(h.size));
	return bool(f);
}

template <typename T>
bool writeSection(std::ofstream& f, uint32_t type, const std::vector<T>& data)
{
	SectionHeader h = {type, 0, 0, uint64_t(data.size() * sizeof(T))};
	uLongf bound = compressBound(uLong(h.uncompressedSize));
	std::vector<unsigned char> packed(bound);
	if (compress2(packed.data(), &bound, reinterpret_cast<const Bytef*>(data.data()), uLong(h.uncompressedSize), 9) != Z_OK) return false;
	h.size = bound;
	f.write(reinterpret_cast<const char*>(&h), sizeof(h));
	f.write(reinterpret_cast<const char*>(packed.data()), std::streamsize(bound));
	return bool(f);
}
} // namespace

bool BinaryModel::loadEnvironmentFile(const std::string& envFile, std::string* error)
{
	std::ifstream in(envFile);
	if (!in) { if (error) *error = "Opening environment JSON failed: " + envFile; return false; }
	std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
	return loadEnvironmentText(text, error);
}

bool BinaryModel::loadEnvironmentText(const std::string& text, std::string* error)
{
	json::Value root;
	std::string perr;
	if (!json::parse(text, root, &perr)) { if (error) *error = "JSON: " + perr; return false; }
	m_materials.clear(); m_materialOrder.clear(); m_bimFile.clear();
	if (const json::Value* mats = root.find("materials"))
	{
		for (const auto& kv : mats->object)
		{
			Material mat(kv.first, "");
			for (const auto& prop : kv.second.object)
			{
				const json::Value& v = prop.second;
				if (v.type == json::Value::STRING)
				{
					if (prop.first == "type") mat.setType(v.str);
					else mat.setTexture(prop.first, v.str);
				}
				else if (v.type == json::Value::ARRAY)
				{
					float vals[4] = {0, 0, 0, 0};
					int n = 0;
					for (size_t i = 0; i < v.array.size() && i < 4; ++i) { vals[i] = float(v.array[i].number); ++n; }
					mat.set(prop.first, vals, n);
				}
				else if (v.type == json::Value::NUMBER)
				{
					float val = float(v.number);
					mat.set(prop.first, &val, 1);
				}
			}
			m_materials.emplace(mat.getName(), mat);
			m_materialOrder.push_back(mat.getName());
		}
	}
	else if (error) *error = "Cannot find 'materials' section in the scene file!";
	if (const json::Value* scene = root.find("scene")) m_bimFile = scene->str;
	if (const json::Value* acc = root.find("accelerator"))
	{
		if (acc->str == "aabox") m_accelerator = Property::AABOX_BVH;
		else if (acc->str == "obox") m_accelerator = Property::OBOX_BVH;
	}
	return true;
}

bool BinaryModel::load(const std::string& envFile, std::string* error)
{
	if (!loadEnvironmentFile(envFile, error)) return false;
	if (m_bimFile.empty()) { if (error) *error = "The Environment-File did not contain a binary file reference!"; return false; }
	std::string path = directoryOf(envFile) + m_bimFile;
	std::ifstream f(path, std::ios::binary);
	if (!f) { if (error) *error = "Cannot open scene file " + path; return false; }
	SectionHeader h;
	f.read(reinterpret_cast<char*>(&h), sizeof(h));
	if (!f || h.type != META_SECTION || h.size != sizeof(MetaSection)) { if (error) *error = "Invalid file. Meta-section not found!"; return false; }
	MetaSection meta;
	f.read(reinterpret_cast<char*>(&meta), sizeof(meta));
	m_boundingBox = meta.boundingBox;
	if (meta.numChunks[0] != 1 || meta.numChunks[1] != 1 || meta.numChunks[2] != 1) { if (error) *error = "Only single-chunk files are supported"; return false; }
	m_materialIndirection.clear();
	m_chunk = Chunk();
	bool haveChunk = false;
	while (f.read(reinterpret_cast<char*>(&h), sizeof(h)))
	{
		if (h.type == MATERIAL_REFERENCE)
		{
			uint32_t num = 0;
			f.read(reinterpret_cast<char*>(&num), 4);
			if (!f || uint64_t(num) * 64 > bytesLeft(f)) { if (error) *error = "Invalid file. Truncated material reference section"; return false; }
			char buf[65]; buf[64] = 0;
			for (uint32_t i = 0; i < num; ++i) { f.read(buf, 64); m_materialIndirection.push_back(buf); }
		}
		else if (h.type == CHUNK_SECTION && !haveChunk)
		{
			haveChunk = true;
			std::streamoff end = std::streamoff(f.tellg()) + std::streamoff(h.size);
			Chunk& c = m_chunk;
			SectionHeader s;
			while (f.tellg() < end && f.read(reinterpret_cast<char*>(&s), sizeof(s)))
			{
				bool ok = true;
				switch (s.type)
				{
				case CHUNK_META_SECTION: {
					ChunkMetaSection cm;
					f.read(reinterpret_cast<char*>(&cm), sizeof(cm));
					c.m_boundingBox = cm.boundingBox; c.m_numTrianglesPerLeaf = cm.numTrianglesPerLeaf; c.m_numTreeLevels = cm.numTreeLevels;
					break; }
				case Property::POSITION: ok = readSection(f, s, c.m_positions); c.m_properties |= Property::POSITION; break;
				case Property::NORMAL: ok = readSection(f, s, c.m_normals); c.m_properties |= Property::NORMAL; break;
				case Property::TEXCOORD0: ok = readSection(f, s, c.m_texCoords0); c.m_properties |= Property::TEXCOORD0; break;
				case Property::TRIANGLE_IDX: ok = readSection(f, s, c.m_triangles); c.m_properties |= Property::TRIANGLE_IDX; break;
				case Property::TRIANGLE_MAT: ok = readSection(f, s, c.m_triangleMaterials); c.m_properties |= Property::TRIANGLE_MAT; break;
				case Property::HIERARCHY: ok = readSection(f, s, c.m_hierarchy); c.m_properties |= Property::HIERARCHY; break;
				case HIERARCHY_PARENTS: ok = readSection(f, s, c.m_hierarchyParents); break;
				case HIERARCHY_LEAVES: ok = readSection(f, s, c.m_hierarchyLeaves); break;
				case Property::AABOX_BVH: ok = readSection(f, s, c.m_aaBoxes); c.m_properties |= Property::AABOX_BVH; break;
				case Property::OBOX_BVH: ok = readSection(f, s, c.m_oBoxes); c.m_properties |= Property::OBOX_BVH; break;
				default: f.seekg(std::streamoff(s.size), std::ios::cur); // tangents, qormals, OBoxes, SGGX, ...: not consumed
				}
				if (!ok) { if (error) *error = "Error in chunk decompression."; return false; }
			}
			f.clear();
			f.seekg(end, std::ios::beg);
		}
		else
			f.seekg(std::streamoff(h.size), std::ios::cur);
	}
	if (!haveChunk) { if (error) *error = "Invalid number of chunks. Expected 1 found 0"; return false; }
	if (m_materialIndirection.empty()) m_materialIndirection = m_materialOrder;
	// Missing per-vertex properties are filled with defaults (bim_io.cpp:297-312, bim_chunk.cpp:244-262).
	Chunk& c = m_chunk;
	if (c.m_normals.size() != c.m_positions.size()) c.m_normals.assign(c.m_positions.size(), Vec3{0.0f, 0.0f, 0.0f});
	if (c.m_texCoords0.size() != c.m_positions.size()) c.m_texCoords0.assign(c.m_positions.size(), Vec2{0.0f, 0.0f});
	if (c.m_triangleMaterials.size() != c.m_triangles.size()) c.m_triangleMaterials.assign(c.m_triangles.size(), 0u);
	std::string why;
	if (!c.validate(&why)) { if (error) *error = "Invalid file: " + why; return false; }
	return true;
}

bool BinaryModel::store(const std::string& envFile, const std::string& bimFileName, std::string* error) const
{
	// --- environment JSON (storeEnvironmentFile): scene reference, accelerator, materials
	{
		std::ofstream out(envFile);
		if (!out) { if (error) *error = "Cannot open " + envFile; return false; }
		out << "{\n\t\"scene\": \"" << bimFileName << "\",\n\t\"accelerator\": \"aabox\",\n\t\"materials\": {\n";
		bool firstMat = true;
		for (const std::string& name : m_materialIndirection)
		{
			auto it = m_materials.find(name);
			if (it == m_materials.end()) continue;
			const Material& m = it->second;
			out << (firstMat ? "" : ",\n") << "\t\t\"" << name << "\": {\n\t\t\t\"type\": \"" << (m.getType().empty() ? "legacy" : m.getType()) << "\"";
			firstMat = false;
			char buf[64];
			for (const auto& kv : m.values())
			{
				out << ",\n\t\t\t\"" << kv.first << "\": [";
				for (int i = 0; i < kv.second.numComponents; ++i)
				{
					snprintf(buf, sizeof(buf), "%.9g", kv.second.values[i]);
					out << (i ? ", " : "") << buf;
				}
				out << "]";
			}
			for (const auto& kv : m.textures()) out << ",\n\t\t\t\"" << kv.first << "\": \"" << kv.second << "\"";
			out << "\n\t\t}";
		}
		out << "\n\t},\n\t\"lights\": {\n\t},\n\t\"cameras\": {\n\t}\n}\n";
	}
	// --- binary (storeBinaryHeader + storeChunk)
	std::string path = directoryOf(envFile) + bimFileName;
	std::ofstream f(path, std::ios::binary);
	if (!f) { if (error) *error = "Cannot open file for writing: " + path; return false; }
	const Chunk& c = m_chunk;
	SectionHeader h = {META_SECTION, 0, sizeof(MetaSection), 0};
	f.write(reinterpret_cast<const char*>(&h), sizeof(h));
	MetaSection meta = {{1, 1, 1}, c.m_boundingBox, 2};
	f.write(reinterpret_cast<const char*>(&meta), sizeof(meta));
	h = {MATERIAL_REFERENCE, 0, m_materialIndirection.size() * 64 + 4, 0};
	f.write(reinterpret_cast<const char*>(&h), sizeof(h));
	uint32_t num = uint32_t(m_materialIndirection.size());
	f.write(reinterpret_cast<const char*>(&num), 4);
	for (const std::string& s : m_materialIndirection)
	{
		char buf[64];
		std::memset(buf, 0, sizeof(buf));
		std::strncpy(buf, s.c_str(), 63);
		f.write(buf, 64);
	}
	std::streamoff chunkHeaderPos = f.tellp();
	h = {CHUNK_SECTION, 0, ~uint64_t(0), ~uint64_t(0)};
	f.write(reinterpret_cast<const char*>(&h), sizeof(h));
	h = {CHUNK_META_SECTION, 0, sizeof(ChunkMetaSection), 0};
	f.write(reinterpret_cast<const char*>(&h), sizeof(h));
	ChunkMetaSection cm = {c.m_boundingBox, c.m_numTrianglesPerLeaf, c.m_numTreeLevels};
	f.write(reinterpret_cast<const char*>(&cm), sizeof(cm));
	bool ok = writeSection(f, Property::POSITION, c.m_positions);
	if (c.m_properties & Property::NORMAL) ok = ok && writeSection(f, Property::NORMAL, c.m_normals);
	if (c.m_properties & Property::TEXCOORD0) ok = ok && writeSection(f, Property::TEXCOORD0, c.m_texCoords0);
	ok = ok && writeSection(f, Property::TRIANGLE_IDX, c.m_triangles);
	if (c.m_properties & Property::TRIANGLE_MAT) ok = ok && writeSection(f, Property::TRIANGLE_MAT, c.m_triangleMaterials);
	if (c.m_properties & Property::HIERARCHY)
	{
		ok = ok && writeSection(f, Property::HIERARCHY, c.m_hierarchy);
		ok = ok && writeSection(f, HIERARCHY_PARENTS, c.m_hierarchyParents);
		ok = ok && writeSection(f, HIERARCHY_LEAVES, c.m_hierarchyLeaves);
	}
	if (c.m_properties & Property::AABOX_BVH) ok = ok && writeSection(f, Property::AABOX_BVH, c.m_aaBoxes);
	if (c.m_properties & Property::OBOX_BVH) ok = ok && writeSection(f, Property::OBOX_BVH, c.m_oBoxes);
	std::streamoff endPos = f.tellp();
	h = {CHUNK_SECTION, 0, uint64_t(endPos - chunkHeaderPos) - sizeof(SectionHeader), 0};
	f.seekp(chunkHeaderPos);
	f.write(reinterpret_cast<const char*>(&h), sizeof(h));
	if (!ok || !f) { if (error) *error = "Failed to write " + path; return false; }
	return true;
}

const Material* BinaryModel::getMaterial(uint32_t index) const
{
	if (index >= m_materialIndirection.size()) return nullptr;
	auto it = m_materials.find(m_materialIndirection[index]);
	return it == m_materials.end() ? nullptr : &it->second;
}

uint32_t BinaryModel::addMaterial(const Material& m)
{
	m_materials[m.getName()] = m;
	m_materialIndirection.push_back(m.getName());
	m_materialOrder.push_back(m.getName());
	return uint32_t(m_materialIndirection.size() - 1);
}

} // namespace bim
} // namespace gi
