// bvhmake.cpp -- see bvhmake.hpp.
#include "bvhmake.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <memory>

namespace gi {
namespace bvhmake {

namespace {

using bim::Box;
using bim::Vec3;

struct NamedArray { char name[32]; uint32_t numElements; uint32_t elementSize; }; // filedef.hpp:11-16
static_assert(sizeof(NamedArray) == 40 && sizeof(Vertex) == 32 && sizeof(Triangle) == 16 && sizeof(Node) == 12 && sizeof(Leaf) == 128 && sizeof(Box) == 24, "file records");

inline Vec3 vmin(Vec3 a, Vec3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline Vec3 vmax(Vec3 a, Vec3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline Box unite(const Box& a, const Box& b) { return {vmin(a.min, b.min), vmax(a.max, b.max)}; }
inline float comp(const Vec3& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
inline bool valid(const Triangle& t) { return t.vertices[0] != t.vertices[1]; } // filedef.hpp:43-49
// ei::surface(Box): 2 (xy + xz + yz) of the side lengths
inline float surface(const Box& b)
{
	const float x = b.max.x - b.min.x, y = b.max.y - b.min.y, z = b.max.z - b.min.z;
	return 2.0f * (x * y + x * z + y * z);
}

// The build state the three methods share: BVHBuilder's node / leaf / bounding-volume pools (bvhmake.hpp:224-232) and the aabox fit
// (fitmethods/aaboxfit.cpp:5-22).  Nodes are allocated in preorder (the root is node 0) with left / right child indices.
struct Builder
{
	const std::vector<Vertex>& vertices;
	const std::vector<Triangle>& triangles;
	struct BNode { uint32_t left, right; };
	std::vector<BNode> nodes;
	std::vector<Box> boxes;
	std::vector<Leaf> leaves;

	Builder(const std::vector<Vertex>& v, const std::vector<Triangle>& t) : vertices(v), triangles(t) {}
	uint32_t numTriangles() const { return uint32_t(triangles.size()); }
	Box triBox(const Triangle& t) const
	{
		const Vec3 &a = vertices[t.vertices[0]].position, &b = vertices[t.vertices[1]].position, &c = vertices[t.vertices[2]].position;
		return {vmin(vmin(a, b), c), vmax(vmax(a, b), c)};
	}
	Vec3 center(uint32_t i) const
	{
		const Triangle& t = triangles[i];
		const Vec3 &a = vertices[t.vertices[0]].position, &b = vertices[t.vertices[1]].position, &c = vertices[t.vertices[2]].position;
		return {(a.x + b.x + c.x) / 3.0f, (a.y + b.y + c.y) / 3.0f, (a.z + b.z + c.z) / 3.0f};
	}
	uint32_t newNode() { nodes.push_back({0u, 0u}); boxes.push_back(Box{{0, 0, 0}, {0, 0, 0}}); return uint32_t(nodes.size() - 1); }
	// a node that points to a fresh leaf holding ids[lo..hi] (at most NUM_PRIMITIVES), the rest of the leaf all-zero
	void makeLeaf(uint32_t node, const uint32_t* ids, uint32_t lo, uint32_t hi, bool fitBox)
	{
		Leaf leaf;
		std::memset(&leaf, 0, sizeof(leaf));
		for (uint32_t i = lo; i <= hi; ++i) leaf.triangles[i - lo] = triangles[ids[i]];
		nodes[node].left = 0x80000000u | uint32_t(leaves.size());
		nodes[node].right = 0u;
		leaves.push_back(leaf);
		if (fitBox)
		{
			// box of the first triangle, the others united one after the other
			Box b = triBox(leaf.triangles[0]);
			for (uint32_t i = 1; i < hi - lo + 1 && valid(leaf.triangles[i]); ++i) b = unite(triBox(leaf.triangles[i]), b);
			boxes[node] = b;
		}
	}
	// box over ids[lo..hi]: the last triangle first, then lo .. hi-1 (sweep.cpp:262-273)
	Box rangeBox(const uint32_t* ids, uint32_t lo, uint32_t hi) const
	{
		Box b = triBox(triangles[ids[hi]]);
		for (uint32_t i = lo; i < hi; ++i) b = unite(b, triBox(triangles[ids[i]]));
		return b;
	}
};

// ---------------------------------------------------------------- kdtree.cpp: median split of the centre along the axis of largest extent
struct KdTree
{
	Builder& b;
	std::vector<uint32_t> sorted[3];
	std::vector<Vec3> centers;

	void split(uint32_t* list, uint32_t size, int dim, float plane)
	{
		// the inverse of a merge: elements <= plane keep their order in the first (size + 1) / 2 places, the others behind them
		std::vector<uint32_t> tmp(size);
		uint32_t l = 0, r = (size + 1) / 2;
		for (uint32_t i = 0; i < size; ++i)
		{
			if (comp(centers[list[i]], dim) <= plane) tmp[l++] = list[i];
			else tmp[r++] = list[i];
		}
		std::memcpy(list, tmp.data(), size * sizeof(uint32_t));
	}

	uint32_t build(uint32_t lo, uint32_t hi)
	{
		const uint32_t node = b.newNode();
		if (hi - lo < NUM_PRIMITIVES)
		{
			b.makeLeaf(node, sorted[0].data(), lo, hi, true);
			return node;
		}
		const Vec3 bbMin = {centers[sorted[0][lo]].x, centers[sorted[1][lo]].y, centers[sorted[2][lo]].z};
		const Vec3 bbMax = {centers[sorted[0][hi]].x, centers[sorted[1][hi]].y, centers[sorted[2][hi]].z};
		const float w[3] = {bbMax.x - bbMin.x, bbMax.y - bbMin.y, bbMax.z - bbMin.z};
		int dim = 0;
		if (w[1] > w[0] && w[1] > w[2]) dim = 1;
		if (w[2] > w[0] && w[2] > w[1]) dim = 2;
		const int co1 = (dim + 1) % 3, co2 = (dim + 2) % 3;
		const uint32_t m = (lo + hi) / 2;
		// elements that share the median's coordinate and sit behind it are lifted above the plane for the duration of the split
		auto coord = [&](uint32_t id) -> float& { Vec3& c = centers[id]; return dim == 0 ? c.x : (dim == 1 ? c.y : c.z); };
		uint32_t changed = 0;
		const float plane = coord(sorted[dim][m]);
		while (m + changed < hi && coord(sorted[dim][m + 1 + changed]) == plane)
		{
			coord(sorted[dim][m + 1 + changed]) += 1.0f;
			++changed;
		}
		split(&sorted[co1][lo], hi - lo + 1, dim, plane);
		split(&sorted[co2][lo], hi - lo + 1, dim, plane);
		for (uint32_t i = m + 1; i < m + 1 + changed; ++i) coord(sorted[dim][i]) = plane;
		const uint32_t left = build(lo, m);
		const uint32_t right = build(m + 1, hi);
		b.nodes[node].left = left; b.nodes[node].right = right;
		b.boxes[node] = unite(b.boxes[left], b.boxes[right]);
		return node;
	}

	explicit KdTree(Builder& builder) : b(builder)
	{
		const uint32_t n = b.numTriangles();
		centers.resize(n);
		for (int k = 0; k < 3; ++k) sorted[k].resize(n);
		for (uint32_t i = 0; i < n; ++i) { sorted[0][i] = sorted[1][i] = sorted[2][i] = i; centers[i] = b.center(i); }
		for (int k = 0; k < 3; ++k)
			std::sort(sorted[k].begin(), sorted[k].end(), [&](uint32_t l, uint32_t r) { return comp(centers[l], k) < comp(centers[r], k); });
		build(0, n - 1);
	}
};

// ---------------------------------------------------------------- sweep.cpp: Hilbert order of the centres, surface-area sweep along that order
inline uint64_t partBy2(uint16_t x)
{
	uint64_t r = x;
	r = (r | (r << 16)) & 0x0000ff0000ffull;
	r = (r | (r << 8)) & 0x00f00f00f00full;
	r = (r | (r << 4)) & 0x0c30c30c30c3ull;
	r = (r | (r << 2)) & 0x249249249249ull;
	return r;
}
inline uint64_t morton(uint16_t a, uint16_t b, uint16_t c) { return partBy2(a) | (partBy2(b) << 1) | (partBy2(c) << 2); }
inline uint64_t grayToBinary(uint64_t n)
{
	n ^= n >> 32; n ^= n >> 16; n ^= n >> 8; n ^= n >> 4; n ^= n >> 2; n ^= n >> 1;
	return n;
}
// centres (shifted to be positive) on a 1/4096 grid: upper 16 bits of the three coordinates interleaved and taken through the inverse Gray
// code, the lower 16 bits break ties (sweep.cpp:136-150)
inline bool hilbertLess(const Vec3& a, const Vec3& b)
{
	const uint32_t ax = uint32_t(a.x * 4096.0f), ay = uint32_t(a.y * 4096.0f), az = uint32_t(a.z * 4096.0f);
	const uint32_t bx = uint32_t(b.x * 4096.0f), by = uint32_t(b.y * 4096.0f), bz = uint32_t(b.z * 4096.0f);
	uint64_t ca = grayToBinary(morton(uint16_t(ax >> 16), uint16_t(ay >> 16), uint16_t(az >> 16)));
	uint64_t cb = grayToBinary(morton(uint16_t(bx >> 16), uint16_t(by >> 16), uint16_t(bz >> 16)));
	if (ca == cb)
	{
		ca = grayToBinary(morton(uint16_t(ax & 0xffff), uint16_t(ay & 0xffff), uint16_t(az & 0xffff)));
		cb = grayToBinary(morton(uint16_t(bx & 0xffff), uint16_t(by & 0xffff), uint16_t(bz & 0xffff)));
	}
	return ca < cb;
}

// the split of ids[lo..hi] that minimises surface(left) / surface(parent) * nLeft + surface(right) / surface(parent) * nRight, boxes grown one
// triangle at a time from both ends (sweep.cpp:218-252 = lds.cpp:127-163).  firstCandidate: sweep.cpp starts its search at split 1, lds.cpp at 0.
uint32_t sweepSplit(const Builder& b, const uint32_t* ids, uint32_t lo, uint32_t hi, const Box& parent, uint32_t firstCandidate)
{
	const uint32_t n = hi - lo;
	std::unique_ptr<float[]> hx(new float[n]), hy(new float[n]);
	const float parentSurface = surface(parent);
	Box left = b.triBox(b.triangles[ids[lo]]), right = b.triBox(b.triangles[ids[hi]]);
	for (uint32_t i = lo; i < hi; ++i)
	{
		left = unite(left, b.triBox(b.triangles[ids[i]]));
		right = unite(right, b.triBox(b.triangles[ids[hi - i + lo]]));
		const int num = int(i - lo + 1);
		hx[i - lo] = surface(left) / parentSurface * num;
		hy[hi - i - 1] = surface(right) / parentSurface * num;
	}
	float minCost = std::numeric_limits<float>::infinity();
	uint32_t split = lo;
	for (uint32_t i = firstCandidate; i < n; ++i)
		if (hx[i] + hy[i] < minCost) { minCost = hx[i] + hy[i]; split = i + lo; }
	return split;
}

struct Sweep
{
	Builder& b;
	std::vector<uint32_t> sorted;
	uint32_t build(uint32_t lo, uint32_t hi)
	{
		const uint32_t node = b.newNode();
		b.boxes[node] = b.rangeBox(sorted.data(), lo, hi);
		if (hi - lo < NUM_PRIMITIVES) { b.makeLeaf(node, sorted.data(), lo, hi, false); return node; }
		const uint32_t split = sweepSplit(b, sorted.data(), lo, hi, b.boxes[node], 1u);
		const uint32_t left = build(lo, split);
		const uint32_t right = build(split + 1, hi);
		b.nodes[node].left = left; b.nodes[node].right = right;
		return node;
	}
	explicit Sweep(Builder& builder) : b(builder)
	{
		const uint32_t n = b.numTriangles();
		sorted.resize(n);
		std::vector<Vec3> centers(n);
		const Triangle& t0 = b.triangles[0];
		Vec3 minCenter = b.vertices[t0.vertices[0]].position;
		for (uint32_t i = 0; i < n; ++i) { sorted[i] = i; centers[i] = b.center(i); minCenter = vmin(centers[i], minCenter); }
		for (uint32_t i = 0; i < n; ++i) centers[i] = {centers[i].x - minCenter.x, centers[i].y - minCenter.y, centers[i].z - minCenter.z};
		std::sort(sorted.begin(), sorted.end(), [&](uint32_t l, uint32_t r) { return hilbertLess(centers[l], centers[r]); });
		build(0, n - 1);
	}
};

// ---------------------------------------------------------------- lds.cpp: sort along the principal axis of the centres, then the same sweep
// symmetric 3 x 3 eigen-decomposition by cyclic Jacobi rotations (the reference calls epsilon's decomposeQl; any exact method yields the same
// axis up to rounding, and the axis only orders the triangles)
void principalAxis(const double c[3][3], double axis[3])
{
	double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
	std::memcpy(a, c, sizeof(a));
	for (int sweep = 0; sweep < 32; ++sweep)
	{
		const double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
		if (off < 1e-300) break;
		for (int p = 0; p < 2; ++p)
			for (int q = p + 1; q < 3; ++q)
			{
				if (std::fabs(a[p][q]) < 1e-300) continue;
				const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
				const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
				const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
				for (int k = 0; k < 3; ++k) { const double akp = a[k][p], akq = a[k][q]; a[k][p] = cs * akp - sn * akq; a[k][q] = sn * akp + cs * akq; }
				for (int k = 0; k < 3; ++k) { const double apk = a[p][k], aqk = a[q][k]; a[p][k] = cs * apk - sn * aqk; a[q][k] = sn * apk + cs * aqk; }
				for (int k = 0; k < 3; ++k) { const double vkp = v[k][p], vkq = v[k][q]; v[k][p] = cs * vkp - sn * vkq; v[k][q] = sn * vkp + cs * vkq; }
			}
	}
	int best = 0;
	if (a[1][1] > a[best][best]) best = 1;
	if (a[2][2] > a[best][best]) best = 2;
	for (int k = 0; k < 3; ++k) axis[k] = v[k][best];
	// a fixed sign, so that the order does not depend on the rotation sequence
	int big = 0;
	for (int k = 1; k < 3; ++k) if (std::fabs(axis[k]) > std::fabs(axis[big])) big = k;
	if (axis[big] < 0) for (int k = 0; k < 3; ++k) axis[k] = -axis[k];
}

struct Lds
{
	Builder& b;
	std::vector<uint32_t> ids;
	std::vector<Vec3> centers;
	std::vector<float> proj;
	uint32_t build(uint32_t lo, uint32_t hi)
	{
		const uint32_t node = b.newNode();
		if (hi - lo < NUM_PRIMITIVES) { b.makeLeaf(node, ids.data(), lo, hi, true); return node; }
		// covariance of the centres in the range (two passes, unbiased)
		double m[3] = {0, 0, 0};
		for (uint32_t i = lo; i <= hi; ++i) { const Vec3& c = centers[ids[i]]; m[0] += c.x; m[1] += c.y; m[2] += c.z; }
		const double cnt = double(hi - lo + 1);
		for (int k = 0; k < 3; ++k) m[k] /= cnt;
		double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
		for (uint32_t i = lo; i <= hi; ++i)
		{
			const Vec3& c = centers[ids[i]];
			const double e[3] = {c.x - m[0], c.y - m[1], c.z - m[2]};
			for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) cov[r][s] += e[r] * e[s];
		}
		for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) cov[r][s] /= double(hi - lo);
		double axis[3];
		principalAxis(cov, axis);
		for (uint32_t i = lo; i <= hi; ++i) { const Vec3& c = centers[ids[i]]; proj[ids[i]] = float(c.x * axis[0] + c.y * axis[1] + c.z * axis[2]); }
		std::stable_sort(ids.begin() + lo, ids.begin() + hi + 1, [&](uint32_t l, uint32_t r) { return proj[l] < proj[r]; });
		// LDS_SPLITMODE_SWEEP (lds.hpp:6): the node's box first, then the sweep over the projected order
		b.boxes[node] = b.rangeBox(ids.data(), lo, hi);
		const uint32_t split = sweepSplit(b, ids.data(), lo, hi, b.boxes[node], 0u);
		const uint32_t left = build(lo, split);
		const uint32_t right = build(split + 1, hi);
		b.nodes[node].left = left; b.nodes[node].right = right;
		return node;
	}
	explicit Lds(Builder& builder) : b(builder)
	{
		const uint32_t n = b.numTriangles();
		ids.resize(n); centers.resize(n); proj.assign(n, 0.0f);
		for (uint32_t i = 0; i < n; ++i) { ids[i] = i; centers[i] = b.center(i); }
		build(0, n - 1);
	}
};

// RecursiveWriteHierarchy (bvhbuilder.cpp:706-718): preorder, parent and escape links filled in on the way
void writeHierarchy(const Builder& b, uint32_t self, uint32_t parent, uint32_t escape, std::vector<Node>& out, std::vector<Box>& boxesOut, std::vector<uint32_t>& order)
{
	struct Item { uint32_t self, parent, escape; };
	std::vector<Item> stack{{self, parent, escape}};
	while (!stack.empty())
	{
		const Item it = stack.back();
		stack.pop_back();
		out.push_back({it.parent, b.nodes[it.self].left, it.escape});
		order.push_back(it.self);
		(void)boxesOut;
		if (!(b.nodes[it.self].left & 0x80000000u))
		{
			stack.push_back({b.nodes[it.self].right, it.self, it.escape});
			stack.push_back({b.nodes[it.self].left, it.self, b.nodes[it.self].right});
		}
	}
}

bool setError(std::string* error, const std::string& text) { if (error) *error = text; return false; }

} // namespace

void BuildHierarchy(RawScene& scene, const std::vector<Triangle>& triangles, BuildMethod method)
{
	scene.nodes.clear(); scene.boxes.clear(); scene.leaves.clear();
	if (triangles.empty()) return;
	Builder b(scene.vertices, triangles);
	if (method == BuildMethod::KDTREE) KdTree k(b);
	else if (method == BuildMethod::SWEEP) Sweep s(b);
	else Lds l(b);
	// The builders allocate their nodes in preorder, and the file stores the node pool as it is (bounding volumes: bvhbuilder.cpp:592-594; the
	// hierarchy is written by a preorder walk whose indices are pool indices), so pool order == file order.
	std::vector<uint32_t> order;
	writeHierarchy(b, 0u, 0u, 0u, scene.nodes, scene.boxes, order);
	scene.boxes = b.boxes;
	scene.leaves = b.leaves;
}

bool StoreRawScene(const std::string& file, const RawScene& s, std::string* error)
{
	std::ofstream out(file, std::ios::binary);
	if (!out) return setError(error, "cannot open '" + file + "' for writing");
	auto put = [&](const char* name, const void* data, size_t count, size_t elementSize)
	{
		NamedArray h;
		std::memset(&h, 0, sizeof(h));
		std::strncpy(h.name, name, sizeof(h.name) - 1);
		h.numElements = uint32_t(count); h.elementSize = uint32_t(elementSize);
		out.write(reinterpret_cast<const char*>(&h), sizeof(h));
		if (count) out.write(static_cast<const char*>(data), std::streamsize(count * elementSize));
	};
	// the order bvhmake's main.cpp writes them in (main.cpp:127-143)
	put("materialref", s.materials.data(), s.materials.size(), sizeof(MaterialRef));
	put("vertices", s.vertices.data(), s.vertices.size(), sizeof(Vertex));
	put("bounding_aabox", s.boxes.data(), s.boxes.size(), sizeof(Box));
	put("hierarchy", s.nodes.data(), s.nodes.size(), sizeof(Node));
	put("triangles", s.leaves.data(), s.leaves.size(), sizeof(Leaf));
	return bool(out) || setError(error, "write error on '" + file + "'");
}

bool LoadRawScene(const std::string& file, RawScene& s, std::string* error)
{
	std::ifstream in(file, std::ios::binary);
	if (!in) return setError(error, "cannot open '" + file + "'");
	s = RawScene();
	in.seekg(0, std::ios::end);
	const uint64_t fileSize = uint64_t(std::max<std::streamoff>(in.tellg(), 0));
	in.seekg(0, std::ios::beg);
	NamedArray h;
	while (in.read(reinterpret_cast<char*>(&h), sizeof(h)))
	{
		h.name[31] = 0;
		const std::string name(h.name);
		const uint64_t bytes = uint64_t(h.numElements) * h.elementSize;
		auto take = [&](auto& vec, size_t elementSize) -> bool
		{
			if (h.elementSize != elementSize) return setError(error, "array '" + name + "': element size " + std::to_string(h.elementSize) + ", expected " + std::to_string(elementSize));
			if (bytes > fileSize - std::min<uint64_t>(fileSize, uint64_t(std::max<std::streamoff>(in.tellg(), 0)))) return setError(error, "array '" + name + "' is truncated");
			vec.resize(h.numElements);
			if (bytes && !in.read(reinterpret_cast<char*>(vec.data()), std::streamsize(bytes))) return setError(error, "array '" + name + "' is truncated");
			return true;
		};
		bool ok = true;
		if (name == "materialref") ok = take(s.materials, sizeof(MaterialRef));
		else if (name == "vertices") ok = take(s.vertices, sizeof(Vertex));
		else if (name == "bounding_aabox") ok = take(s.boxes, sizeof(Box));
		else if (name == "hierarchy") ok = take(s.nodes, sizeof(Node));
		else if (name == "triangles") ok = take(s.leaves, sizeof(Leaf));
		else
		{
			s.skipped.push_back(name);
			in.seekg(std::streamoff(bytes), std::ios::cur);
			if (!in) return setError(error, "array '" + name + "' is truncated");
		}
		if (!ok) return false;
	}
	if (s.vertices.empty() || s.nodes.empty() || s.leaves.empty()) return setError(error, "'" + file + "' holds no vertices / hierarchy / triangles array");
	if (s.boxes.size() != s.nodes.size()) return setError(error, "'" + file + "': bounding_aabox and hierarchy differ in length");
	for (const Leaf& l : s.leaves)
		for (const Triangle& t : l.triangles)
			for (uint32_t v : t.vertices)
				if (v >= s.vertices.size()) return setError(error, "'" + file + "': vertex index out of range");
	for (size_t i = 0; i < s.nodes.size(); ++i)
	{
		const Node& n = s.nodes[i];
		const uint32_t c = n.firstChild & 0x7fffffffu;
		if ((n.firstChild & 0x80000000u) ? c >= s.leaves.size() : c >= s.nodes.size()) return setError(error, "'" + file + "': child index out of range");
		if (n.escape >= s.nodes.size()) return setError(error, "'" + file + "': escape index out of range");
		// preorder: the first child and the escape target of a node lie behind it, and the first child's escape is its sibling
		if (!(n.firstChild & 0x80000000u) && (c <= i || s.nodes[c].escape <= c)) return setError(error, "'" + file + "': hierarchy is not in preorder");
		if (n.escape != 0 && n.escape <= i) return setError(error, "'" + file + "': hierarchy is not in preorder");
	}
	return true;
}

bool ToModel(const RawScene& s, const std::string& envJson, bim::BinaryModel& model, std::string* error)
{
	using namespace bim;
	if (s.nodes.empty()) return setError(error, "raw scene without hierarchy");
	if (!model.loadEnvironmentText(envJson, error)) return false;
	// vertex arrays
	std::vector<Vec3> pos(s.vertices.size()), nrm(s.vertices.size());
	std::vector<Vec2> uv(s.vertices.size());
	for (size_t i = 0; i < s.vertices.size(); ++i) { pos[i] = s.vertices[i].position; nrm[i] = s.vertices[i].normal; uv[i] = s.vertices[i].texcoord; }
	// the triangle list in leaf order
	std::vector<UVec3> tris;
	std::vector<uint32_t> mats;
	for (const Leaf& l : s.leaves)
		for (const Triangle& t : l.triangles)
			if (valid(t)) { tris.push_back({t.vertices[0], t.vertices[1], t.vertices[2]}); mats.push_back(t.material); }
	model.chunk().setMesh(pos.data(), nrm.data(), uv.data(), uint32_t(pos.size()), tris.data(), mats.data(), uint32_t(tris.size()));
	// material index i = materialref entry i: make the model's indirection table follow the file
	for (const MaterialRef& m : s.materials)
	{
		char name[33];
		std::memcpy(name, m.material, 32); name[32] = 0;
		if (!model.useMaterialName(name)) return setError(error, std::string("materialref '") + name + "' is not defined in the environment file");
	}
	if (s.materials.empty()) model.useJsonMaterialOrder();

	// the hierarchy: inner nodes as stored; a leaf node with k > 2 triangles becomes a balanced subtree over ceil(k / 2) two-triangle leaves
	struct TNode { uint32_t left, right; Box box; int leafFirst, leafCount; }; // leafCount > 0: leaf holding tris[leafFirst .. +leafCount)
	std::vector<TNode> tree;
	std::vector<uint32_t> firstTriOfLeaf(s.leaves.size() + 1, 0u);
	for (size_t l = 0; l < s.leaves.size(); ++l)
	{
		uint32_t k = 0;
		for (const Triangle& t : s.leaves[l].triangles) if (valid(t)) ++k;
		firstTriOfLeaf[l + 1] = firstTriOfLeaf[l] + k;
	}
	auto triBoxOf = [&](uint32_t t) { const UVec3& q = tris[t]; return Box{vmin(vmin(pos[q.x], pos[q.y]), pos[q.z]), vmax(vmax(pos[q.x], pos[q.y]), pos[q.z])}; };
	// recursive lambdas over explicit std::function-free helpers
	struct Expander
	{
		std::vector<TNode>& tree;
		decltype(triBoxOf)& boxOf;
		uint32_t leafRange(uint32_t first, uint32_t count, const Box* given)
		{
			const uint32_t self = uint32_t(tree.size());
			tree.push_back({0u, 0u, Box{{0, 0, 0}, {0, 0, 0}}, 0, 0});
			if (count <= 2)
			{
				Box b = boxOf(first);
				if (count == 2) b = unite(boxOf(first + 1), b);
				tree[self].box = given ? *given : b;
				tree[self].leafFirst = int(first); tree[self].leafCount = int(count);
				return self;
			}
			const uint32_t half = ((count / 2 + 1) / 2) * 2; // an even number of triangles goes left, so that only the last leaf can be half full
			const uint32_t l = leafRange(first, half, nullptr);
			const uint32_t r = leafRange(first + half, count - half, nullptr);
			tree[self].left = l; tree[self].right = r;
			tree[self].box = given ? *given : unite(tree[l].box, tree[r].box);
			return self;
		}
	} expand{tree, triBoxOf};
	// walk the file's preorder nodes; children of node i: firstChild, and the node that follows firstChild's subtree (= firstChild's escape)
	struct Walker
	{
		const RawScene& s; Expander& ex; std::vector<TNode>& tree; const std::vector<uint32_t>& firstTri;
		std::vector<uint8_t> visited; bool broken = false;
		uint32_t node(uint32_t i, uint32_t depth = 0)
		{
			// a tree visits every node once; a file whose links share or repeat nodes, or a chain deeper than any real tree, is refused
			if (broken || visited[i] || depth > 4096u) { broken = true; return 0xffffffffu; }
			visited[i] = 1;
			const Node& n = s.nodes[i];
			if (n.firstChild & 0x80000000u)
			{
				const uint32_t leaf = n.firstChild & 0x7fffffffu;
				const uint32_t k = firstTri[leaf + 1] - firstTri[leaf];
				if (k == 0) return 0xffffffffu;
				return ex.leafRange(firstTri[leaf], k, &s.boxes[i]);
			}
			const uint32_t self = uint32_t(tree.size());
			tree.push_back({0u, 0u, s.boxes[i], 0, 0});
			const uint32_t lc = n.firstChild, rc = s.nodes[lc].escape;
			const uint32_t l = node(lc, depth + 1);
			const uint32_t r = node(rc, depth + 1);
			tree[self].left = l; tree[self].right = r;
			return self;
		}
	} walk{s, expand, tree, firstTriOfLeaf, std::vector<uint8_t>(s.nodes.size(), 0), false};
	const uint32_t root = walk.node(0);
	if (walk.broken) return setError(error, "raw scene: the hierarchy is not a tree (a node is reached twice, or it is deeper than 4096 levels)");
	if (root == 0xffffffffu) return setError(error, "raw scene: the root leaf is empty");
	for (const TNode& t : tree)
		if (t.leafCount == 0 && (t.left == 0xffffffffu || t.right == 0xffffffffu)) return setError(error, "raw scene: a leaf without triangles");
	// emit in the chunk's layout: preorder nodes {firstChild, escape}, parents, two slots per leaf
	std::vector<bim::Node> nodes(tree.size());
	std::vector<uint32_t> parents(tree.size(), 0u);
	std::vector<Box> boxes(tree.size());
	std::vector<UVec4> leaves;
	uint32_t levels = 0;
	struct Item { uint32_t self, parent, escape, depth; };
	std::vector<Item> stack{{0u, 0u, 0u, 1u}};
	while (!stack.empty())
	{
		const Item it = stack.back();
		stack.pop_back();
		const TNode& t = tree[it.self];
		boxes[it.self] = t.box; parents[it.self] = it.parent; levels = std::max(levels, it.depth);
		nodes[it.self].escape = it.escape;
		if (t.leafCount > 0)
		{
			nodes[it.self].firstChild = 0x80000000u | uint32_t(leaves.size() / 2);
			for (int k = 0; k < 2; ++k)
			{
				if (k < t.leafCount) { const UVec3& q = tris[size_t(t.leafFirst + k)]; leaves.push_back({q.x, q.y, q.z, mats[size_t(t.leafFirst + k)]}); }
				else leaves.push_back({0u, 0u, 0u, 0u});
			}
		}
		else
		{
			nodes[it.self].firstChild = t.left;
			stack.push_back({t.right, it.self, it.escape, it.depth + 1});
			stack.push_back({t.left, it.self, t.right, it.depth + 1});
		}
	}
	model.chunk().setHierarchy(nodes.data(), parents.data(), boxes.data(), uint32_t(nodes.size()), leaves.data(), uint32_t(leaves.size()), levels);
	return true;
}

} // namespace bvhmake
} // namespace gi
