// gi_layout.h -- host-side half of gi_upload_scene(): validates a GiSceneDesc (the reference's
// Scene::SanityCheck) and re-lays the bim hierarchy out for the device (DESIGN.md "HBM layout"):
//   refNodes   the 32-byte TreeNode<Box> array as is (reference-order walk / exact fallback)
//   fastNodes  one 64-byte record per binary INNER node carrying both child boxes + child links (one-thread-per-path kernels)
//   qNodes     the wavefront walk's records: 4-wide nodes (emitWide4) with the child boxes on a 16-bit grid over the root box;
//              build switches: binary records (WF_WIDE4=0), centre / half-extent records (WF_CENTRE_EXTENT)
//   leafPos    vertex positions gathered per leaf triangle slot (3 x 16 B) + the leaf node's own box in the .w components
//   materials  Scene::Material with 1x1 constant textures folded into the record
// fastNodes / qNodes stand on the reference's topology (buildHostLayout) or on any other binary tree over the reference's leaves
// (applyLeafTree: gi_upload_scene hands in the tree the device clustered, GI_OPT_UPLOAD_TOPOLOGY).
// Pure host C++ (no CUDA calls); gi_device.cu uploads the vectors and patches in texture objects.
#pragma once

#include "gi_core.h"
#include "gi_lane.h"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace gik {

struct HostLayout
{
	std::vector<F4> refNodes, fastNodes, leafPos, vertexInfo;
	std::vector<F4> oboxNodes;              // bvhType = 1: GiOBoxNode array as is (3 x F4 per node), else empty
	std::vector<I4> leafTris;
	std::vector<uint32_t> qNodes;           // 8 words per inner node (quantised child boxes + child links)
	std::vector<uint32_t> leafNode;         // leaf -> index of the node that points to it
	std::vector<DMaterial> materials;
	std::vector<DTexel> texels;             // 4 per material; .pixels still points at HOST memory here
	std::vector<int32_t> materialTexture;   // 4 per material: index into GiSceneDesc::textures
	struct ChildBox { uint32_t word; uint32_t child; float mn[3], mx[3]; }; // test hook: first word of the node record in qNodes, child slot, the exact box
	bool keepChildBoxes = false;            // set by the test emulator before the build
	std::vector<ChildBox> childBoxes;
	bool binary = false;                    // strictly binary tree => fastNodes usable
	DScene scene = {};                      // counts, flags, ambiguity window (pointers unset)
};

// Encoder of the two inner-node records (fastNodes: the reference's floats; qNodes: 16-bit grid over the root box, rounded outwards and
// verified in double against the decode lo + q * scale the device's slab test stands for).
struct FastEmitter
{
	double gridLo[3], gridScale[3];

	void init(const float rootMin[3], const float rootMax[3], HostLayout& L)
	{
		// cell size rounded up so that 65535 cells cover the box
		for (int k = 0; k < 3; ++k)
		{
			const float lo = rootMin[k], hi = rootMax[k];
			float sc = float((double(hi) - double(lo)) / 65535.0 * (1.0 + 1.0e-6));
			if (!(sc > 1.0e-30f)) sc = 1.0e-30f;
			L.scene.gridLo[k] = lo; L.scene.gridScale[k] = sc;
			L.scene.gridAbsMax[k] = std::fmax(std::fabs(lo), std::fabs(hi));
			{
				// centre of the root box as a float; half extent around THAT float, rounded up
				const float c0 = float(0.5 * (double(lo) + double(hi)));
				float he = float(std::fmax(double(hi) - double(c0), double(c0) - double(lo)));
				if (double(he) < std::fmax(double(hi) - double(c0), double(c0) - double(lo))) he = std::nextafterf(he, INFINITY);
				L.scene.ceCentre[k] = c0;
				L.scene.cePmax = k == 0 ? he : std::fmax(L.scene.cePmax, he);
			}
			gridLo[k] = double(lo); gridScale[k] = double(sc);
		}
	}
	static bool nested(const float pmin[3], const float pmax[3], const float cmin[3], const float cmax[3])
	{
		for (int k = 0; k < 3; ++k)
			if (!(pmin[k] <= cmin[k] && cmin[k] <= pmax[k] && pmin[k] <= cmax[k] && cmax[k] <= pmax[k])) return false;
		return true;
	}
	uint32_t quantLo(float x, int k) const
	{
		double q = std::floor((double(x) - gridLo[k]) / gridScale[k]);
		if (!(q > 0.0)) q = 0.0;
		if (q > 65535.0) q = 65535.0;
		while (q > 0.0 && gridLo[k] + q * gridScale[k] > double(x)) q -= 1.0;
		return uint32_t(q);
	}
	uint32_t quantHi(float x, int k) const
	{
		double q = std::ceil((double(x) - gridLo[k]) / gridScale[k]);
		if (!(q > 0.0)) q = 0.0;
		if (q > 65535.0) q = 65535.0;
		while (q < 65535.0 && gridLo[k] + q * gridScale[k] < double(x)) q += 1.0;
		return uint32_t(q);
	}
	// one box -> three words (min plane | max plane << 16 per axis); false if the grid does not cover it
	bool quantBox(const float bmin[3], const float bmax[3], uint32_t out[3]) const
	{
		bool covered = true;
		for (int k = 0; k < 3; ++k)
		{
			const uint32_t lo = quantLo(bmin[k], k), hi = quantHi(bmax[k], k);
			if (gridLo[k] + double(lo) * gridScale[k] > double(bmin[k]) || gridLo[k] + double(hi) * gridScale[k] < double(bmax[k])) covered = false;
			out[k] = lo | (hi << 16);
		}
		return covered;
	}
	static void noteChildBox(HostLayout& L, size_t word, uint32_t child, const float bmin[3], const float bmax[3])
	{
		if (!L.keepChildBoxes) return;
		HostLayout::ChildBox cb; cb.word = uint32_t(word); cb.child = child;
		for (int k = 0; k < 3; ++k) { cb.mn[k] = bmin[k]; cb.mx[k] = bmax[k]; }
		L.childBoxes.push_back(cb);
	}
	// inner node `index` with the child boxes a, b and the child links c0, c1 (inner index >= 0, leaf = ~leafIndex).
	// false: the grid does not cover a box (cannot happen for boxes nested in the root box).
	bool emit(HostLayout& L, uint32_t index, const float amin[3], const float amax[3], const float bmin[3], const float bmax[3], int c0, int c1) const
	{
		F4* q = &L.fastNodes[size_t(index) * 4];
		q[0] = F4{amin[0], amin[1], amin[2], amax[0]};
		q[1] = F4{amax[1], amax[2], bmin[0], bmin[1]};
		q[2] = F4{bmin[2], bmax[0], bmax[1], bmax[2]};
		F4 ch = {0.0f, 0.0f, 0.0f, 0.0f};
		std::memcpy(&ch.x, &c0, 4);
		std::memcpy(&ch.y, &c1, 4);
		q[3] = ch;
		bool covered = true;
		uint32_t v[12];
		for (int k = 0; k < 3; ++k)
		{
			v[k] = quantLo(amin[k], k); v[3 + k] = quantHi(amax[k], k);
			v[6 + k] = quantLo(bmin[k], k); v[9 + k] = quantHi(bmax[k], k);
			if (gridLo[k] + double(v[k]) * gridScale[k] > double(amin[k]) || gridLo[k] + double(v[3 + k]) * gridScale[k] < double(amax[k]) ||
			    gridLo[k] + double(v[6 + k]) * gridScale[k] > double(bmin[k]) || gridLo[k] + double(v[9 + k]) * gridScale[k] < double(bmax[k]))
				covered = false;
		}
		uint32_t* qn = &L.qNodes[size_t(index) * 8];
		for (int k = 0; k < 3; ++k) { qn[k] = v[k] | (v[3 + k] << 16); qn[3 + k] = v[6 + k] | (v[9 + k] << 16); }
		qn[6] = uint32_t(c0); qn[7] = uint32_t(c1);
		noteChildBox(L, size_t(index) * 8, 0u, amin, amax); noteChildBox(L, size_t(index) * 8, 1u, bmin, bmax);
		return covered;
	}
};

// ---- a second topology over the SAME leaves (GI_OPT_UPLOAD_TOPOLOGY).  The ordered walks only need a binary tree whose inner boxes
// contain their children's and whose leaves are the reference's leaves with the reference leaf node's own box (the exact gate of
// DESIGN.md 4): the reference visits a leaf iff the slab test of every ancestor passes, which -- the test being monotone in the box and
// bim's boxes being nested -- it does iff the test of the leaf's own box passes.  So ANY such tree returns the reference's answer; the
// hit ids stay reference leaf indices, ties and near-ties still go to the reference-order re-trace on refNodes.
// Node ids: 0 .. L-1 = leaves (position p holds reference leaf leafOf[p]), L .. 2L-2 = inner nodes.
struct LeafTree
{
	uint32_t numLeaves = 0, root = 0;
	std::vector<uint32_t> left, right;      // per node id (entries of leaves unused)
	std::vector<float> boxMin, boxMax;      // 3 per node id
	std::vector<uint32_t> leafOf;           // per leaf position
};

// WF_WIDE4: the ordered wavefront walk runs on 4-wide nodes collapsed from the binary tree `t` (repeatedly replace the inner child with
// the largest surface area by its two children until the node has four children or only leaves).  Wide nodes are numbered in preorder;
// 16 words each (gi_lane.h: wfInnerStep).  The exactness argument is that of the binary layout: a collapsed node's children are
// descendants of the binary node, their boxes nested in its box.
inline bool emitWide4(const GiSceneDesc* d, HostLayout& L, const LeafTree& t, std::string& why)
{
	const uint32_t nl = t.numLeaves;
	FastEmitter em;
	HostLayout gridOnly;
	em.init(d->nodes[0].min, d->nodes[0].max, gridOnly);
	std::vector<uint32_t> q;
	q.reserve(size_t(nl) * 8);
	if (L.keepChildBoxes) L.childBoxes.clear();
	auto areaOf = [&](uint32_t x) {
		const float* mn = &t.boxMin[size_t(x) * 3]; const float* mx = &t.boxMax[size_t(x) * 3];
		const double dx = double(mx[0]) - mn[0], dy = double(mx[1]) - mn[1], dz = double(mx[2]) - mn[2];
		return dx * dy + dy * dz + dz * dx;
	};
	struct Item { uint32_t node; uint32_t parentWord; }; // parentWord: index into q of the link to patch (0xffffffff for the root)
	std::vector<Item> stack;
	stack.push_back({t.root, 0xffffffffu});
	while (!stack.empty())
	{
		const Item it = stack.back(); stack.pop_back();
		const uint32_t index = uint32_t(q.size() / WF_NODE_WORDS);
		if (it.parentWord != 0xffffffffu) q[it.parentWord] = index;
		uint32_t kids[4] = {t.left[it.node], t.right[it.node], 0, 0};
		int n = 2;
		while (n < 4)
		{
			int pick = -1; double best = -1.0;
			for (int k = 0; k < n; ++k)
				if (kids[k] >= nl) { const double a = areaOf(kids[k]); if (a > best) { best = a; pick = k; } }
			if (pick < 0) break;
			const uint32_t x = kids[pick];
			for (int k = n; k > pick + 1; --k) kids[k] = kids[k - 1];
			kids[pick] = t.left[x]; kids[pick + 1] = t.right[x];
			++n;
		}
		const size_t base = q.size();
#if WF_CENTRE_EXTENT
		// 24 words (gi_lane.h: wfInnerStep): centres (relative to the root box's centre) as floats, half extents as bfloat16 rounded up, slack
		// e = 330 u P included (P = largest half extent of the root box); every box is checked in double against what the device's test stands
		// for: [c - (h - e), c + (h - e)] contains it
		q.resize(base + 24, 0u);
		const double pmax = double(gridOnly.scene.cePmax);
		const double slack = 330.0 * 5.9604644775390625e-08 * pmax;
		auto bf16Up = [](double x, bool& ok) -> uint32_t {   // smallest bfloat16 >= x (x >= 0), as its 16 bits
			float f = float(x);
			if (double(f) < x) f = std::nextafterf(f, INFINITY);
			uint32_t bits; std::memcpy(&bits, &f, 4);
			if (bits & 0xffffu) bits += 0x10000u;
			bits &= 0xffff0000u;
			float back; std::memcpy(&back, &bits, 4);
			if (!(double(back) >= x) || !std::isfinite(back)) ok = false;
			return bits >> 16;
		};
		for (int k = 0; k < 4; ++k)
		{
			float cv[3] = {0.0f, 0.0f, 0.0f};
			uint32_t hv[3];
			uint32_t link = uint32_t(WF_NODE_DONE);
			bool ok = true;
			if (k < n)
			{
				const float* mn = &t.boxMin[size_t(kids[k]) * 3]; const float* mx = &t.boxMax[size_t(kids[k]) * 3];
				for (int a = 0; a < 3; ++a)
				{
					if (!(std::fabs(mn[a]) <= gridOnly.scene.gridAbsMax[a] && std::fabs(mx[a]) <= gridOnly.scene.gridAbsMax[a] && mn[a] <= mx[a])) { why = "box outside the root box"; return false; }
					// primed coordinates: relative to the root box's centre (float - float in double: exact up to 2^-53, far below the slack)
					const double c0 = double(gridOnly.scene.ceCentre[a]);
					const double lo = double(mn[a]) - c0, hi = double(mx[a]) - c0;
					cv[a] = float(0.5 * (lo + hi));
					const double h0 = std::fmax(hi - double(cv[a]), double(cv[a]) - lo);
					hv[a] = bf16Up(h0 + slack, ok);
					uint32_t hb = hv[a] << 16; float hf; std::memcpy(&hf, &hb, 4);
					if (!(double(cv[a]) - (double(hf) - slack) <= lo && double(cv[a]) + (double(hf) - slack) >= hi)) ok = false;
					if (!(std::fabs(double(cv[a])) + double(hf) * (1.0 + 1.0 / 128.0) <= 1.02 * pmax)) ok = false; // P'' of the error bound (the odd child's h is read up to 2^-7 larger)
				}
				if (!ok) { why = "centre / half-extent record does not cover a box"; return false; }
				FastEmitter::noteChildBox(L, base, uint32_t(k), mn, mx);
				if (kids[k] < nl) link = uint32_t(~int(t.leafOf[kids[k]]));
			}
			else
			{
				// unused slot: h = -(65536 Pmax): near > far for every ray inside the domain of the bound
				const uint32_t neg = bf16Up(65536.0 * pmax, ok) | 0x8000u;
				hv[0] = hv[1] = hv[2] = neg;
			}
			for (int a = 0; a < 3; ++a)
			{
				std::memcpy(&q[base + 3 * size_t(k) + size_t(a)], &cv[a], 4);
				uint32_t& w = q[base + 12 + 3 * size_t(k >> 1) + size_t(a)];
				w |= (k & 1) ? (hv[a] << 16) : hv[a];
			}
			q[base + 20 + size_t(k)] = link;
		}
		const size_t linkBase = base + 20;
#else
		q.resize(base + 16, 0u);
		for (int k = 0; k < 4; ++k)
		{
			uint32_t w[3] = {0x0000ffffu, 0x0000ffffu, 0x0000ffffu}; // empty box: min 65535, max 0
			uint32_t link = uint32_t(WF_NODE_DONE);
			if (k < n)
			{
				const float* mn = &t.boxMin[size_t(kids[k]) * 3]; const float* mx = &t.boxMax[size_t(kids[k]) * 3];
				if (!em.quantBox(mn, mx, w)) { why = "quantisation grid does not cover a box"; return false; }
				FastEmitter::noteChildBox(L, base, uint32_t(k), mn, mx);
				if (kids[k] < nl) link = uint32_t(~int(t.leafOf[kids[k]]));
			}
			q[base + 3 * size_t(k)] = w[0]; q[base + 3 * size_t(k) + 1] = w[1]; q[base + 3 * size_t(k) + 2] = w[2];
			q[base + 12 + size_t(k)] = link;
		}
		const size_t linkBase = base + 12;
#endif
		for (int k = n - 1; k >= 0; --k) // reverse push: the first child gets the next preorder index
			if (kids[k] >= nl) stack.push_back({kids[k], uint32_t(linkBase + size_t(k))});
	}
	L.qNodes.swap(q);
	return true;
}

// the reference's own (binary, validated) topology as a LeafTree: leaf position = leaf index, inner id = numLeaves + preorder inner index
inline void referenceLeafTree(const GiSceneDesc* d, const std::vector<int>& fastIndex, LeafTree& t)
{
	const uint32_t nl = d->num_leaf_triangles / 2, N = 2u * nl - 1u;
	t = LeafTree();
	t.numLeaves = nl; t.root = nl;
	t.left.assign(N, 0u); t.right.assign(N, 0u); t.boxMin.assign(size_t(N) * 3, 0.0f); t.boxMax.assign(size_t(N) * 3, 0.0f); t.leafOf.resize(nl);
	for (uint32_t p = 0; p < nl; ++p) t.leafOf[p] = p;
	auto idOf = [&](uint32_t nodeIdx) -> uint32_t {
		const GiNode& n = d->nodes[nodeIdx];
		return (n.first_child & 0x80000000u) ? (n.first_child & 0x7fffffffu) : nl + uint32_t(fastIndex[nodeIdx]);
	};
	for (uint32_t i = 0; i < d->num_nodes; ++i)
	{
		const GiNode& n = d->nodes[i];
		const uint32_t id = idOf(i);
		if (id >= N) continue;
		std::memcpy(&t.boxMin[size_t(id) * 3], n.min, 12); std::memcpy(&t.boxMax[size_t(id) * 3], n.max, 12);
		if (!(n.first_child & 0x80000000u)) { t.left[id] = idOf(n.first_child); t.right[id] = idOf(d->nodes[n.first_child].escape); }
	}
}

inline float srgbDecodeHost(uint8_t v)
{
	float c = float(v) / 255.0f;
	return (c <= 0.04045f) ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f);
}

inline bool buildHostLayout(const GiSceneDesc* d, HostLayout& L, std::string& err)
{
	if (!d->nodes || d->num_nodes == 0 || !d->leaf_triangles || d->num_leaf_triangles == 0 || (d->num_leaf_triangles & 1u) || !d->positions ||
	    !d->vertex_info || d->num_vertices == 0 || !d->materials || d->num_materials == 0 || !d->textures || d->num_textures == 0)
	{
		err = "incomplete scene description";
		return false;
	}
	// ---- index ranges (Scene::SanityCheck, scene.hpp:191)
	const uint32_t numLeaves = d->num_leaf_triangles / 2;
	for (uint32_t i = 0; i < d->num_leaf_triangles; ++i)
	{
		const GiTriangle& t = d->leaf_triangles[i];
		if (t.v[0] == t.v[1]) continue; // unused slot
		if (t.v[0] >= d->num_vertices || t.v[1] >= d->num_vertices || t.v[2] >= d->num_vertices || t.material >= d->num_materials)
		{
			err = "leaf triangle " + std::to_string(i) + " references a missing vertex or material";
			return false;
		}
	}
	for (uint32_t i = 0; i < d->num_nodes; ++i)
	{
		const GiNode& n = d->nodes[i];
		const uint32_t child = n.first_child & 0x7fffffffu;
		if ((n.first_child & 0x80000000u) ? (child >= numLeaves) : (child >= d->num_nodes || child == 0))
		{
			err = "node " + std::to_string(i) + " has an invalid child";
			return false;
		}
		if (n.escape >= d->num_nodes)
		{
			err = "node " + std::to_string(i) + " has an invalid escape pointer";
			return false;
		}
	}
	// ---- the links must not form a cycle: the stackless walk (traceray.glsl:56-175) goes node -> firstChild on a hit, -> escape on a miss or
	// behind a leaf, and ends at escape == 0; on a cyclic file a ray would walk for ever ON THE DEVICE.  Node order is free (any acyclic
	// layout is accepted, the builders here all emit preorder): iterative depth-first colouring over both links of every node.
	{
		std::vector<uint8_t> colour(d->num_nodes, 0); // 0 new, 1 on the stack, 2 done
		std::vector<uint32_t> stack;
		for (uint32_t start = 0; start < d->num_nodes; ++start)
		{
			if (colour[start]) continue;
			stack.push_back(start);
			while (!stack.empty())
			{
				const uint32_t i = stack.back();
				if (colour[i] == 0)
				{
					colour[i] = 1;
					const GiNode& n = d->nodes[i];
					const uint32_t next[2] = {(n.first_child & 0x80000000u) ? 0u : n.first_child, n.escape};
					for (uint32_t e : next)
					{
						if (e == 0) continue; // the root only ever ends a walk
						if (colour[e] == 1) { err = "the child / escape links form a cycle at node " + std::to_string(e); return false; }
						if (colour[e] == 0) stack.push_back(e);
					}
				}
				else
				{
					colour[i] = 2;
					stack.pop_back();
				}
			}
		}
	}
	for (uint32_t ti = 0; ti < d->num_textures; ++ti)
	{
		const GiTexture& t = d->textures[ti];
		if (t.width <= 0 || t.height <= 0 || !t.pixels || t.format < GI_TEX_SRGB8_ALPHA8 || t.format > GI_TEX_RGBA32F)
		{
			err = "texture " + std::to_string(ti) + " is empty or has an unknown format";
			return false;
		}
	}

	std::memset(&L.scene, 0, sizeof(L.scene));

	// ---- reference layout
	L.refNodes.resize(size_t(d->num_nodes) * 2);
	std::memcpy(L.refNodes.data(), d->nodes, size_t(d->num_nodes) * sizeof(GiNode));
	L.leafTris.resize(d->num_leaf_triangles);
	std::memcpy(L.leafTris.data(), d->leaf_triangles, size_t(d->num_leaf_triangles) * sizeof(GiTriangle));
	L.vertexInfo.resize(d->num_vertices);
	std::memcpy(L.vertexInfo.data(), d->vertex_info, size_t(d->num_vertices) * sizeof(GiVertexInfo));

	// ---- leaf -> node map, and whether leaf indices increase along the reference's preorder walk (node -> firstChild,
	// leaf/skip -> escape; with every box "hit" this enumerates all leaves in tie-break order)
	L.leafNode.assign(numLeaves, 0u);
	bool preorder = true;
	{
		std::vector<uint8_t> seen(numLeaves, 0);
		uint32_t node = 0, steps = 0;
		int64_t lastLeaf = -1;
		do
		{
			const GiNode& n = d->nodes[node];
			if (n.first_child & 0x80000000u)
			{
				const uint32_t leaf = n.first_child & 0x7fffffffu;
				if (seen[leaf] || int64_t(leaf) <= lastLeaf) preorder = false;
				seen[leaf] = 1; lastLeaf = leaf;
				L.leafNode[leaf] = node;
				node = n.escape;
			}
			else
				node = n.first_child;
			if (++steps > 2u * d->num_nodes) { preorder = false; break; } // malformed escape chain
		} while (node != 0);
	}

	// ---- leaf block: per leaf the gathered vertex positions of both slots + the leaf node's own box in the .w components
	for (uint32_t v = 0; v < d->num_vertices; ++v)
		if (!std::isfinite(d->positions[3 * size_t(v)]) || !std::isfinite(d->positions[3 * size_t(v) + 1]) || !std::isfinite(d->positions[3 * size_t(v) + 2]))
		{
			err = "vertex " + std::to_string(v) + " has a non-finite position";
			return false;
		}
	L.leafPos.assign(size_t(numLeaves) * 6, F4{0.0f, 0.0f, 0.0f, 0.0f});
	for (uint32_t leaf = 0; leaf < numLeaves; ++leaf)
	{
		F4* rec = &L.leafPos[size_t(leaf) * 6];
		const GiNode& ln = d->nodes[L.leafNode[leaf]];
		for (uint32_t slot = 0; slot < 2; ++slot)
		{
			const GiTriangle& t = d->leaf_triangles[2 * size_t(leaf) + slot];
			const bool valid = t.v[0] != t.v[1];
			for (int k = 0; k < 3; ++k)
			{
				F4 v = {0.0f, 0.0f, 0.0f, 0.0f};
				if (valid) { const float* p = d->positions + 3 * size_t(t.v[k]); v.x = p[0]; v.y = p[1]; v.z = p[2]; }
				else if (k == 0) v.x = std::nanf("");
				v.w = slot == 0 ? ln.min[k] : ln.max[k];
				rec[3 * slot + k] = v;
			}
		}
	}

	// ---- fast layout.  Children of an inner node: firstChild, then its escape (the right sibling),
	// whose own escape must be the parent's escape (i.e. it is the last child).
	std::vector<int> fastIndex(d->num_nodes, -1);
	uint32_t numInner = 0;
	for (uint32_t i = 0; i < d->num_nodes; ++i)
		if (!(d->nodes[i].first_child & 0x80000000u)) fastIndex[i] = int(numInner++);
	L.binary = true;
	L.fastNodes.assign(size_t(numInner ? numInner : 1) * 4, F4{0.0f, 0.0f, 0.0f, 0.0f});
	L.qNodes.assign(size_t(numInner ? numInner : 1) * 8, 0u);
	FastEmitter em;
	em.init(d->nodes[0].min, d->nodes[0].max, L);
	for (uint32_t i = 0; i < d->num_nodes; ++i)
	{
		const GiNode& n = d->nodes[i];
		if (n.first_child & 0x80000000u) continue;
		const uint32_t left = n.first_child;
		const uint32_t right = d->nodes[left].escape;
		if (right == 0 || right >= d->num_nodes || right == n.escape || d->nodes[right].escape != n.escape) { L.binary = false; break; }
		const GiNode& a = d->nodes[left];
		const GiNode& b = d->nodes[right];
		// The wavefront walk tests inner nodes conservatively and applies the reference's exact cull test to a leaf's own
		// box only; that equals the reference's chain of tests iff every child box lies inside its parent's box (the slab
		// test is monotone in the box).  bim's bottom-up box pass guarantees it; other inputs fall back to the reference walk.
		if (!FastEmitter::nested(n.min, n.max, a.min, a.max) || !FastEmitter::nested(n.min, n.max, b.min, b.max)) { L.binary = false; break; }
		auto encode = [&](uint32_t nodeIdx) -> int {
			const GiNode& ch = d->nodes[nodeIdx];
			if (ch.first_child & 0x80000000u) return ~int(ch.first_child & 0x7fffffffu);
			return fastIndex[nodeIdx];
		};
		if (!em.emit(L, uint32_t(fastIndex[i]), a.min, a.max, b.min, b.max, encode(left), encode(right))) L.binary = false;
	}
#if WF_WIDE4
	if (L.binary && numInner > 0)
	{
		LeafTree rt;
		referenceLeafTree(d, fastIndex, rt);
		std::string why;
		if (!emitWide4(d, L, rt, why)) L.binary = false;
	}
#endif

	// ---- materials
	L.materials.resize(d->num_materials);
	L.texels.resize(size_t(d->num_materials) * 4);
	L.materialTexture.resize(size_t(d->num_materials) * 4);
	auto constantOf = [&](int32_t texIndex) -> F4 {
		const GiTexture& t = d->textures[texIndex];
		if (t.format == GI_TEX_RGBA32F) { const float* p = static_cast<const float*>(t.pixels); return F4{p[0], p[1], p[2], p[3]}; }
		const uint8_t* p = static_cast<const uint8_t*>(t.pixels);
		if (t.format == GI_TEX_SRGB8_ALPHA8) return F4{srgbDecodeHost(p[0]), srgbDecodeHost(p[1]), srgbDecodeHost(p[2]), float(p[3]) / 255.0f};
		return F4{float(p[0]) / 255.0f, float(p[1]) / 255.0f, float(p[2]) / 255.0f, float(p[3]) / 255.0f};
	};
	for (uint32_t mi = 0; mi < d->num_materials; ++mi)
	{
		const GiMaterial& m = d->materials[mi];
		const int32_t slots[4] = {m.reflectiveness_tex, m.opacity_tex, m.diffuse_tex, m.emissivity_tex};
		DMaterial dm;
		std::memset(&dm, 0, sizeof(dm));
		F4 consts[4];
		for (int s = 0; s < 4; ++s)
		{
			if (slots[s] < 0 || uint32_t(slots[s]) >= d->num_textures)
			{
				err = "material " + std::to_string(mi) + " references a missing texture";
				return false;
			}
			const GiTexture& t = d->textures[slots[s]];
			DTexel& tx = L.texels[size_t(mi) * 4 + s];
			tx.pixels = t.pixels; tx.width = t.width; tx.height = t.height; tx.format = t.format; tx.pad = 0;
			L.materialTexture[size_t(mi) * 4 + s] = slots[s];
			const bool constant = t.width == 1 && t.height == 1;
			consts[s] = constant ? constantOf(slots[s]) : F4{0.0f, 0.0f, 0.0f, 0.0f};
			dm.tex[s] = constant ? 0ull : 1ull; // non-zero = "sample the texture"; the device patches in the texture object
		}
		dm.cReflectiveness = consts[0];
		dm.cOpacity = consts[1];
		dm.cDiffuse = consts[2];
		dm.cDiffuse.w = m.refraction_index_avg;
		dm.cEmissivity = consts[3];
		dm.fresnel0 = F4{m.fresnel0[0], m.fresnel0[1], m.fresnel0[2], 0.0f};
		dm.fresnel1 = F4{m.fresnel1[0], m.fresnel1[1], m.fresnel1[2], 0.0f};
		L.materials[mi] = dm;
	}

	L.scene.numNodes = d->num_nodes;
	L.scene.numLeafTris = d->num_leaf_triangles;
	L.scene.numMaterials = d->num_materials;
	L.scene.rootIsLeaf = (d->nodes[0].first_child & 0x80000000u) ? 1 : 0;
	if (d->obox_nodes)
	{
		// bvhType = 1: the oriented boxes replace the AABoxes in the reference-order walk (intersectiontests.glsl:114-128).  Their tree
		// pointers must be the ones of `nodes`; the ordered walk (built on the AABoxes) does not apply.
		for (uint32_t i = 0; i < d->num_nodes; ++i)
		{
			const GiOBoxNode& o = d->obox_nodes[i];
			if (o.first_child != d->nodes[i].first_child || o.escape != d->nodes[i].escape) { err = "obox_nodes: tree pointers differ from nodes"; return false; }
			for (int k = 0; k < 3; ++k)
				if (!std::isfinite(o.center[k]) || !std::isfinite(o.sides_half[k]) || o.sides_half[k] < 0.0f) { err = "obox_nodes: non-finite or negative box"; return false; }
		}
		L.oboxNodes.resize(size_t(d->num_nodes) * 3);
		std::memcpy(L.oboxNodes.data(), d->obox_nodes, size_t(d->num_nodes) * sizeof(GiOBoxNode));
		L.binary = false;
	}
	// structure only (no box conditions): what the ordered walk of an oriented-box hierarchy needs
	{
		bool two = true;
		for (uint32_t i = 0; i < d->num_nodes && two; ++i)
		{
			const GiNode& n = d->nodes[i];
			if (n.first_child & 0x80000000u) continue;
			const uint32_t left = n.first_child;
			if (left == 0 || left >= d->num_nodes) { two = false; break; }
			const uint32_t right = d->nodes[left].escape;
			if (right == 0 || right >= d->num_nodes || right == n.escape || d->nodes[right].escape != n.escape) two = false;
		}
		L.scene.refTreeBinary = two ? 1 : 0;
	}
	L.scene.oboxOrdered = 0;
	L.scene.fastOrderValid = L.binary ? 1 : 0;
	L.scene.leafPreorder = preorder ? 1 : 0;
	{
		const GiNode& r = d->nodes[0];
		float dx = r.max[0] - r.min[0], dy = r.max[1] - r.min[1], dz = r.max[2] - r.min[2];
		L.scene.ambiguityAbs = 1.0e-6f * sqrtf(dx * dx + dy * dy + dz * dz);
	}
	return true;
}

// Rewrites fastNodes / qNodes from `t`, inner nodes numbered in preorder (the memory order the walks like).  Leaves L untouched and
// returns false (with the reason) if `t` is not a valid tree over the reference's leaves with nested boxes.
inline bool applyLeafTree(const GiSceneDesc* d, HostLayout& L, const LeafTree& t, std::string& why)
{
	const uint32_t nl = d->num_leaf_triangles / 2;
	if (!L.binary || nl < 2 || t.numLeaves != nl) { why = "no ordered layout to replace or leaf count mismatch"; return false; }
	const uint32_t numNodes = 2u * nl - 1u;
	if (t.left.size() < numNodes || t.right.size() < numNodes || t.boxMin.size() < size_t(numNodes) * 3 || t.boxMax.size() < size_t(numNodes) * 3 ||
	    t.leafOf.size() < nl || t.root < nl || t.root >= numNodes)
	{ why = "incomplete tree"; return false; }
	// leaves: a permutation of the reference's leaves carrying exactly the reference leaf node's box
	{
		std::vector<uint8_t> seen(nl, 0);
		for (uint32_t p = 0; p < nl; ++p)
		{
			const uint32_t leaf = t.leafOf[p];
			if (leaf >= nl || seen[leaf]) { why = "leaf list is not a permutation"; return false; }
			seen[leaf] = 1;
			const GiNode& ln = d->nodes[L.leafNode[leaf]];
			if (std::memcmp(&t.boxMin[size_t(p) * 3], ln.min, 12) != 0 || std::memcmp(&t.boxMax[size_t(p) * 3], ln.max, 12) != 0) { why = "leaf box differs from the reference's"; return false; }
		}
	}
	// preorder numbering of the inner nodes (iterative), every node reached exactly once
	std::vector<uint32_t> rank(numNodes, 0xffffffffu);
	std::vector<uint32_t> order; order.reserve(nl - 1);
	{
		std::vector<uint8_t> seen(numNodes, 0);
		std::vector<uint32_t> stack; stack.push_back(t.root);
		uint32_t leavesSeen = 0;
		while (!stack.empty())
		{
			const uint32_t x = stack.back(); stack.pop_back();
			if (x >= numNodes || seen[x]) { why = "child link out of range or node reached twice"; return false; }
			seen[x] = 1;
			if (x < nl) { ++leavesSeen; continue; }
			rank[x] = uint32_t(order.size()); order.push_back(x);
			stack.push_back(t.right[x]); stack.push_back(t.left[x]);
		}
		if (leavesSeen != nl || order.size() != size_t(nl - 1)) { why = "tree does not span all leaves"; return false; }
	}
	// root box must lie inside the reference's root box (the quantisation grid) and every child inside its parent
	if (!FastEmitter::nested(d->nodes[0].min, d->nodes[0].max, &t.boxMin[size_t(t.root) * 3], &t.boxMax[size_t(t.root) * 3])) { why = "root box outside the reference's"; return false; }
	for (uint32_t x : order)
		for (uint32_t ch : {t.left[x], t.right[x]})
			if (!FastEmitter::nested(&t.boxMin[size_t(x) * 3], &t.boxMax[size_t(x) * 3], &t.boxMin[size_t(ch) * 3], &t.boxMax[size_t(ch) * 3])) { why = "child box not inside its parent's"; return false; }
	// encode into fresh vectors, swap in on success
	HostLayout tmp;
	tmp.keepChildBoxes = L.keepChildBoxes;
	tmp.fastNodes.assign(size_t(nl - 1) * 4, F4{0.0f, 0.0f, 0.0f, 0.0f});
	tmp.qNodes.assign(size_t(nl - 1) * 8, 0u);
	FastEmitter em;
	em.init(d->nodes[0].min, d->nodes[0].max, tmp);
	auto link = [&](uint32_t ch) -> int { return ch < nl ? ~int(t.leafOf[ch]) : int(rank[ch]); };
	for (uint32_t x : order)
	{
		const uint32_t a = t.left[x], b = t.right[x];
		if (!em.emit(tmp, rank[x], &t.boxMin[size_t(a) * 3], &t.boxMax[size_t(a) * 3], &t.boxMin[size_t(b) * 3], &t.boxMax[size_t(b) * 3], link(a), link(b)))
		{ why = "quantisation grid does not cover a box"; return false; }
	}
#if WF_WIDE4
	{
		HostLayout wide;
		wide.keepChildBoxes = L.keepChildBoxes;
		if (!emitWide4(d, wide, t, why)) return false;
		L.fastNodes.swap(tmp.fastNodes);
		L.qNodes.swap(wide.qNodes);
		L.childBoxes.swap(wide.childBoxes);
		return true;
	}
#endif
	if (L.keepChildBoxes) L.childBoxes.swap(tmp.childBoxes);
	L.fastNodes.swap(tmp.fastNodes);
	L.qNodes.swap(tmp.qNodes);
	return true;
}

} // namespace gik
