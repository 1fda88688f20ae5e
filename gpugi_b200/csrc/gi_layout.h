// gi_layout.h -- host-side half of gi_upload_scene(): validates a GiSceneDesc (the reference's
// Scene::SanityCheck) and re-lays the bim hierarchy out for the device (DESIGN.md "HBM layout"):
//   refNodes   the 32-byte TreeNode<Box> array as is (reference-order walk / exact fallback)
//   fastNodes  one 64-byte record per INNER node carrying both child boxes + child links
//   leafPos    vertex positions gathered per leaf triangle slot (3 x 16 B), validity flag in .w
//   materials  Scene::Material with 1x1 constant textures folded into the record
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
	bool binary = false;                    // strictly binary tree => fastNodes usable
	DScene scene = {};                      // counts, flags, ambiguity window (pointers unset)
};

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
	// 16-bit grid over the root box; cell size rounded up so that 65535 cells cover the box
	double gridLo[3], gridScale[3];
	for (int k = 0; k < 3; ++k)
	{
		const float lo = d->nodes[0].min[k], hi = d->nodes[0].max[k];
		float sc = float((double(hi) - double(lo)) / 65535.0 * (1.0 + 1.0e-6));
		if (!(sc > 1.0e-30f)) sc = 1.0e-30f;
		L.scene.gridLo[k] = lo; L.scene.gridScale[k] = sc;
		gridLo[k] = double(lo); gridScale[k] = double(sc);
	}
	// outward rounding, verified in double against the decode lo + q * scale the device's slab test stands for
	auto quantLo = [&](float x, int k) -> uint32_t {
		double q = std::floor((double(x) - gridLo[k]) / gridScale[k]);
		if (!(q > 0.0)) q = 0.0;
		if (q > 65535.0) q = 65535.0;
		while (q > 0.0 && gridLo[k] + q * gridScale[k] > double(x)) q -= 1.0;
		return uint32_t(q);
	};
	auto quantHi = [&](float x, int k) -> uint32_t {
		double q = std::ceil((double(x) - gridLo[k]) / gridScale[k]);
		if (!(q > 0.0)) q = 0.0;
		if (q > 65535.0) q = 65535.0;
		while (q < 65535.0 && gridLo[k] + q * gridScale[k] < double(x)) q += 1.0;
		return uint32_t(q);
	};
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
		for (int k = 0; k < 3; ++k)
		{
			const float lo = n.min[k], hi = n.max[k];
			if (!(lo <= a.min[k] && a.min[k] <= hi && lo <= a.max[k] && a.max[k] <= hi && lo <= b.min[k] && b.min[k] <= hi && lo <= b.max[k] && b.max[k] <= hi))
				L.binary = false;
		}
		if (!L.binary) break;
		auto encode = [&](uint32_t nodeIdx) -> int {
			const GiNode& ch = d->nodes[nodeIdx];
			if (ch.first_child & 0x80000000u) return ~int(ch.first_child & 0x7fffffffu);
			return fastIndex[nodeIdx];
		};
		const int c0 = encode(left), c1 = encode(right);
		F4* q = &L.fastNodes[size_t(fastIndex[i]) * 4];
		q[0] = F4{a.min[0], a.min[1], a.min[2], a.max[0]};
		q[1] = F4{a.max[1], a.max[2], b.min[0], b.min[1]};
		q[2] = F4{b.min[2], b.max[0], b.max[1], b.max[2]};
		F4 ch = {0.0f, 0.0f, 0.0f, 0.0f};
		std::memcpy(&ch.x, &c0, 4);
		std::memcpy(&ch.y, &c1, 4);
		q[3] = ch;
		{
			uint32_t v[12];
			for (int k = 0; k < 3; ++k)
			{
				v[k] = quantLo(a.min[k], k); v[3 + k] = quantHi(a.max[k], k);
				v[6 + k] = quantLo(b.min[k], k); v[9 + k] = quantHi(b.max[k], k);
				// the grid must cover the box (guaranteed for boxes nested in the root box)
				if (gridLo[k] + double(v[k]) * gridScale[k] > double(a.min[k]) || gridLo[k] + double(v[3 + k]) * gridScale[k] < double(a.max[k]) ||
				    gridLo[k] + double(v[6 + k]) * gridScale[k] > double(b.min[k]) || gridLo[k] + double(v[9 + k]) * gridScale[k] < double(b.max[k]))
					L.binary = false;
			}
			uint32_t* qn = &L.qNodes[size_t(fastIndex[i]) * 8];
			for (int k = 0; k < 3; ++k) { qn[k] = v[k] | (v[3 + k] << 16); qn[3 + k] = v[6 + k] | (v[9 + k] << 16); }
			qn[6] = uint32_t(c0); qn[7] = uint32_t(c1);
		}
	}

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
	L.scene.fastOrderValid = L.binary ? 1 : 0;
	L.scene.leafPreorder = preorder ? 1 : 0;
	{
		const GiNode& r = d->nodes[0];
		float dx = r.max[0] - r.min[0], dy = r.max[1] - r.min[1], dz = r.max[2] - r.min[2];
		L.scene.ambiguityAbs = 1.0e-6f * sqrtf(dx * dx + dy * dy + dz * dz);
	}
	return true;
}

} // namespace gik
