// gi_lbvh.cuh -- device-side hierarchy build: replaces Chunk::buildHierarchy + remapNodePointers + computeBVHAABoxes
// (bim_chunk.cpp:227-242, bim_sahspilt.cpp:27-151, bim_hierarchy.cpp:8-50,238-252) for meshes that arrive without a hierarchy.
//
// The OUTPUT is the reference's layout, so everything downstream (upload, both walks, the oracle, the .bim writer) is unchanged:
//   nodes in preorder, 32 bytes each (min, firstChild)(max, escape); a leaf node has firstChild = 0x80000000 | leaf, a leaf = 2 triangle
//   slots (an unused slot is all-zero, bim_sahspilt.cpp:45-46); escape = the next node in preorder after the subtree, 0 at the end;
//   parent of the root = 0; tree levels counted from 1.
// The TREE is an LBVH, not the reference's SAH sweep: triangles in Morton order of their centroids (63-bit codes), consecutive pairs
// form the leaves, Karras' binary radix tree (HPG 2012) over the leaf codes, boxes by a bottom-up pass.  All steps are data-parallel:
// one radix sort (cub) + five kernels, ~ms for millions of triangles where the host sweep takes seconds.
//
// Preorder numbering without a traversal: leaves are in tree order, and every inner node has two children, so for a node X whose
// first leaf is a and that lies in the LEFT subtree of m of its ancestors:  pre(X) = 2a + m   (the a leaves before X fill k complete
// subtrees with 2a - k nodes, k = ancestors with X on the right; ancestors = depth = k + m).  size(X) = 2 leaves(X) - 1.
#pragma once

#include "gi_core.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace gik {

#define LBVH_LEAF_FLAG 0x80000000u

struct LbvhTree
{
	// inner node i in [0, L-1): children (inner index, or LBVH_LEAF_FLAG | leaf), parent (inner index, root: 0xffffffff), leaf range
	uint32_t* left; uint32_t* right; uint32_t* parent; uint32_t* first; uint32_t* last;
	uint32_t* leafParent;   // per leaf: inner index
	float* boxMin; float* boxMax;       // per inner node, 3 floats each (bottom-up)
	float* leafMin; float* leafMax;     // per leaf
	uint32_t* visits;       // per inner node: arrival counter of the bottom-up pass
};

GI_D unsigned floatOrdered(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
GI_D float orderedFloat(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// bounding box of all vertices: ordered-uint atomics, one per warp and component
__global__ void __launch_bounds__(256) k_lbvh_bounds(const float* __restrict__ positions, uint32_t numVertices, unsigned* __restrict__ bounds /*min xyz, max xyz (ordered)*/)
{
	float mn[3] = {GI_INF, GI_INF, GI_INF}, mx[3] = {-GI_INF, -GI_INF, -GI_INF};
	for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < numVertices; v += gridDim.x * blockDim.x)
		for (int k = 0; k < 3; ++k) { const float p = positions[3 * size_t(v) + k]; mn[k] = fminf(mn[k], p); mx[k] = fmaxf(mx[k], p); }
	for (int k = 0; k < 3; ++k)
	{
		unsigned a = floatOrdered(mn[k]), b = floatOrdered(mx[k]);
		a = __reduce_min_sync(0xffffffffu, a); b = __reduce_max_sync(0xffffffffu, b);
		if ((threadIdx.x & 31) == 0) { atomicMin(bounds + k, a); atomicMax(bounds + 3 + k, b); }
	}
}

GI_D unsigned long long expandBits21(unsigned long long v)
{
	v &= 0x1fffffull;
	v = (v | v << 32) & 0x1f00000000ffffull;
	v = (v | v << 16) & 0x1f0000ff0000ffull;
	v = (v | v << 8) & 0x100f00f00f00f00full;
	v = (v | v << 4) & 0x10c30c30c30c30c3ull;
	v = (v | v << 2) & 0x1249249249249249ull;
	return v;
}

__global__ void __launch_bounds__(256) k_lbvh_morton(const float* __restrict__ positions, const uint32_t* __restrict__ triangles, uint32_t numTriangles,
                                                      const unsigned* __restrict__ bounds, unsigned long long* __restrict__ codes, uint32_t* __restrict__ ids)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= numTriangles) return;
	float c[3];
	for (int k = 0; k < 3; ++k)
	{
		const float lo = orderedFloat(bounds[k]), hi = orderedFloat(bounds[3 + k]);
		const float p = (positions[3 * size_t(triangles[3 * size_t(t)]) + k] + positions[3 * size_t(triangles[3 * size_t(t) + 1]) + k] +
		                 positions[3 * size_t(triangles[3 * size_t(t) + 2]) + k]) * (1.0f / 3.0f);
		const float ext = hi - lo;
		c[k] = ext > 0.0f ? fminf(fmaxf((p - lo) / ext, 0.0f), 1.0f) : 0.0f;
	}
	const unsigned long long x = (unsigned long long)(c[0] * 2097151.0f), y = (unsigned long long)(c[1] * 2097151.0f), z = (unsigned long long)(c[2] * 2097151.0f);
	codes[t] = (expandBits21(x) << 2) | (expandBits21(y) << 1) | expandBits21(z);
	ids[t] = t;
}

// leaf l = sorted triangles 2l, 2l+1: triangle records in the reference's slot format, leaf box, leaf code
__global__ void __launch_bounds__(256) k_lbvh_leaves(const float* __restrict__ positions, const uint32_t* __restrict__ triangles, const uint32_t* __restrict__ materials,
                                                      const uint32_t* __restrict__ sortedIds, const unsigned long long* __restrict__ sortedCodes, uint32_t numTriangles,
                                                      uint32_t numLeaves, I4* __restrict__ leafTris, unsigned long long* __restrict__ leafCodes, LbvhTree tr)
{
	const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
	if (l >= numLeaves) return;
	float mn[3] = {GI_INF, GI_INF, GI_INF}, mx[3] = {-GI_INF, -GI_INF, -GI_INF};
	for (uint32_t s = 0; s < 2u; ++s)
	{
		I4 rec; rec.x = rec.y = rec.z = rec.w = 0;
		const uint32_t k = 2u * l + s;
		if (k < numTriangles)
		{
			const uint32_t t = sortedIds[k];
			rec.x = int(triangles[3 * size_t(t)]); rec.y = int(triangles[3 * size_t(t) + 1]); rec.z = int(triangles[3 * size_t(t) + 2]);
			rec.w = materials ? int(materials[t]) : 0;
			const int v[3] = {rec.x, rec.y, rec.z};
			for (int j = 0; j < 3; ++j)
				for (int a = 0; a < 3; ++a) { const float p = positions[3 * size_t(v[j]) + a]; mn[a] = fminf(mn[a], p); mx[a] = fmaxf(mx[a], p); }
		}
		*reinterpret_cast<int4*>(leafTris + k) = make_int4(rec.x, rec.y, rec.z, rec.w);
	}
	for (int a = 0; a < 3; ++a) { tr.leafMin[3 * size_t(l) + a] = mn[a]; tr.leafMax[3 * size_t(l) + a] = mx[a]; }
	leafCodes[l] = sortedCodes[2u * l];
}

// length of the common prefix of the keys of leaves i and j (ties broken by the index), -1 outside the array
GI_D int lbvhDelta(const unsigned long long* codes, int numLeaves, int i, long long j)
{
	if (j < 0 || j >= numLeaves) return -1;
	const unsigned long long a = codes[i], b = codes[j];
	if (a == b) return 64 + __clz(unsigned(i) ^ unsigned(j));
	return __clzll(a ^ b);
}

// Karras 2012, algorithm of section 4: inner node i covers the leaves [first, last] and splits after `split`
__global__ void __launch_bounds__(256) k_lbvh_tree(const unsigned long long* __restrict__ codes, int numLeaves, LbvhTree tr)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= numLeaves - 1) return;
	const int d = (lbvhDelta(codes, numLeaves, i, i + 1) - lbvhDelta(codes, numLeaves, i, i - 1)) >= 0 ? 1 : -1;
	const int deltaMin = lbvhDelta(codes, numLeaves, i, i - d);
	int lmax = 2;
	while (lbvhDelta(codes, numLeaves, i, i + (long long)lmax * d) > deltaMin) lmax *= 2; // (lmax <= 2^30 for < 2^30 leaves)
	int l = 0;
	for (int t = lmax / 2; t >= 1; t /= 2)
		if (lbvhDelta(codes, numLeaves, i, i + (l + t) * d) > deltaMin) l += t;
	const int j = i + l * d;
	const int deltaNode = lbvhDelta(codes, numLeaves, i, j);
	int s = 0;
	for (int t = (l + 1) / 2;; t = (t + 1) / 2)
	{
		if (lbvhDelta(codes, numLeaves, i, i + (s + t) * d) > deltaNode) s += t;
		if (t == 1) break;
	}
	const int split = i + s * d + min(d, 0);
	const int lo = min(i, j), hi = max(i, j);
	uint32_t left, right;
	if (lo == split) { left = LBVH_LEAF_FLAG | uint32_t(split); tr.leafParent[split] = uint32_t(i); }
	else { left = uint32_t(split); tr.parent[split] = uint32_t(i); }
	if (hi == split + 1) { right = LBVH_LEAF_FLAG | uint32_t(split + 1); tr.leafParent[split + 1] = uint32_t(i); }
	else { right = uint32_t(split + 1); tr.parent[split + 1] = uint32_t(i); }
	tr.left[i] = left; tr.right[i] = right; tr.first[i] = uint32_t(lo); tr.last[i] = uint32_t(hi);
	if (i == 0) tr.parent[0] = 0xffffffffu;
}

// bottom-up boxes: the second thread to arrive at a node unites its children's boxes and goes on (the first one stops)
__global__ void __launch_bounds__(256) k_lbvh_boxes(int numLeaves, LbvhTree tr)
{
	const int l = blockIdx.x * blockDim.x + threadIdx.x;
	if (l >= numLeaves) return;
	uint32_t node = tr.leafParent[l];
	for (;;)
	{
		__threadfence();
		if (atomicAdd(tr.visits + node, 1u) == 0u) return;
		const uint32_t c[2] = {tr.left[node], tr.right[node]};
		float mn[3], mx[3];
		for (int a = 0; a < 3; ++a)
		{
			float lo[2], hi[2];
			for (int k = 0; k < 2; ++k)
			{
				const bool leaf = (c[k] & LBVH_LEAF_FLAG) != 0u;
				const size_t idx = 3 * size_t(c[k] & ~LBVH_LEAF_FLAG) + a;
				lo[k] = leaf ? tr.leafMin[idx] : __ldcg(tr.boxMin + idx);
				hi[k] = leaf ? tr.leafMax[idx] : __ldcg(tr.boxMax + idx);
			}
			mn[a] = fminf(lo[0], lo[1]); mx[a] = fmaxf(hi[0], hi[1]);
		}
		for (int a = 0; a < 3; ++a) { __stcg(tr.boxMin + 3 * size_t(node) + a, mn[a]); __stcg(tr.boxMax + 3 * size_t(node) + a, mx[a]); }
		if (tr.parent[node] == 0xffffffffu) return;
		node = tr.parent[node];
	}
}

// preorder index of every inner node (pre = 2 * first + left turns on the root path) and the tree depth
__global__ void __launch_bounds__(256) k_lbvh_preorder(int numLeaves, LbvhTree tr, uint32_t* __restrict__ preInner, uint32_t* __restrict__ preLeaf, uint32_t* __restrict__ maxDepth)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int numInner = numLeaves - 1;
	if (x >= numInner + numLeaves) return;
	const bool isLeaf = x >= numInner;
	const uint32_t self = isLeaf ? uint32_t(x - numInner) : uint32_t(x);
	const uint32_t firstLeaf = isLeaf ? self : tr.first[self];
	uint32_t child = isLeaf ? (LBVH_LEAF_FLAG | self) : self;
	uint32_t node = isLeaf ? tr.leafParent[self] : tr.parent[self];
	uint32_t m = 0, depth = 0;
	while (node != 0xffffffffu)
	{
		if (tr.left[node] == child) ++m;
		++depth;
		child = node;
		node = tr.parent[node];
	}
	const uint32_t pre = 2u * firstLeaf + m;
	if (isLeaf) preLeaf[self] = pre; else preInner[self] = pre;
	const unsigned mask = __activemask();
	depth = __reduce_max_sync(mask, depth);
	if (int(threadIdx.x & 31) == __ffs(mask) - 1) atomicMax(maxDepth, depth + 1u); // levels are counted from 1 (remapNodePointers)
}

// the reference's node array: one 32-byte record per inner node and per leaf, at its preorder position
__global__ void __launch_bounds__(256) k_lbvh_emit(int numLeaves, LbvhTree tr, const uint32_t* __restrict__ preInner, const uint32_t* __restrict__ preLeaf,
                                                    F4* __restrict__ nodes /*2 per node*/, uint32_t* __restrict__ parents)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	const int numInner = numLeaves - 1;
	const uint32_t numNodes = uint32_t(2 * numLeaves - 1);
	if (x >= numInner + numLeaves) return;
	const bool isLeaf = x >= numInner;
	const uint32_t self = isLeaf ? uint32_t(x - numInner) : uint32_t(x);
	const uint32_t pre = isLeaf ? preLeaf[self] : preInner[self];
	const uint32_t leaves = isLeaf ? 1u : tr.last[self] - tr.first[self] + 1u;
	uint32_t escape = pre + 2u * leaves - 1u;
	if (escape >= numNodes) escape = 0u;
	const uint32_t par = isLeaf ? tr.leafParent[self] : tr.parent[self];
	const float* mn = isLeaf ? tr.leafMin + 3 * size_t(self) : tr.boxMin + 3 * size_t(self);
	const float* mx = isLeaf ? tr.leafMax + 3 * size_t(self) : tr.boxMax + 3 * size_t(self);
	const uint32_t firstChild = isLeaf ? (LBVH_LEAF_FLAG | self) : pre + 1u;
	*reinterpret_cast<float4*>(nodes + 2 * size_t(pre)) = make_float4(mn[0], mn[1], mn[2], __uint_as_float(firstChild));
	*reinterpret_cast<float4*>(nodes + 2 * size_t(pre) + 1) = make_float4(mx[0], mx[1], mx[2], __uint_as_float(escape));
	parents[pre] = par == 0xffffffffu ? 0u : preInner[par];
}

// ====================================================================================================================================
// PLOC -- parallel locally-ordered clustering (Meister & Bittner 2018) over the same Morton-ordered leaves: bottom-up, in every pass each
// cluster looks for the neighbour (within PLOC_RADIUS positions) whose union with it has the smallest surface area, mutual nearest
// neighbours merge, the cluster list is compacted in order.  The trees come close to the host SAH sweep in traversal cost (the LBVH above
// splits by Morton bits only), at a few more milliseconds of build time.
// Node ids: leaves 0 .. L-1, inner nodes L .. 2L-2 (allocated at merge time).  Per node: box, children, parent, subtree sizes.
// ====================================================================================================================================
#define PLOC_RADIUS 16      // window of the triangle-pairing pass and default of the clustering passes
#define PLOC_MAX_RADIUS 64  // the clustering passes take the radius as an argument (GI_OPT_PLOC_RADIUS)

struct PlocTree
{
	float* boxMin; float* boxMax;   // 3 floats per node (leaf boxes first)
	uint32_t* left; uint32_t* right; // per node (leaves: unused)
	uint32_t* parent;                // per node, root: 0xffffffff
	uint32_t* size;                  // nodes in the subtree
	uint32_t* leaves;                // leaves in the subtree
};

__global__ void __launch_bounds__(256) k_ploc_init(uint32_t numLeaves, LbvhTree lt, PlocTree t, uint32_t* __restrict__ clusters)
{
	const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
	if (l >= numLeaves) return;
	for (int a = 0; a < 3; ++a) { t.boxMin[3 * size_t(l) + a] = lt.leafMin[3 * size_t(l) + a]; t.boxMax[3 * size_t(l) + a] = lt.leafMax[3 * size_t(l) + a]; }
	t.parent[l] = 0xffffffffu; t.size[l] = 1u; t.leaves[l] = 1u; t.left[l] = 0u; t.right[l] = 0u;
	clusters[l] = l;
}

GI_D float plocUnionArea(const float* amn, const float* amx, const float* bmn, const float* bmx)
{
	const float dx = fmaxf(amx[0], bmx[0]) - fminf(amn[0], bmn[0]);
	const float dy = fmaxf(amx[1], bmx[1]) - fminf(amn[1], bmn[1]);
	const float dz = fmaxf(amx[2], bmx[2]) - fminf(amn[2], bmn[2]);
	return dx * dy + dy * dz + dz * dx;
}

// nearest neighbour of every cluster inside the window; ties go to the smaller position, so that mutual pairs always exist
__global__ void __launch_bounds__(256) k_ploc_nearest(uint32_t n, const uint32_t* __restrict__ clusters, PlocTree t, uint32_t* __restrict__ nearest, int radius = PLOC_RADIUS)
{
	__shared__ float smn[256 + 2 * PLOC_MAX_RADIUS][3], smx[256 + 2 * PLOC_MAX_RADIUS][3];
	const int base = int(blockIdx.x * blockDim.x) - radius;
	for (int k = threadIdx.x; k < 256 + 2 * radius; k += blockDim.x)
	{
		const int j = base + k;
		if (j >= 0 && j < int(n))
		{
			const uint32_t node = clusters ? clusters[j] : uint32_t(j); // (no cluster list: position = node, the leaf-pairing pass)
			for (int a = 0; a < 3; ++a) { smn[k][a] = t.boxMin[3 * size_t(node) + a]; smx[k][a] = t.boxMax[3 * size_t(node) + a]; }
		}
	}
	__syncthreads();
	const int i = int(blockIdx.x * blockDim.x + threadIdx.x);
	if (i >= int(n)) return;
	const int self = int(threadIdx.x) + radius;
	float best = GI_INF; int bestJ = -1;
	for (int o = -radius; o <= radius; ++o)
	{
		const int j = i + o;
		if (o == 0 || j < 0 || j >= int(n)) continue;
		const float a = plocUnionArea(smn[self], smx[self], smn[self + o], smx[self + o]);
		if (a < best) { best = a; bestJ = j; }
	}
	nearest[i] = uint32_t(bestJ);
}

// mutual nearest neighbours merge into a new inner node (kept at the smaller position)
__global__ void __launch_bounds__(256) k_ploc_merge(uint32_t n, const uint32_t* __restrict__ clusters, const uint32_t* __restrict__ nearest, PlocTree t,
                                                     uint32_t* __restrict__ nodeCounter, uint32_t* __restrict__ merged, uint32_t* __restrict__ valid)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t j = nearest[i];
	uint32_t out = clusters[i], keep = 1u;
	if (j < n && nearest[j] == i)
	{
		if (i < j)
		{
			const uint32_t l = clusters[i], r = clusters[j];
			const uint32_t id = atomicAdd(nodeCounter, 1u);
			for (int a = 0; a < 3; ++a)
			{
				t.boxMin[3 * size_t(id) + a] = fminf(t.boxMin[3 * size_t(l) + a], t.boxMin[3 * size_t(r) + a]);
				t.boxMax[3 * size_t(id) + a] = fmaxf(t.boxMax[3 * size_t(l) + a], t.boxMax[3 * size_t(r) + a]);
			}
			t.left[id] = l; t.right[id] = r; t.parent[id] = 0xffffffffu;
			t.parent[l] = id; t.parent[r] = id;
			t.size[id] = t.size[l] + t.size[r] + 1u;
			t.leaves[id] = t.leaves[l] + t.leaves[r];
			out = id;
		}
		else keep = 0u;
	}
	merged[i] = out;
	valid[i] = keep;
}

__global__ void __launch_bounds__(256) k_ploc_compact(uint32_t n, const uint32_t* __restrict__ merged, const uint32_t* __restrict__ valid, const uint32_t* __restrict__ offset,
                                                       uint32_t* __restrict__ clustersOut, uint32_t* __restrict__ newCount)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (valid[i]) clustersOut[offset[i]] = merged[i];
	if (i == n - 1u) *newCount = offset[i] + valid[i];
}

// safety net: a pass that merged nothing (cannot happen with the tie rule above, but must never hang) pairs up neighbours
__global__ void __launch_bounds__(256) k_ploc_force_pairs(uint32_t n, uint32_t* __restrict__ nearest)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint32_t partner = i ^ 1u;
	nearest[i] = partner < n ? partner : 0xffffffffu; // the last cluster of an odd list stays as it is (k_ploc_merge: no partner)
}

// ---- leaves by mutual nearest neighbours: one clustering pass over the single triangles.  Two Morton-neighbouring triangles share a leaf
// only if each is the other's best partner inside the window; everything else gets a leaf of its own (second slot unused), as the host
// SAH sweep does when a pair would not pay.  Pairing Morton neighbours blindly instead costs ~40 % in traversal time.
__global__ void __launch_bounds__(256) k_pair_boxes(const float* __restrict__ positions, const uint32_t* __restrict__ triangles, const uint32_t* __restrict__ sortedIds,
                                                     uint32_t numTriangles, float* __restrict__ mn, float* __restrict__ mx)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= numTriangles) return;
	const uint32_t t = sortedIds[k];
	for (int a = 0; a < 3; ++a)
	{
		const float p0 = positions[3 * size_t(triangles[3 * size_t(t)]) + a], p1 = positions[3 * size_t(triangles[3 * size_t(t) + 1]) + a],
		            p2 = positions[3 * size_t(triangles[3 * size_t(t) + 2]) + a];
		mn[3 * size_t(k) + a] = fminf(p0, fminf(p1, p2)); mx[3 * size_t(k) + a] = fmaxf(p0, fmaxf(p1, p2));
	}
}

__global__ void __launch_bounds__(256) k_pair_flags(uint32_t n, const uint32_t* __restrict__ nearest, uint32_t* __restrict__ valid)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const uint32_t j = nearest[k];
	valid[k] = (j < n && nearest[j] == k && k > j) ? 0u : 1u; // the later triangle of a mutual pair moves into its partner's leaf
}

__global__ void __launch_bounds__(256) k_pair_emit(uint32_t n, const uint32_t* __restrict__ triangles, const uint32_t* __restrict__ materials,
                                                    const uint32_t* __restrict__ sortedIds, const unsigned long long* __restrict__ sortedCodes,
                                                    const uint32_t* __restrict__ nearest, const uint32_t* __restrict__ valid, const uint32_t* __restrict__ offset,
                                                    const float* __restrict__ mn, const float* __restrict__ mx, I4* __restrict__ leafTris,
                                                    unsigned long long* __restrict__ leafCodes, LbvhTree tr, uint32_t* __restrict__ numLeaves)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	if (k == n - 1u) *numLeaves = offset[k] + valid[k];
	if (!valid[k]) return;
	const uint32_t l = offset[k];
	const uint32_t j = nearest[k];
	const bool paired = j < n && nearest[j] == k;
	const uint32_t member[2] = {k, paired ? j : 0xffffffffu};
	float bmn[3] = {GI_INF, GI_INF, GI_INF}, bmx[3] = {-GI_INF, -GI_INF, -GI_INF};
	for (uint32_t s = 0; s < 2u; ++s)
	{
		int4 rec = make_int4(0, 0, 0, 0);
		if (member[s] != 0xffffffffu)
		{
			const uint32_t t = sortedIds[member[s]];
			rec = make_int4(int(triangles[3 * size_t(t)]), int(triangles[3 * size_t(t) + 1]), int(triangles[3 * size_t(t) + 2]), materials ? int(materials[t]) : 0);
			for (int a = 0; a < 3; ++a) { bmn[a] = fminf(bmn[a], mn[3 * size_t(member[s]) + a]); bmx[a] = fmaxf(bmx[a], mx[3 * size_t(member[s]) + a]); }
		}
		*reinterpret_cast<int4*>(leafTris + 2u * l + s) = rec;
	}
	for (int a = 0; a < 3; ++a) { tr.leafMin[3 * size_t(l) + a] = bmn[a]; tr.leafMax[3 * size_t(l) + a] = bmx[a]; }
	leafCodes[l] = sortedCodes[k];
}

// preorder position of every node: its depth + the sizes of all left siblings on its root path; leaf order likewise
__global__ void __launch_bounds__(256) k_ploc_preorder(uint32_t numNodes, uint32_t numLeaves, PlocTree t, uint32_t* __restrict__ pre, uint32_t* __restrict__ leafIndex,
                                                        uint32_t* __restrict__ maxDepth)
{
	const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x >= numNodes) return;
	uint32_t p = 0, li = 0, depth = 0;
	uint32_t child = x, node = t.parent[x];
	while (node != 0xffffffffu)
	{
		if (t.right[node] == child) { p += t.size[t.left[node]]; li += t.leaves[t.left[node]]; }
		++p; ++depth;
		child = node; node = t.parent[node];
	}
	pre[x] = p;
	if (x < numLeaves) leafIndex[x] = li;
	const unsigned mask = __activemask();
	depth = __reduce_max_sync(mask, depth);
	if (int(threadIdx.x & 31) == __ffs(mask) - 1) atomicMax(maxDepth, depth + 1u);
}

__global__ void __launch_bounds__(256) k_ploc_emit(uint32_t numNodes, uint32_t numLeaves, PlocTree t, const uint32_t* __restrict__ pre, const uint32_t* __restrict__ leafIndex,
                                                    const I4* __restrict__ leafTrisIn, F4* __restrict__ nodes, uint32_t* __restrict__ parents, I4* __restrict__ leafTrisOut)
{
	const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x >= numNodes) return;
	const uint32_t p = pre[x];
	uint32_t escape = p + t.size[x];
	if (escape >= numNodes) escape = 0u;
	uint32_t firstChild;
	if (x < numLeaves)
	{
		const uint32_t li = leafIndex[x];
		firstChild = LBVH_LEAF_FLAG | li;
		for (uint32_t s = 0; s < 2u; ++s)
		{
			const int4 rec = *reinterpret_cast<const int4*>(leafTrisIn + 2u * x + s);
			*reinterpret_cast<int4*>(leafTrisOut + 2u * li + s) = rec;
		}
	}
	else
		firstChild = p + 1u;
	*reinterpret_cast<float4*>(nodes + 2 * size_t(p)) = make_float4(t.boxMin[3 * size_t(x)], t.boxMin[3 * size_t(x) + 1], t.boxMin[3 * size_t(x) + 2], __uint_as_float(firstChild));
	*reinterpret_cast<float4*>(nodes + 2 * size_t(p) + 1) = make_float4(t.boxMax[3 * size_t(x)], t.boxMax[3 * size_t(x) + 1], t.boxMax[3 * size_t(x) + 2], __uint_as_float(escape));
	parents[p] = t.parent[x] == 0xffffffffu ? 0u : pre[t.parent[x]];
}

// ---- re-clustering of an existing hierarchy's leaves (GI_OPT_UPLOAD_TOPOLOGY, gi_layout.h: applyLeafTree): Morton codes of the leaf
// boxes' centres over the root box, then the PLOC passes above over the sorted leaves
__global__ void __launch_bounds__(256) k_leafbox_morton(const float* __restrict__ leafMin, const float* __restrict__ leafMax, uint32_t numLeaves, F3 rootLo, F3 rootHi,
                                                         unsigned long long* __restrict__ codes, uint32_t* __restrict__ ids)
{
	const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
	if (l >= numLeaves) return;
	const float lo[3] = {rootLo.x, rootLo.y, rootLo.z}, hi[3] = {rootHi.x, rootHi.y, rootHi.z};
	float c[3];
	for (int k = 0; k < 3; ++k)
	{
		const float p = 0.5f * leafMin[3 * size_t(l) + k] + 0.5f * leafMax[3 * size_t(l) + k];
		const float ext = hi[k] - lo[k];
		c[k] = ext > 0.0f ? fminf(fmaxf((p - lo[k]) / ext, 0.0f), 1.0f) : 0.0f; // (fminf / fmaxf drop a NaN)
	}
	const unsigned long long x = (unsigned long long)(c[0] * 2097151.0f), y = (unsigned long long)(c[1] * 2097151.0f), z = (unsigned long long)(c[2] * 2097151.0f);
	codes[l] = (expandBits21(x) << 2) | (expandBits21(y) << 1) | expandBits21(z);
	ids[l] = l;
}

__global__ void __launch_bounds__(256) k_ploc_init_gather(uint32_t numLeaves, const uint32_t* __restrict__ sortedIds, const float* __restrict__ leafMin,
                                                           const float* __restrict__ leafMax, PlocTree t, uint32_t* __restrict__ clusters)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= numLeaves) return;
	const uint32_t l = sortedIds[p];
	for (int a = 0; a < 3; ++a) { t.boxMin[3 * size_t(p) + a] = leafMin[3 * size_t(l) + a]; t.boxMax[3 * size_t(p) + a] = leafMax[3 * size_t(l) + a]; }
	t.parent[p] = 0xffffffffu; t.size[p] = 1u; t.leaves[p] = 1u; t.left[p] = 0u; t.right[p] = 0u;
	clusters[p] = p;
}

} // namespace gik
