// gi_core.h -- per-thread device logic of the light-transport hot path (sm_100a).
//
// Everything here is GI_HD (__host__ __device__) inline code operating on a DScene view of
// device memory.  nvcc compiles it into the kernels of gi_kernels.cu; tests/emul compiles the very
// same header with g++ to single-step the device logic on the CPU (test infrastructure only -- the
// product has no CPU path, the C ABI never calls the host instantiation).
//
// Exact zone: ray/box and ray/triangle tests use explicitly rounded single operations
// (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn, never contracted to FMA) in the reference's operation
// order (intersectiontests.glsl:9-20,63-79) so that hit flag + triangle id are bit-exact against
// the CPU oracle.  Shading uses ordinary float arithmetic.
#pragma once

#include "../../include/gpugi_b200.h"

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GI_HD __host__ __device__ __forceinline__
#define GI_D __device__ __forceinline__
#else
#define GI_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define GI_UNROLL _Pragma("unroll")
#ifndef GI_USE_FRCP
#define GI_USE_FRCP 0
#endif
#define GI_MUL(a, b) __fmul_rn((a), (b))
#define GI_ADD(a, b) __fadd_rn((a), (b))
#define GI_SUB(a, b) __fsub_rn((a), (b))
#define GI_DIV(a, b) __fdiv_rn((a), (b))
#if GI_USE_FRCP
#define GI_RCP(a) __frcp_rn((a))          // correctly rounded reciprocal == the IEEE quotient 1.0f / a, a shorter instruction sequence than div.rn
#else
#define GI_RCP(a) __fdiv_rn(1.0f, (a))
#endif
#else
// host emulation: compiled with -ffp-contract=off, so each operator is one rounded operation
#define GI_UNROLL
#define GI_MUL(a, b) ((a) * (b))
#define GI_ADD(a, b) ((a) + (b))
#define GI_SUB(a, b) ((a) - (b))
#define GI_DIV(a, b) ((a) / (b))
#define GI_RCP(a) (1.0f / (a))
#endif

namespace gik {

// ------------------------------------------------------------------ constants (stdheader.glsl:5-6 ...)
#define GI_RAY_HIT_EPSILON 0.0001f
#define GI_RAY_MAX 3.402823466e+38f
#define GI_INTERSECT_EPSILON 0.0001f
#define GI_DIVISOR_EPSILON 1e-5f
#define GI_PI 3.14159265358979f
#define GI_PI_2 6.28318530717958f
#define GI_INF __builtin_huge_valf()

struct alignas(16) F4 { float x, y, z, w; };
struct alignas(16) I4 { int x, y, z, w; };
struct F3 { float x, y, z; };
struct F2 { float x, y; };

GI_HD F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
GI_HD F3 f3(float s) { return f3(s, s, s); }
GI_HD F3 f3(const float* p) { return f3(p[0], p[1], p[2]); }
GI_HD F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
GI_HD F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
GI_HD F3 operator*(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
GI_HD F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
GI_HD F3 operator*(float s, F3 a) { return f3(s * a.x, s * a.y, s * a.z); }
GI_HD F3 operator/(F3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
GI_HD F3 operator/(F3 a, F3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
GI_HD F3 operator-(F3 a) { return f3(-a.x, -a.y, -a.z); }
GI_HD float dot(F3 a, F3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
GI_HD F3 cross(F3 a, F3 b) { return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
GI_HD F3 normalize(F3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
GI_HD float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
GI_HD F3 saturate(F3 a) { return f3(saturate(a.x), saturate(a.y), saturate(a.z)); }
GI_HD float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

struct alignas(32) U8 { uint32_t v[8]; };

#if defined(__CUDA_ARCH__)
// 256-bit global load (sm_100: LDG.E.256): a scattered 32-byte record costs one L1 wavefront per lane instead of two
// (tools/gpu/micro/gather.cu: 31 vs 70 cycles per warp for an L2-resident table).  p must be 32-byte aligned.
GI_HD U8 ld8(const void* p)
{
	U8 r;
	asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
	             : "l"(p));
	return r;
}
#else
GI_HD U8 ld8(const void* p) { U8 r; __builtin_memcpy(&r, p, 32); return r; }
#endif

#if defined(__CUDA_ARCH__)
GI_HD F4 ld4(const F4* p) { float4 v = __ldg(reinterpret_cast<const float4*>(p)); F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
GI_HD I4 ld4(const I4* p) { int4 v = __ldg(reinterpret_cast<const int4*>(p)); I4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
GI_HD int f2i(float f) { return __float_as_int(f); }
GI_HD float i2f(int i) { return __int_as_float(i); }
#else
GI_HD F4 ld4(const F4* p) { return *p; }
GI_HD I4 ld4(const I4* p) { return *p; }
GI_HD int f2i(float f) { int i; __builtin_memcpy(&i, &f, 4); return i; }
GI_HD float i2f(int i) { float f; __builtin_memcpy(&f, &i, 4); return f; }
#endif

// 96 bytes (6 x F4) with three 256-bit loads
GI_HD void ld8x3(const F4* p, F4* out)
{
	const U8 a = ld8(p), b = ld8(p + 2), c = ld8(p + 4);
	out[0].x = i2f(int(a.v[0])); out[0].y = i2f(int(a.v[1])); out[0].z = i2f(int(a.v[2])); out[0].w = i2f(int(a.v[3]));
	out[1].x = i2f(int(a.v[4])); out[1].y = i2f(int(a.v[5])); out[1].z = i2f(int(a.v[6])); out[1].w = i2f(int(a.v[7]));
	out[2].x = i2f(int(b.v[0])); out[2].y = i2f(int(b.v[1])); out[2].z = i2f(int(b.v[2])); out[2].w = i2f(int(b.v[3]));
	out[3].x = i2f(int(b.v[4])); out[3].y = i2f(int(b.v[5])); out[3].z = i2f(int(b.v[6])); out[3].w = i2f(int(b.v[7]));
	out[4].x = i2f(int(c.v[0])); out[4].y = i2f(int(c.v[1])); out[4].z = i2f(int(c.v[2])); out[4].w = i2f(int(c.v[3]));
	out[5].x = i2f(int(c.v[4])); out[5].y = i2f(int(c.v[5])); out[5].z = i2f(int(c.v[6])); out[5].w = i2f(int(c.v[7]));
}

// ------------------------------------------------------------------ device scene view (DESIGN.md "HBM layout")
// Material record: the 64-byte Scene::Material (scene.hpp:79-89) with the bindless handles replaced by
// CUDA texture objects and the 1x1 constant textures folded in as values (flag bit per slot).
struct alignas(16) DMaterial
{
	unsigned long long tex[4];  // cudaTextureObject_t: reflectiveness, opacity, diffuse, emissivity (0 if constant)
	F4 cReflectiveness;         // constants used when the matching tex[] is 0
	F4 cOpacity;                // .w unused
	F4 cDiffuse;                // .w = RefractionIndexAvg
	F4 cEmissivity;             // .w unused
	F4 fresnel0;                // .w unused
	F4 fresnel1;                // .w unused
};

// Software-visible copy of a texture for the fp32 ("exact") bilinear filter and the host emulation.
struct DTexel
{
	const void* pixels; // device pointer, RGBA8 or RGBA32F
	int width, height, format, pad;
};

struct DScene
{
	// reference layout (escape-pointer walk; exact fallback)
	const F4* refNodes;      // 2 x F4 per node: (min.xyz, firstChild bits), (max.xyz, escape bits)
	const F4* oboxNodes;     // bvhType = 1: 3 x F4 per node (center.xyz, firstChild) (sidesHalf.xyz, escape) (rotationInv.xyzw); else null
	// fast layout: one 64-byte record per inner node holding BOTH child boxes
	const F4* fastNodes;     // 4 x F4: (c0.min.xyz, c0.max.x) (c0.max.yz, c1.min.xy) (c1.min.z, c1.max.xyz) (child0, child1, -, -)
	const uint32_t* qNodes;  // wavefront walk: child boxes on a 16-bit grid over the root box, rounded outwards, min | max << 16 per axis.
	                         // WF_WIDE4 (default): 16 words (64 bytes, two 256-bit loads) per 4-wide node -- 4 x (x, y, z) + 4 child links,
	                         // unused slot = empty box; WF_WIDE4=0: 8 words per binary inner node (c0.xyz, c1.xyz, child0, child1)
	float gridLo[3];         // grid origin (root box min) and cell size per axis
	float gridScale[3];
	float gridAbsMax[3];     // max(|root box min|, |root box max|) per axis: bound of every plane coordinate
	float ceCentre[3];       // WF_CENTRE_EXTENT: centre of the root box (the node records' coordinates are relative to it)
	float cePmax;            // ... and the largest half extent of the root box around it (error bound of the conservative node test)
	const F4* leafPos;       // 6 x F4 per leaf: slot s -> [3s] = (p0.xyz, .), [3s+1] = (p1.xyz, .), [3s+2] = (p2.xyz, .); an unused slot has
	                         // p0.x = NaN (vertex positions are checked to be finite at upload).  The six .w hold the leaf node's own box:
	                         // [0..2].w = box.min.xyz, [3..5].w = box.max.xyz
	const I4* leafTris;      // GiTriangle per leaf slot (v0, v1, v2, material)
	const F4* vertexInfo;    // GiVertexInfo per vertex
	const DMaterial* materials;
	const DTexel* texels;    // 4 per material, same order as DMaterial::tex (fp32 filter / emulation)
	int leafPreorder;        // leaf indices increase in the reference's preorder walk (tie-break order == leaf slot order)
	uint32_t numNodes;
	uint32_t numLeafTris;
	uint32_t numMaterials;
	int rootIsLeaf;          // scene with a single leaf: no fast nodes
	int fastOrderValid;      // leaf indices increase in preorder (needed by nothing but asserted at upload)
	int refTreeBinary;       // every inner node of the reference tree has exactly two children (firstChild, and firstChild's escape)
	int oboxOrdered;         // per frame: bvhType = 1 scene walked nearest child first (traceOrderedOBox) instead of in the reference's order
	float ambiguityAbs;      // absolute part of the near-tie window (scaled with the scene extent)
	int textureFilterExact;  // GI_OPT_TEXTURE_FILTER
};

// ------------------------------------------------------------------ exact-zone primitives
struct BoxT { float first, last; };

// IntersectBox, intersectiontests.glsl:9-20.  inv = 1.0/dir computed once per ray (traceray.glsl:53).
GI_HD bool intersectBox(F3 o, F3 inv, float minx, float miny, float minz, float maxx, float maxy, float maxz, float& firstHit, float& lastHit)
{
	float tbx = GI_MUL(inv.x, GI_SUB(minx, o.x)), tby = GI_MUL(inv.y, GI_SUB(miny, o.y)), tbz = GI_MUL(inv.z, GI_SUB(minz, o.z));
	float ttx = GI_MUL(inv.x, GI_SUB(maxx, o.x)), tty = GI_MUL(inv.y, GI_SUB(maxy, o.y)), ttz = GI_MUL(inv.z, GI_SUB(maxz, o.z));
	float tminx = fminf(ttx, tbx), tminy = fminf(tty, tby), tminz = fminf(ttz, tbz);
	float tmaxx = fmaxf(ttx, tbx), tmaxy = fmaxf(tty, tby), tmaxz = fmaxf(ttz, tbz);
	float t0 = fmaxf(tminx, tminy), t1 = fmaxf(tminx, tminz);
	firstHit = fmaxf(0.0f, fmaxf(t0, t1));
	t0 = fminf(tmaxx, tmaxy); t1 = fminf(tmaxx, tmaxz);
	lastHit = fminf(t0, t1);
	return firstHit <= lastHit;
}

GI_HD float dotx(F3 a, F3 b) { return GI_ADD(GI_ADD(GI_MUL(a.x, b.x), GI_MUL(a.y, b.y)), GI_MUL(a.z, b.z)); }
GI_HD F3 crossx(F3 a, F3 b)
{
	return f3(GI_SUB(GI_MUL(a.y, b.z), GI_MUL(b.y, a.z)), GI_SUB(GI_MUL(a.z, b.x), GI_MUL(b.z, a.x)), GI_SUB(GI_MUL(a.x, b.y), GI_MUL(b.x, a.y)));
}
GI_HD F3 subx(F3 a, F3 b) { return f3(GI_SUB(a.x, b.x), GI_SUB(a.y, b.y), GI_SUB(a.z, b.z)); }

// IntersectTriangle, intersectiontests.glsl:63-79 (Moeller-Trumbore, branchless OptiX ordering).
GI_HD bool intersectTriangle(F3 o, F3 d, F3 p0, F3 p1, F3 p2, float& hit, F3& bary, F3& n)
{
	F3 e0 = subx(p1, p0);
	F3 e1 = subx(p0, p2);
	n = crossx(e1, e0);
	float r = GI_RCP(dotx(n, d));
	F3 po = subx(p0, o);
	F3 e2 = f3(GI_MUL(r, po.x), GI_MUL(r, po.y), GI_MUL(r, po.z));
	F3 i = crossx(d, e2);
	bary.y = dotx(i, e1);
	bary.z = dotx(i, e0);
	bary.x = GI_SUB(1.0f, GI_ADD(bary.z, bary.y));
	hit = dotx(n, e2);
	return (hit > GI_INTERSECT_EPSILON) && (bary.x >= 0.0f) && (bary.y >= 0.0f) && (bary.z >= 0.0f);
}

GI_HD F3 invDir(F3 d) { return f3(GI_RCP(d.x), GI_RCP(d.y), GI_RCP(d.z)); }

// rotate(), intersectiontests.glsl:82-88: x rotated by the quaternion q
GI_HD F3 rotateByQuaternion(F3 x, F4 q)
{
	F3 t = crossx(f3(q.x, q.y, q.z), x);
	return f3(GI_ADD(x.x, GI_MUL(2.0f, GI_SUB(GI_ADD(GI_MUL(q.w, t.x), GI_MUL(q.y, t.z)), GI_MUL(q.z, t.y)))),
	          GI_ADD(x.y, GI_MUL(2.0f, GI_SUB(GI_ADD(GI_MUL(q.w, t.y), GI_MUL(q.z, t.x)), GI_MUL(q.x, t.z)))),
	          GI_ADD(x.z, GI_MUL(2.0f, GI_SUB(GI_ADD(GI_MUL(q.w, t.z), GI_MUL(q.x, t.y)), GI_MUL(q.y, t.x)))));
}

// FetchIntersectBoxNode / FetchIntersectOBoxNode, intersectiontests.glsl:100-128: node test + tree pointers of node `index`
GI_HD bool fetchIntersectNode(const DScene& s, F3 o, F3 d, F3 inv, int index, float& firstHit, float& lastHit, uint32_t& childCode, int& escape)
{
	if (s.oboxNodes)
	{
		const F4 a = ld4(s.oboxNodes + 3 * size_t(index)), b = ld4(s.oboxNodes + 3 * size_t(index) + 1), q = ld4(s.oboxNodes + 3 * size_t(index) + 2);
		childCode = uint32_t(f2i(a.w));
		escape = f2i(b.w);
		const F3 ro = rotateByQuaternion(subx(o, f3(a.x, a.y, a.z)), q);
		const F3 rd = rotateByQuaternion(d, q);
		return intersectBox(ro, invDir(rd), -b.x, -b.y, -b.z, b.x, b.y, b.z, firstHit, lastHit);
	}
	const F4 a = ld4(s.refNodes + 2 * size_t(index)), b = ld4(s.refNodes + 2 * size_t(index) + 1);
	childCode = uint32_t(f2i(a.w));
	escape = f2i(b.w);
	return intersectBox(o, inv, a.x, a.y, a.z, b.x, b.y, b.z, firstHit, lastHit);
}

struct TraceStats { uint32_t nodes, tris; }; // traversal work actually done by this thread (not the reference counts)

struct Hit
{
	float t;        // rayLength after the call
	int leafTri;    // HIT_INDEX_OUTPUT .y, -1 = miss
	F3 bary;
	F3 normal;      // TRINORMAL_OUTPUT (un-normalised)
	I4 triangle;    // (v0, v1, v2, material)
};

GI_HD void loadLeafTri(const DScene& s, int leafTri, F3& p0, F3& p1, F3& p2, int& valid)
{
	const F4* p = s.leafPos + 3 * size_t(leafTri);
	F4 a = ld4(p), b = ld4(p + 1), c = ld4(p + 2);
	p0 = f3(a.x, a.y, a.z); p1 = f3(b.x, b.y, b.z); p2 = f3(c.x, c.y, c.z);
	valid = (a.x == a.x) ? 1 : 0;
}

// ------------------------------------------------------------------ reference-order traversal (traceray.glsl:25-180)
// Stackless escape-pointer walk in exactly the reference's visit order.  Used (a) as
// GI_TRAVERSAL_REFERENCE and (b) as the exact fallback for rays the fast traversal flags as
// ambiguous.  Vertex positions come from the pre-gathered leaf block (identical floats).
template <bool ANY_HIT>
GI_HD bool traceReference(const DScene& s, F3 o, F3 d, float& rayLength, int& hitLeafTri, TraceStats* st)
{
	int currentNodeIndex = 0;
	int currentLeafIndex = 0;
	hitLeafTri = -1;
	const F3 inv = invDir(d);
	bool nextIsLeafNode = false;
	do
	{
		if (!nextIsLeafNode)
		{
			if (st) ++st->nodes;
			uint32_t childCode;
			int escape;
			float newHit, exitDist;
			if (fetchIntersectNode(s, o, d, inv, currentNodeIndex, newHit, exitDist, childCode, escape) && newHit <= rayLength)
			{
				currentNodeIndex = int(childCode & 0x7FFFFFFFu);
				nextIsLeafNode = uint32_t(currentNodeIndex) != childCode;
				if (nextIsLeafNode)
				{
					currentLeafIndex = 2 * currentNodeIndex;
					currentNodeIndex = escape;
				}
			}
			else
				currentNodeIndex = escape;
		}
		if (nextIsLeafNode)
		{
			F3 p0, p1, p2; int valid;
			loadLeafTri(s, currentLeafIndex, p0, p1, p2, valid);
			if (!valid)
				nextIsLeafNode = false;
			else
			{
				if (st) ++st->tris;
				float newHit; F3 nb, nn;
				if (intersectTriangle(o, d, p0, p1, p2, newHit, nb, nn) && newHit < rayLength)
				{
					if (ANY_HIT) return true;
					rayLength = newHit;
					hitLeafTri = currentLeafIndex;
				}
				++currentLeafIndex;
				nextIsLeafNode = (currentLeafIndex % 2) != 0;
			}
		}
	} while (currentNodeIndex != 0 || nextIsLeafNode);
	return false;
}

// ------------------------------------------------------------------ fast traversal
// Same tree, same per-node box tests, same per-triangle tests as the reference, but children are
// visited near-to-far from a short per-thread stack.  A different visit order can only change the
// result when two candidate triangles have (nearly) the same distance; those rays are detected
// (second-closest hit inside a relative+absolute window of the closest, or a stack overflow) and
// re-traced with traceReference().  DESIGN.md "Why the fast traversal is exact" has the argument.
#ifndef GI_STACK_SIZE
#define GI_STACK_SIZE 48
#endif
#define GI_AMBIGUITY_REL 1.0e-4f

struct StackEntry { int node; float first; };

template <bool ANY_HIT>
GI_HD bool traceFast(const DScene& s, F3 o, F3 d, float& rayLength, int& hitLeafTri, TraceStats* st, bool* usedFallback)
{
	const F3 inv = invDir(d);
	const float tMaxInit = rayLength;
	hitLeafTri = -1;
	// root box, tested first by the reference too (traceray.glsl:56-69)
	{
		F4 a = ld4(s.refNodes), b = ld4(s.refNodes + 1);
		float f, l;
		if (st) ++st->nodes;
		if (!(intersectBox(o, inv, a.x, a.y, a.z, b.x, b.y, b.z, f, l) && f <= rayLength))
			return false;
	}
	float best = rayLength;     // closest accepted hit so far
	float second = GI_INF;      // second-closest candidate hit (closest-hit only)
	bool ambiguous = false;
	StackEntry stack[GI_STACK_SIZE];
	int sp = 0;
	int node = s.rootIsLeaf ? ~0 : 0;
	// culling limit: exact for any-hit; widened for closest-hit so that every candidate inside the
	// ambiguity window of the final hit is seen.
	const float margin = (1.0f + GI_AMBIGUITY_REL) * (1.0f + GI_AMBIGUITY_REL);
	for (;;)
	{
		if (node >= 0)
		{
			const F4* p = s.fastNodes + 4 * size_t(node);
			F4 n0 = ld4(p), n1 = ld4(p + 1), n2 = ld4(p + 2), n3 = ld4(p + 3);
			if (st) st->nodes += 2;
			float f0, l0, f1, l1;
			bool h0 = intersectBox(o, inv, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, f0, l0);
			bool h1 = intersectBox(o, inv, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, f1, l1);
			const float limit = ANY_HIT ? rayLength : (best * margin + s.ambiguityAbs);
			h0 = h0 && f0 <= limit;
			h1 = h1 && f1 <= limit;
			int c0 = f2i(n3.x), c1 = f2i(n3.y);
			if (h0 && h1)
			{
				if (f1 < f0) { int t = c0; c0 = c1; c1 = t; float tf = f0; f0 = f1; f1 = tf; }
				if (sp < GI_STACK_SIZE) { stack[sp].node = c1; stack[sp].first = f1; ++sp; }
				else ambiguous = true; // overflow: the skipped subtree is covered by the exact re-trace
				node = c0;
				continue;
			}
			if (h0) { node = c0; continue; }
			if (h1) { node = c1; continue; }
		}
		else
		{
			// leaf: two triangle slots (TRIANGLES_PER_LEAF 2, scenedata.glsl:5)
			const int leaf = ~node;
GI_UNROLL
			for (int slot = 0; slot < 2; ++slot)
			{
				const int lt = 2 * leaf + slot;
				F3 p0, p1, p2; int valid;
				loadLeafTri(s, lt, p0, p1, p2, valid);
				if (!valid) break;
				if (st) ++st->tris;
				float t; F3 nb, nn;
				if (intersectTriangle(o, d, p0, p1, p2, t, nb, nn) && t < tMaxInit)
				{
					if (ANY_HIT) return true;
					if (t < best)
					{
						if (hitLeafTri >= 0) second = best;
						best = t;
						hitLeafTri = lt;
					}
					else
						second = fminf(second, t);
				}
			}
		}
		// pop
		for (;;)
		{
			if (sp == 0) goto done;
			--sp;
			node = stack[sp].node;
			const float limit = ANY_HIT ? rayLength : (best * margin + s.ambiguityAbs);
			if (stack[sp].first <= limit) break;
		}
	}
done:
	if (ANY_HIT)
	{
		// a stack overflow skipped a subtree: decide exactly with the stackless walk
		if (ambiguous) { float len = tMaxInit; return traceReference<true>(s, o, d, len, hitLeafTri, st); }
		return false;
	}
	if (hitLeafTri >= 0)
	{
		const float window = best * GI_AMBIGUITY_REL + s.ambiguityAbs;
		if (second <= best + window) ambiguous = true;          // another triangle (almost) as close: order matters
		if (best + window >= tMaxInit) ambiguous = true;        // hit (almost) at the caller's length limit
	}
	if (ambiguous)
	{
		if (usedFallback) *usedFallback = true;
		rayLength = tMaxInit;
		return traceReference<false>(s, o, d, rayLength, hitLeafTri, st);
	}
	rayLength = best;
	return false;
}

// ------------------------------------------------------------------ ordered walk of an oriented-box hierarchy (bvhType = 1)
// The reference reaches a node of an OBox hierarchy iff the oriented-box test of EVERY ancestor passed (intersectiontests.glsl:114-128; unlike bim's
// AABoxes, independently fitted oriented boxes are not nested, so the leaf's own box does not imply its ancestors').  This walk therefore stays on
// the reference's own tree and applies the reference's own test (fetchIntersectNode: same rotation, same slab arithmetic) to every node it reaches --
// but takes the nearer child first from a short stack and culls by entry distance, where the stackless reference walk tests every node whose
// ancestors were hit in file order.  The set of leaves reachable without the distance cull is identical by construction; the cull and the visit
// order can only matter among candidates inside the near-tie window, which are detected exactly as in traceFast() and re-traced in reference order.
template <bool ANY_HIT>
GI_HD bool traceOrderedOBox(const DScene& s, F3 o, F3 d, float& rayLength, int& hitLeafTri, TraceStats* st, bool* usedFallback)
{
	const F3 inv = invDir(d);
	const float tMaxInit = rayLength;
	const float margin = (1.0f + GI_AMBIGUITY_REL) * (1.0f + GI_AMBIGUITY_REL);
	hitLeafTri = -1;
	float best = tMaxInit, second = GI_INF;
	bool ambiguous = false;
	StackEntry stack[GI_STACK_SIZE];
	int sp = 0;
	// `node` is a node whose own box test has passed; the root's comes first, as in the reference (traceray.glsl:56-69)
	int node = 0;
	uint32_t childCode; int escape;
	{
		float f, l;
		if (st) ++st->nodes;
		if (!(fetchIntersectNode(s, o, d, inv, 0, f, l, childCode, escape) && f <= tMaxInit)) { return false; }
	}
	for (;;)
	{
		const F4 a = ld4(s.oboxNodes + 3 * size_t(node));
		childCode = uint32_t(f2i(a.w));
		if (childCode & 0x80000000u)
		{
			// leaf node: both triangle slots (traceray.glsl:122-170)
			const int leaf = int(childCode & 0x7fffffffu);
			for (int slot = 0; slot < 2; ++slot)
			{
				const int lt = 2 * leaf + slot;
				F3 p0, p1, p2; int valid;
				loadLeafTri(s, lt, p0, p1, p2, valid);
				if (!valid) break;
				if (st) ++st->tris;
				float t; F3 nb, nn;
				if (intersectTriangle(o, d, p0, p1, p2, t, nb, nn) && t < tMaxInit)
				{
					if (ANY_HIT) return true;
					if (t < best) { if (hitLeafTri >= 0) second = best; best = t; hitLeafTri = lt; }
					else second = fminf(second, t);
				}
			}
		}
		else
		{
			const float limit = ANY_HIT ? tMaxInit : (best * margin + s.ambiguityAbs);
			const int left = int(childCode);
			uint32_t c0, c1; int e0, e1;
			float f0, l0, f1, l1;
			if (st) st->nodes += 2;
			const bool h0 = fetchIntersectNode(s, o, d, inv, left, f0, l0, c0, e0) && f0 <= limit;
			const int right = e0; // the second child is the node behind the first child's subtree
			const bool h1 = fetchIntersectNode(s, o, d, inv, right, f1, l1, c1, e1) && f1 <= limit;
			if (h0 && h1)
			{
				const bool swap = f1 < f0;
				const int nearNode = swap ? right : left, farNode = swap ? left : right;
				const float farFirst = swap ? f0 : f1;
				if (sp < GI_STACK_SIZE) { stack[sp].node = farNode; stack[sp].first = farFirst; ++sp; }
				else ambiguous = true; // the subtree is skipped: the ray is answered by the stackless reference walk
				node = nearNode;
				continue;
			}
			if (h0) { node = left; continue; }
			if (h1) { node = right; continue; }
		}
		// pop
		for (;;)
		{
			if (sp == 0) goto done;
			--sp;
			node = stack[sp].node;
			const float limit = ANY_HIT ? tMaxInit : (best * margin + s.ambiguityAbs);
			if (stack[sp].first <= limit) break;
		}
	}
done:
	if (ANY_HIT)
	{
		if (ambiguous) { float len = tMaxInit; return traceReference<true>(s, o, d, len, hitLeafTri, st); }
		return false;
	}
	if (hitLeafTri >= 0)
	{
		const float window = best * GI_AMBIGUITY_REL + s.ambiguityAbs;
		if (second <= best + window) ambiguous = true;
		if (best + window >= tMaxInit) ambiguous = true;
	}
	if (ambiguous)
	{
		if (usedFallback) *usedFallback = true;
		rayLength = tMaxInit;
		return traceReference<false>(s, o, d, rayLength, hitLeafTri, st);
	}
	rayLength = best;
	return false;
}

// Closest hit with all outputs of TraceRay (traceray.glsl:25-48): the winning triangle is re-tested
// once to produce barycentrics and the geometric normal (same inputs => same bits as during the walk).
GI_HD void traceClosest(const DScene& s, int traversalMode, F3 o, F3 d, float tMax, Hit& h, TraceStats* st, bool* usedFallback)
{
	h.t = tMax;
	h.leafTri = -1;
	if (traversalMode == GI_TRAVERSAL_REFERENCE)
	{
		if (s.oboxOrdered) traceOrderedOBox<false>(s, o, d, h.t, h.leafTri, st, usedFallback);
		else traceReference<false>(s, o, d, h.t, h.leafTri, st);
	}
	else traceFast<false>(s, o, d, h.t, h.leafTri, st, usedFallback);
	if (h.leafTri >= 0)
	{
		F3 p0, p1, p2; int valid;
		loadLeafTri(s, h.leafTri, p0, p1, p2, valid);
		float t;
		intersectTriangle(o, d, p0, p1, p2, t, h.bary, h.normal);
		h.triangle = ld4(s.leafTris + h.leafTri);
	}
	else
	{
		h.bary = f3(0.0f); h.normal = f3(0.0f);
		h.triangle.x = h.triangle.y = h.triangle.z = h.triangle.w = 0;
	}
}

GI_HD bool traceAny(const DScene& s, int traversalMode, F3 o, F3 d, float rayLength, TraceStats* st)
{
	int dummy; float len = rayLength;
	if (traversalMode == GI_TRAVERSAL_REFERENCE)
	{
		if (s.oboxOrdered) return traceOrderedOBox<true>(s, o, d, len, dummy, st, nullptr);
		return traceReference<true>(s, o, d, len, dummy, st);
	}
	return traceFast<true>(s, o, d, len, dummy, st, nullptr);
}

// ------------------------------------------------------------------ RNG (random.glsl:1-43)
GI_HD uint32_t wanghash(uint32_t x)
{
	x = (x ^ 61u) ^ (x >> 16);
	x *= 9u;
	x = x ^ (x >> 4);
	x *= 0x27d4eb2du;
	x = x ^ (x >> 15);
	return x;
}
GI_HD uint32_t randomUInt(uint32_t& seed)
{
	seed ^= (seed << 13);
	seed ^= (seed >> 17);
	seed ^= (seed << 5);
	return seed;
}
GI_HD uint32_t initRandomSeed(uint32_t frameSeed, uint32_t invocationID) { return wanghash(frameSeed + invocationID); }
// random.glsl:38-43: float(u) / 4294967297.0 with a float literal == 2^32 (can return exactly 1.0)
GI_HD float randomFloat(uint32_t& seed) { return float(randomUInt(seed)) * 2.3283064365386963e-10f; }
GI_HD F2 random2(uint32_t& seed) { F2 r; r.x = randomFloat(seed); r.y = randomFloat(seed); return r; }

GI_HD void sincosPhi(float phi, float& s, float& c)
{
#if defined(__CUDA_ARCH__)
	sincosf(phi, &s, &c);
#else
	s = sinf(phi); c = cosf(phi);
#endif
}

// random.glsl:51-87
GI_HD F3 sampleDirection(F2 rs)
{
	float cosTheta = rs.x * 2.0f - 1.0f;
	float sinTheta = sqrtf((1.0f - cosTheta) * (1.0f + cosTheta));
	float sp, cp; sincosPhi(rs.y * GI_PI_2, sp, cp);
	return normalize(f3(sinTheta * sp, sinTheta * cp, cosTheta));
}
GI_HD F3 sampleHemisphereCosine(F2 rs, F3 U, F3 V, F3 W)
{
	float sp, cp; sincosPhi(GI_PI_2 * rs.x, sp, cp);
	float sinTheta = sqrtf(rs.y);
	float x = sinTheta * cp, y = sinTheta * sp, z = sqrtf(1.0f - rs.y);
	return x * U + y * V + z * W;
}
GI_HD F3 samplePhongLobe(F2 rs, float exponent, F3 U, F3 V, F3 W)
{
	float sp, cp; sincosPhi(GI_PI_2 * rs.x, sp, cp);
	float power = powf(rs.y, 1.0f / (exponent + 1.0f));
	float scale = sqrtf(1.0f - power * power);
	return (cp * scale) * U + (sp * scale) * V + power * W;
}
// helper.glsl:22-35
GI_HD void createONB(F3 n, F3& U, F3& V)
{
	U = cross(n, f3(0.0f, 1.0f, 0.0f));
	if (fabsf(U.x) < 0.0001f && fabsf(U.y) < 0.0001f && fabsf(U.z) < 0.0001f)
		U = cross(n, f3(1.0f, 0.0f, 0.0f));
	U = normalize(U);
	V = cross(n, U);
}
GI_HD float luminance(F3 rgb) { return dot(rgb, f3(0.212671f, 0.715160f, 0.072169f)); }
// helper.glsl:38-63
GI_HD F3 unpackNormal(float p0, float p1)
{
	float sinPhi = sqrtf(saturate((1.0f - p1) * (1.0f + p1)));
	float sp, cp; sincosPhi(p0, sp, cp);
	return f3(cp * sinPhi, sp * sinPhi, p1);
}
GI_HD F3 unpackNormalNoSat(float p0, float p1)
{
	float sinPhi = sqrtf(1.0f - p1 * p1);
	float sp, cp; sincosPhi(p0, sp, cp);
	return f3(cp * sinPhi, sp * sinPhi, p1);
}
GI_HD F2 packNormal(F3 n)
{
	F2 r;
	r.x = (n.x == 0.0f) ? sign(n.y) * GI_PI / 2 : atan2f(n.y, n.x);
	r.y = n.z;
	return r;
}

// ------------------------------------------------------------------ materials (materials.glsl)
struct MaterialData
{
	F4 Reflectiveness;
	F3 Opacity, Diffuse, Emissivity, Fresnel0, Fresnel1;
	float RefractionIndexAvg;
};

GI_HD float srgbToLinear(float c) { return (c <= 0.04045f) ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f); }

GI_HD F4 fetchTexel(const DTexel& t, int x, int y)
{
	size_t i = size_t(y) * t.width + x;
	F4 r;
	if (t.format == GI_TEX_RGBA32F)
		r = ld4(static_cast<const F4*>(t.pixels) + i);
	else
	{
		uint32_t p = static_cast<const uint32_t*>(t.pixels)[i];
		r.x = float(p & 255u) / 255.0f; r.y = float((p >> 8) & 255u) / 255.0f; r.z = float((p >> 16) & 255u) / 255.0f; r.w = float(p >> 24) / 255.0f;
		if (t.format == GI_TEX_SRGB8_ALPHA8) { r.x = srgbToLinear(r.x); r.y = srgbToLinear(r.y); r.z = srgbToLinear(r.z); }
	}
	return r;
}
GI_HD int wrapIndex(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

// fp32 bilinear filter (GL 4.5 spec 8.14.2, REPEAT): the GI_OPT_TEXTURE_FILTER=1 path.
GI_HD F4 sampleBilinearExact(const DTexel& t, float u, float v)
{
	float fu = u * float(t.width) - 0.5f, fv = v * float(t.height) - 0.5f;
	float iu = floorf(fu), iv = floorf(fv);
	float a = fu - iu, b = fv - iv;
	int i0 = wrapIndex(int(iu), t.width), i1 = wrapIndex(int(iu) + 1, t.width);
	int j0 = wrapIndex(int(iv), t.height), j1 = wrapIndex(int(iv) + 1, t.height);
	F4 t00 = fetchTexel(t, i0, j0), t10 = fetchTexel(t, i1, j0), t01 = fetchTexel(t, i0, j1), t11 = fetchTexel(t, i1, j1);
	float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	F4 r;
	r.x = t00.x * w00 + t10.x * w10 + t01.x * w01 + t11.x * w11;
	r.y = t00.y * w00 + t10.y * w10 + t01.y * w01 + t11.y * w11;
	r.z = t00.z * w00 + t10.z * w10 + t01.z * w01 + t11.z * w11;
	r.w = t00.w * w00 + t10.w * w10 + t01.w * w01 + t11.w * w11;
	return r;
}

GI_HD F4 sampleMaterialTexture(const DScene& s, const DMaterial* m, int materialID, int slot, F4 constant, F2 tc)
{
#if defined(__CUDA_ARCH__)
	unsigned long long tex = m->tex[slot];
	if (tex == 0ull) return constant;
	if (s.textureFilterExact) return sampleBilinearExact(s.texels[4 * materialID + slot], tc.x, tc.y);
	float4 v = tex2D<float4>(cudaTextureObject_t(tex), tc.x, tc.y);
	F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
	return r;
#else
	if (m->tex[slot] == 0ull) return constant;
	return sampleBilinearExact(s.texels[4 * materialID + slot], tc.x, tc.y);
#endif
}

// SampleMaterialData, materials.glsl:19-30
GI_HD MaterialData sampleMaterialData(const DScene& s, int materialID, F2 tc)
{
	const DMaterial* m = s.materials + materialID;
	MaterialData d;
	F4 cR = ld4(&m->cReflectiveness), cO = ld4(&m->cOpacity), cD = ld4(&m->cDiffuse), cE = ld4(&m->cEmissivity);
	F4 f0 = ld4(&m->fresnel0), f1 = ld4(&m->fresnel1);
	d.Reflectiveness = sampleMaterialTexture(s, m, materialID, 0, cR, tc);
	F4 o = sampleMaterialTexture(s, m, materialID, 1, cO, tc);
	F4 df = sampleMaterialTexture(s, m, materialID, 2, cD, tc);
	F4 e = sampleMaterialTexture(s, m, materialID, 3, cE, tc);
	d.Opacity = f3(o.x, o.y, o.z);
	d.Diffuse = f3(df.x, df.y, df.z);
	d.Emissivity = f3(e.x, e.y, e.z);
	d.Fresnel0 = f3(f0.x, f0.y, f0.z);
	d.Fresnel1 = f3(f1.x, f1.y, f1.z);
	d.RefractionIndexAvg = cD.w;
	return d;
}

// GetTriangleHitInfo, scenedata.glsl:94-105
GI_HD void getTriangleHitInfo(const DScene& s, I4 tri, F3 bary, F3& normal, F2& tc)
{
	F4 a = ld4(s.vertexInfo + tri.x), b = ld4(s.vertexInfo + tri.y), c = ld4(s.vertexInfo + tri.z);
	normal = normalize(unpackNormal(a.x, a.y) * bary.x + unpackNormal(b.x, b.y) * bary.y + unpackNormal(c.x, c.y) * bary.z);
	tc.x = a.z * bary.x + b.z * bary.y + c.z * bary.z;
	tc.y = a.w * bary.x + b.w * bary.y + c.w * bary.z;
}

GI_HD float avgProbability(F3 p) { return (p.x + p.y + p.z + GI_DIVISOR_EPSILON) / (3.0f + GI_DIVISOR_EPSILON); } // materials.glsl:33-36
GI_HD float avg(F3 p) { return (p.x + p.y + p.z) / 3.0f; }                                                       // :38-41
// materials.glsl:43-63, USE_CONSTANT_REFLECTION
GI_HD void bsdfDecisionProbabilities(const MaterialData& md, F3& preflect, F3& prefract, F3& pdiffuse)
{
	preflect = md.Fresnel0;
	prefract = f3(1.0f) - saturate(preflect);
	pdiffuse = prefract * md.Opacity;
	prefract = prefract - pdiffuse;
}

// __SampleBSDF, materials.glsl:118-277 (rejection loop disabled :232 => counter == 1)
// stopOnDiffuse = the STOP_ON_DIFFUSE_BOUNCE build of the Whitted ray tracer (:164-167): the diffuse decision ends the path
GI_HD F3 sampleBSDF(F3 incident, const MaterialData& md, uint32_t& seed, F3 N, F3& throughput, bool adjoint, float& pdf, bool stopOnDiffuse = false)
{
	float cosTheta = dot(N, incident);
	float cosThetaAbs = saturate(fabsf(cosTheta));
	F3 rrDir = incident - (2.0f * cosTheta) * N;
	F2 rs = random2(seed);
	F3 preflect, prefract, pdiffuse;
	bsdfDecisionProbabilities(md, preflect, prefract, pdiffuse);
	float avgPReflect = avgProbability(preflect), avgPRefract = avgProbability(prefract), avgPDiffuse = avg(pdiffuse);
	F3 pphong = preflect;
	float avgPPhong = avgPReflect;
	float decision = randomFloat(seed);
	if (avgPDiffuse > decision)
	{
		if (stopOnDiffuse) { throughput = f3(-1.0f, 0.0f, 0.0f); pdf = 0.0f; return f3(0.0f); }
		throughput = throughput * ((md.Diffuse * pdiffuse) / avgPDiffuse);
		if (cosTheta > 0.0f) N = -N;
		F3 U, V;
		createONB(N, U, V);
		F3 outDir = sampleHemisphereCosine(rs, U, V, N);
		pdf = saturate(dot(N, outDir)) / GI_PI;
		return outDir;
	}
	else if (avgPDiffuse + avgPRefract > decision)
	{
		pphong = prefract;
		avgPPhong = avgPRefract;
		float eta = cosTheta < 0.0f ? 1.0f / md.RefractionIndexAvg : md.RefractionIndexAvg;
		float etaSq = eta * eta;
		float sinTheta2Sq = etaSq * (1.0f - cosTheta * cosTheta);
		if (sinTheta2Sq < 1.0f)
		{
			rrDir = eta * incident + (sign(cosTheta) * (sqrtf(saturate(1.0f - sinTheta2Sq)) - eta * cosThetaAbs)) * N;
			if (!adjoint) throughput = throughput / etaSq;
		}
	}
	rrDir = normalize(rrDir);
	F3 RU, RV;
	createONB(rrDir, RU, RV);
	F3 outDir = samplePhongLobe(rs, md.Reflectiveness.w, RU, RV, rrDir);
	if (!adjoint) throughput = throughput * (fabsf(dot(N, outDir)) / fabsf(dot(N, rrDir)));
	float cosOutAbs = fabsf(dot(outDir, rrDir));
	pdf = (md.Reflectiveness.w + 1.0f) / GI_PI_2 * powf(cosOutAbs, md.Reflectiveness.w);
	throughput = throughput * (pphong / avgPPhong);
	return outDir;
}

// __BSDF, materials.glsl:306-391
// stopOnDiffuse: STOP_ON_DIFFUSE_BOUNCE, :355-358 (no refractive next event estimation)
GI_HD F3 evalBSDF(F3 incident, F3 excident, const MaterialData& md, F3 N, float& pdf, bool adjoint, bool stopOnDiffuse = false)
{
	float cosTheta = dot(N, incident);
	float cosThetaAbs = saturate(fabsf(cosTheta));
	float cosThetaOut = dot(N, excident);
	F3 preflect, prefract, pdiffuse;
	bsdfDecisionProbabilities(md, preflect, prefract, pdiffuse);
	float phongFactor = (md.Reflectiveness.w + 1.0f) / (2 * GI_PI_2 * fabsf(cosThetaOut) + GI_DIVISOR_EPSILON);
	float phongNormPdf = (md.Reflectiveness.w + 1.0f) / GI_PI_2;
	F3 rrDir = normalize(incident - (2.0f * cosTheta) * N);
	float powN = powf(saturate(dot(rrDir, excident)), md.Reflectiveness.w);
	float brdfFactor = phongFactor * powN;
	float pdfFactor = phongNormPdf * powN;
	F3 bsdf = preflect * brdfFactor;
	bsdf = bsdf * (fabsf(cosThetaOut) / cosThetaAbs);
	pdf = avgProbability(preflect) * pdfFactor;
	if (cosTheta * cosThetaOut < 0.0f)
	{
		bsdf = bsdf + pdiffuse * md.Diffuse / GI_PI;
		pdf += avg(pdiffuse) * saturate(cosThetaOut) / GI_PI;
	}
	if (stopOnDiffuse) return bsdf;
	float eta = cosTheta < 0.0f ? 1.0f / md.RefractionIndexAvg : md.RefractionIndexAvg;
	float etaSq = eta * eta;
	float sinTheta2Sq = etaSq * (1.0f - cosTheta * cosTheta);
	if (sinTheta2Sq < 1.0f)
	{
		rrDir = normalize(eta * incident + (sign(cosTheta) * (sqrtf(saturate(1.0f - sinTheta2Sq)) - eta * cosThetaAbs)) * N);
		float powR = powf(saturate(dot(rrDir, excident)), md.Reflectiveness.w);
		brdfFactor = phongFactor * powR;
		pdfFactor = phongNormPdf * powR;
		if (!adjoint) brdfFactor /= etaSq;
		brdfFactor *= fabsf(dot(N, excident)) / fabsf(dot(N, rrDir));
	}
	bsdf = bsdf + prefract * brdfFactor;
	pdf += avgProbability(prefract) * pdfFactor;
	return bsdf;
}
GI_HD F3 BSDF(F3 in, F3 out, const MaterialData& md, F3 N, float& pdf) { return evalBSDF(in, out, md, N, pdf, false); }              // materials.glsl:393-396
GI_HD F3 adjointBSDF(F3 in, F3 out, const MaterialData& md, F3 N, float& pdf) { return evalBSDF(-out, -in, md, N, pdf, true); }      // :405-408

// materials.glsl:415-436
GI_HD float adjointShadingNormalCorrectedOutDotN(F3 incomingLight, F3 toCamera, F3 gN, F3 sN)
{
	if (dot(incomingLight, gN) * dot(incomingLight, sN) < 0 || dot(toCamera, gN) * dot(toCamera, sN) < 0) return 0.0f;
	return saturate(fabsf(dot(incomingLight, sN) * dot(toCamera, gN)) / fmaxf(fabsf(dot(incomingLight, gN)), GI_DIVISOR_EPSILON));
}
GI_HD float adjointShadingNormalCorrection(F3 inDir, F3 outDir, F3 gN, F3 sN)
{
	if (dot(inDir, gN) * dot(inDir, sN) < 0 || dot(outDir, gN) * dot(outDir, sN) < 0) return 0.0f;
	return saturate(fabsf(dot(inDir, sN) * dot(outDir, gN)) / fmaxf(fabsf(dot(inDir, gN) * dot(outDir, sN)), GI_DIVISOR_EPSILON));
}

// ------------------------------------------------------------------ frame state (the UBOs of globaluniforms.glsl)
struct DFrame
{
	DScene scene;
	GiCamera cam;
	int width, height;                 // BackbufferSize
	const F4* lightSamples;            // 2 x F4 per GiLightSample
	int numLightSamples;               // NumInitialLightSamples
	uint32_t frameSeed;                // FrameSeed
	int maxPathLength;                 // MAX_PATHLENGTH
	int traversalMode;
	float* accum;                      // RGBA32F sum image
	F3 envColor;                       // constant environment (1x1 cube) ...
	unsigned long long envCube;        // ... or a cube map texture object
	// light tracer / BPT
	int numRaysPerLightSample;
	F4* lightCache;                    // 4 x F4 per LightCacheEntry (lightcache.glsl:1-16)
	uint32_t lightCacheCapacity;
	uint32_t* lightCacheCount;         // NumLightCacheEntries
	uint32_t* lightPathLengthSum;      // LightPathLengthSum
	// pipelined BPT batches (gi_wavefront.cuh): iteration i uses cache (i % cacheRing) of a ring of caches `lightCacheCapacity`
	// entries apart, and the counter pair 2 * (i % cacheRing) behind lightCacheCount.  cacheRing <= 1: the single cache above.
	uint32_t cacheRing;
	// shader #define variants (off by default, as in the reference's build): RECONSTRUCTION_FILTER (screenreproject.glsl:14-17) and CHAINED_MIS
	int reconFilter;                   // 0 box, 1 Mitchell 1/3 1/3, 2 cone, 3 B-spline
	int chainedMIS;                    // one-thread-per-path kernels only (the eye path carries g_mis[MAX_PATHLENGTH])
};

struct PathCounters
{
	uint32_t closestRays, anyRays, cameraPaths, lightPaths, shadedVertices, splats, fallbackRays;
};

#if defined(__CUDACC__)
// WriteAtomic (screenreproject.glsl:19-48) as one 128-bit reduction (red.global.add.v4.f32, sm_90+, instead of the reference's three scalar
// imageAtomicAdd; alpha += 0 keeps LPT's "alpha untouched").  WF_SPLAT_AGGREGATE = 1 additionally groups the lanes that splat in one instruction by
// target pixel (match.any on the address): a group of one issues its own reduction, a larger group sums its values in lane order and lets its first
// lane issue it.  Measured on B200 (ms per iteration, aggregated / plain): C4 LPT 4K 6.502 / 6.310, C3 BPT 7.355 / 7.293 -- camera connections land
// all over the frame, groups of more than one lane are rare, and match.any + the votes cost more than the saved reductions: off.
#ifndef WF_SPLAT_AGGREGATE
#define WF_SPLAT_AGGREGATE 0
#endif
__device__ __forceinline__ void splatReduce(float* p, F3 v)
{
#if WF_SPLAT_AGGREGATE
	unsigned lane; asm("mov.u32 %0, %%laneid;" : "=r"(lane));
	const unsigned self = 1u << lane;
	const unsigned peers = __match_any_sync(__activemask(), reinterpret_cast<unsigned long long>(p));
	if (peers != self)
	{
		float sx = 0.0f, sy = 0.0f, sz = 0.0f;
		for (unsigned m = peers; m; m &= m - 1u)
		{
			const int src = __ffs(m) - 1;
			sx += __shfl_sync(peers, v.x, src); sy += __shfl_sync(peers, v.y, src); sz += __shfl_sync(peers, v.z, src);
		}
		if ((peers & (self - 1u)) != 0u) return; // not the first lane of its group
		v.x = sx; v.y = sy; v.z = sz;
	}
#endif
	atomicAdd(reinterpret_cast<float4*>(p), make_float4(v.x, v.y, v.z, 0.0f));
}
#endif

// The accumulator update primitives differ between device (atomics) and emulation.
GI_HD void splatAtomic(float* accum, int width, int height, int px, int py, F3 v, PathCounters& pc)
{
	if (px < 0 || py < 0 || px >= width || py >= height) return; // round() may yield W or H
	++pc.splats;
	float* p = accum + 4 * (size_t(py) * width + px);
#if defined(__CUDA_ARCH__)
	splatReduce(p, v);
#else
	p[0] += v.x; p[1] += v.y; p[2] += v.z;
#endif
}

GI_HD F3 cameraRayDirection(const DFrame& f, uint32_t& seed, int gx, int gy)
{
	// pathtracer.comp:39-46
	F2 r = random2(seed);
	float sx = (r.x - 0.5f + float(gx)) / float(f.width) * 2.0f - 1.0f;
	float sy = (r.y - 0.5f + float(gy)) / float(f.height) * 2.0f - 1.0f;
	return normalize(sx * f3(f.cam.u) + sy * f3(f.cam.v) + f3(f.cam.w));
}

GI_HD F3 environmentRadiance(const DFrame& f, F3 dir)
{
#if defined(__CUDA_ARCH__)
	if (f.envCube != 0ull)
	{
		float4 v = texCubemap<float4>(cudaTextureObject_t(f.envCube), dir.x, dir.y, dir.z);
		return f3(v.x, v.y, v.z);
	}
#endif
	return f.envColor;
}

// EstimateDirectLight, lightsource.glsl:3-93, split in two: everything up to the shadow ray (prepareDirectLight) and the
// any-hit query.  The value returned for an unoccluded sample does not depend on the query, so the wavefront scheduler
// evaluates it before the shadow ray is traced; the megakernel calls both halves back to back (estimateDirectLight).
struct DirectLightQuery
{
	F3 origin, dir;   // shadow ray
	float len;        // rayLength of the any-hit query
	F3 radiance;      // EstimateDirectLight's result when the ray is free
};

GI_HD bool prepareDirectLight(const F4* lightSamples, F3 pos, F3 normal, int lightSampleIndex, F3 viewDir, const MaterialData& md, bool sampleAreaLights,
                              DirectLightQuery& q, bool stopOnDiffuse = false)
{
	F4 l0 = ld4(lightSamples + 2 * lightSampleIndex), l1 = ld4(lightSamples + 2 * lightSampleIndex + 1);
	F3 dir = f3(l0.x, l0.y, l0.z) - pos;
	float lightDistSq = dot(dir, dir) + GI_DIVISOR_EPSILON;
	float lightDist = sqrtf(lightDistSq);
	dir = dir / lightDist;
	float surfaceCos = saturate(dot(dir, normal));
	if (surfaceCos <= 0.0f) return false;
	F3 intensity = f3(l1.x, l1.y, l1.z);
	if (l1.w > 1.5f)
	{
		// point light (lightsource.glsl:25-47)
	}
	else if (sampleAreaLights)
	{
		F3 lightNormal = unpackNormal(l0.w, l1.w);
		float factor = saturate(dot(lightNormal, -dir));
		if (!(factor > 0.0f)) return false;
		intensity = intensity * factor;
	}
	else
		return false;
	q.origin = GI_RAY_HIT_EPSILON * dir + pos;
	q.dir = dir;
	q.len = lightDist - GI_RAY_HIT_EPSILON * 2;
	float pdf;
	F3 bsdf = evalBSDF(viewDir, dir, md, normal, pdf, false, stopOnDiffuse);
	F3 irradiance = (surfaceCos / lightDistSq) * intensity;
	q.radiance = irradiance * bsdf;
	return true;
}

GI_HD F3 estimateDirectLight(const DFrame& f, F3 pos, F3 normal, int lightSampleIndex, F3 viewDir, const MaterialData& md, bool sampleAreaLights,
                             PathCounters& pc, TraceStats* st, bool stopOnDiffuse = false)
{
	DirectLightQuery q;
	if (!prepareDirectLight(f.lightSamples, pos, normal, lightSampleIndex, viewDir, md, sampleAreaLights, q, stopOnDiffuse)) return f3(0.0f);
	++pc.anyRays;
	if (traceAny(f.scene, f.traversalMode, q.origin, q.dir, q.len, st)) return f3(0.0f);
	return q.radiance;
}

// pathtracer.comp:16-110, one invocation (pixel gx,gy)
GI_HD void pathtracePixel(const DFrame& f, int gx, int gy, PathCounters& pc, TraceStats* st)
{
	const float lightSampleWeight = float(f.numLightSamples);
	uint32_t seed = initRandomSeed(f.frameSeed, uint32_t(gx + gy * f.width));
	F3 color = f3(0.0f), pathThroughput = f3(1.0f);
	bool didNEE = false;
	F3 origin = f3(f.cam.position);
	F3 dir = cameraRayDirection(f, seed, gx, gy);
	++pc.cameraPaths;
	for (int i = 0; i < f.maxPathLength; ++i)
	{
		Hit h; bool fb = false;
		++pc.closestRays;
		traceClosest(f.scene, f.traversalMode, origin, dir, GI_RAY_MAX, h, st, &fb);
		if (fb) ++pc.fallbackRays;
		if (h.t == GI_RAY_MAX)
		{
			color = color + pathThroughput * environmentRadiance(f, dir);
			break;
		}
		origin = origin + h.t * dir;
		F3 hitNormal; F2 tc;
		getTriangleHitInfo(f.scene, h.triangle, h.bary, hitNormal, tc);
		MaterialData md = sampleMaterialData(f.scene, h.triangle.w, tc);
		++pc.shadedVertices;
		if (!didNEE) color = color + pathThroughput * md.Emissivity;
		didNEE = md.Reflectiveness.w < 500.0f;
		{
			int sampleIndex = int(randomUInt(seed) % uint32_t(f.numLightSamples));
			F3 radiance = estimateDirectLight(f, origin, hitNormal, sampleIndex, dir, md, didNEE, pc, st);
			color = color + radiance * pathThroughput * lightSampleWeight;
		}
		F3 throughput = f3(1.0f);
		float pdf;
		dir = sampleBSDF(dir, md, seed, hitNormal, throughput, false, pdf);
		origin = origin + dir * GI_RAY_HIT_EPSILON;
		pathThroughput = pathThroughput * throughput;
		float cont = saturate(luminance(throughput));
		if (randomFloat(seed) >= cont) break;
		pathThroughput = pathThroughput / cont;
	}
	// pixel-exclusive read-modify-write (pathtracer.comp:107-109), alpha += 0
	float* px = f.accum + 4 * (size_t(gy) * f.width + gx);
#if defined(__CUDA_ARCH__)
	float4 v = *reinterpret_cast<float4*>(px);
	v.x += color.x; v.y += color.y; v.z += color.z;
	*reinterpret_cast<float4*>(px) = v;
#else
	px[0] += color.x; px[1] += color.y; px[2] += color.z;
#endif
}

// whittedraytracer.comp:16-133 (STOP_ON_DIFFUSE_BOUNCE build, whittedraytracer.cpp:38-49), one invocation: specular chain from the camera,
// two next-event samples at the first diffuse decision, alpha += 1
GI_HD void whittedPixel(const DFrame& f, int gx, int gy, PathCounters& pc, TraceStats* st)
{
	const int NUM_LIGHT_SAMPLES = 2;
	const float lightSampleWeight = float(f.numLightSamples) / float(NUM_LIGHT_SAMPLES);
	uint32_t seed = wanghash(f.frameSeed); // InitCoherentRandomSeed (random.glsl:23-28): one stream for all pixels
	F2 r = random2(seed);
	float sx = (r.x + float(gx)) / float(f.width) * 2.0f - 1.0f;
	float sy = (r.y + float(gy)) / float(f.height) * 2.0f - 1.0f;
	F3 origin = f3(f.cam.position);
	F3 dir = normalize(sx * f3(f.cam.u) + sy * f3(f.cam.v) + f3(f.cam.w));
	++pc.cameraPaths;
	F3 color = f3(0.0f), pathThroughput = f3(1.0f);
	for (int i = 0; i < f.maxPathLength; ++i)
	{
		Hit h; bool fb = false;
		++pc.closestRays;
		traceClosest(f.scene, f.traversalMode, origin, dir, GI_RAY_MAX, h, st, &fb);
		if (fb) ++pc.fallbackRays;
		if (h.t == GI_RAY_MAX) break;
		origin = origin + h.t * dir;
		F3 hitNormal; F2 tc;
		getTriangleHitInfo(f.scene, h.triangle, h.bary, hitNormal, tc);
		MaterialData md = sampleMaterialData(f.scene, h.triangle.w, tc);
		++pc.shadedVertices;
		color = color + pathThroughput * md.Emissivity;
		F3 pdiffuse = (f3(1.0f) - saturate(md.Fresnel0)) * md.Opacity; // GetDiffuseProbability, materials.glsl:96-101
		float pdiffuseAvg = luminance(pdiffuse);
		F3 throughput = f3(1.0f);
		float pdf;
		F3 sampleDirection = sampleBSDF(dir, md, seed, hitNormal, throughput, false, pdf, true);
		if (throughput.x == -1.0f)
		{
			for (int ls = 0; ls < NUM_LIGHT_SAMPLES; ++ls)
			{
				int sampleIndex = int(randomUInt(seed) % uint32_t(f.numLightSamples));
				F3 radiance = estimateDirectLight(f, origin, hitNormal, sampleIndex, dir, md, true, pc, st, true);
				color = color + radiance * pathThroughput * (lightSampleWeight / (pdiffuseAvg + GI_DIVISOR_EPSILON));
			}
			break;
		}
		dir = sampleDirection;
		origin = origin + dir * GI_RAY_HIT_EPSILON;
		pathThroughput = pathThroughput * (throughput / (luminance(throughput) + GI_DIVISOR_EPSILON));
	}
	float* px = f.accum + 4 * (size_t(gy) * f.width + gx);
#if defined(__CUDA_ARCH__)
	float4 v = *reinterpret_cast<float4*>(px);
	v.x += color.x; v.y += color.y; v.z += color.z; v.w += 1.0f;
	*reinterpret_cast<float4*>(px) = v;
#else
	px[0] += color.x; px[1] += color.y; px[2] += color.z; px[3] += 1.0f;
#endif
}

// lightcache.glsl:61-88
GI_HD float misHeuristic(float p) { return fminf(1e18f, p); }
GI_HD float distanceMISHeuristic(float d) { return misHeuristic(d * d); }

// bidirpathtracer/chained_mis.glsl:1-34 (g_mis: the eye path's array of per-segment MIS weights)
#define GI_MAX_PATHLENGTH 64
GI_HD void updateQuality(float& prevQuality, float mis) { prevQuality = prevQuality * (mis + 1.0f); }
GI_HD float qualityWeightedMIS_NEE(float mis_nee, float quality_nee, float quality_trace)
{
	return mis_nee * quality_nee / fmaxf(mis_nee * quality_nee + (1.0f - mis_nee) * quality_trace, GI_DIVISOR_EPSILON);
}
GI_HD float qualityWeightedMIS_Trace(float mis_trace, float quality_nee, float quality_trace)
{
	return mis_trace * quality_trace / fmaxf((1.0f - mis_trace) * quality_nee + mis_trace * quality_trace, GI_DIVISOR_EPSILON);
}
// cmisPathThroughputWeight with g_mis[pathSegment - 1] taken from `last` (the shader writes it into the array right before the call)
GI_HD float cmisPathThroughputWeight(const float* gMis, float last, int pathSegment, float quality_nee, float quality_trace)
{
	float w = 1.0f;
	for (int i = pathSegment - 1; i >= 0; --i)
	{
		const float m = (i == pathSegment - 1) ? last : gMis[i];
		w *= qualityWeightedMIS_Trace(m, quality_nee, quality_trace);
		updateQuality(quality_nee, m);
		updateQuality(quality_trace, 1.0f - m);
	}
	return w;
}

// reconstruction filter kernels of screenreproject.glsl:186-223 over the distance to the pixel centre
GI_HD float reconstructionFilterValue(int filter, float d)
{
	float v = 0.0f;
	if (filter == 1)
	{
		if (d < 1.0f) v = (7.0f * d * d * d - 12.0f * d * d + 16.0f / 3.0f) / 6.0f;
		else if (d < 2.0f) v = (-7.0f / 3.0f * d * d * d + 12.0f * d * d - 20.0f * d + 32.0f / 3.0f) / 6.0f;
	}
	else if (filter == 2)
	{
		if (d < 1.5f) v = 1.0f - d / 1.5f;
	}
	else if (filter == 3)
	{
		if (d < 1.0f) v = (3.0f * d * d * d - 6.0f * d * d + 4.0f) / 6.0f;
		else if (d < 2.0f) v = (-d * d * d + 6.0f * d * d - 12.0f * d + 8.0f) / 6.0f;
	}
	return v;
}

// the splat of one camera connection: one pixel (box) or the 5 x 5 pixels around the rounded position weighted by their distance to it
GI_HD void splatFiltered(float* accum, int width, int height, int filter, float posX, float posY, F3 v, PathCounters& pc)
{
	const int pcx = int(roundf(posX)), pcy = int(roundf(posY));
	if (filter == 0) { splatAtomic(accum, width, height, pcx, pcy, v, pc); return; }
	for (int y = -2; y <= 2; ++y)
		for (int x = -2; x <= 2; ++x)
		{
			const int px = pcx + x, py = pcy + y;
			const float dx = posX - float(px), dy = posY - float(py);
			const float w = reconstructionFilterValue(filter, sqrtf(dx * dx + dy * dy));
			if (w != 0.0f) splatAtomic(accum, width, height, px, py, w * v, pc);
		}
}

// ProjectToScreen, screenreproject.glsl:55-227, split in two like EstimateDirectLight: everything
// but the any-hit query towards the camera (prepareScreenSplat; the splatted value does not depend on the query), and the
// query + atomic splat (projectToScreen, the megakernel's form; the wavefront scheduler queues the query instead).
struct ScreenSplatQuery
{
	F3 origin, dir;   // any-hit ray towards the camera
	float len;
	int px, py;       // target pixel; round() may yield W or H -> bounds-checked by the consumer
	float posX, posY; // pixelPos before rounding (the reconstruction filters weight by the distance to it)
	F3 value;         // what WriteAtomic adds when the ray is free
};

// cmisQualityN .. lastDistMIS: the CHAINED_MIS arguments of screenreproject.glsl:64-69, read only when f.chainedMIS && computeMIS
GI_HD bool prepareScreenSplat(const DFrame& f, F3 pos, F3 inDir, F3 geometryNormal, F3 shadingNormal, const MaterialData& md, F3 pathThroughput, bool computeMIS,
                              ScreenSplatQuery& q, float cmisQualityN = 1.0f, float cmisQualityT = 1.0f, float lastConnectionProbability = 0.0f, float lastDistMIS = 0.0f)
{
	const F3 camPos = f3(f.cam.position), camU = f3(f.cam.u), camV = f3(f.cam.v), camW = f3(f.cam.w);
	F3 toCam = camPos - pos;
	float camViewDotRay = -dot(camW, toCam);
	if (!(camViewDotRay > 0.0f)) return false;
	float sx = -dot(camU, toCam) / camViewDotRay, sy = -dot(camV, toCam) / camViewDotRay;
	sx /= dot(camU, camU);
	sy /= dot(camV, camV);
	if (!(sx >= -1.0f && sx < 1.0f && sy > -1.0f && sy <= 1.0f)) return false;
	float cameraDistSq = dot(toCam, toCam);
	float cameraDist = sqrtf(cameraDistSq);
	toCam = toCam / cameraDist;
	q.origin = toCam * GI_RAY_HIT_EPSILON + pos;
	q.dir = toCam;
	q.len = cameraDist;
	q.posX = (sx * 0.5f + 0.5f) * float(f.width);
	q.posY = (sy * 0.5f + 0.5f) * float(f.height);
	q.px = int(roundf(q.posX));
	q.py = int(roundf(q.posY));
	float cosAtCamera = saturate(camViewDotRay / cameraDist);
	float fluxToIrradiance = adjointShadingNormalCorrectedOutDotN(inDir, toCam, geometryNormal, shadingNormal) /
	                         (cosAtCamera * cosAtCamera * cosAtCamera * f.cam.pixel_area * cameraDistSq);
	float pdf;
	F3 bsdf = adjointBSDF(inDir, toCam, md, shadingNormal, pdf);
	if (computeMIS)
	{
		float eyeToLight = 1.0f; // InitialMIS()
		float lightToEye = misHeuristic(pdf);
		float mis = eyeToLight / fmaxf(lightToEye + eyeToLight, GI_DIVISOR_EPSILON);
		if (f.chainedMIS)
		{
			// screenreproject.glsl:160-170
			float lastPTmis = lastConnectionProbability / fmaxf(lastConnectionProbability + lightToEye * lastDistMIS, GI_DIVISOR_EPSILON);
			updateQuality(cmisQualityN, lastPTmis);
			updateQuality(cmisQualityT, 1.0f - lastPTmis);
			mis = qualityWeightedMIS_NEE(mis, cmisQualityN, cmisQualityT);
		}
		bsdf = bsdf * mis;
	}
	q.value = pathThroughput * fluxToIrradiance * bsdf;
	return true;
}

GI_HD void projectToScreen(const DFrame& f, F3 pos, F3 inDir, F3 geometryNormal, F3 shadingNormal, const MaterialData& md, F3 pathThroughput, bool computeMIS,
                           PathCounters& pc, TraceStats* st, float cmisQualityN = 1.0f, float cmisQualityT = 1.0f, float lastConnectionProbability = 0.0f,
                           float lastDistMIS = 0.0f)
{
	ScreenSplatQuery q;
	if (!prepareScreenSplat(f, pos, inDir, geometryNormal, shadingNormal, md, pathThroughput, computeMIS, q, cmisQualityN, cmisQualityT, lastConnectionProbability, lastDistMIS))
		return;
	++pc.anyRays;
	if (traceAny(f.scene, f.traversalMode, q.origin, q.dir, q.len, st)) return;
	splatFiltered(f.accum, f.width, f.height, f.reconFilter, q.posX, q.posY, q.value, pc);
}

// CHAINED_MIS: the two quality weights of a stored light vertex (lighttracer.comp:67-68,174-177)
GI_HD void setLightCacheQuality(const DFrame& f, uint32_t idx, float weightN, float weightT)
{
	if (idx >= f.lightCacheCapacity) return;
	F4* e = f.lightCache + 4 * size_t(idx) + 3;
	e->z = weightN; e->w = weightT;
}

// saveLightCache, lighttracer.comp:23-45.  Entry = 4 x F4 (64 bytes):
//  (Flux.xyz, MaterialIndex) (Position.xyz, Normal0) (Texcoord.xy, Normal1, PathLength) (IncidentDirectionPacked.xy, 0, 0)
GI_HD void saveLightCache(const DFrame& f, uint32_t idx, F3 flux, int materialIndex, F3 position, F2 packedNormal, F2 tc, F3 incidentDir, int pathLength)
{
	if (idx >= f.lightCacheCapacity) return;
	F4* e = f.lightCache + 4 * size_t(idx);
	F2 pi = packNormal(incidentDir);
	F4 a, b, c, d;
	a.x = flux.x; a.y = flux.y; a.z = flux.z; a.w = i2f(materialIndex);
	b.x = position.x; b.y = position.y; b.z = position.z; b.w = packedNormal.x;
	c.x = tc.x; c.y = tc.y; c.z = packedNormal.y; c.w = i2f(pathLength);
	d.x = pi.x; d.y = pi.y; d.z = 0.0f; d.w = 0.0f;
	e[0] = a; e[1] = b; e[2] = c; e[3] = d;
}

// Slot allocation for a stored light vertex: atomicAdd(NumLightCacheEntries, 1) (lighttracer.comp:120),
// warp-aggregated on the device (one atomic per warp instead of one per lane).
GI_HD uint32_t allocLightCacheSlot(const DFrame& f, bool warmup)
{
#if defined(__CUDA_ARCH__)
	// lanes of one warp may belong to different iterations of a pipelined batch (different caches): aggregate per counter
	const unsigned mask = __match_any_sync(__activemask(), reinterpret_cast<unsigned long long>(f.lightCacheCount));
	const int leader = __ffs(mask) - 1;
	const int lane = threadIdx.x & 31;
	const unsigned count = __popc(mask);
	uint32_t base = 0;
	if (lane == leader)
	{
		base = atomicAdd(f.lightCacheCount, count);
		if (warmup) atomicAdd(f.lightPathLengthSum, count);
	}
	base = __shfl_sync(mask, base, leader);
	return base + __popc(mask & ((1u << lane) - 1u));
#else
	uint32_t r = (*f.lightCacheCount)++;
	if (warmup) ++(*f.lightPathLengthSum);
	return r;
#endif
}

// lighttracer.comp:49-193, one invocation.  SAVE: SAVE_LIGHT_CACHE; WARMUP: SAVE_LIGHT_CACHE_WARMUP.
template <bool SAVE, bool WARMUP>
GI_HD void lighttracePath(const DFrame& f, uint32_t gid, PathCounters& pc, TraceStats* st)
{
	uint32_t seed = initRandomSeed(f.frameSeed, gid);
	int lightSampleIndex = int(gid / uint32_t(f.numRaysPerLightSample));
	F4 l0 = ld4(f.lightSamples + 2 * lightSampleIndex), l1 = ld4(f.lightSamples + 2 * lightSampleIndex + 1);
	F3 intensity = f3(l1.x, l1.y, l1.z);
	++pc.lightPaths;
	if (SAVE && !WARMUP)
	{
		F2 pn; pn.x = l0.w; pn.y = l1.w;
		F2 z; z.x = 0.0f; z.y = 0.0f;
		saveLightCache(f, gid, intensity / GI_PI_2 / float(f.numRaysPerLightSample), -1, f3(l0.x, l0.y, l0.z), pn, z, f3(0.0f), 0);
	}
	// CHAINED_MIS state of the light path (lighttracer.comp:62-69)
	const bool cmis = SAVE && !WARMUP && f.chainedMIS != 0;
	float lastConnectionProbability = 0.0f, lastCosAbs = 1.0f, chainedMISWeightN = 1.0f, chainedMISWeightT = 1.0f;
	if (cmis) setLightCacheQuality(f, gid, chainedMISWeightN, chainedMISWeightT);
	F3 pathThroughput = intensity * (GI_PI / float(f.numRaysPerLightSample));
	F3 origin = f3(l0.x, l0.y, l0.z), dir;
	if (l1.w > 1.5f)
	{
		dir = sampleDirection(random2(seed));
		pathThroughput = pathThroughput * 4.0f;
	}
	else
	{
		F3 normal = unpackNormal(l0.w, l1.w);
		F3 U, V;
		createONB(normal, U, V);
		dir = sampleHemisphereCosine(random2(seed), U, V, normal);
		if (cmis)
		{
			lastCosAbs = saturate(dot(normal, dir)); // lighttracer.comp:89-92
			lastConnectionProbability = misHeuristic(lastCosAbs / GI_PI);
		}
	}
	for (int i = 0; i < f.maxPathLength; ++i)
	{
		Hit h; bool fb = false;
		++pc.closestRays;
		traceClosest(f.scene, f.traversalMode, origin, dir, GI_RAY_MAX, h, st, &fb);
		if (fb) ++pc.fallbackRays;
		if (h.t == GI_RAY_MAX) break;
		origin = origin + h.t * dir;
		F3 shadingNormal; F2 tc;
		getTriangleHitInfo(f.scene, h.triangle, h.bary, shadingNormal, tc);
		MaterialData md = sampleMaterialData(f.scene, h.triangle.w, tc);
		++pc.shadedVertices;
		uint32_t slot = 0xffffffffu;
		if (SAVE)
		{
			if (md.Reflectiveness.w < 100.0f)
			{
				slot = allocLightCacheSlot(f, WARMUP);
				if (!WARMUP) saveLightCache(f, slot, pathThroughput, h.triangle.w, origin, packNormal(shadingNormal), tc, dir, i + 1);
			}
		}
		if (!WARMUP)
		{
			if (cmis) // lighttracer.comp:141-144
				projectToScreen(f, origin, dir, h.normal, shadingNormal, md, pathThroughput, true, pc, st, chainedMISWeightN, chainedMISWeightT, lastConnectionProbability,
				                distanceMISHeuristic(h.t));
			else
				projectToScreen(f, origin, dir, h.normal, shadingNormal, md, pathThroughput, SAVE, pc, st);
		}
		F3 throughput = f3(1.0f);
		float samplePDF;
		F3 newDir = sampleBSDF(dir, md, seed, shadingNormal, throughput, true, samplePDF);
		throughput = throughput * adjointShadingNormalCorrection(dir, newDir, h.normal, shadingNormal);
		pathThroughput = pathThroughput * throughput;
		if (cmis)
		{
			// lighttracer.comp:165-181
			float reverseProb = misHeuristic(samplePDF);
			float distMIS = distanceMISHeuristic(h.t);
			float mis = lastConnectionProbability / fmaxf(lastConnectionProbability + distMIS * reverseProb, GI_DIVISOR_EPSILON);
			updateQuality(chainedMISWeightN, mis);
			updateQuality(chainedMISWeightT, 1.0f - mis);
			if (slot != 0xffffffffu) setLightCacheQuality(f, slot, chainedMISWeightN, chainedMISWeightT);
			lastConnectionProbability = reverseProb;
			lastCosAbs = fabsf(dot(dir, h.normal)); // (the old direction and the un-normalised geometry normal, as written)
		}
		float cont = saturate(luminance(throughput));
		if (randomFloat(seed) >= cont) break;
		pathThroughput = pathThroughput / cont;
		dir = newDir;
		origin = origin + dir * GI_RAY_HIT_EPSILON;
	}
}

// showlightcaches.comp:10-64 with exactly NumInitialLightSamples invocations and an atomic splat
GI_HD void showLightCache(const DFrame& f, int index, PathCounters& pc, TraceStats* st)
{
	F4 l0 = ld4(f.lightSamples + 2 * index), l1 = ld4(f.lightSamples + 2 * index + 1);
	const F3 camPos = f3(f.cam.position), camU = f3(f.cam.u), camV = f3(f.cam.v), camW = f3(f.cam.w);
	F3 pos = f3(l0.x, l0.y, l0.z);
	F3 toCam = camPos - pos;
	float camViewDotRay = -dot(camW, toCam);
	if (!(camViewDotRay > 0.0f)) return;
	float sx = -dot(camU, toCam) / camViewDotRay, sy = -dot(camV, toCam) / camViewDotRay;
	sx /= dot(camU, camU);
	sy /= dot(camV, camV);
	if (!(sx >= -1.0f && sx < 1.0f && sy > -1.0f && sy <= 1.0f)) return;
	float cameraDistSq = dot(toCam, toCam);
	float cameraDist = sqrtf(cameraDistSq);
	toCam = toCam / cameraDist;
	F3 origin = toCam * GI_RAY_HIT_EPSILON + pos;
	++pc.anyRays;
	if (traceAny(f.scene, f.traversalMode, origin, toCam, cameraDist, st)) return;
	F3 normal = unpackNormal(l0.w, l1.w);
	float cosTheta = dot(normal, toCam);
	float cosAtCamera = saturate(camViewDotRay / cameraDist);
	float intensityToRadiance = cosTheta / (GI_PI_2 * cosAtCamera * cosAtCamera * cosAtCamera * f.cam.pixel_area * cameraDistSq);
	int pcx = int((sx * 0.5f + 0.5f) * float(f.width));
	int pcy = int((sy * 0.5f + 0.5f) * float(f.height));
	splatAtomic(f.accum, f.width, f.height, pcx, pcy, f3(l1.x, l1.y, l1.z) * intensityToRadiance, pc);
}

// The connection of an eye-path vertex with one random light-cache entry (pathtracer_bidir.comp:97-205), without the any-hit
// query: returns false when no query is made (combined path too long), else the shadow ray and the radiance the
// connection adds when the ray is free.  Draws one random number in either case.
struct ConnectionQuery
{
	F3 origin, dir;
	float len;
	F3 value;
};

GI_HD bool prepareConnection(const DFrame& f, uint32_t actualNumEntries, float lightSampleWeight, F3 origin, F3 dir, F3 hitNormal, const MaterialData& md, int i,
                             float rayHit, float lastConnectionProbability, F3 pathThroughput, uint32_t& seed, ConnectionQuery& q, const float* gMis = nullptr)
{
	uint32_t idx = randomUInt(seed) % actualNumEntries;
	const F4* e = f.lightCache + 4 * size_t(idx);
	F4 e2 = ld4(e + 2);
	int pathLength = f2i(e2.w) + i + 1;
	if (pathLength > f.maxPathLength) return false;
	F4 e1 = ld4(e + 1);
	F3 ldir = f3(e1.x, e1.y, e1.z) - origin;
	float lightDistSq = dot(ldir, ldir);
	float lightDist = sqrtf(lightDistSq);
	ldir = ldir / fmaxf(lightDist, GI_DIVISOR_EPSILON);
	q.origin = GI_RAY_HIT_EPSILON * ldir + origin;
	q.dir = ldir;
	q.len = lightDist - GI_RAY_HIT_EPSILON * 2;
	F4 e0 = ld4(e), e3 = ld4(e + 3);
	int lightMaterialIndex = f2i(e0.w);
	F3 lightFlux = f3(e0.x, e0.y, e0.z);
	F3 lightNormal = f3(0.0f, 1.0f, 0.0f);
	float lightPathToEye;
	float cosLight = 0.0f;
	F3 lightIntensity;
	float cosCamera = fabsf(dot(ldir, hitNormal));
	const float normal0 = e1.w, normal1 = e2.z;
	if (normal1 < 1.5f)
	{
		lightNormal = unpackNormalNoSat(normal0, normal1);
		cosLight = fabsf(dot(ldir, lightNormal));
	}
	float pdf;
	F3 bsdf = BSDF(dir, ldir, md, hitNormal, pdf);
	float eyeToLight = misHeuristic(pdf);
	float mis;
	if (i > 0)
	{
		eyeToLight *= distanceMISHeuristic(rayHit);
		mis = lastConnectionProbability / fmaxf(lastConnectionProbability + eyeToLight, GI_DIVISOR_EPSILON);
		if (md.Reflectiveness.w >= 100.0f) mis = 1.0f;
	}
	else
		mis = eyeToLight / fmaxf(lastConnectionProbability + eyeToLight, GI_DIVISOR_EPSILON);
	float misLight = 1.0f;
	if (normal1 < 1.5f)
	{
		if (lightMaterialIndex != -1)
		{
			F2 ltc; ltc.x = e2.x; ltc.y = e2.y;
			MaterialData lmd = sampleMaterialData(f.scene, lightMaterialIndex, ltc);
			F3 incidentDir = unpackNormal(e3.x, e3.y);
			F3 lightBSDF = adjointBSDF(ldir, -incidentDir, lmd, lightNormal, pdf);
			lightIntensity = lightFlux * lightBSDF * cosLight;
			lightPathToEye = misHeuristic(pdf) * distanceMISHeuristic(lightDist);
		}
		else
		{
			lightIntensity = lightFlux * (cosLight * GI_PI_2);
			lightPathToEye = misHeuristic(cosLight / GI_PI) * distanceMISHeuristic(lightDist);
		}
		misLight = lightPathToEye / fmaxf(lightPathToEye + eyeToLight, GI_DIVISOR_EPSILON);
	}
	else
		lightIntensity = lightFlux * GI_PI_2;
	if (gMis)
	{
		// CHAINED_MIS, pathtracer_bidir.comp:187-199 (g_mis[i] = mis is visible inside this block only: the end of the loop body overwrites it)
		float quality_nee = e3.z, quality_trace = e3.w;
		misLight = qualityWeightedMIS_NEE(misLight, quality_nee, quality_trace);
		updateQuality(quality_nee, 1.0f - misLight);
		updateQuality(quality_trace, misLight);
		mis = cmisPathThroughputWeight(gMis, mis, i + 1, quality_nee, quality_trace);
	}
	mis *= misLight;
	F3 irradiance = (mis * lightSampleWeight * cosCamera / lightDistSq) * lightIntensity;
	q.value = irradiance * pathThroughput * bsdf;
	return true;
}

// bidirpathtracer/pathtracer_bidir.comp:16-251, one invocation
GI_HD void bidirPathtracePixel(const DFrame& f, int gx, int gy, PathCounters& pc, TraceStats* st)
{
	uint32_t seed = initRandomSeed(f.frameSeed, uint32_t(gx + gy * f.width));
	F3 origin = f3(f.cam.position);
	F3 dir = cameraRayDirection(f, seed, gx, gy);
	++pc.cameraPaths;
	F3 color = f3(0.0f), pathThroughput = f3(1.0f);
	uint32_t numEntries = *f.lightCacheCount;
	uint32_t actualNumEntries = numEntries < f.lightCacheCapacity ? numEntries : f.lightCacheCapacity;
	float lightSampleWeight = float(actualNumEntries);
	float lastConnectionProbability = 1.0f; // InitialMIS()
	float mis = 0.0f;
	const bool cmis = f.chainedMIS != 0;
	float gMis[GI_MAX_PATHLENGTH];
	if (cmis) for (int i = 0; i < GI_MAX_PATHLENGTH; ++i) gMis[i] = 0.0f; // pathtracer_bidir.comp:49-52
	for (int i = 0; i < f.maxPathLength; ++i)
	{
		Hit h; bool fb = false;
		++pc.closestRays;
		traceClosest(f.scene, f.traversalMode, origin, dir, GI_RAY_MAX, h, st, &fb);
		if (fb) ++pc.fallbackRays;
		if (h.t == GI_RAY_MAX) break;
		const float rayHit = h.t;
		origin = origin + rayHit * dir;
		F3 hitNormal; F2 tc;
		getTriangleHitInfo(f.scene, h.triangle, h.bary, hitNormal, tc);
		MaterialData md = sampleMaterialData(f.scene, h.triangle.w, tc);
		++pc.shadedVertices;
		{
			float incidentCosSat = saturate(-dot(hitNormal, dir));
			float lightToEye = distanceMISHeuristic(rayHit) * misHeuristic(incidentCosSat / GI_PI);
			mis = (i == 0) ? 1.0f : (lastConnectionProbability / fmaxf(lastConnectionProbability + lightToEye, GI_DIVISOR_EPSILON));
			if (cmis)
			{
				gMis[i] = mis; // pathtracer_bidir.comp:85-89
				mis = cmisPathThroughputWeight(gMis, mis, i + 1, 1.0f, 1.0f);
			}
			color = color + mis * pathThroughput * md.Emissivity;
		}
		{
			ConnectionQuery cq;
			if (prepareConnection(f, actualNumEntries, lightSampleWeight, origin, dir, hitNormal, md, i, rayHit, lastConnectionProbability, pathThroughput, seed, cq,
			                      cmis ? gMis : nullptr))
			{
				++pc.anyRays;
				if (!traceAny(f.scene, f.traversalMode, cq.origin, cq.dir, cq.len, st)) color = color + cq.value;
			}
		}
		float samplePDF;
		F3 throughput = f3(1.0f);
		dir = sampleBSDF(dir, md, seed, hitNormal, throughput, false, samplePDF);
		origin = origin + dir * GI_RAY_HIT_EPSILON;
		pathThroughput = pathThroughput * throughput;
		float cont = saturate(luminance(throughput));
		if (randomFloat(seed) >= cont) break;
		pathThroughput = pathThroughput / cont;
		float reverseProb = misHeuristic(samplePDF);
		if (i == 0)
			mis = reverseProb / fmaxf(reverseProb + lastConnectionProbability, GI_DIVISOR_EPSILON);
		else
		{
			float distMIS = distanceMISHeuristic(rayHit);
			reverseProb *= distMIS;
			mis = lastConnectionProbability / fmaxf(lastConnectionProbability + distMIS * reverseProb, GI_DIVISOR_EPSILON);
			if (md.Reflectiveness.w >= 100.0f) mis = 1.0f;
		}
		if (cmis) gMis[i] = mis; // pathtracer_bidir.comp:236-242: the weight changes with the future of the path
		else pathThroughput = pathThroughput * mis;
		lastConnectionProbability = reverseProb;
	}
	float* px = f.accum + 4 * (size_t(gy) * f.width + gx);
#if defined(__CUDA_ARCH__)
	float4 v = *reinterpret_cast<float4*>(px);
	v.x += color.x; v.y += color.y; v.z += color.z; v.w += 1.0f;
	*reinterpret_cast<float4*>(px) = v;
#else
	px[0] += color.x; px[1] += color.y; px[2] += color.z; px[3] += 1.0f;
#endif
}

} // namespace gik
