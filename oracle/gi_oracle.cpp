// gi_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  NOT part of the product.
//
// CPU restatement of the light-transport hot path of Jojendersie/gpugi (GLSL compute
// shaders + the few host routines that feed them).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library; the product
// (gpugi_b200/) never links, imports or calls it.
//
// PARITY PINNED.  The reference ships no golden vectors, reference images or tests for this path and cannot
// be built or run as a whole here (GLSL needing an NVIDIA GL 4.5 driver; MSVC-only host).  The restatement is pinned
// (a) against the reference's OWN SHADER SOURCE run on the CPU -- oracle/ref_glsl/ compiles gpugi/shader/*.glsl / *.comp,
// read where they lie, as C++ (oracle/_ref/libref_glsl_*.so); every function below that restates a shader equals it bit for
// bit on ray batches and whole iterations of every renderer and variant (tests/test_ref_glsl.py, tests/golden/ref_glsl.npz);
// (b) against the reference's C / C++ files compiled where they lie (random, hdrimage, bim + epsilon, bvhmake builders,
// stb_image: tests/test_ref_pinned.py, test_bvhmake.py, test_imageio.py), and (c) the ray / box and ray / triangle KATs of
// dependencies/epsilon/unittest/intersection3d.cpp:236-298 (tests/test_oracle_kat.py).  Each function cites the reference
// file:line it follows.
//
// Arithmetic: every GLSL float op is one IEEE-754 binary32 operation, no contraction
// (compile with -ffp-contract=off), min/max with fminf/fmaxf NaN semantics, normalize(v) =
// v * (1/sqrt(dot(v,v))).  GLSL leaves these details to the driver; this is the definition
// the CUDA path is held to.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC).

#include "../include/gpugi_b200.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>

#ifdef _OPENMP
#include <omp.h>
#endif

#define OR_API extern "C" __attribute__((visibility("default")))

namespace {

// ---------------------------------------------------------------- constants
// stdheader.glsl:5-6, intersectiontests.glsl:1, materials.glsl:1, helper.glsl:1-2
const float RAY_HIT_EPSILON = 0.0001f;
const float RAY_MAX = 3.402823466e+38f;
const float INTERSECT_EPSILON = 0.0001f;
const float DIVISOR_EPSILON = 1e-5f;
const float PI = 3.14159265358979f;
const float PI_2 = 6.28318530717958f;
const int TRIANGLES_PER_LEAF = 2; // scenedata.glsl:5

// ---------------------------------------------------------------- vec helpers
struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 v3(float s) { V3 r = {s, s, s}; return r; }
inline V3 v3(const float* p) { V3 r = {p[0], p[1], p[2]}; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator/(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
inline V3& operator*=(V3& a, V3 b) { a = a * b; return a; }
inline V3& operator*=(V3& a, float s) { a = a * s; return a; }
inline V3& operator/=(V3& a, float s) { a = a / s; return a; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// GLSL spec 8.5: cross(x,y) = (x[1]*y[2]-y[1]*x[2], x[2]*y[0]-y[2]*x[0], x[0]*y[1]-y[0]*x[1])
inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline V3 normalize(V3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline V3 einormalize(V3 a) { return a / length(a); } // ei::normalize (epsilon vector.hpp:1096-1099), host code only
inline float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); } // helper.glsl:4-7
inline V3 saturate(V3 a) { return v3(saturate(a.x), saturate(a.y), saturate(a.z)); }
inline float signf_(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// ---------------------------------------------------------------- RNG (random.glsl:1-43)
inline uint32_t wanghash(uint32_t x)
{
	x = (x ^ 61u) ^ (x >> 16);
	x *= 9u;
	x = x ^ (x >> 4);
	x *= 0x27d4eb2du;
	x = x ^ (x >> 15);
	return x;
}
inline uint32_t RandomUInt(uint32_t& seed)
{
	seed ^= (seed << 13);
	seed ^= (seed >> 17);
	seed ^= (seed << 5);
	return seed;
}
inline uint32_t InitRandomSeed(uint32_t frameSeed, uint32_t invocationID) { return wanghash(frameSeed + invocationID); }
// random.glsl:38-43: the literal 4294967297.0 is a float literal == 2^32; can return 1.0.
inline float Random(uint32_t& seed) { return float(RandomUInt(seed)) / 4294967296.0f; }
inline V2 Random2(uint32_t& seed) { V2 r; r.x = Random(seed); r.y = Random(seed); return r; }

// random.glsl:51-87
inline V3 SampleDirection(V2 rs)
{
	float cosTheta = rs.x * 2.0f - 1.0f;
	float sinTheta = sqrtf((1.0f - cosTheta) * (1.0f + cosTheta));
	float phi = rs.y * PI_2;
	return normalize(v3(sinTheta * sinf(phi), sinTheta * cosf(phi), cosTheta));
}
inline V3 SampleHemisphereCosine(V2 rs, V3 U, V3 V, V3 W)
{
	float phi = PI_2 * rs.x;
	float sinTheta = sqrtf(rs.y);
	float x = sinTheta * cosf(phi);
	float y = sinTheta * sinf(phi);
	float z = sqrtf(1.0f - rs.y);
	return x * U + y * V + z * W;
}
inline V3 SamplePhongLobe(V2 rs, float exponent, V3 U, V3 V, V3 W)
{
	float phi = PI_2 * rs.x;
	float power = powf(rs.y, 1.0f / (exponent + 1.0f));
	float scale = sqrtf(1.0f - power * power);
	float x = cosf(phi) * scale;
	float y = sinf(phi) * scale;
	float z = power;
	return x * U + y * V + z * W;
}

// helper.glsl:22-35
inline void CreateONB(V3 n, V3& U, V3& V)
{
	U = cross(n, v3(0.0f, 1.0f, 0.0f));
	if (fabsf(U.x) < 0.0001f && fabsf(U.y) < 0.0001f && fabsf(U.z) < 0.0001f)
		U = cross(n, v3(1.0f, 0.0f, 0.0f));
	U = normalize(U);
	V = cross(n, U);
}
inline float GetLuminance(V3 rgb) { return dot(rgb, v3(0.212671f, 0.715160f, 0.072169f)); }

// helper.glsl:38-63
inline V3 UnpackNormal(float p0, float p1)
{ // vec2 overload, helper.glsl:38-42
	float sinPhi = sqrtf(saturate((1.0f - p1) * (1.0f + p1)));
	return v3(cosf(p0) * sinPhi, sinf(p0) * sinPhi, p1);
}
inline V3 UnpackNormalNoSat(float p0, float p1)
{ // (float,float) overload, helper.glsl:43-47
	float sinPhi = sqrtf(1.0f - p1 * p1);
	return v3(cosf(p0) * sinPhi, sinf(p0) * sinPhi, p1);
}
inline V2 PackNormal(V3 n)
{
	V2 r;
	r.x = (n.x == 0.0f) ? signf_(n.y) * PI / 2 : atan2f(n.y, n.x);
	r.y = n.z;
	return r;
}

// ---------------------------------------------------------------- scene view
struct Texture
{
	int w, h, format;
	const void* px;
};

struct Scene
{
	const GiNode* nodes;
	const GiOBoxNode* oboxNodes; // bvhType = 1 (Scene::TreeNode<ei::OBox>), else null
	uint32_t numNodes;
	const GiTriangle* leafTris;
	uint32_t numLeafTris;
	const float* positions;
	const GiVertexInfo* vinfo;
	uint32_t numVertices;
	const GiMaterial* materials;
	uint32_t numMaterials;
	std::vector<Texture> textures;
};

float g_srgbLut[256];
bool g_srgbInit = false;
void initSrgb()
{
	if (g_srgbInit) return;
	for (int i = 0; i < 256; ++i)
	{
		float c = float(i) / 255.0f;
		// GL 4.5 spec 8.24 sRGB decode
		g_srgbLut[i] = (c <= 0.04045f) ? c / 12.92f : powf((c + 0.055f) / 1.055f, 2.4f);
	}
	g_srgbInit = true;
}

Scene makeScene(const GiSceneDesc* d)
{
	initSrgb();
	Scene s;
	s.nodes = d->nodes; s.numNodes = d->num_nodes;
	s.oboxNodes = d->obox_nodes;
	s.leafTris = d->leaf_triangles; s.numLeafTris = d->num_leaf_triangles;
	s.positions = d->positions; s.vinfo = d->vertex_info; s.numVertices = d->num_vertices;
	s.materials = d->materials; s.numMaterials = d->num_materials;
	for (uint32_t i = 0; i < d->num_textures; ++i)
	{
		Texture t = {d->textures[i].width, d->textures[i].height, d->textures[i].format, d->textures[i].pixels};
		s.textures.push_back(t);
	}
	return s;
}

inline V4 texel(const Texture& t, int x, int y)
{
	size_t i = size_t(y) * t.w + x;
	V4 r;
	if (t.format == GI_TEX_RGBA32F)
	{
		const float* p = static_cast<const float*>(t.px) + 4 * i;
		r.x = p[0]; r.y = p[1]; r.z = p[2]; r.w = p[3];
	}
	else
	{
		const uint8_t* p = static_cast<const uint8_t*>(t.px) + 4 * i;
		if (t.format == GI_TEX_SRGB8_ALPHA8) { r.x = g_srgbLut[p[0]]; r.y = g_srgbLut[p[1]]; r.z = g_srgbLut[p[2]]; }
		else { r.x = float(p[0]) / 255.0f; r.y = float(p[1]) / 255.0f; r.z = float(p[2]) / 255.0f; }
		r.w = float(p[3]) / 255.0f;
	}
	return r;
}
inline int wrapi(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }

// textureLod(sampler2D, uv, 0) with the sampler of scene.cpp:16-21 (mag LINEAR, REPEAT):
// GL 4.5 spec 8.14.2 linear filtering, texel centres at (i+0.5)/size.
V4 sampleTexture(const Texture& t, float u, float v)
{
	if (t.w == 1 && t.h == 1) return texel(t, 0, 0);
	float fu = u * float(t.w) - 0.5f;
	float fv = v * float(t.h) - 0.5f;
	float iu = floorf(fu), iv = floorf(fv);
	float a = fu - iu, b = fv - iv;
	int i0 = wrapi(int(iu), t.w), i1 = wrapi(int(iu) + 1, t.w);
	int j0 = wrapi(int(iv), t.h), j1 = wrapi(int(iv) + 1, t.h);
	V4 t00 = texel(t, i0, j0), t10 = texel(t, i1, j0), t01 = texel(t, i0, j1), t11 = texel(t, i1, j1);
	float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	V4 r;
	r.x = t00.x * w00 + t10.x * w10 + t01.x * w01 + t11.x * w11;
	r.y = t00.y * w00 + t10.y * w10 + t01.y * w01 + t11.y * w11;
	r.z = t00.z * w00 + t10.z * w10 + t01.z * w01 + t11.z * w11;
	r.w = t00.w * w00 + t10.w * w10 + t01.w * w01 + t11.w * w11;
	return r;
}

// ---------------------------------------------------------------- intersection tests
struct Ray { V3 Origin; V3 Direction; };

// intersectiontests.glsl:9-20
inline bool IntersectBox(V3 rayOrigin, V3 invRayDir, V3 aabbMin, V3 aabbMax, float& firstHit, float& lastHit)
{
	V3 tbot = invRayDir * (aabbMin - rayOrigin);
	V3 ttop = invRayDir * (aabbMax - rayOrigin);
	V3 tmin = v3(fminf(ttop.x, tbot.x), fminf(ttop.y, tbot.y), fminf(ttop.z, tbot.z));
	V3 tmax = v3(fmaxf(ttop.x, tbot.x), fmaxf(ttop.y, tbot.y), fmaxf(ttop.z, tbot.z));
	float tx = fmaxf(tmin.x, tmin.y), ty = fmaxf(tmin.x, tmin.z);
	firstHit = fmaxf(0.0f, fmaxf(tx, ty));
	tx = fminf(tmax.x, tmax.y); ty = fminf(tmax.x, tmax.z);
	lastHit = fminf(tx, ty);
	return firstHit <= lastHit;
}

// intersectiontests.glsl:63-79
inline bool IntersectTriangle(const Ray& ray, V3 p0, V3 p1, V3 p2, float& hit, V3& bary, V3& triangleNormal)
{
	const V3 e0 = p1 - p0;
	const V3 e1 = p0 - p2;
	triangleNormal = cross(e1, e0);
	const V3 e2 = (1.0f / dot(triangleNormal, ray.Direction)) * (p0 - ray.Origin);
	const V3 i = cross(ray.Direction, e2);
	bary.y = dot(i, e1);
	bary.z = dot(i, e0);
	bary.x = 1.0f - (bary.z + bary.y);
	hit = dot(triangleNormal, e2);
	return (hit > INTERSECT_EPSILON) && (bary.x >= 0.0f && bary.y >= 0.0f && bary.z >= 0.0f);
}

struct TraceCounters { uint64_t boxes, tris; };

// intersectiontests.glsl:82-88
inline V3 rotate(V3 x, const float q[4])
{
	V3 t = cross(v3(q[0], q[1], q[2]), x);
	return x + 2.0f * v3(q[3] * t.x + q[1] * t.z - q[2] * t.y,
	                     q[3] * t.y + q[2] * t.x - q[0] * t.z,
	                     q[3] * t.z + q[0] * t.y - q[1] * t.x);
}

// FetchIntersectBoxNode (intersectiontests.glsl:100-111) / FetchIntersectOBoxNode (:114-128), selected by the scene's bvhType
// (the AABOX_BVH / OBOX_BVH define of traceray.glsl:62-66)
inline bool FetchIntersectNode(const Scene& s, const Ray& ray, V3 invRayDir, int nodeIdx, float& firstHit, float& lastHit, uint32_t& childCode, int& escape)
{
	if (s.oboxNodes)
	{
		const GiOBoxNode& n = s.oboxNodes[nodeIdx];
		childCode = n.first_child;
		escape = int(n.escape);
		V3 bbSidesHalf = v3(n.sides_half);
		V3 rayOrigin = rotate(ray.Origin - v3(n.center), n.rotation_inv);
		V3 rayDir = rotate(ray.Direction, n.rotation_inv);
		return IntersectBox(rayOrigin, v3(1.0f) / rayDir, -bbSidesHalf, bbSidesHalf, firstHit, lastHit);
	}
	const GiNode& n = s.nodes[nodeIdx];
	childCode = n.first_child;
	escape = int(n.escape);
	return IntersectBox(ray.Origin, invRayDir, v3(n.min), v3(n.max), firstHit, lastHit);
}

inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// traceray.glsl:25-180 without ANY_HIT; HIT_INDEX_OUTPUT + TRINORMAL_OUTPUT always on.
void TraceRay(const Scene& s, const Ray& ray, float& rayLength, int hitIndex[2], V3& outBary, GiTriangle& outTriangle,
              V3& triangleNormal, TraceCounters* cnt)
{
	int currentNodeIndex = 0;
	int currentLeafIndex = 0;
	hitIndex[0] = -1; hitIndex[1] = -1; // 0xFFFFFFFF
	int lastNodeIndex = 0;
	V3 invRayDir = v3(1.0f) / ray.Direction;
	bool nextIsLeafNode = false;
	do
	{
		if (!nextIsLeafNode)
		{
			if (cnt) ++cnt->boxes;
			float newHit, exitDist;
			uint32_t childCode;
			int escape;
			if (FetchIntersectNode(s, ray, invRayDir, currentNodeIndex, newHit, exitDist, childCode, escape) && newHit <= rayLength)
			{
				lastNodeIndex = currentNodeIndex;
				currentNodeIndex = int(childCode & 0x7FFFFFFFu);
				nextIsLeafNode = uint32_t(currentNodeIndex) != childCode;
				if (nextIsLeafNode)
				{
					currentLeafIndex = TRIANGLES_PER_LEAF * currentNodeIndex;
					currentNodeIndex = escape;
				}
			}
			else
				currentNodeIndex = escape;
		}
		if (nextIsLeafNode)
		{
			GiTriangle triangle = s.leafTris[currentLeafIndex];
			if (triangle.v[0] == triangle.v[1])
				nextIsLeafNode = false;
			else
			{
				if (cnt) ++cnt->tris;
				V3 p0 = v3(s.positions + 3 * size_t(triangle.v[0]));
				V3 p1 = v3(s.positions + 3 * size_t(triangle.v[1]));
				V3 p2 = v3(s.positions + 3 * size_t(triangle.v[2]));
				V3 newNormal, newBary; float newHit;
				if (IntersectTriangle(ray, p0, p1, p2, newHit, newBary, newNormal) && newHit < rayLength)
				{
					rayLength = newHit;
					outTriangle = triangle;
					outBary = newBary;
					hitIndex[0] = lastNodeIndex;
					hitIndex[1] = currentLeafIndex;
					triangleNormal = newNormal;
				}
				++currentLeafIndex;
				nextIsLeafNode = (currentLeafIndex % TRIANGLES_PER_LEAF) != 0;
			}
		}
	} while (currentNodeIndex != 0 || nextIsLeafNode);
}

// traceray.glsl:19-23 + body under ANY_HIT
bool TraceRayAnyHit(const Scene& s, const Ray& ray, float rayLength, TraceCounters* cnt)
{
	int currentNodeIndex = 0;
	int currentLeafIndex = 0;
	V3 invRayDir = v3(1.0f) / ray.Direction;
	bool nextIsLeafNode = false;
	do
	{
		if (!nextIsLeafNode)
		{
			if (cnt) ++cnt->boxes;
			float newHit, exitDist;
			uint32_t childCode;
			int escape;
			if (FetchIntersectNode(s, ray, invRayDir, currentNodeIndex, newHit, exitDist, childCode, escape) && newHit <= rayLength)
			{
				currentNodeIndex = int(childCode & 0x7FFFFFFFu);
				nextIsLeafNode = uint32_t(currentNodeIndex) != childCode;
				if (nextIsLeafNode)
				{
					currentLeafIndex = TRIANGLES_PER_LEAF * currentNodeIndex;
					currentNodeIndex = escape;
				}
			}
			else
				currentNodeIndex = escape;
		}
		if (nextIsLeafNode)
		{
			GiTriangle triangle = s.leafTris[currentLeafIndex];
			if (triangle.v[0] == triangle.v[1])
				nextIsLeafNode = false;
			else
			{
				if (cnt) ++cnt->tris;
				V3 p0 = v3(s.positions + 3 * size_t(triangle.v[0]));
				V3 p1 = v3(s.positions + 3 * size_t(triangle.v[1]));
				V3 p2 = v3(s.positions + 3 * size_t(triangle.v[2]));
				V3 newNormal, newBary; float newHit;
				if (IntersectTriangle(ray, p0, p1, p2, newHit, newBary, newNormal) && newHit < rayLength)
					return true;
				++currentLeafIndex;
				nextIsLeafNode = (currentLeafIndex % TRIANGLES_PER_LEAF) != 0;
			}
		}
	} while (currentNodeIndex != 0 || nextIsLeafNode);
	return false;
}

// ---------------------------------------------------------------- shading
// materials.glsl:9-30
struct MaterialData
{
	V4 Reflectiveness;
	V3 Opacity;
	V3 Diffuse;
	V3 Emissivity;
	V3 Fresnel0;
	float RefractionIndexAvg;
	V3 Fresnel1;
};
inline V3 xyz(V4 v) { return v3(v.x, v.y, v.z); }
MaterialData SampleMaterialData(const Scene& s, int materialID, V2 tc)
{
	const GiMaterial& m = s.materials[materialID];
	MaterialData d;
	d.Reflectiveness = sampleTexture(s.textures[m.reflectiveness_tex], tc.x, tc.y);
	d.Opacity = xyz(sampleTexture(s.textures[m.opacity_tex], tc.x, tc.y));
	d.Diffuse = xyz(sampleTexture(s.textures[m.diffuse_tex], tc.x, tc.y));
	d.Emissivity = xyz(sampleTexture(s.textures[m.emissivity_tex], tc.x, tc.y));
	d.Fresnel0 = v3(m.fresnel0);
	d.RefractionIndexAvg = m.refraction_index_avg;
	d.Fresnel1 = v3(m.fresnel1);
	return d;
}

// scenedata.glsl:94-105
void GetTriangleHitInfo(const Scene& s, const GiTriangle& tri, V3 bary, V3& normal, V2& texcoord)
{
	const GiVertexInfo& a = s.vinfo[tri.v[0]];
	const GiVertexInfo& b = s.vinfo[tri.v[1]];
	const GiVertexInfo& c = s.vinfo[tri.v[2]];
	normal = normalize(UnpackNormal(a.normal_phi, a.normal_z) * bary.x + UnpackNormal(b.normal_phi, b.normal_z) * bary.y +
	                   UnpackNormal(c.normal_phi, c.normal_z) * bary.z);
	texcoord.x = a.texcoord[0] * bary.x + b.texcoord[0] * bary.y + c.texcoord[0] * bary.z;
	texcoord.y = a.texcoord[1] * bary.x + b.texcoord[1] * bary.y + c.texcoord[1] * bary.z;
}

// materials.glsl:33-41
inline float AvgProbability(V3 p) { return (p.x + p.y + p.z + DIVISOR_EPSILON) / (3.0f + DIVISOR_EPSILON); }
inline float Avg(V3 p) { return (p.x + p.y + p.z) / 3.0f; }

// materials.glsl:43-63 with USE_CONSTANT_REFLECTION (materials.glsl:4)
inline void GetBSDFDecisionPropabilities(const MaterialData& md, float /*cosThetaAbs*/, V3& preflect, V3& prefract, V3& pdiffuse)
{
	preflect = md.Fresnel0;
	prefract = v3(1.0f) - saturate(preflect);
	pdiffuse = prefract * md.Opacity;
	prefract = prefract - pdiffuse;
}
// materials.glsl:87-103
inline V3 GetDiffuseProbability(const MaterialData& md, float /*cosThetaAbs*/)
{
	V3 preflect = md.Fresnel0;
	return (v3(1.0f) - saturate(preflect)) * md.Opacity;
}

// materials.glsl:118-277 (__SampleBSDF); rejection loop disabled by `plausible = true` (:232),
// so counter == 1; Fresnel corrections return vec3(1.0) (:70-83).
V3 SampleBSDFImpl(V3 incidentDirection, const MaterialData& md, uint32_t& seed, V3 N, V3& pathThroughput, bool adjoint, float& pdf,
                  bool& isReflectedDiffuse, bool stopOnDiffuseBounce = false /* STOP_ON_DIFFUSE_BOUNCE, materials.glsl:164-167 */)
{
	float cosTheta = dot(N, incidentDirection);
	float cosThetaAbs = saturate(fabsf(cosTheta));
	V3 reflectRefractDir = incidentDirection - (2.0f * cosTheta) * N;
	V2 randomSamples = Random2(seed);
	V3 preflect, prefract, pdiffuse;
	GetBSDFDecisionPropabilities(md, cosThetaAbs, preflect, prefract, pdiffuse);
	float avgPReflect = AvgProbability(preflect);
	float avgPRefract = AvgProbability(prefract);
	float avgPDiffuse = Avg(pdiffuse);
	V3 pphong = preflect;
	float avgPPhong = avgPReflect;
	float pathDecisionVar = Random(seed);
	if (avgPDiffuse > pathDecisionVar)
	{
		if (stopOnDiffuseBounce) { pathThroughput = v3(-1.0f, 0.0f, 0.0f); return v3(0.0f); }
		isReflectedDiffuse = true;
		pathThroughput *= (md.Diffuse * pdiffuse) / avgPDiffuse;
		if (cosTheta > 0.0f) N = -N;
		V3 U, V;
		CreateONB(N, U, V);
		V3 outDir = SampleHemisphereCosine(randomSamples, U, V, N);
		float cosThetaOut = dot(N, outDir);
		pdf = saturate(cosThetaOut) / PI;
		return outDir;
	}
	else if (avgPDiffuse + avgPRefract > pathDecisionVar)
	{
		pphong = prefract;
		avgPPhong = avgPRefract;
		float eta = cosTheta < 0.0f ? 1.0f / md.RefractionIndexAvg : md.RefractionIndexAvg;
		float etaSq = eta * eta;
		float sinTheta2Sq = etaSq * (1.0f - cosTheta * cosTheta);
		if (sinTheta2Sq < 1.0f)
		{
			reflectRefractDir = eta * incidentDirection + (signf_(cosTheta) * (sqrtf(saturate(1.0f - sinTheta2Sq)) - eta * cosThetaAbs)) * N;
			if (!adjoint)
				pathThroughput /= etaSq;
		}
	}
	reflectRefractDir = normalize(reflectRefractDir);
	V3 RU, RV;
	CreateONB(reflectRefractDir, RU, RV);
	V3 outDir = SamplePhongLobe(randomSamples, md.Reflectiveness.w, RU, RV, reflectRefractDir);
	if (!adjoint)
		pathThroughput *= fabsf(dot(N, outDir)) / fabsf(dot(N, reflectRefractDir));
	float cosThetaOutAbs = fabsf(dot(outDir, reflectRefractDir));
	pdf = (md.Reflectiveness.w + 1.0f) / PI_2 * powf(cosThetaOutAbs, md.Reflectiveness.w);
	pdf *= 1.0f; // float(counter), counter == 1
	pathThroughput *= pphong / avgPPhong;
	isReflectedDiffuse = false;
	return outDir;
}

// materials.glsl:306-391 (__BSDF)
V3 BSDFImpl(V3 incidentDirection, V3 excidentDirection, const MaterialData& md, V3 N, float& pdf, bool adjoint, bool specularOnly,
           bool stopOnDiffuseBounce = false /* materials.glsl:355-358: no refractive next event estimation */)
{
	V3 bsdf = v3(0.0f);
	float cosTheta = dot(N, incidentDirection);
	float cosThetaAbs = saturate(fabsf(cosTheta));
	float cosThetaOut = dot(N, excidentDirection);
	V3 preflect, prefract, pdiffuse;
	GetBSDFDecisionPropabilities(md, cosThetaAbs, preflect, prefract, pdiffuse);
	float phongFactor_brdf = (md.Reflectiveness.w + 1.0f) / (2 * PI_2 * fabsf(cosThetaOut) + DIVISOR_EPSILON);
	float phongNormalization_pdf = (md.Reflectiveness.w + 1.0f) / PI_2;
	V3 reflectRefractDir = normalize(incidentDirection - (2.0f * cosTheta) * N);
	float reflectionDotExcident_powN = powf(saturate(dot(reflectRefractDir, excidentDirection)), md.Reflectiveness.w);
	float reflrefrBRDFFactor = phongFactor_brdf * reflectionDotExcident_powN;
	float reflrefrPDFFactor = phongNormalization_pdf * reflectionDotExcident_powN;
	bsdf = preflect * reflrefrBRDFFactor;
	bsdf *= fabsf(cosThetaOut) / cosThetaAbs;
	pdf = AvgProbability(preflect) * reflrefrPDFFactor;
	if (!specularOnly)
		if (cosTheta * cosThetaOut < 0.0f)
		{
			bsdf += pdiffuse * md.Diffuse / PI;
			pdf += Avg(pdiffuse) * saturate(cosThetaOut) / PI;
		}
	if (stopOnDiffuseBounce) return bsdf;
	float eta = cosTheta < 0.0f ? 1.0f / md.RefractionIndexAvg : md.RefractionIndexAvg;
	float etaSq = eta * eta;
	float sinTheta2Sq = etaSq * (1.0f - cosTheta * cosTheta);
	if (sinTheta2Sq < 1.0f)
	{
		reflectRefractDir = normalize(eta * incidentDirection + (signf_(cosTheta) * (sqrtf(saturate(1.0f - sinTheta2Sq)) - eta * cosThetaAbs)) * N);
		float refractionDotExcident_powN = powf(saturate(dot(reflectRefractDir, excidentDirection)), md.Reflectiveness.w);
		reflrefrBRDFFactor = phongFactor_brdf * refractionDotExcident_powN;
		reflrefrPDFFactor = phongNormalization_pdf * refractionDotExcident_powN;
		if (!adjoint)
			reflrefrBRDFFactor /= etaSq;
		reflrefrBRDFFactor *= fabsf(dot(N, excidentDirection)) / fabsf(dot(N, reflectRefractDir));
	}
	bsdf += prefract * reflrefrBRDFFactor;
	pdf += AvgProbability(prefract) * reflrefrPDFFactor;
	return bsdf;
}
inline V3 BSDF(V3 in, V3 out, const MaterialData& md, V3 N, float& pdf) { return BSDFImpl(in, out, md, N, pdf, false, false); }       // :393-396
inline V3 AdjointBSDF(V3 in, V3 out, const MaterialData& md, V3 N, float& pdf) { return BSDFImpl(-out, -in, md, N, pdf, true, false); } // :405-408

// materials.glsl:415-436
float AdjointBSDFShadingNormalCorrectedOutDotN(V3 incomingLight, V3 toCamera, V3 geometryNormal, V3 shadingNormal)
{
	if (dot(incomingLight, geometryNormal) * dot(incomingLight, shadingNormal) < 0 || dot(toCamera, geometryNormal) * dot(toCamera, shadingNormal) < 0)
		return 0.0f;
	return saturate(fabsf(dot(incomingLight, shadingNormal) * dot(toCamera, geometryNormal)) / fmaxf(fabsf(dot(incomingLight, geometryNormal)), DIVISOR_EPSILON));
}
float AdjointBSDFShadingNormalCorrection(V3 inDir, V3 outDir, V3 geometryNormal, V3 shadingNormal)
{
	if (dot(inDir, geometryNormal) * dot(inDir, shadingNormal) < 0 || dot(outDir, geometryNormal) * dot(outDir, shadingNormal) < 0)
		return 0.0f;
	return saturate(fabsf(dot(inDir, shadingNormal) * dot(outDir, geometryNormal)) /
	                fmaxf(fabsf(dot(inDir, geometryNormal) * dot(outDir, shadingNormal)), DIVISOR_EPSILON));
}

// ---------------------------------------------------------------- per-render state
struct Counters
{
	uint64_t closest_rays, any_rays, box_tests, tri_tests, camera_paths, light_paths, shaded_vertices, nee, connections, splats, stored_light_vertices;
};

struct Frame
{
	const Scene* scene;
	GiCamera cam;
	int width, height;
	const GiLightSample* lightSamples;
	int numLightSamples;
	uint32_t frameSeed;
	int maxPathLength;
	float* accum; // RGBA32F sum image
	int reconFilter; // RECONSTRUCTION_FILTER (screenreproject.glsl:14-17): 0 box, 1 Mitchell 1/3 1/3, 2 cone, 3 B-spline
	bool chainedMIS; // CHAINED_MIS (bidirpathtracer/chained_mis.glsl)
};

// lightsource.glsl:3-93 (EstimateDirectLight, no TRACERAY_IMPORTANCE_BREAK / CACHE_DIRECT_DIFFUSE)
V3 EstimateDirectLight(const Frame& f, V3 pos, V3 normal, int lightSampleIndex, V3 viewDir, const MaterialData& md, bool sampleAreaLights, Counters& c,
                       TraceCounters& tc, bool stopOnDiffuseBounce = false)
{
	const GiLightSample& ls = f.lightSamples[lightSampleIndex];
	Ray lightRay;
	lightRay.Direction = v3(ls.position) - pos;
	float lightDistSq = dot(lightRay.Direction, lightRay.Direction) + DIVISOR_EPSILON;
	float lightDist = sqrtf(lightDistSq);
	lightRay.Direction /= lightDist;
	float surfaceCos = saturate(dot(lightRay.Direction, normal));
	if (surfaceCos <= 0.0f) return v3(0.0f);
	if (ls.normal_theta_cos > 1.5f)
	{
		lightRay.Origin = RAY_HIT_EPSILON * lightRay.Direction + pos;
		++c.any_rays;
		if (!TraceRayAnyHit(*f.scene, lightRay, lightDist - RAY_HIT_EPSILON * 2, &tc))
		{
			V3 irradiance = (surfaceCos / lightDistSq) * v3(ls.intensity);
			float pdf;
			V3 bsdf = BSDFImpl(viewDir, lightRay.Direction, md, normal, pdf, false, false, stopOnDiffuseBounce);
			return irradiance * bsdf;
		}
	}
	else if (sampleAreaLights)
	{
		V3 lightNormal = UnpackNormal(ls.normal_phi, ls.normal_theta_cos);
		float lightSampleIntensityFactor = saturate(dot(lightNormal, -lightRay.Direction));
		if (lightSampleIntensityFactor > 0.0f)
		{
			lightRay.Origin = RAY_HIT_EPSILON * lightRay.Direction + pos;
			++c.any_rays;
			if (!TraceRayAnyHit(*f.scene, lightRay, lightDist - RAY_HIT_EPSILON * 2, &tc))
			{
				float pdf;
				V3 lightSampleIntensity = v3(ls.intensity) * lightSampleIntensityFactor;
				V3 bsdf = BSDFImpl(viewDir, lightRay.Direction, md, normal, pdf, false, false, stopOnDiffuseBounce);
				V3 irradiance = (surfaceCos / lightDistSq) * lightSampleIntensity;
				return irradiance * bsdf;
			}
		}
	}
	return v3(0.0f);
}

inline V3 cameraRayDirection(const Frame& f, uint32_t& seed, int gx, int gy)
{
	// pathtracer.comp:39-46
	V2 r = Random2(seed);
	V2 sub; sub.x = r.x - 0.5f; sub.y = r.y - 0.5f;
	V2 sc;
	sc.x = (sub.x + float(gx)) / float(f.width) * 2.0f - 1.0f;
	sc.y = (sub.y + float(gy)) / float(f.height) * 2.0f - 1.0f;
	return normalize(sc.x * v3(f.cam.u) + sc.y * v3(f.cam.v) + v3(f.cam.w));
}

// textureLod(EnvironmentMap, vec4(dir, 0), 0) (pathtracer.comp:57) for a size x size SRGB8 cube (renderersystem.cpp:173-186): face
// selection and (s, t) of the GL 4.5 specification, table 8.19 / section 8.13; linear filter inside the selected face with
// clamp-to-edge (no seamless filtering), texels decoded from sRGB before filtering.  Faces in the order -X,+X,-Y,+Y,-Z,+Z, RGBA8.
struct EnvCube { const uint8_t* const* faces; int size; };

inline V3 sampleCube(const EnvCube& e, V3 r)
{
	const float ax = fabsf(r.x), ay = fabsf(r.y), az = fabsf(r.z);
	int face; float sc, tc, ma;
	if (ax >= ay && ax >= az) { ma = ax; if (r.x > 0) { face = 1; sc = -r.z; tc = -r.y; } else { face = 0; sc = r.z; tc = -r.y; } }
	else if (ay >= az) { ma = ay; if (r.y > 0) { face = 3; sc = r.x; tc = r.z; } else { face = 2; sc = r.x; tc = -r.z; } }
	else { ma = az; if (r.z > 0) { face = 5; sc = r.x; tc = -r.y; } else { face = 4; sc = -r.x; tc = -r.y; } }
	const float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
	const float fu = s * float(e.size) - 0.5f, fv = t * float(e.size) - 0.5f;
	const float iu = floorf(fu), iv = floorf(fv);
	const float a = fu - iu, b = fv - iv;
	auto clampi = [&](int i) { return i < 0 ? 0 : (i >= e.size ? e.size - 1 : i); };
	const int i0 = clampi(int(iu)), i1 = clampi(int(iu) + 1), j0 = clampi(int(iv)), j1 = clampi(int(iv) + 1);
	auto tex = [&](int i, int j) { const uint8_t* p = e.faces[face] + 4 * (size_t(j) * e.size + i); return v3(g_srgbLut[p[0]], g_srgbLut[p[1]], g_srgbLut[p[2]]); };
	const V3 t00 = tex(i0, j0), t10 = tex(i1, j0), t01 = tex(i0, j1), t11 = tex(i1, j1);
	const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
	return v3(t00.x * w00 + t10.x * w10 + t01.x * w01 + t11.x * w11, t00.y * w00 + t10.y * w10 + t01.y * w01 + t11.y * w11,
	          t00.z * w00 + t10.z * w10 + t01.z * w01 + t11.z * w11);
}

// pathtracer.comp:16-110, one invocation. envColor: textureLod(EnvironmentMap,..) of the
// default 1x1 cube (renderersystem.cpp:39-47) generalised to a constant colour.
void PathtracerPixel(const Frame& f, int gx, int gy, V3 envColor, Counters& c, TraceCounters& tc, const EnvCube* cube = nullptr)
{
	const Scene& s = *f.scene;
	const float lightSampleWeight = float(f.numLightSamples) / 1;
	uint32_t randomSeed = InitRandomSeed(f.frameSeed, uint32_t(gx + gy * f.width));
	V3 color = v3(0.0f);
	V3 pathThroughput = v3(1.0f);
	bool didNextEventEstimation = false;
	Ray ray;
	ray.Origin = v3(f.cam.position);
	ray.Direction = cameraRayDirection(f, randomSeed, gx, gy);
	++c.camera_paths;
	for (int i = 0; i < f.maxPathLength; ++i)
	{
		GiTriangle triangle; V3 bary; float rayHit = RAY_MAX; V3 geometryNormal; int hitIndex[2];
		++c.closest_rays;
		TraceRay(s, ray, rayHit, hitIndex, bary, triangle, geometryNormal, &tc);
		if (rayHit == RAY_MAX)
		{
			color += pathThroughput * (cube ? sampleCube(*cube, ray.Direction) : envColor);
			break;
		}
		geometryNormal = normalize(geometryNormal);
		ray.Origin = ray.Origin + rayHit * ray.Direction;
		V3 hitNormal; V2 hitTexcoord;
		GetTriangleHitInfo(s, triangle, bary, hitNormal, hitTexcoord);
		MaterialData md = SampleMaterialData(s, int(triangle.material), hitTexcoord);
		++c.shaded_vertices;
		if (!didNextEventEstimation)
			color += pathThroughput * md.Emissivity;
		didNextEventEstimation = md.Reflectiveness.w < 500.0f;
		{
			int sampleIndex = int(RandomUInt(randomSeed) % uint32_t(f.numLightSamples));
			++c.nee;
			V3 radiance = EstimateDirectLight(f, ray.Origin, hitNormal, sampleIndex, ray.Direction, md, didNextEventEstimation, c, tc);
			color += radiance * pathThroughput * lightSampleWeight;
		}
		V3 throughput = v3(1.0f);
		float pdf; bool dummy;
		ray.Direction = SampleBSDFImpl(ray.Direction, md, randomSeed, hitNormal, throughput, false, pdf, dummy);
		ray.Origin += ray.Direction * RAY_HIT_EPSILON;
		pathThroughput *= throughput;
		float continuationPropability = saturate(GetLuminance(throughput));
		if (Random(randomSeed) >= continuationPropability)
			break;
		pathThroughput /= continuationPropability;
	}
	float* px = f.accum + 4 * (size_t(gy) * f.width + gx);
	px[0] += color.x; px[1] += color.y; px[2] += color.z; px[3] += 0.0f;
}

// screenreproject.glsl:19-49 (WriteAtomic) with an explicit bounds check.
inline void WriteAtomic(const Frame& f, V3 value, int px, int py, Counters& c)
{
	if (px < 0 || py < 0 || px >= f.width || py >= f.height) return;
	float* p = f.accum + 4 * (size_t(py) * f.width + px);
	++c.splats;
#pragma omp atomic
	p[0] += value.x;
#pragma omp atomic
	p[1] += value.y;
#pragma omp atomic
	p[2] += value.z;
}

// lightcache.glsl:61-88
inline float MISHeuristic(float p) { return fminf(1e18f, p); }
inline float InitialMIS() { return 1.0f; }
inline float DistanceMISHeuristic(float d, float /*c*/) { return MISHeuristic(d * d); }

// bidirpathtracer/chained_mis.glsl:1-34.  g_mis is the shader's global array, here owned by the eye path.
#define OR_MAX_PATHLENGTH 64
inline void updateQuality(float& prevQuality, float mis) { prevQuality = prevQuality * (mis + 1.0f); }
inline float qualityWeightedMIS_NEE(float mis_nee, float quality_nee, float quality_trace)
{
	return mis_nee * quality_nee / fmaxf(mis_nee * quality_nee + (1 - mis_nee) * quality_trace, DIVISOR_EPSILON);
}
inline float qualityWeightedMIS_Trace(float mis_trace, float quality_nee, float quality_trace)
{
	return mis_trace * quality_trace / fmaxf((1 - mis_trace) * quality_nee + mis_trace * quality_trace, DIVISOR_EPSILON);
}
inline float cmisPathThroughputWeight(const float* g_mis, int pathSegment, float quality_nee, float quality_trace)
{
	float w = 1.0f; // (chained_mis.glsl:25-26: the first assignment is dead)
	for (int i = pathSegment - 1; i >= 0; --i)
	{
		w *= qualityWeightedMIS_Trace(g_mis[i], quality_nee, quality_trace);
		updateQuality(quality_nee, g_mis[i]);
		updateQuality(quality_trace, 1.0f - g_mis[i]);
	}
	return w;
}

// the reconstruction filter kernels of screenreproject.glsl:186-223, as a function of the distance to the pixel centre
inline float reconstructionFilterValue(int filter, float d)
{
	float filterValue = 0.0f;
	if (filter == 1)
	{
		if (d < 1.0f) filterValue = (7.0f * d * d * d - 12.0f * d * d + 16.0f / 3.0f) / 6.0f;
		else if (d < 2.0f) filterValue = (-7.0f / 3.0f * d * d * d + 12.0f * d * d - 20.0f * d + 32.0f / 3.0f) / 6.0f;
	}
	else if (filter == 2)
	{
		if (d < 1.5f) filterValue = 1.0f - d / 1.5f;
	}
	else if (filter == 3)
	{
		if (d < 1.0f) filterValue = (3.0f * d * d * d - 6.0f * d * d + 4.0f) / 6.0f;
		else if (d < 2.0f) filterValue = (-d * d * d + 6.0f * d * d - 12.0f * d + 8.0f) / 6.0f;
	}
	return filterValue;
}

// screenreproject.glsl:55-227 (ProjectToScreen).  f.reconFilter selects RECONSTRUCTION_FILTER, f.chainedMIS the CHAINED_MIS variant, whose four
// extra arguments (:64-69) are ignored otherwise.
void ProjectToScreen(const Frame& f, const Ray& ray, V3 geometryNormal, V3 shadingNormal, const MaterialData& md, V3 pathThroughput, bool computeFullBSDF,
                     bool computeMIS, Counters& c, TraceCounters& tc, float cmisQualityN = 1.0f, float cmisQualityT = 1.0f, float lastConnectionProbability = 0.0f,
                     float lastDistMIS = 0.0f)
{
	V3 CameraPosition = v3(f.cam.position), CameraU = v3(f.cam.u), CameraV = v3(f.cam.v), CameraW = v3(f.cam.w);
	Ray cameraRay;
	cameraRay.Direction = CameraPosition - ray.Origin;
	float camViewDotRay = -dot(CameraW, cameraRay.Direction);
	if (camViewDotRay > 0.0f)
	{
		V3 proj = v3(-dot(CameraU, cameraRay.Direction), -dot(CameraV, cameraRay.Direction), camViewDotRay);
		V2 screenCoord; screenCoord.x = proj.x / proj.z; screenCoord.y = proj.y / proj.z;
		screenCoord.x /= dot(CameraU, CameraU);
		screenCoord.y /= dot(CameraV, CameraV);
		if (screenCoord.x >= -1.0f && screenCoord.x < 1.0f && screenCoord.y > -1.0f && screenCoord.y <= 1.0f)
		{
			float cameraDistSq = dot(cameraRay.Direction, cameraRay.Direction);
			float cameraDist = sqrtf(cameraDistSq);
			cameraRay.Direction /= cameraDist;
			cameraRay.Origin = cameraRay.Direction * RAY_HIT_EPSILON + ray.Origin;
			V2 pixelPos; pixelPos.x = (screenCoord.x * 0.5f + 0.5f) * float(f.width); pixelPos.y = (screenCoord.y * 0.5f + 0.5f) * float(f.height);
			int pcx = int(roundf(pixelPos.x)), pcy = int(roundf(pixelPos.y));
			++c.any_rays;
			if (!TraceRayAnyHit(*f.scene, cameraRay, cameraDist, &tc))
			{
				float cosAtCamera = saturate(camViewDotRay / cameraDist);
				float fluxToIrradiance = AdjointBSDFShadingNormalCorrectedOutDotN(ray.Direction, cameraRay.Direction, geometryNormal, shadingNormal) /
				                         (cosAtCamera * cosAtCamera * cosAtCamera * f.cam.pixel_area * cameraDistSq);
				V3 bsdf = v3(1.0f);
				if (computeFullBSDF)
				{
					float pdf;
					bsdf = AdjointBSDF(ray.Direction, cameraRay.Direction, md, shadingNormal, pdf);
					if (computeMIS)
					{
						float connectionEyeToLightPath = InitialMIS();
						float connectionLightPathToEye = MISHeuristic(pdf);
						float mis = connectionEyeToLightPath / fmaxf(connectionLightPathToEye + connectionEyeToLightPath, DIVISOR_EPSILON);
						if (f.chainedMIS)
						{
							// screenreproject.glsl:160-170
							float lastPTmis = lastConnectionProbability / fmaxf(lastConnectionProbability + connectionLightPathToEye * lastDistMIS, DIVISOR_EPSILON);
							updateQuality(cmisQualityN, lastPTmis);
							updateQuality(cmisQualityT, 1.0f - lastPTmis);
							mis = qualityWeightedMIS_NEE(mis, cmisQualityN, cmisQualityT);
						}
						bsdf *= mis;
					}
				}
				else
				{
					float surfaceInCos = dot(ray.Direction, shadingNormal);
					float surfaceOutCos = dot(cameraRay.Direction, shadingNormal);
					if (surfaceInCos * surfaceOutCos < 0.0f)
					{
						V3 pDiffuse = GetDiffuseProbability(md, fabsf(surfaceInCos));
						bsdf = pDiffuse * md.Diffuse * Avg(pDiffuse) / PI;
					}
					else bsdf = v3(0.0f);
				}
				V3 color = pathThroughput * fluxToIrradiance * bsdf;
				if (f.reconFilter == 0)
					WriteAtomic(f, color, pcx, pcy, c);
				else
				{
					// screenreproject.glsl:186-223: 5 x 5 pixels around the rounded position, weighted by the distance to pixelPos
					for (int y = -2; y <= 2; ++y)
						for (int x = -2; x <= 2; ++x)
						{
							int px = pcx + x, py = pcy + y;
							float ddx = pixelPos.x - float(px), ddy = pixelPos.y - float(py);
							float d = sqrtf(ddx * ddx + ddy * ddy);
							float filterValue = reconstructionFilterValue(f.reconFilter, d);
							if (filterValue != 0.0f) WriteAtomic(f, filterValue * color, px, py, c);
						}
				}
			}
		}
	}
}

// lightcache.glsl:1-16
struct LightCacheEntry
{
	float Flux[3]; int32_t MaterialIndex;
	float Position[3]; float Normal0;
	float Texcoord[2]; float Normal1; int32_t PathLength;
	float IncidentDirectionPacked[2]; float ChainedMISWeightN; float ChainedMISWeightT;
};

struct LightCache
{
	LightCacheEntry* entries;
	uint32_t capacity;          // LightCacheCapacity
	uint32_t numEntries;        // NumLightCacheEntries (atomic)
	uint32_t pathLengthSum;     // LightPathLengthSum (warm-up)
	bool warmup;
};

// lighttracer.comp:23-45
inline void saveLightCache(LightCache& lc, uint32_t idx, V3 flux, int materialIndex, V3 position, V2 packedShadingNormal, V2 hitTexcoord, V3 incidentDirection,
                           int pathLength, Counters& c)
{
	if (idx < lc.capacity)
	{
		LightCacheEntry& e = lc.entries[idx];
		e.Flux[0] = flux.x; e.Flux[1] = flux.y; e.Flux[2] = flux.z;
		e.MaterialIndex = materialIndex;
		e.Position[0] = position.x; e.Position[1] = position.y; e.Position[2] = position.z;
		e.Normal0 = packedShadingNormal.x; e.Normal1 = packedShadingNormal.y;
		e.Texcoord[0] = hitTexcoord.x; e.Texcoord[1] = hitTexcoord.y;
		V2 pi = PackNormal(incidentDirection);
		e.IncidentDirectionPacked[0] = pi.x; e.IncidentDirectionPacked[1] = pi.y;
		e.PathLength = pathLength;
		e.ChainedMISWeightN = 0.0f; e.ChainedMISWeightT = 0.0f;
		++c.stored_light_vertices;
	}
}

// lighttracer.comp:49-193, one invocation; lc == nullptr => plain LPT (no SAVE_LIGHT_CACHE).
void LighttracerPath(const Frame& f, uint32_t gid, int numRaysPerLightSample, LightCache* lc, Counters& c, TraceCounters& tc)
{
	const Scene& s = *f.scene;
	uint32_t randomSeed = InitRandomSeed(f.frameSeed, gid);
	int lightSampleIndex = int(gid / uint32_t(numRaysPerLightSample));
	const GiLightSample& ls = f.lightSamples[lightSampleIndex];
	++c.light_paths;
	if (lc && !lc->warmup)
	{
		V2 pn; pn.x = ls.normal_phi; pn.y = ls.normal_theta_cos;
		V2 zero2; zero2.x = 0.0f; zero2.y = 0.0f;
		saveLightCache(*lc, gid, v3(ls.intensity) / PI_2 / float(numRaysPerLightSample), -1, v3(ls.position), pn, zero2, v3(0.0f), 0, c);
	}
	// CHAINED_MIS state of the light path (lighttracer.comp:62-69)
	const bool cmis = f.chainedMIS && lc && !lc->warmup;
	float lastConnectionProbability = 0.0f; // Point light init (cannot be traced by a PT)
	float lastCosAbs = 1.0f;
	float chainedMISWeightN = 1.0f, chainedMISWeightT = 1.0f;
	if (cmis && gid < lc->capacity) { lc->entries[gid].ChainedMISWeightN = chainedMISWeightN; lc->entries[gid].ChainedMISWeightT = chainedMISWeightT; }
	V3 pathThroughput = v3(ls.intensity) * (PI / float(numRaysPerLightSample));
	Ray ray;
	ray.Origin = v3(ls.position);
	if (ls.normal_theta_cos > 1.5f)
	{
		ray.Direction = SampleDirection(Random2(randomSeed));
		pathThroughput *= 4.0f;
	}
	else
	{
		V3 normal = UnpackNormal(ls.normal_phi, ls.normal_theta_cos);
		V3 U, V;
		CreateONB(normal, U, V);
		ray.Direction = SampleHemisphereCosine(Random2(randomSeed), U, V, normal);
		if (cmis)
		{
			// lighttracer.comp:89-92
			lastCosAbs = saturate(dot(normal, ray.Direction));
			lastConnectionProbability = MISHeuristic(lastCosAbs / PI);
		}
	}
	for (int i = 0; i < f.maxPathLength; ++i)
	{
		GiTriangle triangle; V3 bary; float rayHit = RAY_MAX; V3 geometryNormal; int hitIndex[2];
		++c.closest_rays;
		TraceRay(s, ray, rayHit, hitIndex, bary, triangle, geometryNormal, &tc);
		if (rayHit == RAY_MAX)
			break;
		ray.Origin = ray.Origin + rayHit * ray.Direction;
		V3 shadingNormal; V2 hitTexcoord;
		GetTriangleHitInfo(s, triangle, bary, shadingNormal, hitTexcoord);
		MaterialData md = SampleMaterialData(s, int(triangle.material), hitTexcoord);
		++c.shaded_vertices;
		uint32_t lightCacheEntryIndex = ~0u;
		if (lc)
		{
			if (md.Reflectiveness.w < 100.0f)
			{
#pragma omp atomic capture
				lightCacheEntryIndex = lc->numEntries++;
				if (lc->warmup)
				{
#pragma omp atomic
					lc->pathLengthSum += 1;
				}
				else
					saveLightCache(*lc, lightCacheEntryIndex, pathThroughput, int(triangle.material), ray.Origin, PackNormal(shadingNormal), hitTexcoord,
					               ray.Direction, i + 1, c);
			}
		}
		if (!(lc && lc->warmup))
		{
			if (cmis) // lighttracer.comp:141-144
				ProjectToScreen(f, ray, geometryNormal, shadingNormal, md, pathThroughput, true, true, c, tc, chainedMISWeightN, chainedMISWeightT, lastConnectionProbability,
				                DistanceMISHeuristic(rayHit, lastCosAbs));
			else
				ProjectToScreen(f, ray, geometryNormal, shadingNormal, md, pathThroughput, true, lc != nullptr, c, tc);
		}
		V3 throughput = v3(1.0f);
		float samplePDF; bool isReflectedDiffuse;
		V3 newDirection = SampleBSDFImpl(ray.Direction, md, randomSeed, shadingNormal, throughput, true, samplePDF, isReflectedDiffuse);
		throughput *= AdjointBSDFShadingNormalCorrection(ray.Direction, newDirection, geometryNormal, shadingNormal);
		pathThroughput *= throughput;
		if (cmis)
		{
			// lighttracer.comp:165-181: how likely a PT could have generated the last path segment
			float reverseConnectionProbability = MISHeuristic(samplePDF);
			float distMIS = DistanceMISHeuristic(rayHit, lastCosAbs);
			float mis = lastConnectionProbability / fmaxf(lastConnectionProbability + distMIS * reverseConnectionProbability, DIVISOR_EPSILON);
			updateQuality(chainedMISWeightN, mis);
			updateQuality(chainedMISWeightT, 1.0f - mis);
			if (lightCacheEntryIndex != ~0u && lightCacheEntryIndex < lc->capacity)
			{
				lc->entries[lightCacheEntryIndex].ChainedMISWeightN = chainedMISWeightN;
				lc->entries[lightCacheEntryIndex].ChainedMISWeightT = chainedMISWeightT;
			}
			lastConnectionProbability = reverseConnectionProbability;
			lastCosAbs = fabsf(dot(ray.Direction, geometryNormal)); // (the old direction and the un-normalised geometry normal, as written)
		}
		float continuationPropability = saturate(GetLuminance(throughput));
		if (Random(randomSeed) >= continuationPropability)
			break;
		pathThroughput /= continuationPropability;
		ray.Direction = newDirection;
		ray.Origin += ray.Direction * RAY_HIT_EPSILON;
	}
}

// showlightcaches.comp:10-64, restated with exactly NumInitialLightSamples invocations and
// an atomic splat (the reference over-dispatches 64x and races).
void ShowLightCache(const Frame& f, int index, Counters& c, TraceCounters& tc)
{
	const GiLightSample& ls = f.lightSamples[index];
	V3 CameraPosition = v3(f.cam.position), CameraU = v3(f.cam.u), CameraV = v3(f.cam.v), CameraW = v3(f.cam.w);
	Ray cameraRay;
	cameraRay.Direction = CameraPosition - v3(ls.position);
	float camViewDotRay = -dot(CameraW, cameraRay.Direction);
	if (camViewDotRay > 0.0f)
	{
		V3 proj = v3(-dot(CameraU, cameraRay.Direction), -dot(CameraV, cameraRay.Direction), camViewDotRay);
		V2 screenCoord; screenCoord.x = proj.x / proj.z; screenCoord.y = proj.y / proj.z;
		screenCoord.x /= dot(CameraU, CameraU);
		screenCoord.y /= dot(CameraV, CameraV);
		if (screenCoord.x >= -1.0f && screenCoord.x < 1.0f && screenCoord.y > -1.0f && screenCoord.y <= 1.0f)
		{
			float cameraDistSq = dot(cameraRay.Direction, cameraRay.Direction);
			float cameraDist = sqrtf(cameraDistSq);
			cameraRay.Direction /= cameraDist;
			cameraRay.Origin = cameraRay.Direction * RAY_HIT_EPSILON + v3(ls.position);
			++c.any_rays;
			if (!TraceRayAnyHit(*f.scene, cameraRay, cameraDist, &tc))
			{
				V3 normal = UnpackNormal(ls.normal_phi, ls.normal_theta_cos);
				float cosTheta = dot(normal, cameraRay.Direction);
				float cosAtCamera = saturate(camViewDotRay / cameraDist);
				float intensityToRadiance = cosTheta / (PI_2 * cosAtCamera * cosAtCamera * cosAtCamera * f.cam.pixel_area * cameraDistSq);
				V3 color = v3(ls.intensity) * intensityToRadiance;
				int pcx = int((screenCoord.x * 0.5f + 0.5f) * float(f.width));
				int pcy = int((screenCoord.y * 0.5f + 0.5f) * float(f.height));
				WriteAtomic(f, color, pcx, pcy, c);
			}
		}
	}
}

// bidirpathtracer/pathtracer_bidir.comp:16-251, one invocation (no CHAINED_MIS).
void BidirPathtracerPixel(const Frame& f, int gx, int gy, const LightCache& lc, Counters& c, TraceCounters& tc)
{
	const Scene& s = *f.scene;
	uint32_t randomSeed = InitRandomSeed(f.frameSeed, uint32_t(gx + gy * f.width));
	Ray ray;
	ray.Origin = v3(f.cam.position);
	ray.Direction = cameraRayDirection(f, randomSeed, gx, gy);
	++c.camera_paths;
	V3 color = v3(0.0f);
	V3 pathThroughput = v3(1.0f);
	uint32_t actualNumLightCacheEntries = std::min(lc.numEntries, lc.capacity);
	float lightSampleWeight = float(actualNumLightCacheEntries);
	float lastConnectionProbability = InitialMIS();
	float lastCosAbs = 1.0f;
	float mis = 0.0f;
	float g_mis[OR_MAX_PATHLENGTH];
	for (int i = 0; i < OR_MAX_PATHLENGTH; ++i) g_mis[i] = 0.0f; // pathtracer_bidir.comp:49-52
	for (int i = 0; i < f.maxPathLength; ++i)
	{
		GiTriangle triangle; V3 bary; float rayHit = RAY_MAX; V3 geometryNormal; int hitIndex[2];
		++c.closest_rays;
		TraceRay(s, ray, rayHit, hitIndex, bary, triangle, geometryNormal, &tc);
		if (rayHit == RAY_MAX)
			break;
		geometryNormal = normalize(geometryNormal);
		ray.Origin = ray.Origin + rayHit * ray.Direction;
		V3 hitNormal; V2 hitTexcoord;
		GetTriangleHitInfo(s, triangle, bary, hitNormal, hitTexcoord);
		MaterialData md = SampleMaterialData(s, int(triangle.material), hitTexcoord);
		++c.shaded_vertices;
		{
			float incidentSurfaceCosSat = saturate(-dot(hitNormal, ray.Direction));
			float connectionLightToEyePath = DistanceMISHeuristic(rayHit, lastCosAbs) * MISHeuristic(incidentSurfaceCosSat / PI);
			mis = (i == 0) ? 1.0f : (lastConnectionProbability / fmaxf(lastConnectionProbability + connectionLightToEyePath, DIVISOR_EPSILON));
			if (f.chainedMIS)
			{
				// pathtracer_bidir.comp:85-89
				g_mis[i] = mis;
				mis = cmisPathThroughputWeight(g_mis, i + 1, 1.0f, 1.0f);
			}
			color += mis * pathThroughput * md.Emissivity;
		}
		Ray lightRay;
		for (int lightSample = 0; lightSample < 1; ++lightSample)
		{
			uint32_t lightCacheSampleIndex = RandomUInt(randomSeed) % actualNumLightCacheEntries;
			const LightCacheEntry& e = lc.entries[lightCacheSampleIndex];
			int pathLength = e.PathLength + i + 1;
			if (pathLength > f.maxPathLength)
				continue;
			++c.connections;
			lightRay.Direction = v3(e.Position) - ray.Origin;
			float lightDistSq = dot(lightRay.Direction, lightRay.Direction);
			float lightDist = sqrtf(lightDistSq);
			lightRay.Direction /= fmaxf(lightDist, DIVISOR_EPSILON);
			lightRay.Origin = RAY_HIT_EPSILON * lightRay.Direction + ray.Origin;
			++c.any_rays;
			if (!TraceRayAnyHit(s, lightRay, lightDist - RAY_HIT_EPSILON * 2, &tc))
			{
				int lightSampleMaterialIndex = e.MaterialIndex;
				V3 lightSampleFlux = v3(e.Flux);
				V3 lightSampleHitNormal = v3(0.0f, 1.0f, 0.0f);
				float connectionLightPathToEyePath;
				float cosLight = 0.0f;
				V3 lightSampleIntensity;
				float cosCamera = fabsf(dot(lightRay.Direction, hitNormal));
				if (e.Normal1 < 1.5f)
				{
					lightSampleHitNormal = UnpackNormalNoSat(e.Normal0, e.Normal1);
					cosLight = fabsf(dot(lightRay.Direction, lightSampleHitNormal));
				}
				float pdf;
				V3 bsdf = BSDF(ray.Direction, lightRay.Direction, md, hitNormal, pdf);
				float connectionEyePathToLightPath = MISHeuristic(pdf);
				if (i > 0)
				{
					float distMIS = DistanceMISHeuristic(rayHit, lastCosAbs);
					connectionEyePathToLightPath *= distMIS;
					mis = lastConnectionProbability / fmaxf(lastConnectionProbability + connectionEyePathToLightPath, DIVISOR_EPSILON);
					if (md.Reflectiveness.w >= 100.0f)
						mis = 1.0f;
				}
				else
					mis = connectionEyePathToLightPath / fmaxf(lastConnectionProbability + connectionEyePathToLightPath, DIVISOR_EPSILON);
				float misLight = 1.0f;
				if (e.Normal1 < 1.5f)
				{
					if (lightSampleMaterialIndex != -1)
					{
						V2 ltc; ltc.x = e.Texcoord[0]; ltc.y = e.Texcoord[1];
						MaterialData lmd = SampleMaterialData(s, lightSampleMaterialIndex, ltc);
						V3 incidentDir = UnpackNormal(e.IncidentDirectionPacked[0], e.IncidentDirectionPacked[1]);
						V3 lightBSDF = AdjointBSDF(lightRay.Direction, -incidentDir, lmd, lightSampleHitNormal, pdf);
						lightSampleIntensity = lightSampleFlux * lightBSDF * cosLight;
						connectionLightPathToEyePath = MISHeuristic(pdf) * DistanceMISHeuristic(lightDist, cosCamera);
					}
					else
					{
						lightSampleIntensity = lightSampleFlux * (cosLight * PI_2);
						connectionLightPathToEyePath = MISHeuristic(cosLight / PI) * DistanceMISHeuristic(lightDist, cosCamera);
					}
					misLight = connectionLightPathToEyePath / fmaxf(connectionLightPathToEyePath + connectionEyePathToLightPath, DIVISOR_EPSILON);
				}
				else
				{
					lightSampleIntensity = lightSampleFlux * (PI_2);
					misLight = 1.0f;
				}
				if (f.chainedMIS)
				{
					// pathtracer_bidir.comp:187-199
					float quality_nee = e.ChainedMISWeightN;
					float quality_trace = e.ChainedMISWeightT;
					misLight = qualityWeightedMIS_NEE(misLight, quality_nee, quality_trace);
					updateQuality(quality_nee, 1.0f - misLight);
					updateQuality(quality_trace, misLight);
					g_mis[i] = mis;
					mis = cmisPathThroughputWeight(g_mis, i + 1, quality_nee, quality_trace);
				}
				mis *= misLight;
				V3 irradiance = (mis * lightSampleWeight * cosCamera / lightDistSq) * lightSampleIntensity;
				color += irradiance * pathThroughput * bsdf;
			}
		}
		float samplePDF; bool dummy;
		V3 throughput = v3(1.0f);
		ray.Direction = SampleBSDFImpl(ray.Direction, md, randomSeed, hitNormal, throughput, false, samplePDF, dummy);
		ray.Origin += ray.Direction * RAY_HIT_EPSILON;
		pathThroughput *= throughput;
		float continuationPropability = saturate(GetLuminance(throughput));
		if (Random(randomSeed) >= continuationPropability)
			break;
		pathThroughput /= continuationPropability;
		float reverseConnectionProbability = MISHeuristic(samplePDF);
		if (i == 0)
			mis = reverseConnectionProbability / fmaxf(reverseConnectionProbability + lastConnectionProbability, DIVISOR_EPSILON);
		else
		{
			float distMIS = DistanceMISHeuristic(rayHit, lastCosAbs);
			reverseConnectionProbability *= distMIS;
			mis = lastConnectionProbability / fmaxf(lastConnectionProbability + distMIS * reverseConnectionProbability, DIVISOR_EPSILON);
			if (md.Reflectiveness.w >= 100.0f)
				mis = 1.0f;
		}
		if (f.chainedMIS) g_mis[i] = mis; // pathtracer_bidir.comp:236-242: the weight changes with the future of the path
		else pathThroughput *= mis;
		lastConnectionProbability = reverseConnectionProbability;
		lastCosAbs = fabsf(dot(ray.Direction, hitNormal));
	}
	float* px = f.accum + 4 * (size_t(gy) * f.width + gx);
	px[0] += color.x; px[1] += color.y; px[2] += color.z; px[3] += 1.0f;
}

// ---------------------------------------------------------------- host-side restatements
// utilities/random.cpp:5-13,25-43
inline uint64_t Xorshift64(uint64_t s)
{
	s ^= s >> 12;
	s ^= s << 25;
	s ^= s >> 27;
	return s * UINT64_C(2685821657736338717);
}
inline float XorshiftF(uint64_t& s)
{
	s = Xorshift64(s);
	return static_cast<float>(s % 8388593) / 8388593.0f;
}

struct LightTriangle { V3 v0, v1, v2; V3 luminance; };
struct PointLight { V3 position; V3 intensity; };

struct LightSamplerState
{
	std::vector<LightTriangle> lightTriangles;
	std::vector<float> lightSummedArea;
	std::vector<PointLight> pointLights;
	std::vector<float> pointLightSummedFlux;
	float lightAreaSum, totalAreaLightFlux, totalPointLightFlux;
	uint64_t randomSeed;
};

} // namespace

// ================================================================= C API
struct OrCounters
{
	uint64_t closest_rays, any_rays, box_tests, tri_tests, camera_paths, light_paths, shaded_vertices, nee, connections, splats, stored_light_vertices;
};

static void addCounters(OrCounters* out, const Counters& c, const TraceCounters& tc)
{
	if (!out) return;
#pragma omp critical(or_counters)
	{
		out->closest_rays += c.closest_rays; out->any_rays += c.any_rays; out->box_tests += tc.boxes; out->tri_tests += tc.tris;
		out->camera_paths += c.camera_paths; out->light_paths += c.light_paths; out->shaded_vertices += c.shaded_vertices; out->nee += c.nee;
		out->connections += c.connections; out->splats += c.splats; out->stored_light_vertices += c.stored_light_vertices;
	}
}

OR_API int or_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

// launchers such as torchrun export OMP_NUM_THREADS=1 to every rank; the CPU baseline is meant to use all host cores
OR_API void or_set_num_threads(int n)
{
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#endif
}

OR_API uint32_t or_wanghash(uint32_t x) { return wanghash(x); }
// the remaining scalar generators, exported so that tests can pin them against oracle/_ref (the reference's own utilities/random.cpp,
// compiled where it lies) and the golden vectors generated from it: random.glsl:11-19 == random.cpp:15-22 (xorshift32), random.cpp:25-43
OR_API uint32_t or_xorshift32(uint32_t s) { uint32_t t = s; return RandomUInt(t); }
OR_API uint64_t or_xorshift64(uint64_t s) { return Xorshift64(s); }
OR_API float or_xorshift_f(uint64_t* state) { return XorshiftF(*state); }

// TraceRay / TraceRayAnyHit over a batch.  node_visits / tri_visits (optional, per ray) are the
// TRACERAY_DEBUG_VARS counters (traceray.glsl:59-61,131-133) that define the algorithmic bytes.
OR_API int or_trace_batch(const GiSceneDesc* desc, const GiRay* rays, uint32_t n, GiHit* hits, uint32_t* node_visits, uint32_t* tri_visits)
{
	Scene s = makeScene(desc);
#pragma omp parallel for schedule(dynamic, 256)
	for (int64_t i = 0; i < int64_t(n); ++i)
	{
		Ray r; r.Origin = v3(rays[i].origin); r.Direction = v3(rays[i].direction);
		TraceCounters tc = {0, 0};
		GiHit h;
		if (rays[i].flags & 1u)
		{
			bool occluded = TraceRayAnyHit(s, r, rays[i].t_max, &tc);
			h.t = rays[i].t_max; h.leaf_tri = occluded ? 1 : 0; h.bary_y = 0.0f; h.bary_z = 0.0f;
		}
		else
		{
			float len = rays[i].t_max; int hi[2]; V3 bary = v3(0.0f); GiTriangle tri; V3 nrm;
			TraceRay(s, r, len, hi, bary, tri, nrm, &tc);
			h.t = len; h.leaf_tri = hi[1]; h.bary_y = hi[1] >= 0 ? bary.y : 0.0f; h.bary_z = hi[1] >= 0 ? bary.z : 0.0f;
		}
		hits[i] = h;
		if (node_visits) node_visits[i] = uint32_t(tc.boxes);
		if (tri_visits) tri_visits[i] = uint32_t(tc.tris);
	}
	return 0;
}

// Single primitive tests for the known-answer checks.
OR_API int or_intersect_box(const float o[3], const float d[3], const float bmin[3], const float bmax[3], float* first, float* last)
{
	V3 inv = v3(1.0f) / v3(d);
	return IntersectBox(v3(o), inv, v3(bmin), v3(bmax), *first, *last) ? 1 : 0;
}
OR_API int or_intersect_triangle(const float o[3], const float d[3], const float p0[3], const float p1[3], const float p2[3], float* t, float bary[3], float nrm[3])
{
	Ray r; r.Origin = v3(o); r.Direction = v3(d);
	V3 b, n; float h;
	bool hit = IntersectTriangle(r, v3(p0), v3(p1), v3(p2), h, b, n);
	*t = h; bary[0] = b.x; bary[1] = b.y; bary[2] = b.z; nrm[0] = n.x; nrm[1] = n.y; nrm[2] = n.z;
	return hit ? 1 : 0;
}

struct OrRegion { int x0, y0, x1, y1; }; // pixel rectangle [x0,x1) x [y0,y1); all-zero => full frame

// shader #defines that select variants (RECONSTRUCTION_FILTER: screenreproject.glsl:14-17, CHAINED_MIS: bidirectionalpathtracer.cpp:142); they apply
// to the render calls that follow
static int g_reconFilter = 0;
static int g_chainedMIS = 0;
OR_API void or_set_variant(int reconstruction_filter, int chained_mis) { g_reconFilter = reconstruction_filter; g_chainedMIS = chained_mis; }

static Frame makeFrame(const Scene& s, const GiCamera* cam, int w, int h, const GiLightSample* ls, int nls, uint32_t frameSeed, int maxLen, float* accum)
{
	Frame f;
	f.scene = &s; f.cam = *cam; f.width = w; f.height = h; f.lightSamples = ls; f.numLightSamples = nls; f.frameSeed = frameSeed; f.maxPathLength = maxLen;
	f.accum = accum;
	f.reconFilter = g_reconFilter; f.chainedMIS = g_chainedMIS != 0;
	return f;
}

// Pathtracer::Draw (pathtracer.cpp:47-50): (W/8) x (H/8) groups of 8x8 => pixels beyond the last
// full group are not rendered.
static const uint8_t* const* g_envFaces = nullptr;
static int g_envSize = 0;
// Environment cube for the following or_render_pt calls (size 0 / NULL: back to the constant colour argument).
OR_API void or_set_environment_cube(int size, const uint8_t* const* faces) { g_envFaces = size > 0 ? faces : nullptr; g_envSize = size > 0 ? size : 0; }

OR_API int or_render_pt(const GiSceneDesc* desc, const GiCamera* cam, int w, int h, const GiLightSample* ls, int nls, uint32_t frameSeed, int maxLen,
                        const float env_rgb[3], const OrRegion* region, float* accum, OrCounters* counters)
{
	EnvCube cubeStorage = {g_envFaces, g_envSize};
	const EnvCube* cube = g_envFaces ? &cubeStorage : nullptr;
	Scene s = makeScene(desc);
	Frame f = makeFrame(s, cam, w, h, ls, nls, frameSeed, maxLen, accum);
	V3 env = env_rgb ? v3(env_rgb) : v3(0.0f);
	int gw = w / 8, gh = h / 8;
	int x0 = 0, y0 = 0, x1 = gw * 8, y1 = gh * 8;
	if (region && (region->x1 > region->x0)) { x0 = std::max(x0, region->x0); y0 = std::max(y0, region->y0); x1 = std::min(x1, region->x1); y1 = std::min(y1, region->y1); }
	int tx0 = x0 / 8, ty0 = y0 / 8, tx1 = (x1 + 7) / 8, ty1 = (y1 + 7) / 8;
	int ntx = tx1 - tx0, nty = ty1 - ty0;
#pragma omp parallel
	{
		Counters c; memset(&c, 0, sizeof(c)); TraceCounters tc = {0, 0};
#pragma omp for schedule(dynamic, 1)
		for (int t = 0; t < ntx * nty; ++t)
		{
			int tx = tx0 + t % ntx, ty = ty0 + t / ntx;
			for (int ly = 0; ly < 8; ++ly)
				for (int lx = 0; lx < 8; ++lx)
				{
					int gx = tx * 8 + lx, gy = ty * 8 + ly;
					if (gx < x0 || gx >= x1 || gy < y0 || gy >= y1) continue;
					PathtracerPixel(f, gx, gy, env, c, tc, cube);
				}
		}
		addCounters(counters, c, tc);
	}
	return 0;
}

// whittedraytracer.comp:16-133 (compiled with STOP_ON_DIFFUSE_BOUNCE, whittedraytracer.cpp:38-49), one invocation
void WhittedPixel(const Frame& f, int gx, int gy, Counters& c, TraceCounters& tc)
{
	const int NUM_LIGHT_SAMPLES = 2;
	const float lightSampleWeight = float(f.numLightSamples) / NUM_LIGHT_SAMPLES;
	const Scene& s = *f.scene;
	uint32_t randomSeed = wanghash(f.frameSeed); // InitCoherentRandomSeed, random.glsl:23-28: the same stream for every pixel
	V2 r = Random2(randomSeed);
	float sx = (r.x + float(gx)) / float(f.width) * 2.0f - 1.0f;
	float sy = (r.y + float(gy)) / float(f.height) * 2.0f - 1.0f;
	Ray ray;
	ray.Origin = v3(f.cam.position);
	ray.Direction = normalize(sx * v3(f.cam.u) + sy * v3(f.cam.v) + v3(f.cam.w));
	++c.camera_paths;
	V3 color = v3(0.0f), pathThroughput = v3(1.0f);
	for (int i = 0; i < f.maxPathLength; ++i)
	{
		GiTriangle triangle; V3 bary, geometryNormal;
		float rayHit = RAY_MAX;
		int hitIndex[2];
		++c.closest_rays;
		TraceRay(s, ray, rayHit, hitIndex, bary, triangle, geometryNormal, &tc);
		if (rayHit == RAY_MAX) break;
		ray.Origin = ray.Origin + rayHit * ray.Direction;
		V3 hitNormal; V2 hitTexcoord;
		GetTriangleHitInfo(s, triangle, bary, hitNormal, hitTexcoord);
		MaterialData md = SampleMaterialData(s, int(triangle.material), hitTexcoord);
		++c.shaded_vertices;
		color += pathThroughput * md.Emissivity;
		// GetDiffuseProbability, materials.glsl:96-101 (constant reflection: preflect = Fresnel0)
		V3 pdiffuse = (v3(1.0f) - saturate(md.Fresnel0)) * md.Opacity;
		float pdiffuseAvg = GetLuminance(pdiffuse);
		V3 throughput = v3(1.0f);
		float pdf; bool dummy;
		V3 sampleDirection = SampleBSDFImpl(ray.Direction, md, randomSeed, hitNormal, throughput, false, pdf, dummy, true);
		if (throughput.x == -1.0f)
		{
			for (int lightSample = 0; lightSample < NUM_LIGHT_SAMPLES; ++lightSample)
			{
				int sampleIndex = int(RandomUInt(randomSeed) % uint32_t(f.numLightSamples));
				V3 radiance = EstimateDirectLight(f, ray.Origin, hitNormal, sampleIndex, ray.Direction, md, true, c, tc, true);
				color += radiance * pathThroughput * (lightSampleWeight / (pdiffuseAvg + DIVISOR_EPSILON));
			}
			break;
		}
		ray.Direction = sampleDirection;
		ray.Origin += ray.Direction * RAY_HIT_EPSILON;
		pathThroughput *= throughput / (GetLuminance(throughput) + DIVISOR_EPSILON);
	}
	float* px = f.accum + 4 * (size_t(gy) * f.width + gx);
	px[0] += color.x; px[1] += color.y; px[2] += color.z; px[3] += 1.0f; // vec4(color, 1.0) (:130-132)
}

// WhittedRayTracer::Draw (whittedraytracer.cpp:31-36): (W/8) x (H/8) groups of 8x8
OR_API int or_render_whitted(const GiSceneDesc* desc, const GiCamera* cam, int w, int h, const GiLightSample* ls, int nls, uint32_t frameSeed, int maxLen,
                             float* accum, OrCounters* counters)
{
	Scene s = makeScene(desc);
	Frame f = makeFrame(s, cam, w, h, ls, nls, frameSeed, maxLen, accum);
	const int gw = w / 8, gh = h / 8;
#pragma omp parallel
	{
		Counters c; memset(&c, 0, sizeof(c)); TraceCounters tc = {0, 0};
#pragma omp for schedule(dynamic, 1)
		for (int t = 0; t < gw * gh; ++t)
			for (int l = 0; l < 64; ++l) WhittedPixel(f, (t % gw) * 8 + (l & 7), (t / gw) * 8 + (l >> 3), c, tc);
		addCounters(counters, c, tc);
	}
	return 0;
}

// LightPathtracer::Draw (lightpathtracer.cpp:45-51): light paths [first_path, first_path+num_paths)
// of the NumInitialLightSamples*NumRaysPerLightSample dispatched; num_paths == 0 => all.
OR_API int or_render_lpt(const GiSceneDesc* desc, const GiCamera* cam, int w, int h, const GiLightSample* ls, int nls, uint32_t frameSeed, int maxLen,
                         int numRaysPerLightSample, uint32_t first_path, uint32_t num_paths, int show_light_caches, float* accum, OrCounters* counters)
{
	Scene s = makeScene(desc);
	Frame f = makeFrame(s, cam, w, h, ls, nls, frameSeed, maxLen, accum);
	uint32_t total = uint32_t(nls) * uint32_t(numRaysPerLightSample);
	if (num_paths == 0) { first_path = 0; num_paths = total; }
	uint32_t end = std::min(total, first_path + num_paths);
#pragma omp parallel
	{
		Counters c; memset(&c, 0, sizeof(c)); TraceCounters tc = {0, 0};
#pragma omp for schedule(dynamic, 64)
		for (int64_t gid = first_path; gid < int64_t(end); ++gid)
			LighttracerPath(f, uint32_t(gid), numRaysPerLightSample, nullptr, c, tc);
		if (show_light_caches)
		{
#pragma omp for schedule(static)
			for (int i = 0; i < nls; ++i)
				ShowLightCache(f, i, c, tc);
		}
		addCounters(counters, c, tc);
	}
	return 0;
}

// BidirectionalPathtracer::Draw (bidirectionalpathtracer.cpp:182-201): resetcounter, light pass with
// SAVE_LIGHT_CACHE, eye pass.  warmup != 0 runs only the SAVE_LIGHT_CACHE_WARMUP light pass (:90-93).
// cache: capacity * 64 bytes scratch (may be NULL for warm-up).  out_counts[0] = NumLightCacheEntries,
// out_counts[1] = LightPathLengthSum.
static int renderBpt(const GiSceneDesc* desc, const GiCamera* cam, int w, int h, const GiLightSample* ls, int nls, uint32_t frameSeed, int maxLen,
                     int numRaysPerLightSample, uint32_t capacity, void* cache, int warmup, const OrRegion* region, float* accum, uint32_t out_counts[2],
                     OrCounters* counters, const void* eyeCache, uint32_t eyeCacheEntries);

OR_API int or_render_bpt(const GiSceneDesc* desc, const GiCamera* cam, int w, int h, const GiLightSample* ls, int nls, uint32_t frameSeed, int maxLen,
                         int numRaysPerLightSample, uint32_t capacity, void* cache, int warmup, const OrRegion* region, float* accum, uint32_t out_counts[2],
                         OrCounters* counters)
{
	return renderBpt(desc, cam, w, h, ls, nls, frameSeed, maxLen, numRaysPerLightSample, capacity, cache, warmup, region, accum, out_counts, counters, nullptr, 0u);
}

// The same iteration, but the eye pass connects to a light cache GIVEN by the caller (capacity x 64 bytes in LightCacheEntry layout, lightcache.glsl:1-16,
// + the NumLightCacheEntries counter that goes with it) instead of the one the light pass above just wrote.  The light pass still runs (its
// camera-connection splats belong to the image).  Which slot a light vertex lands in depends on the scheduling of the atomic counter
// (lighttracer.comp:23-45), so two implementations only produce the same eye pass when they read the same cache: this is the deterministic
// half of the BPT parity test.
OR_API int or_render_bpt_with_cache(const GiSceneDesc* desc, const GiCamera* cam, int w, int h, const GiLightSample* ls, int nls, uint32_t frameSeed, int maxLen,
                                    int numRaysPerLightSample, uint32_t capacity, void* cache, const void* eye_cache, uint32_t eye_cache_entries,
                                    const OrRegion* region, float* accum, uint32_t out_counts[2], OrCounters* counters)
{
	return renderBpt(desc, cam, w, h, ls, nls, frameSeed, maxLen, numRaysPerLightSample, capacity, cache, 0, region, accum, out_counts, counters, eye_cache, eye_cache_entries);
}

static int renderBpt(const GiSceneDesc* desc, const GiCamera* cam, int w, int h, const GiLightSample* ls, int nls, uint32_t frameSeed, int maxLen,
                     int numRaysPerLightSample, uint32_t capacity, void* cache, int warmup, const OrRegion* region, float* accum, uint32_t out_counts[2],
                     OrCounters* counters, const void* eyeCache, uint32_t eyeCacheEntries)
{
	Scene s = makeScene(desc);
	Frame f = makeFrame(s, cam, w, h, ls, nls, frameSeed, maxLen, accum);
	LightCache lc;
	lc.entries = static_cast<LightCacheEntry*>(cache);
	lc.capacity = warmup ? 0x7fffffffu : capacity; // bidirectionalpathtracer.cpp:39 (INT_MAX until known)
	lc.warmup = warmup != 0;
	// resetcounter.comp:10-11
	lc.numEntries = uint32_t(numRaysPerLightSample) * uint32_t(nls);
	lc.pathLengthSum = 0;
	uint32_t total = uint32_t(nls) * uint32_t(numRaysPerLightSample);
#pragma omp parallel
	{
		Counters c; memset(&c, 0, sizeof(c)); TraceCounters tc = {0, 0};
#pragma omp for schedule(dynamic, 64)
		for (int64_t gid = 0; gid < int64_t(total); ++gid)
			LighttracerPath(f, uint32_t(gid), numRaysPerLightSample, &lc, c, tc);
		addCounters(counters, c, tc);
	}
	if (out_counts) { out_counts[0] = lc.numEntries; out_counts[1] = lc.pathLengthSum; }
	if (warmup) return 0;
	if (eyeCache)
	{
		lc.entries = static_cast<LightCacheEntry*>(const_cast<void*>(eyeCache));
		lc.numEntries = eyeCacheEntries;
	}
	int gw = w / 8, gh = h / 8;
	int x0 = 0, y0 = 0, x1 = gw * 8, y1 = gh * 8;
	if (region && (region->x1 > region->x0)) { x0 = std::max(x0, region->x0); y0 = std::max(y0, region->y0); x1 = std::min(x1, region->x1); y1 = std::min(y1, region->y1); }
	int tx0 = x0 / 8, ty0 = y0 / 8, tx1 = (x1 + 7) / 8, ty1 = (y1 + 7) / 8;
	int ntx = tx1 - tx0, nty = ty1 - ty0;
#pragma omp parallel
	{
		Counters c; memset(&c, 0, sizeof(c)); TraceCounters tc = {0, 0};
#pragma omp for schedule(dynamic, 1)
		for (int t = 0; t < ntx * nty; ++t)
		{
			int tx = tx0 + t % ntx, ty = ty0 + t / ntx;
			for (int ly = 0; ly < 8; ++ly)
				for (int lx = 0; lx < 8; ++lx)
				{
					int gx = tx * 8 + lx, gy = ty * 8 + ly;
					if (gx < x0 || gx >= x1 || gy < y0 || gy >= y1) continue;
					BidirPathtracerPixel(f, gx, gy, lc, c, tc);
				}
		}
		addCounters(counters, c, tc);
	}
	return 0;
}

// ---- host routines -------------------------------------------------------------------------
// Camera::ComputeCameraParams (camera.cpp:76-85) + PixelArea (renderersystem.cpp:208).
OR_API void or_camera(const float pos[3], const float lookat[3], const float up[3], float yfov_deg, float aspect, int w, int h, GiCamera* out)
{
	V3 W = einormalize(v3(lookat) - v3(pos));
	V3 U = einormalize(cross(W, v3(up)));
	V3 V = einormalize(cross(U, W));
	float f = tanf(yfov_deg / 2.0f * (3.14159265358979323846f / 180.0f));
	U *= f * aspect;
	V *= f;
	out->position[0] = pos[0]; out->position[1] = pos[1]; out->position[2] = pos[2];
	out->u[0] = U.x; out->u[1] = U.y; out->u[2] = U.z;
	out->v[0] = V.x; out->v[1] = V.y; out->v[2] = V.z;
	out->w[0] = W.x; out->w[1] = W.y; out->w[2] = W.z;
	out->pixel_area = length(U) * length(V) * 4.0f / float(w * h);
}

// Scene::LoadLightSources (scene.cpp:244-293) + ComputePointLightTable (:295-307).
// triangles: num_tris x 3 vertex indices (the chunk's triangle list, not the leaf list);
// emissivity: per material, constant emissivity (0 for textured emissivity, scene.cpp:173-177).
OR_API void* or_light_sampler_create(const float* positions, const uint32_t* triangles, const uint32_t* tri_materials, uint32_t num_tris, const float* emissivity,
                                     const float* point_lights /* n x 6: pos, intensity */, uint32_t num_point_lights)
{
	LightSamplerState* st = new LightSamplerState();
	st->randomSeed = 123; // lightsampler.cpp:8
	st->totalAreaLightFlux = 0.0f; st->totalPointLightFlux = 0.0f; st->lightAreaSum = 0.0f;
	float sum = 0.0f;
	for (uint32_t i = 0; i < num_tris; ++i)
	{
		const uint32_t* tri = triangles + 3 * size_t(i);
		V3 e = v3(emissivity + 3 * size_t(tri_materials[i]));
		if (tri[0] != tri[1] && (e.x != 0.0f || e.y != 0.0f || e.z != 0.0f))
		{
			LightTriangle lt;
			lt.luminance = e;
			lt.v0 = v3(positions + 3 * size_t(tri[0])); lt.v1 = v3(positions + 3 * size_t(tri[1])); lt.v2 = v3(positions + 3 * size_t(tri[2]));
			st->lightTriangles.push_back(lt);
			float area = length(cross(lt.v1 - lt.v0, lt.v2 - lt.v0)) * 0.5f; // ei::surface(Triangle)
			st->totalAreaLightFlux += dot(v3(0.2126f, 0.7152f, 0.0722f), lt.luminance) * area * 3.14159265358979323846f;
			sum += area;
			st->lightSummedArea.push_back(sum);
		}
	}
	if (!st->lightSummedArea.empty())
	{
		st->lightAreaSum = st->lightSummedArea.back();
		for (size_t i = 0; i < st->lightSummedArea.size(); ++i) st->lightSummedArea[i] /= st->lightAreaSum;
	}
	for (uint32_t i = 0; i < num_point_lights; ++i)
	{
		PointLight pl; pl.position = v3(point_lights + 6 * size_t(i)); pl.intensity = v3(point_lights + 6 * size_t(i) + 3);
		st->pointLights.push_back(pl);
	}
	for (size_t i = 0; i < st->pointLights.size(); ++i)
	{
		float flux = dot(v3(0.2126f, 0.7152f, 0.0722f), st->pointLights[i].intensity) * 4 * 3.14159265358979323846f;
		st->totalPointLightFlux += logf(1.0f + flux);
		st->pointLightSummedFlux.push_back(st->totalPointLightFlux);
	}
	for (size_t i = 0; i < st->pointLightSummedFlux.size(); ++i) st->pointLightSummedFlux[i] /= st->totalPointLightFlux;
	return st;
}
OR_API void or_light_sampler_destroy(void* h) { delete static_cast<LightSamplerState*>(h); }
OR_API void or_light_sampler_seed(void* h, uint64_t seed) { static_cast<LightSamplerState*>(h)->randomSeed = seed; }
OR_API uint32_t or_light_sampler_num_light_triangles(void* h) { return uint32_t(static_cast<LightSamplerState*>(h)->lightTriangles.size()); }

// LightSampler::GenerateRandomSamples (lightsampler.cpp:18-95)
OR_API void or_light_sampler_generate(void* h, GiLightSample* dst, uint32_t numSamples, float positionBias)
{
	LightSamplerState& st = *static_cast<LightSamplerState*>(h);
	if (st.totalAreaLightFlux + st.totalPointLightFlux == 0.0f)
	{
		memset(dst, 0, sizeof(GiLightSample) * numSamples);
		return;
	}
	float pointLightPercentage = st.totalPointLightFlux / (st.totalAreaLightFlux + st.totalPointLightFlux);
	unsigned numPointSamples = unsigned(pointLightPercentage * numSamples);
	unsigned numAreaSamples = numSamples - numPointSamples;
	float sampleWeight = st.lightAreaSum / numAreaSamples;
	for (unsigned i = 0; i < numAreaSamples; ++i, ++dst)
	{
		float randomTriangle = XorshiftF(st.randomSeed);
		const float* tab = st.lightSummedArea.data();
		const float* it = std::lower_bound(tab, tab + st.lightTriangles.size(), randomTriangle);
		unsigned triangleIdx = unsigned(it - tab);
		float xi0 = XorshiftF(st.randomSeed);
		float xi1 = XorshiftF(st.randomSeed);
		xi0 = sqrtf(xi0);
		float alpha = (1.0f - xi0);
		float beta = xi0 * (1.0f - xi1);
		float gamma = xi0 * xi1;
		const LightTriangle& lt = st.lightTriangles[triangleIdx];
		V3 position = lt.v0 * alpha + lt.v1 * beta + lt.v2 * gamma;
		V3 normal = einormalize(cross(lt.v0 - lt.v1, lt.v0 - lt.v2));
		position += normal * positionBias;
		dst->position[0] = position.x; dst->position[1] = position.y; dst->position[2] = position.z;
		dst->normal_phi = atan2f(normal.y, normal.x);
		V3 inten = lt.luminance * sampleWeight;
		dst->intensity[0] = inten.x; dst->intensity[1] = inten.y; dst->intensity[2] = inten.z;
		dst->normal_theta_cos = normal.z;
	}
	for (unsigned i = 0; i < numPointSamples; ++i, ++dst)
	{
		float randomPointLight = XorshiftF(st.randomSeed);
		const float* tab = st.pointLightSummedFlux.data();
		const float* it = std::lower_bound(tab, tab + st.pointLights.size(), randomPointLight);
		unsigned lightIdx = unsigned(it - tab);
		sampleWeight = *it;
		if (lightIdx > 0) sampleWeight -= tab[lightIdx - 1];
		const PointLight& pl = st.pointLights[lightIdx];
		dst->position[0] = pl.position.x; dst->position[1] = pl.position.y; dst->position[2] = pl.position.z;
		dst->normal_phi = 0.0f;
		dst->normal_theta_cos = 2.0f;
		V3 inten = pl.intensity / (numPointSamples * sampleWeight);
		dst->intensity[0] = inten.x; dst->intensity[1] = inten.y; dst->intensity[2] = inten.z;
	}
}

// TextureMSE::ComputeCurrentMSE (texturemse.cpp:39-66, sqerror.frag:16-17)
OR_API double or_mse(const float* accum_rgba, const float* ref_rgb, int w, int h, uint32_t divisor)
{
	double sum = 0.0;
	for (size_t i = 0; i < size_t(w) * h; ++i)
	{
		float d0 = accum_rgba[4 * i] / float(divisor) - ref_rgb[3 * i];
		float d1 = accum_rgba[4 * i + 1] / float(divisor) - ref_rgb[3 * i + 1];
		float d2 = accum_rgba[4 * i + 2] / float(divisor) - ref_rgb[3 * i + 2];
		sum += double(d0 * d0 + d1 * d1 + d2 * d2);
	}
	return sum / (double(w) * h);
}
