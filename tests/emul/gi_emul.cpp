// gi_emul.cpp -- TEST INFRASTRUCTURE ONLY: compiles the product's device logic
// (gpugi_b200/csrc/gi_core.h + gi_layout.h) for the host so that `-m "not gpu"` tests can
// single-step exactly the code the CUDA kernels run and compare it with the oracle without a GPU.
// It is NOT a fallback: nothing in gpugi_b200/ links or loads this file, and the C ABI
// (libgpugi_b200.so) fails without a device.
#include "../../gpugi_b200/csrc/gi_core.h"
#include "../../gpugi_b200/csrc/gi_layout.h"
#include "../../gpugi_b200/csrc/gi_lane.h"
#include "../../gpugi_b200/csrc/gi_shade.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

using namespace gik;

#define EM_API extern "C" __attribute__((visibility("default")))

namespace {
struct EmScene
{
	HostLayout L;
	DScene s;
};

// em_set_topology: 0 = the reference's topology, 1 = median split of the leaf centres, 2 = random balanced pairing (a terrible tree: the
// exactness of the ordered walks must not depend on tree quality).  Host stand-ins for the device's PLOC pass, fed to the same
// applyLeafTree encoder gi_upload_scene uses.
int g_topology = 0;
uint32_t g_topologySeed = 1;
int g_topologyApplied = 0;

LeafTree g_userTree;

uint32_t emRand(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

void uniteBoxes(LeafTree& t, uint32_t id, uint32_t a, uint32_t b)
{
	for (int k = 0; k < 3; ++k)
	{
		t.boxMin[size_t(id) * 3 + k] = std::fmin(t.boxMin[size_t(a) * 3 + k], t.boxMin[size_t(b) * 3 + k]);
		t.boxMax[size_t(id) * 3 + k] = std::fmax(t.boxMax[size_t(a) * 3 + k], t.boxMax[size_t(b) * 3 + k]);
	}
	t.left[id] = a; t.right[id] = b;
}

// median split over positions [lo, hi) of the (reordered) leaf list; returns the node id of the subtree
uint32_t medianSplit(LeafTree& t, std::vector<uint32_t>& pos, uint32_t lo, uint32_t hi, uint32_t& nextInner)
{
	if (hi - lo == 1) return pos[lo];
	float mn[3] = {GI_INF, GI_INF, GI_INF}, mx[3] = {-GI_INF, -GI_INF, -GI_INF};
	auto centre = [&](uint32_t p, int k) { return 0.5f * t.boxMin[size_t(p) * 3 + k] + 0.5f * t.boxMax[size_t(p) * 3 + k]; };
	for (uint32_t i = lo; i < hi; ++i)
		for (int k = 0; k < 3; ++k) { mn[k] = std::fmin(mn[k], centre(pos[i], k)); mx[k] = std::fmax(mx[k], centre(pos[i], k)); }
	int axis = 0;
	if (mx[1] - mn[1] > mx[axis] - mn[axis]) axis = 1;
	if (mx[2] - mn[2] > mx[axis] - mn[axis]) axis = 2;
	const uint32_t mid = lo + (hi - lo) / 2;
	std::nth_element(pos.begin() + lo, pos.begin() + mid, pos.begin() + hi, [&](uint32_t a, uint32_t b) { return centre(a, axis) < centre(b, axis); });
	const uint32_t a = medianSplit(t, pos, lo, mid, nextInner), b = medianSplit(t, pos, mid, hi, nextInner);
	const uint32_t id = nextInner++;
	uniteBoxes(t, id, a, b);
	return id;
}

bool applyTestTopology(const GiSceneDesc* d, HostLayout& L)
{
	const uint32_t nl = d->num_leaf_triangles / 2;
	if (!L.binary || nl < 2) return false;
	const uint32_t N = 2 * nl - 1;
	LeafTree t;
	t.numLeaves = nl; t.left.assign(N, 0); t.right.assign(N, 0); t.boxMin.assign(size_t(N) * 3, 0.0f); t.boxMax.assign(size_t(N) * 3, 0.0f); t.leafOf.resize(nl);
	uint32_t seed = g_topologySeed ? g_topologySeed : 1u;
	for (uint32_t p = 0; p < nl; ++p) t.leafOf[p] = p;
	for (uint32_t p = nl - 1; p > 0; --p) std::swap(t.leafOf[p], t.leafOf[emRand(seed) % (p + 1)]);
	for (uint32_t p = 0; p < nl; ++p)
	{
		const GiNode& ln = d->nodes[L.leafNode[t.leafOf[p]]];
		std::memcpy(&t.boxMin[size_t(p) * 3], ln.min, 12); std::memcpy(&t.boxMax[size_t(p) * 3], ln.max, 12);
	}
	uint32_t nextInner = nl;
	if (g_topology == 3)
	{
		// tree handed in by em_set_tree (experiments with builders in Python): leaf positions as given, boxes bottom-up
		if (g_userTree.numLeaves != nl) return false;
		t.leafOf = g_userTree.leafOf; t.left = g_userTree.left; t.right = g_userTree.right; t.root = g_userTree.root;
		for (uint32_t p = 0; p < nl; ++p)
		{
			if (t.leafOf[p] >= nl) return false;
			const GiNode& ln = d->nodes[L.leafNode[t.leafOf[p]]];
			std::memcpy(&t.boxMin[size_t(p) * 3], ln.min, 12); std::memcpy(&t.boxMax[size_t(p) * 3], ln.max, 12);
		}
		std::vector<std::pair<uint32_t, int>> stack; stack.push_back({t.root, 0});
		size_t guard = 0;
		while (!stack.empty())
		{
			auto [x, state] = stack.back(); stack.pop_back();
			if (x >= N || ++guard > size_t(4) * N) return false;
			if (x < nl) continue;
			if (state == 0) { stack.push_back({x, 1}); stack.push_back({t.left[x], 0}); stack.push_back({t.right[x], 0}); }
			else uniteBoxes(t, x, t.left[x], t.right[x]);
		}
	}
	else if (g_topology == 1)
	{
		std::vector<uint32_t> pos(nl);
		for (uint32_t p = 0; p < nl; ++p) pos[p] = p;
		t.root = medianSplit(t, pos, 0, nl, nextInner);
	}
	else
	{
		std::vector<uint32_t> cur(nl), nxt;
		for (uint32_t p = 0; p < nl; ++p) cur[p] = p;
		while (cur.size() > 1)
		{
			nxt.clear();
			for (size_t i = 0; i + 1 < cur.size(); i += 2) { const uint32_t id = nextInner++; uniteBoxes(t, id, cur[i], cur[i + 1]); nxt.push_back(id); }
			if (cur.size() & 1) nxt.push_back(cur.back());
			cur.swap(nxt);
		}
		t.root = cur[0];
	}
	std::string why;
	return applyLeafTree(d, L, t, why);
}

bool makeScene(const GiSceneDesc* d, EmScene& e, std::string& err, bool keepChildBoxes = false)
{
	e.L.keepChildBoxes = keepChildBoxes;
	if (!buildHostLayout(d, e.L, err)) return false;
	g_topologyApplied = 0;
	if (g_topology != 0 && applyTestTopology(d, e.L)) g_topologyApplied = 1;
	e.s = e.L.scene;
	e.s.refNodes = e.L.refNodes.data();
	e.s.oboxNodes = e.L.oboxNodes.empty() ? nullptr : e.L.oboxNodes.data();
	e.s.fastNodes = e.L.fastNodes.data();
	e.s.qNodes = e.L.qNodes.data();
	e.s.leafPos = e.L.leafPos.data();
	e.s.leafTris = e.L.leafTris.data();
	e.s.vertexInfo = e.L.vertexInfo.data();
	e.s.materials = e.L.materials.data();
	e.s.texels = e.L.texels.data();
	e.s.textureFilterExact = 1;
	return true;
}
}

EM_API void em_set_topology(int kind, uint32_t seed) { g_topology = kind; g_topologySeed = seed; }
EM_API void em_set_tree(uint32_t num_leaves, const uint32_t* left, const uint32_t* right, const uint32_t* leaf_of, uint32_t root)
{
	const uint32_t N = 2 * num_leaves - 1;
	g_userTree = LeafTree();
	g_userTree.numLeaves = num_leaves; g_userTree.root = root;
	g_userTree.left.assign(left, left + N); g_userTree.right.assign(right, right + N); g_userTree.leafOf.assign(leaf_of, leaf_of + num_leaves);
	g_topology = 3;
}
EM_API int em_topology_applied(void) { return g_topologyApplied; }

// the shader #define variants of the frame (GI_OPT_RECONSTRUCTION_FILTER 0-3, GI_OPT_CHAINED_MIS) for every following em_render* call
static int g_reconFilter = 0, g_chainedMIS = 0;
EM_API void em_set_variant(int reconstruction_filter, int chained_mis) { g_reconFilter = reconstruction_filter; g_chainedMIS = chained_mis ? 1 : 0; }

struct EmCounters
{
	uint64_t closest_rays, any_rays, camera_paths, light_paths, shaded_vertices, splats, fallback_rays, node_tests, tri_tests;
	uint32_t light_cache_entries, light_path_length_sum;
};

static void addCounters(EmCounters* out, const PathCounters& pc, const TraceStats& st)
{
	if (!out) return;
	out->closest_rays += pc.closestRays; out->any_rays += pc.anyRays; out->camera_paths += pc.cameraPaths; out->light_paths += pc.lightPaths;
	out->shaded_vertices += pc.shadedVertices; out->splats += pc.splats; out->fallback_rays += pc.fallbackRays;
	out->node_tests += st.nodes; out->tri_tests += st.tris;
}

#define EM_MODE_LANE 2 // the wavefront lane state machine (gi_lane.h) instead of traceFast()

// k_trace, one "thread" per ray.  mode: GI_TRAVERSAL_* or EM_MODE_LANE.  Returns 0, or 1 if the scene is rejected.
EM_API int em_trace(const GiSceneDesc* desc, const GiRay* rays, uint32_t n, GiHit* hits, int mode, int* fast_valid, EmCounters* counters)
{
	EmScene e; std::string err;
	if (!makeScene(desc, e, err)) return 1;
	if (fast_valid) *fast_valid = e.L.binary ? 1 : 0;
	// bvhType = 1: GI_TRAVERSAL_FAST selects the ordered walk on the oriented boxes, hosted by the reference-order entry points (gi_device.cu: makeFrame)
	e.s.oboxOrdered = (mode == GI_TRAVERSAL_FAST && e.s.oboxNodes && e.s.refTreeBinary) ? 1 : 0;
	if (!e.L.binary) mode = GI_TRAVERSAL_REFERENCE;
	for (uint32_t i = 0; i < n; ++i)
	{
		PathCounters pc; std::memset(&pc, 0, sizeof(pc));
		TraceStats st = {0, 0};
		const F3 o = f3(rays[i].origin), d = f3(rays[i].direction);
		GiHit out;
		if (mode == EM_MODE_LANE)
		{
			WfCounters<true> cnt;
			WfLane L;
			const bool isAny = (rays[i].flags & 1u) != 0u;
			wfWalkSingle(L, e.s, o, d, rays[i].t_max, isAny, cnt);
			if (isAny) { ++pc.anyRays; out.t = rays[i].t_max; out.leaf_tri = (L.flags & WFL_OCCLUDED) ? 1 : 0; out.bary_y = 0.0f; out.bary_z = 0.0f; }
			else
			{
				++pc.closestRays;
				out.t = L.best; out.leaf_tri = L.hitLeaf; out.bary_y = 0.0f; out.bary_z = 0.0f;
				if (L.hitLeaf >= 0)
				{
					F3 p0, p1, p2; int valid;
					loadLeafTri(e.s, L.hitLeaf, p0, p1, p2, valid);
					float tt; F3 bary, nn;
					intersectTriangle(o, d, p0, p1, p2, tt, bary, nn);
					out.bary_y = bary.y; out.bary_z = bary.z;
				}
			}
			pc.splats += cnt.v[WFC_SPLATS]; pc.fallbackRays += cnt.v[WFC_FALLBACK]; st.nodes += cnt.v[WFC_NODES]; st.tris += cnt.v[WFC_TRIS];
		}
		else if (rays[i].flags & 1u)
		{
			++pc.anyRays;
			bool occ = traceAny(e.s, mode, o, d, rays[i].t_max, &st);
			out.t = rays[i].t_max; out.leaf_tri = occ ? 1 : 0; out.bary_y = 0.0f; out.bary_z = 0.0f;
		}
		else
		{
			++pc.closestRays;
			Hit h; bool fb = false;
			traceClosest(e.s, mode, o, d, rays[i].t_max, h, &st, &fb);
			if (fb) ++pc.fallbackRays;
			out.t = h.t; out.leaf_tri = h.leafTri; out.bary_y = h.leafTri >= 0 ? h.bary.y : 0.0f; out.bary_z = h.leafTri >= 0 ? h.bary.z : 0.0f;
		}
		hits[i] = out;
		addCounters(counters, pc, st);
	}
	return 0;
}

// gi_render_iteration (+ gi_show_light_caches when show != 0) executed thread by thread.
EM_API int em_render(const GiSceneDesc* desc, const GiCamera* cam, int w, int h, const GiLightSample* ls, int nls, const GiIterationParams* p, int mode,
                     const float env_rgb[3], int show_light_caches, void* light_cache, float* accum, EmCounters* counters)
{
	EmScene e; std::string err;
	if (!makeScene(desc, e, err)) return 1;
	if (!e.L.binary) mode = GI_TRAVERSAL_REFERENCE;
	uint32_t bptCounters[2] = {0, 0};
	DFrame f;
	std::memset(&f, 0, sizeof(f));
	f.reconFilter = g_reconFilter; f.chainedMIS = g_chainedMIS;
	f.scene = e.s; f.cam = *cam; f.width = w; f.height = h;
	f.lightSamples = reinterpret_cast<const F4*>(ls); f.numLightSamples = nls;
	f.frameSeed = p->frame_seed; f.maxPathLength = p->max_path_length; f.traversalMode = mode; f.accum = accum;
	f.envColor = env_rgb ? f3(env_rgb) : f3(0.0f); f.envCube = 0;
	f.numRaysPerLightSample = p->num_rays_per_light_sample;
	f.lightCache = static_cast<F4*>(light_cache);
	f.lightCacheCapacity = uint32_t(p->light_cache_capacity);
	f.lightCacheCount = &bptCounters[0]; f.lightPathLengthSum = &bptCounters[1];
	PathCounters pc; std::memset(&pc, 0, sizeof(pc));
	TraceStats st = {0, 0};
	const int gw = w / 8, gh = h / 8;
	const uint32_t numLightPaths = uint32_t(nls) * uint32_t(p->num_rays_per_light_sample);
	switch (p->renderer)
	{
	case GI_RENDERER_PT:
		for (int by = 0; by < gh; ++by) for (int bx = 0; bx < gw; ++bx) for (int t = 0; t < 64; ++t)
			pathtracePixel(f, bx * 8 + (t & 7), by * 8 + (t >> 3), pc, &st);
		break;
	case GI_RENDERER_WHITTED:
		for (int by = 0; by < gh; ++by) for (int bx = 0; bx < gw; ++bx) for (int t = 0; t < 64; ++t)
			whittedPixel(f, bx * 8 + (t & 7), by * 8 + (t >> 3), pc, &st);
		break;
	case GI_RENDERER_LPT:
		for (uint32_t gid = 0; gid < (numLightPaths / 64u) * 64u; ++gid) lighttracePath<false, false>(f, gid, pc, &st);
		break;
	case GI_RENDERER_BPT:
		bptCounters[0] = numLightPaths; bptCounters[1] = 0;
		if (p->warmup)
		{
			f.lightCacheCapacity = 0x7fffffffu;
			for (uint32_t gid = 0; gid < (numLightPaths / 64u) * 64u; ++gid) lighttracePath<true, true>(f, gid, pc, &st);
			break;
		}
		for (uint32_t gid = 0; gid < (numLightPaths / 64u) * 64u; ++gid) lighttracePath<true, false>(f, gid, pc, &st);
		for (int by = 0; by < gh; ++by) for (int bx = 0; bx < gw; ++bx) for (int t = 0; t < 64; ++t)
			bidirPathtracePixel(f, bx * 8 + (t & 7), by * 8 + (t >> 3), pc, &st);
		break;
	default: return 2;
	}
	if (show_light_caches)
		for (int i = 0; i < nls; ++i) showLightCache(f, i, pc, &st);
	addCounters(counters, pc, st);
	if (counters) { counters->light_cache_entries = bptCounters[0]; counters->light_path_length_sum = bptCounters[1]; }
	return 0;
}

// ---- the wavefront schedule of the light tracer / BPT (gi_wavefront.cuh) executed sequentially: the same lane functions
// (gi_shade.h), queues as vectors, one wave after the other.  phase: 0 = whole iteration, 1 = BPT light pass only,
// 2 = BPT eye pass only (light_cache / cache_count are inputs then).  cache_count: in/out NumLightCacheEntries.
namespace {
struct EmRay { F3 o, d; int bounce; uint32_t id; };
struct EmShadow { F3 o, d; float len; F3 v; uint32_t target; bool splat; float posX, posY; };

void emTraceRays(const DFrame& f, const std::vector<EmRay>& rays, std::vector<float>& t, std::vector<int>& leaf, PathCounters& pc, TraceStats& st)
{
	t.resize(rays.size()); leaf.resize(rays.size());
	for (size_t i = 0; i < rays.size(); ++i)
	{
		Hit h; bool fb = false;
		traceClosest(f.scene, f.traversalMode, rays[i].o, rays[i].d, GI_RAY_MAX, h, &st, &fb);
		if (fb) ++pc.fallbackRays;
		t[i] = h.t; leaf[i] = h.leafTri;
	}
}

void emTraceShadows(const DFrame& f, const std::vector<EmShadow>& sh, std::vector<F4>& stateC, PathCounters& pc, TraceStats& st)
{
	for (const EmShadow& r : sh)
	{
		if (traceAny(f.scene, f.traversalMode, r.o, r.d, r.len, &st)) continue;
		if (r.splat && f.reconFilter)
		{
			// wfFinishShadow (gi_wavefront.cuh): 5 x 5 weighted splats around the unrounded position carried beside the shadow record
			++pc.splats;
			PathCounters dummy; std::memset(&dummy, 0, sizeof(dummy));
			splatFiltered(f.accum, f.width, f.height, f.reconFilter, r.posX, r.posY, r.v, dummy);
		}
		else if (r.splat) { ++pc.splats; float* p = f.accum + 4 * size_t(r.target); p[0] += r.v.x; p[1] += r.v.y; p[2] += r.v.z; }
		else { F4& c = stateC[r.target]; c.x += r.v.x; c.y += r.v.y; c.z += r.v.z; }
	}
}
}

EM_API int em_render_wavefront(const GiSceneDesc* desc, const GiCamera* cam, int w, int h, const GiLightSample* ls, int nls, const GiIterationParams* p, int mode,
                               int phase, void* light_cache, uint32_t* cache_count, float* accum, EmCounters* counters)
{
	EmScene e; std::string err;
	if (!makeScene(desc, e, err)) return 1;
	if (!e.L.binary) mode = GI_TRAVERSAL_REFERENCE;
	uint32_t bptCounters[2] = {cache_count ? *cache_count : 0u, 0u};
	DFrame f;
	std::memset(&f, 0, sizeof(f));
	f.reconFilter = g_reconFilter; f.chainedMIS = g_chainedMIS;
	f.scene = e.s; f.cam = *cam; f.width = w; f.height = h;
	f.lightSamples = reinterpret_cast<const F4*>(ls); f.numLightSamples = nls;
	f.frameSeed = p->frame_seed; f.maxPathLength = p->max_path_length; f.traversalMode = mode; f.accum = accum;
	f.envColor = f3(0.0f); f.envCube = 0;
	f.numRaysPerLightSample = p->num_rays_per_light_sample;
	f.lightCache = static_cast<F4*>(light_cache);
	f.lightCacheCapacity = uint32_t(p->light_cache_capacity);
	f.lightCacheCount = &bptCounters[0]; f.lightPathLengthSum = &bptCounters[1];
	PathCounters pc; std::memset(&pc, 0, sizeof(pc));
	TraceStats st = {0, 0};
	const uint32_t numLightPaths = (uint32_t(nls) * uint32_t(p->num_rays_per_light_sample) / 64u) * 64u;
	const bool bpt = p->renderer == GI_RENDERER_BPT;
	if (p->renderer != GI_RENDERER_LPT && !bpt) return 2;
	std::vector<float> t; std::vector<int> leaf;
	std::vector<F4> noC;
	if (phase != 2)
	{
		// ---- light pass
		uint32_t flags = bpt ? uint32_t(WF_LIGHT_SAVE) : 0u;
		if (bpt && p->warmup) { flags |= WF_LIGHT_WARMUP; f.lightCacheCapacity = 0x7fffffffu; }
		if (bpt) { bptCounters[0] = uint32_t(nls) * uint32_t(p->num_rays_per_light_sample); bptCounters[1] = 0; }
		std::vector<F4> stateT(numLightPaths);
		std::vector<EmRay> cur(numLightPaths), next;
		std::vector<EmShadow> shCur, shNext;
		for (uint32_t gid = 0; gid < numLightPaths; ++gid)
		{
			wfGenerateLightLane(f, flags, gid, cur[gid].o, cur[gid].d, stateT[gid], pc);
			++pc.closestRays;
			cur[gid].bounce = 0; cur[gid].id = gid;
		}
		for (int wave = 0; wave <= f.maxPathLength; ++wave)
		{
			emTraceRays(f, cur, t, leaf, pc, st);
			emTraceShadows(f, shCur, noC, pc, st);
			next.clear(); shNext.clear();
			for (size_t i = 0; i < cur.size(); ++i)
			{
				WfPush out;
				wfShadeLightLane(f, flags, cur[i].o, cur[i].d, cur[i].bounce, t[i], leaf[i], stateT[cur[i].id], out, pc);
				if (out.shadow) shNext.push_back(EmShadow{out.sOrigin, out.sDir, out.sLen, out.sValue, uint32_t(out.py) * uint32_t(w) + uint32_t(out.px), true, out.posX, out.posY});
				if (out.ray) next.push_back(EmRay{out.origin, out.dir, cur[i].bounce + 1, cur[i].id});
			}
			cur.swap(next); shCur.swap(shNext);
		}
		if (!cur.empty() || !shCur.empty()) return 3; // the schedule must drain within maxPathLength + 1 waves
	}
	if (bpt && !p->warmup && phase != 1)
	{
		// ---- eye pass
		const int gw = w / 8, gh = h / 8;
		const uint32_t numIds = uint32_t(w) * uint32_t(h);
		std::vector<F4> stateT(numIds), stateC(numIds);
		std::vector<float> stateM(numIds, 1.0f);
		std::memset(stateC.data(), 0, numIds * sizeof(F4));
		std::vector<EmRay> cur, next;
		std::vector<EmShadow> shCur, shNext;
		for (int by = 0; by < gh; ++by) for (int bx = 0; bx < gw; ++bx) for (int l = 0; l < 64; ++l)
		{
			const int gx = bx * 8 + (l & 7), gy = by * 8 + (l >> 3);
			const uint32_t pixel = uint32_t(gx + gy * w);
			uint32_t seed = initRandomSeed(f.frameSeed, pixel);
			const F3 dir = cameraRayDirection(f, seed, gx, gy);
			cur.push_back(EmRay{f3(f.cam.position), dir, 0, pixel});
			++pc.cameraPaths; ++pc.closestRays;
		}
		const uint32_t numEntries = bptCounters[0];
		for (int wave = 0; wave <= f.maxPathLength; ++wave)
		{
			emTraceRays(f, cur, t, leaf, pc, st);
			emTraceShadows(f, shCur, stateC, pc, st);
			next.clear(); shNext.clear();
			for (size_t i = 0; i < cur.size(); ++i)
			{
				WfPush out;
				const uint32_t id = cur[i].id;
				wfShadeEyeLane(f, numEntries, cur[i].o, cur[i].d, cur[i].bounce, id, t[i], leaf[i], stateT[id], stateC[id], stateM[id], out, pc);
				if (out.shadow) shNext.push_back(EmShadow{out.sOrigin, out.sDir, out.sLen, out.sValue, id, false, 0.0f, 0.0f});
				if (out.ray) next.push_back(EmRay{out.origin, out.dir, cur[i].bounce + 1, id});
			}
			cur.swap(next); shCur.swap(shNext);
		}
		if (!cur.empty() || !shCur.empty()) return 3;
		for (uint32_t i = 0; i < numIds; ++i) { accum[4 * size_t(i)] += stateC[i].x; accum[4 * size_t(i) + 1] += stateC[i].y; accum[4 * size_t(i) + 2] += stateC[i].z; accum[4 * size_t(i) + 3] += stateC[i].w; }
	}
	addCounters(counters, pc, st);
	if (counters) { counters->light_cache_entries = bptCounters[0]; counters->light_path_length_sum = bptCounters[1]; }
	if (cache_count) *cache_count = bptCounters[0];
	return 0;
}

// k_bidir_eye alone on a given light cache (the megakernel form), for a like-for-like check of the eye waves.
EM_API int em_bidir_eye(const GiSceneDesc* desc, const GiCamera* cam, int w, int h, const GiIterationParams* p, int mode, void* light_cache, uint32_t cache_count,
                        float* accum, EmCounters* counters)
{
	EmScene e; std::string err;
	if (!makeScene(desc, e, err)) return 1;
	if (!e.L.binary) mode = GI_TRAVERSAL_REFERENCE;
	uint32_t bptCounters[2] = {cache_count, 0u};
	DFrame f;
	std::memset(&f, 0, sizeof(f));
	f.reconFilter = g_reconFilter; f.chainedMIS = g_chainedMIS;
	f.scene = e.s; f.cam = *cam; f.width = w; f.height = h;
	f.frameSeed = p->frame_seed; f.maxPathLength = p->max_path_length; f.traversalMode = mode; f.accum = accum;
	f.envColor = f3(0.0f);
	f.numRaysPerLightSample = p->num_rays_per_light_sample;
	f.lightCache = static_cast<F4*>(light_cache);
	f.lightCacheCapacity = uint32_t(p->light_cache_capacity);
	f.lightCacheCount = &bptCounters[0]; f.lightPathLengthSum = &bptCounters[1];
	PathCounters pc; std::memset(&pc, 0, sizeof(pc));
	TraceStats st = {0, 0};
	for (int by = 0; by < h / 8; ++by) for (int bx = 0; bx < w / 8; ++bx) for (int t = 0; t < 64; ++t)
		bidirPathtracePixel(f, bx * 8 + (t & 7), by * 8 + (t >> 3), pc, &st);
	addCounters(counters, pc, st);
	return 0;
}

// Conservativeness of the quantised inner-node test (gi_lane.h: wfSlabTest) against the exact-zone test on the real child
// box (intersectBox): for every ray and every `stride`-th inner node, both children.  out[0] = pairs tested,
// out[1] = pairs the exact test accepts, out[2] = pairs the conservative test accepts, out[3] = VIOLATIONS (exact accepts
// with first <= limit but the conservative test rejects, or its entry bound exceeds the exact entry distance),
// out[4] = rays outside the domain of the bound (answered by the reference walk).
EM_API int em_slab_check(const GiSceneDesc* desc, const GiRay* rays, uint32_t n, uint32_t stride, uint64_t* out)
{
	EmScene e; std::string err;
	if (!makeScene(desc, e, err, true)) return 1;
	if (!e.L.binary) return 2;
	for (int i = 0; i < 5; ++i) out[i] = 0;
	// every child box the layout encoded (binary or 4-wide nodes, reference or re-clustered topology), with its exact floats
	const std::vector<HostLayout::ChildBox>& boxes = e.L.childBoxes;
	for (uint32_t r = 0; r < n; ++r)
	{
		WfLane L;
		const F3 o = f3(rays[r].origin), d = f3(rays[r].direction);
		const F3 inv = invDir(d);
#if WF_CACHE_INV
		L.inv = inv;
#endif
		if (!wfSlabSetup(L, e.s, o, inv)) { ++out[4]; continue; }
		L.limit = rays[r].t_max;
		for (size_t bi = r % stride; bi < boxes.size(); bi += stride)
		{
			const HostLayout::ChildBox& cb = boxes[bi];
			float first, last, nearLow;
			const bool exact = intersectBox(o, inv, cb.mn[0], cb.mn[1], cb.mn[2], cb.mx[0], cb.mx[1], cb.mx[2], first, last) && first <= L.limit;
			const bool cons = wfSlabTestChild(L, e.s.qNodes + cb.word, int(cb.child), nearLow);
			++out[0];
			if (exact) ++out[1];
			if (cons) ++out[2];
			if (exact && (!cons || nearLow > first)) ++out[3];
		}
	}
	return 0;
}
