// gi_emul.cpp -- TEST INFRASTRUCTURE ONLY: compiles the product's device logic
// (gpugi_b200/csrc/gi_core.h + gi_layout.h) for the host so that `-m "not gpu"` tests can
// single-step exactly the code the CUDA kernels run and compare it with the oracle without a GPU.
// It is NOT a fallback: nothing in gpugi_b200/ links or loads this file, and the C ABI
// (libgpugi_b200.so) fails without a device.
#include "../../gpugi_b200/csrc/gi_core.h"
#include "../../gpugi_b200/csrc/gi_layout.h"
#include "../../gpugi_b200/csrc/gi_lane.h"

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

bool makeScene(const GiSceneDesc* d, EmScene& e, std::string& err)
{
	if (!buildHostLayout(d, e.L, err)) return false;
	e.s = e.L.scene;
	e.s.refNodes = e.L.refNodes.data();
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

// Conservativeness of the quantised inner-node test (gi_lane.h: wfSlabTest) against the exact-zone test on the real child
// box (intersectBox): for every ray and every `stride`-th inner node, both children.  out[0] = pairs tested,
// out[1] = pairs the exact test accepts, out[2] = pairs the conservative test accepts, out[3] = VIOLATIONS (exact accepts
// with first <= limit but the conservative test rejects, or its entry bound exceeds the exact entry distance),
// out[4] = rays outside the domain of the bound (answered by the reference walk).
EM_API int em_slab_check(const GiSceneDesc* desc, const GiRay* rays, uint32_t n, uint32_t stride, uint64_t* out)
{
	EmScene e; std::string err;
	if (!makeScene(desc, e, err)) return 1;
	if (!e.L.binary) return 2;
	for (int i = 0; i < 5; ++i) out[i] = 0;
	std::vector<uint32_t> inner;
	for (uint32_t i = 0; i < desc->num_nodes; ++i)
		if (!(desc->nodes[i].first_child & 0x80000000u)) inner.push_back(i);
	for (uint32_t r = 0; r < n; ++r)
	{
		WfLane L;
		const F3 o = f3(rays[r].origin), d = f3(rays[r].direction);
		const F3 inv = invDir(d);
		if (!wfSlabSetup(L, e.s, o, inv)) { ++out[4]; continue; }
		L.limit = rays[r].t_max;
		for (size_t fi = r % stride; fi < inner.size(); fi += stride)
		{
			const GiNode& nd = desc->nodes[inner[fi]];
			const uint32_t kids[2] = {nd.first_child, desc->nodes[nd.first_child].escape};
			const U8 q = ld8(e.s.qNodes + 8 * fi);
			for (int c = 0; c < 2; ++c)
			{
				const GiNode& ch = desc->nodes[kids[c]];
				float first, last, nearLow;
				const bool exact = intersectBox(o, inv, ch.min[0], ch.min[1], ch.min[2], ch.max[0], ch.max[1], ch.max[2], first, last) && first <= L.limit;
				const bool cons = wfSlabTest(L, q.v[3 * c], q.v[3 * c + 1], q.v[3 * c + 2], nearLow);
				++out[0];
				if (exact) ++out[1];
				if (cons) ++out[2];
				if (exact && (!cons || nearLow > first)) ++out[3];
			}
		}
	}
	return 0;
}
